// Overlap-save block of the dereverberation operator (host/device phases, see fft_core.cuh for the convention).
//
// Reference arithmetic (diffmusic/inverse_problem/operator.py:244-250): y[i] = sum_{k<K} xz[i+k] ir[k] with xz the
// input zero-padded by K/2 on both sides (F.conv1d = cross-correlation), i = 0 .. L + 2*(K/2) - K.
// VJP: xbar[j] = sum_k ybar[j + K/2 - k] ir[k].
//
// One block = one 8192-point real FFT done as a 4096-point complex FFT by 512 threads:
//   load 8192 input samples -> 4 Stockham passes -> per-bin multiply by conj(H) (correlation) or H (adjoint), done on
//   the packed spectrum pair (k, 4096-k) in registers -> 4 inverse passes -> the valid 8192-K+1 samples are stored.
#pragma once
#include "fft_core.cuh"

namespace dm {

constexpr int kRirN = 8192;      // real FFT length
constexpr int kRirH = 4096;      // complex FFT length
constexpr int kRirThreads = 512;

struct RirSmem {
    float* a_re;  // [swz_len(4096)]
    float* a_im;
    float* b_re;
    float* b_im;
};
constexpr int kRirSmemFloats = 4 * swz_len(kRirH);

// spectrum of the impulse response: ir zero-padded to 8192, H[k], k = 0..4096, from the packed FFT Z (in `z`)
DM_HD void rir_unpack_spectrum(int tid, SwzLoad Z, const cf* w8192, cf* spec) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = tid + kRirThreads * i;  // 0..2047
        if (k == 0) {
            cf z0 = Z(0);
            spec[0] = cf{z0.x + z0.y, 0.f};
            spec[kRirH] = cf{z0.x - z0.y, 0.f};
            spec[kRirH / 2] = cconj(Z(kRirH / 2));
        } else {
            cf xk, xc;
            rfft_unpack_pair(Z(k), Z(kRirH - k), w8192[k], xk, xc);
            spec[k] = xk;
            spec[kRirH - k] = xc;
        }
    }
}

// Pointwise spectral product on packed data: Zin (FFT of the packed real block) -> Zout (packed spectrum of the
// product, ready for the unnormalised inverse).  CONJ = true multiplies by conj(H) (cross-correlation).
template <bool CONJ>
DM_HD void rir_pointwise(int tid, SwzLoad Zin, SwzStore Zout, const cf* spec, const cf* w8192) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = tid + kRirThreads * i;  // 0..2047
        if (k == 0) {
            cf z0 = Zin(0);
            float y0 = (z0.x + z0.y) * spec[0].x;          // X[0], H[0] real
            float yh = (z0.x - z0.y) * spec[kRirH].x;      // X[4096], H[4096] real
            Zout(0, cf{y0 + yh, y0 - yh});
            cf hq = spec[kRirH / 2];
            if (CONJ) hq = cconj(hq);
            cf yq = cmul(cconj(Zin(kRirH / 2)), hq);       // X[2048] = conj(Z[2048])
            Zout(kRirH / 2, cf{2.f * yq.x, -2.f * yq.y});
        } else {
            cf xk, xc;
            rfft_unpack_pair(Zin(k), Zin(kRirH - k), w8192[k], xk, xc);
            cf hk = spec[k], hc = spec[kRirH - k];
            if (CONJ) {
                hk = cconj(hk);
                hc = cconj(hc);
            }
            cf zk, zc;
            irfft_pack_pair(cmul(xk, hk), cmul(xc, hc), w8192[k], zk, zc);
            Zout(k, zk);
            Zout(kRirH - k, zc);
        }
    }
}

// Input loader of the first pass: complex element i = (sample[2i], sample[2i+1]); `Src` maps a block-local sample
// index to a value (zero outside the signal).
template <class Src>
struct RirLoad {
    Src src;
    DM_HD cf operator()(int i) const { return cf{src(2 * i), src(2 * i + 1)}; }
};
// Output store of the last inverse pass: keeps samples shift <= s < shift + valid, scaled by 1/8192.
struct RirStore {
    float* out;        // already offset to the first output sample of this block (a 16-bit pointer when io != 0)
    int shift, valid;  // valid is clipped to the end of the signal by the caller
    float scale;
    int io = 0;        // DM_IO_* of the destination (the adjoint writes dLoss/dwav in the waveform's dtype)
    DM_HD void put(int s, float v) const {
        int m = s - shift;
        if (m < 0 || m >= valid) return;
#if defined(__CUDA_ARCH__)
        if (io != 0) {
            st_wave(out, io, m, v * scale);
            return;
        }
#endif
        out[m] = v * scale;
    }
    DM_HD void operator()(int i, cf c) const {
        put(2 * i, c.x);
        put(2 * i + 1, c.y);
    }
};

// The eight FFT passes + pointwise product of one block, phase by phase (ph = 0..8), for thread `tid`.
template <bool CONJ, class Src>
DM_HD void rir_block_phase(int ph, int tid, const cf* tw, const cf* w8192, const cf* spec, RirSmem s, Src src,
                           RirStore st) {
    SwzLoad la{s.a_re, s.a_im}, lb{s.b_re, s.b_im};
    SwzStore sa{s.a_re, s.a_im}, sb{s.b_re, s.b_im};
    switch (ph) {
        case 0: stockham_pass<kRirH, 1, -1>(tid, tw, RirLoad<Src>{src}, sa); break;
        case 1: stockham_pass_rec_swz<kRirH, 8, -1>(tid, tw, s.a_re, s.a_im, s.b_re, s.b_im); break;
        case 2: stockham_pass_rec_swz<kRirH, 64, -1>(tid, tw, s.b_re, s.b_im, s.a_re, s.a_im); break;
        case 3: stockham_pass_rec_swz<kRirH, 512, -1>(tid, tw, s.a_re, s.a_im, s.b_re, s.b_im); break;
        case 4: rir_pointwise<CONJ>(tid, lb, sa, spec, w8192); break;
        case 5: stockham_pass<kRirH, 1, +1>(tid, tw, la, sb); break;
        case 6: stockham_pass_rec_swz<kRirH, 8, +1>(tid, tw, s.b_re, s.b_im, s.a_re, s.a_im); break;
        case 7: stockham_pass_rec_swz<kRirH, 64, +1>(tid, tw, s.a_re, s.a_im, s.b_re, s.b_im); break;
        default: {  // last inverse pass: swizzled loads, recurrence twiddles, outputs go through the store functor
            cf v[8];
            load8_swz<kRirH>(s.b_re, s.b_im, tid, v);
            twiddle8_rec<kRirH, 512, +1>(tid, tw, v);
            dft8<+1>(v);
#pragma unroll
            for (int q = 0; q < 8; ++q) st(tid + q * 512, v[q]);
            break;
        }
    }
}
constexpr int kRirPhases = 9;

// block geometry shared by host emulation and the kernels
struct RirGeom {
    int K, pad, valid;  // taps, K/2, 8192 - K + 1
    long long L, nout;  // input length, output length L + 2*pad - K + 1
};
DM_HD RirGeom rir_geom(long long L, int K) {
    RirGeom g;
    g.K = K;
    g.pad = K / 2;
    g.valid = kRirN - K + 1;
    g.L = L;
    g.nout = L + 2 * (long long)g.pad - K + 1;
    return g;
}

}  // namespace dm
