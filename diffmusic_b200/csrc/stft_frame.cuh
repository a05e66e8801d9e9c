// Per-frame STFT guidance pipeline (one 1024-sample frame handled by a group of 64 threads), as host/device phases.
//
//   forward : windowed frame -> 512-pt complex FFT -> X[0..512] -> |X|^2 (or |X|) -> sparse mel -> dB / clamp
//   residual: d = ref - out ; sum d^2 ; Gbar = -d (the 1/loss factor is applied later, per clip)
//   backward: melbar -> Pbar[k] (<= 2 mel bands per bin) -> Xbar -> Hermitian pack -> 512-pt inverse FFT
//             -> frame gradient (still to be multiplied by the window and overlap-added by the caller)
//
// Arithmetic restated from torchaudio (functional.spectrogram / MelScale.forward / amplitude_to_DB) as used by
// diffmusic/inverse_problem/operator.py:24-36,123-124,153-154,162-171 (reference); VJP edge rules follow
// autograd: clamp passes gradient at equality, |0| has zero gradient.
#pragma once
#include <math.h>

#include "fft_core.cuh"

namespace dm {

constexpr int kNfft = 1024;
constexpr int kH = 512;     // complex FFT size
constexpr int kBins = 513;  // one-sided bins
constexpr int kMels = 64;
constexpr int kGroupThreads = 64;

// clamp that propagates NaN like torch.clamp (fminf/fmaxf would swallow it and hide a diverged sample from the
// pipelines' NaN guard, pipeline_musicldm.py:742)
DM_HD float clamp_nan(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

enum StftMode { kModeMelDb = 0, kModePhaseMel = 1, kModePhaseWav = 2 };

// Device/host constant tables (built on the host by diffmusic_b200/tables.py from torch / torchaudio tensors).
struct StftTables {
    const float* window;     // [1024] analysis window (hann periodic, or ones for phase retrieval)
    const cf* tw512;         // [512]  exp(-2 pi i m / 512)
    const cf* w1024;         // [257]  exp(-2 pi i k / 1024)
    const int* mel_kstart;   // [64]   first bin with non-zero weight
    const int* mel_klen;     // [64]   number of consecutive non-zero bins
    const float* mel_w;      // [mel_wstride][64] banded filterbank, transposed: mel_w[i * 64 + m] = fb[kstart[m] + i, m]
    int mel_wstride;
    const int* bin_m0;       // [513]  first mel band touching bin k
    const float* bin_w0;     // [513]  fb[k, m0]
    const float* bin_w1;     // [513]  fb[k, m0+1] (0 if none)
    // optional (warp-per-frame-pair kernel): host-built shared-memory image of its tables, diffmusic_b200/tables.py
    // warp_image(); na / nb = bin-pair rows of the two mel bands a lane sums
    const float* warp_image = nullptr;
    int warp_image_floats = 0, warp_na = 0, warp_nb = 0;
};

// Shared-memory working set of one frame group (all float): two swizzled complex buffers and the mel cotangent.
// The spectrum and the per-bin energy do not get arrays of their own, they live in the FFT buffers while those are dead:
//   X[k], k = 0..511 (forward spectrum, kept for the VJP) in b_re/b_im at the swizzled slot of k, with the two real
//        bins packed into slot 0 = (Re X[0], Re X[512]);
//   P[k] (|X|^2 or |X|; later the magnitude cotangent in phase_wav mode) in a_re at the swizzled slot of k, P[512] in aux.
// Both are written by the thread that has just read the same slots (unpack / pack work on the pair k, 512-k in place),
// so no other thread's data is clobbered.  8.8 KB per group instead of 15 KB -> 3 CTAs per SM.
struct FrameSmem {
    float* a_re;    // [512]
    float* a_im;
    float* b_re;
    float* b_im;
    float* melbar;  // [72] (index 64 must read as 0)
    float* aux;     // [8]  aux[0] = P[512]
};
constexpr int kFrameSmemFloats = 4 * swz_len(kH) + 72 + 8;

DM_HD float& p_at(const FrameSmem& s, int k) { return k == kH ? s.aux[0] : s.a_re[swz(k)]; }
DM_HD cf x_at(const FrameSmem& s, int k) {
    if (k == 0) return cf{s.b_re[swz(0)], 0.f};
    if (k == kH) return cf{s.b_im[swz(0)], 0.f};
    const int a = swz(k);
    return cf{s.b_re[a], s.b_im[a]};
}

// Per-thread constants of the frame pipeline (thread gt of a 64-thread group always owns the same twiddles and the
// same mel band), loaded once per CTA.
struct ThreadConsts {
    cf w8[7];    // pass NS = 8 twiddles
    cf w64[7];   // pass NS = 64 twiddles
    int mel_k0, mel_n;
};
DM_HD void load_thread_consts(int gt, const StftTables& t, ThreadConsts& c) {
    thread_twiddles<kH, 8>(gt, t.tw512, c.w8);
    thread_twiddles<kH, 64>(gt, t.tw512, c.w64);
    c.mel_k0 = t.mel_kstart[gt];
    c.mel_n = t.mel_klen[gt];
}

// ---- forward FFT: three Stockham passes. frame[] is the (already masked) signal tile in shared memory. ----
// ALIGNED8: frame and window are 8-byte aligned (even hop), so sample pairs are fetched with one 64-bit access.
template <bool ALIGNED8>
DM_HD void fwd_pass1(int tid, const float* frame, const float* window, FrameSmem s) {
    auto load = [&](int i) {
        if (ALIGNED8) {
            const f2 f = reinterpret_cast<const f2*>(frame)[i];
            const f2 w = reinterpret_cast<const f2*>(window)[i];
            return cf{f.x * w.x, f.y * w.y};
        }
        return cf{frame[2 * i] * window[2 * i], frame[2 * i + 1] * window[2 * i + 1]};
    };
    cf v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = load(tid + r * (kH / 8));
    dft8<-1>(v);
    store8_swz<kH, 1>(s.a_re, s.a_im, tid, v);
}
DM_HD void fwd_pass2(int tid, const ThreadConsts& c, FrameSmem s) {
    stockham_pass_swz<kH, 8, -1>(tid, c.w8, s.a_re, s.a_im, s.b_re, s.b_im);
}
DM_HD void fwd_pass3(int tid, const ThreadConsts& c, FrameSmem s) {
    stockham_pass_swz<kH, 64, -1>(tid, c.w64, s.b_re, s.b_im, s.a_re, s.a_im);
}

// ---- unpack Z (in a_re/a_im, natural order) to the real-FFT spectrum X (-> b) and the per-bin energy (-> a_re) ----
template <int MODE>
DM_HD void fwd_unpack(int tid, const cf* w1024, FrameSmem s) {
    SwzLoad Z{s.a_re, s.a_im};
    auto energy = [](cf x) {
        float e = x.x * x.x + x.y * x.y;
        return (MODE == kModeMelDb) ? e : sqrtf(e);
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = tid + 64 * i;  // 0..255
        if (k == 0) {
            const cf z0 = Z(0), zq = Z(kH / 2);
            const cf x0 = cf{z0.x + z0.y, 0.f}, xh = cf{z0.x - z0.y, 0.f}, xq = cconj(zq);
            s.b_re[swz(0)] = x0.x;  // slot 0 packs the two real bins
            s.b_im[swz(0)] = xh.x;
            s.b_re[swz(kH / 2)] = xq.x;
            s.b_im[swz(kH / 2)] = xq.y;
            s.a_re[swz(0)] = energy(x0);
            s.aux[0] = energy(xh);
            s.a_re[swz(kH / 2)] = energy(xq);
        } else {
            const int ak = swz(k), ac = swz(kH - k);
            cf xk, xc;
            rfft_unpack_pair(cf{s.a_re[ak], s.a_im[ak]}, cf{s.a_re[ac], s.a_im[ac]}, w1024[k], xk, xc);
            s.b_re[ak] = xk.x;
            s.b_im[ak] = xk.y;
            s.b_re[ac] = xc.x;
            s.b_im[ac] = xc.y;
            s.a_re[ak] = energy(xk);
            s.a_re[ac] = energy(xc);
        }
    }
}

// ---- mel projection + dB + clamp + residual (threads 0..63, one mel band each). Returns d^2 (0 if no ref). ----
// out_val receives the transformed value (what operator.transform returns); melbar gets dLoss_unscaled/dMel.
template <int MODE>
DM_HD float mel_residual(int m, const ThreadConsts& c, const float* melw_t, FrameSmem s, bool clamp, bool has_ref,
                         float ref_val, float* out_val) {
    // melw_t is the banded filterbank TRANSPOSED to [i][64] so the 64 band-threads read consecutive words
    const int k0 = c.mel_k0, n = c.mel_n;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < n; ++i) acc = fmaf(melw_t[i * kMels + m], s.a_re[swz(k0 + i)], acc);  // bin 512 has no weight
    float val, dval_dmel;  // transformed value and its derivative w.r.t. the mel energy
    if (MODE == kModeMelDb) {
        float c = acc < 1e-10f ? 1e-10f : acc;  // torch.clamp(min=amin): NaN stays NaN
        float db = 10.0f * log10f(c);
        dval_dmel = (acc >= 1e-10f) ? (4.342944819032518f / c) : 0.f;  // 10 / ln(10) / mel
        val = db;
        if (clamp) {
            val = clamp_nan(db, -80.f, 80.f);
            if (!(db >= -80.f && db <= 80.f)) dval_dmel = 0.f;
        }
    } else {  // phase_mel: clamp(mel of magnitude, +-80), no log
        val = acc;
        dval_dmel = 1.f;
        if (clamp) {
            val = clamp_nan(acc, -80.f, 80.f);
            if (!(acc >= -80.f && acc <= 80.f)) dval_dmel = 0.f;
        }
    }
    *out_val = val;
    float d = 0.f;
    if (has_ref) {
        d = ref_val - val;
        s.melbar[m] = -d * dval_dmel;
    }
    return d * d;
}

// ---- backward: spectrum cotangent -> Hermitian-packed Z for the inverse FFT (written to b_re/b_im) ----
// For MODE == kModePhaseWav, P[] must already hold magbar[k] = -(ref - |X|) (written by the caller through p_at).
// In place: a thread reads X at the slots of its pair (k, 512-k) from b and writes the packed Z to the same slots.
template <int MODE>
DM_HD cf xbar_of_bin(int k, const StftTables& t, const FrameSmem& s) {
    const cf x = x_at(s, k);
    float g;
    if (MODE == kModePhaseWav) {
        g = p_at(s, k);
    } else {
        int m0 = t.bin_m0[k];
        g = t.bin_w0[k] * s.melbar[m0] + t.bin_w1[k] * s.melbar[m0 + 1];
    }
    float scale;
    if (MODE == kModeMelDb) {
        scale = 2.f * g;  // d|X|^2 = 2 X
    } else {
        float mag = sqrtf(x.x * x.x + x.y * x.y);
        scale = mag > 0.f ? g / mag : 0.f;  // d|X| = X/|X|, 0 at X = 0
    }
    return cf{scale * x.x, scale * x.y};
}

template <int MODE>
DM_HD void bwd_pack(int tid, const StftTables& t, FrameSmem s) {
    SwzStore Zb{s.b_re, s.b_im};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = tid + 64 * i;  // 0..255
        if (k == 0) {
            float y0 = xbar_of_bin<MODE>(0, t, s).x;     // only the real part of the DC / Nyquist cotangent acts
            float yh = xbar_of_bin<MODE>(kH, t, s).x;
            Zb(0, cf{y0 + yh, y0 - yh});
            cf yq = xbar_of_bin<MODE>(kH / 2, t, s);     // Y[256] = Xbar/2 ; Z[256] = 2 conj(Y)
            Zb(kH / 2, cf{yq.x, -yq.y});
        } else {
            cf yk = xbar_of_bin<MODE>(k, t, s), yc = xbar_of_bin<MODE>(kH - k, t, s);
            yk = cf{0.5f * yk.x, 0.5f * yk.y};
            yc = cf{0.5f * yc.x, 0.5f * yc.y};
            cf zk, zc;
            irfft_pack_pair(yk, yc, t.w1024[k], zk, zc);
            Zb(k, zk);
            Zb(kH - k, zc);
        }
    }
}

// ---- inverse FFT: b -> a -> b -> a ; afterwards frame gradient n is a_re[pad(n/2)] (n even) / a_im (n odd) ----
DM_HD void inv_pass1(int tid, FrameSmem s) {
    cf v[8];
    load8_swz<kH>(s.b_re, s.b_im, tid, v);
    dft8<+1>(v);
    store8_swz<kH, 1>(s.a_re, s.a_im, tid, v);
}
DM_HD void inv_pass2(int tid, const ThreadConsts& c, FrameSmem s) {
    stockham_pass_swz<kH, 8, +1>(tid, c.w8, s.a_re, s.a_im, s.b_re, s.b_im);
}
DM_HD void inv_pass3(int tid, const ThreadConsts& c, FrameSmem s) {
    stockham_pass_swz<kH, 64, +1>(tid, c.w64, s.b_re, s.b_im, s.a_re, s.a_im);
}
DM_HD float frame_grad_sample(const FrameSmem& s, int n) {
    int p = swz(n >> 1);
    return (n & 1) ? s.a_im[p] : s.a_re[p];
}

// reflect padding (center=True, pad_mode="reflect", pad = 512): padded index -> source index
DM_HD long long reflect_src(long long i, long long len) {
    long long j = i - 512;
    if (j < 0) j = -j;
    if (j >= len) j = 2 * (len - 1) - j;
    return j;
}

}  // namespace dm
