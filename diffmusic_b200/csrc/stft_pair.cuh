// Frame-PAIR STFT guidance pipeline: a group of 64 threads carries TWO 1024-sample frames (A, B) through the same
// FFT -> |X|^2 -> sparse mel -> dB/clamp -> residual -> VJP -> inverse FFT chain as stft_frame.cuh, with the two frames
// interleaved element by element in shared memory:
//
//   cell i = (Re zA[i], Im zA[i], Re zB[i], Im zB[i])      one 128-bit shared-memory access moves both frames
//
// Thread j owns butterfly j of BOTH frames in every pass, so the twiddles, the swizzled addresses, the real-FFT
// unpack/pack factors and the loop control are paid once per two frames, and every FFT load / store is 128 bits wide
// (a quarter of the LDS/STS instructions of the split re/im layout).  The forward spectrum X of the bins a thread owns
// (k = j + 64 i and 512 - k) stays in REGISTERS between the forward unpack and the backward pack: only the per-bin
// energy goes through shared memory (for the band-parallel mel projection) -- X is never stored.
//
// Bank conflicts: a 128-bit access is served per quarter-warp (8 lanes x 16 B = all 32 banks).  With the cell swizzle
// sw4(i) = i ^ ((i >> 3) & 7) every access pattern of the radix-8 Stockham passes hits 8 distinct 16-byte bank groups
// per quarter-warp (checked exhaustively in tests/cpu_emul), and the swizzled addresses reduce to `base + constant`
// (loads, last-pass stores, unpack / pack) or `base + (x ^ q)` (first two passes' stores).
//
// Same arithmetic as stft_frame.cuh (reference: diffmusic/inverse_problem/operator.py:24-36,123-124,153-154,162-171 via
// torchaudio functional.spectrogram / MelScale / amplitude_to_DB); everything is __host__ __device__ phase code that
// tests/cpu_emul runs thread by thread on the host.
#pragma once
#include "stft_frame.cuh"

namespace dm {

struct alignas(16) c2 {  // one cell: element of frame A and of frame B
    float ax, ay, bx, by;
};

DM_HD int sw4(int i) { return i ^ ((i >> 3) & 7); }

// Shared memory of one group: two cell buffers (ping-pong), nothing else.  Between the forward unpack and the first
// inverse pass `b` is dead as an FFT buffer and holds, as f2 = (frame A, frame B):
//   [0, 513)    per-bin energies P[k], natural order
//   [520, 592)  mel cotangent melbar[72] (entries 64.. are zero: a bin's second band may be "band 64")
//   [600, 664)  partial mel sums handed from warp 0 to warp 1
struct PairSmem {
    c2* a;  // [512]
    c2* b;  // [512]
};
constexpr int kPairSmemFloats = 2 * 4 * kH;

DM_HD f2* pair_energy(const PairSmem& s) { return reinterpret_cast<f2*>(s.b); }
DM_HD f2* pair_melbar(const PairSmem& s) { return reinterpret_cast<f2*>(s.b) + 520; }
DM_HD f2* pair_scratch(const PairSmem& s) { return reinterpret_cast<f2*>(s.b) + 600; }

// Per-thread constants (identical for every frame pair the thread will process)
struct PairConsts {
    cf w8[7];   // pass NS = 8 twiddles
    cf w64[7];  // pass NS = 64 twiddles
    cf wu[4];   // exp(-2 pi i k / 1024) for the owned pairs k = j + 64 i
    // mel projection, balanced over the two warps of a group (band lengths grow from 3 to 41 bins): thread j sums
    // weights [0, mel_n) of its own band j; threads j < 32 additionally sum the tail [x_i0, x_i0 + x_n) of band 63 - j,
    // whose owner (thread 63 - j, in the other warp) only sums the head.
    int mel_k0, mel_n;
    int x_k0, x_i0, x_n;
};
DM_HD int mel_head_len(int n) { return (n + 1) >> 1; }
DM_HD void load_pair_consts(int j, const StftTables& t, PairConsts& c) {
    thread_twiddles<kH, 8>(j, t.tw512, c.w8);
    thread_twiddles<kH, 64>(j, t.tw512, c.w64);
#pragma unroll
    for (int i = 0; i < 4; ++i) c.wu[i] = t.w1024[j + 64 * i];
    c.mel_k0 = t.mel_kstart[j];
    c.mel_n = t.mel_klen[j];
    c.x_k0 = c.x_i0 = c.x_n = 0;
    if (j >= 32) {
        c.mel_n = mel_head_len(c.mel_n);
    } else {
        const int e = 63 - j, n = t.mel_klen[e];
        c.x_i0 = mel_head_len(n);
        c.x_n = n - c.x_i0;
        c.x_k0 = t.mel_kstart[e] + c.x_i0;
    }
}

// Forward spectrum (later: its cotangent factor) of the bins a thread owns, for both frames [.][0] = A, [.][1] = B.
// lo[i] = bin j + 64 i, hi[i] = bin 512 - (j + 64 i).  Thread 0, i = 0: lo = (X[0], 0), hi = (X[512], 0), q = X[256].
struct PairX {
    cf lo[4][2], hi[4][2], q[2];
};

// ---- 8 cells in / out of a swizzled buffer --------------------------------------------------------------------------
DM_HD void load8_c2(const c2* __restrict__ buf, int j, cf (&va)[8], cf (&vb)[8]) {
    const c2* p = buf + sw4(j);  // sw4(j + 64 r) = sw4(j) + 64 r
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const c2 c = p[64 * r];
        va[r] = cf{c.ax, c.ay};
        vb[r] = cf{c.bx, c.by};
    }
}
// logical output index of butterfly j, slot q:  NS = 1: 8 j + q ; NS = 8: 64 (j >> 3) + (j & 7) + 8 q ; NS = 64: j + 64 q
template <int NS>
DM_HD int st_c2_addr(int j, int q) {
    if (NS == 1) return 8 * j + ((j & 7) ^ q);                    // sw4(8 j + q): the mask is j & 7
    if (NS == 8) return 64 * (j >> 3) + 8 * q + ((j & 7) ^ q);    // sw4(64 a + k + 8 q): the mask is q
    return sw4(j) + 64 * q;
}
template <int NS>
DM_HD void store8_c2(c2* __restrict__ buf, int j, const cf (&va)[8], const cf (&vb)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) buf[st_c2_addr<NS>(j, q)] = c2{va[q].x, va[q].y, vb[q].x, vb[q].y};
}

// ---- forward FFT of both frames ---------------------------------------------------------------------------------------
// fa / fb: the two (already masked) 1024-sample frames in shared memory, 8-byte aligned (even hop).
DM_HD void pair_fwd_pass1(int j, const float* fa, const float* fb, const float* window, PairSmem s) {
    cf va[8], vb[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = j + r * (kH / 8);
        const f2 w = reinterpret_cast<const f2*>(window)[i];
        const f2 a = reinterpret_cast<const f2*>(fa)[i];
        const f2 b = reinterpret_cast<const f2*>(fb)[i];
        va[r] = cf{a.x * w.x, a.y * w.y};
        vb[r] = cf{b.x * w.x, b.y * w.y};
    }
    dft8<-1>(va);
    dft8<-1>(vb);
    store8_c2<1>(s.a, j, va, vb);
}
template <int NS, int SIGN>
DM_HD void pair_pass(int j, const cf (&w)[7], const c2* in, c2* out) {
    cf va[8], vb[8];
    load8_c2(in, j, va, vb);
    twiddle8<SIGN>(va, w);
    twiddle8<SIGN>(vb, w);
    dft8<SIGN>(va);
    dft8<SIGN>(vb);
    store8_c2<NS>(out, j, va, vb);
}
DM_HD void pair_fwd_pass2(int j, const PairConsts& c, PairSmem s) { pair_pass<8, -1>(j, c.w8, s.a, s.b); }
DM_HD void pair_fwd_pass3(int j, const PairConsts& c, PairSmem s) { pair_pass<64, -1>(j, c.w64, s.b, s.a); }

// 1/sqrt(e): one MUFU.RSQ on the device (<= 2 ulp, far inside the 1e-4 parity bound) instead of the ~20-instruction
// IEEE sqrt + divide sequences -- the phase-retrieval modes take a magnitude and a reciprocal magnitude per bin and frame
DM_HD float fast_rsqrt(float e) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(e);
#else
    return 1.0f / sqrtf(e);
#endif
}
template <int MODE>
DM_HD float pair_bin_energy(cf x) {
    const float e = x.x * x.x + x.y * x.y;
    if (MODE == kModeMelDb) return e;
    return e == 0.f ? 0.f : e * fast_rsqrt(e);  // |X|; NaN stays NaN
}

// ---- unpack Z (in a) -> X of the owned bins (registers) and their energies (-> f2 P[513] in b) ----------------------
template <int MODE>
DM_HD void pair_unpack(int j, const PairConsts& c, PairSmem s, PairX& x) {
    f2* P = pair_energy(s);
    if (j < 8) pair_melbar(s)[64 + j] = f2{0.f, 0.f};
    const c2* lo = s.a + sw4(j);                // cell of k = j + 64 i     : lo[64 i]
    const c2* hi = s.a + sw4((kH - j) & (kH - 1));  // cell of 512 - k, j >= 1  : hi[-64 i]  (j = 0: 512 - 64 i, i >= 1)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = j + 64 * i;
        if (k == 0) {
            const c2 z0 = s.a[0], zq = s.a[kH / 2];  // sw4(0) = 0, sw4(256) = 256
            x.lo[0][0] = cf{z0.ax + z0.ay, 0.f};
            x.hi[0][0] = cf{z0.ax - z0.ay, 0.f};
            x.lo[0][1] = cf{z0.bx + z0.by, 0.f};
            x.hi[0][1] = cf{z0.bx - z0.by, 0.f};
            x.q[0] = cf{zq.ax, -zq.ay};
            x.q[1] = cf{zq.bx, -zq.by};
            P[0] = f2{pair_bin_energy<MODE>(x.lo[0][0]), pair_bin_energy<MODE>(x.lo[0][1])};
            P[kH] = f2{pair_bin_energy<MODE>(x.hi[0][0]), pair_bin_energy<MODE>(x.hi[0][1])};
            P[kH / 2] = f2{pair_bin_energy<MODE>(x.q[0]), pair_bin_energy<MODE>(x.q[1])};
        } else {
            const c2 zk = lo[64 * i];
            const c2 zc = (j == 0) ? s.a[sw4(kH - k)] : hi[-64 * i];
            rfft_unpack_pair(cf{zk.ax, zk.ay}, cf{zc.ax, zc.ay}, c.wu[i], x.lo[i][0], x.hi[i][0]);
            rfft_unpack_pair(cf{zk.bx, zk.by}, cf{zc.bx, zc.by}, c.wu[i], x.lo[i][1], x.hi[i][1]);
            P[k] = f2{pair_bin_energy<MODE>(x.lo[i][0]), pair_bin_energy<MODE>(x.lo[i][1])};
            P[kH - k] = f2{pair_bin_energy<MODE>(x.hi[i][0]), pair_bin_energy<MODE>(x.hi[i][1])};
        }
    }
}

// ---- mel projection of both frames for band m (threads 0..63), then dB / clamp and the derivative -----------------
// partial sums of both frames over weights [i0, i0 + n) of band m, bins starting at k0 (k0 already includes i0)
DM_HD f2 pair_mel_segment(const float* __restrict__ melw_t, const f2* __restrict__ P, int m, int k0, int i0, int n) {
    const f2* p = P + k0;
    const float* w = melw_t + i0 * kMels + m;
    float a = 0.f, b = 0.f;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const float wi = w[i * kMels];
        const f2 v = p[i];
        a = fmaf(wi, v.x, a);
        b = fmaf(wi, v.y, b);
    }
    return f2{a, b};
}
// phase 1 (all 64 threads): own band (head only for j >= 32); j < 32 also leave the tail of band 63 - j in scratch[63 - j]
DM_HD f2 pair_mel_project(int j, const PairConsts& c, const float* __restrict__ melw_t, const PairSmem& s) {
    const f2* P = pair_energy(s);
    f2* scratch = pair_scratch(s);
    const f2 own = pair_mel_segment(melw_t, P, j, c.mel_k0, 0, c.mel_n);
    if (j < 32) scratch[63 - j] = pair_mel_segment(melw_t, P, 63 - j, c.x_k0, c.x_i0, c.x_n);
    return own;
}
// phase 2 (after a group barrier): complete the long bands
DM_HD f2 pair_mel_combine(int j, f2 own, const PairSmem& s) {
    if (j >= 32) {
        const f2 t = pair_scratch(s)[j];
        own.x += t.x;
        own.y += t.y;
    }
    return own;
}
// transformed value (what operator.transform returns) and d value / d mel energy
template <int MODE>
DM_HD void mel_value(float acc, bool clamp, float& val, float& dval) {
    if (MODE == kModeMelDb) {
        const float c = acc < 1e-10f ? 1e-10f : acc;  // torch.clamp(min=amin): NaN stays NaN
        const float db = 10.0f * log10f(c);
        dval = (acc >= 1e-10f) ? (4.342944819032518f / c) : 0.f;  // 10 / ln(10) / mel
        val = db;
        if (clamp) {
            val = clamp_nan(db, -80.f, 80.f);
            if (!(db >= -80.f && db <= 80.f)) dval = 0.f;
        }
    } else {  // phase_mel: clamp(mel of magnitude, +-80), no log
        val = acc;
        dval = 1.f;
        if (clamp) {
            val = clamp_nan(acc, -80.f, 80.f);
            if (!(acc >= -80.f && acc <= 80.f)) dval = 0.f;
        }
    }
}

// ---- backward: cotangent of the owned bins -> Hermitian-packed Z cells (-> a) -----------------------------------------
template <int MODE>
DM_HD cf pair_xbar(cf x, float g) {
    float scale;
    if (MODE == kModeMelDb) {
        scale = 2.f * g;  // d|X|^2 = 2 X
    } else {
        const float e = x.x * x.x + x.y * x.y;
        scale = e > 0.f ? g * fast_rsqrt(e) : 0.f;  // d|X| = X / |X|, 0 at X = 0 (and for NaN, like `mag > 0 ? ... : 0`)
    }
    return cf{scale * x.x, scale * x.y};
}
// Per-bin rows of the filterbank (<= 2 bands touch a bin), packed for the pack phase: binw[k] = (fb[k, m0], fb[k, m0+1]),
// binm[k] = m0.  The kernel keeps one copy per CTA in shared memory (4.6 KB).
struct PairBinTab {
    const f2* binw;              // [513]
    const unsigned char* binm;   // [513]
};
// energy cotangent of bin k for both frames from the mel cotangent
DM_HD f2 pair_bin_cotangent(int k, const PairBinTab& t, const PairSmem& s) {
    const int m0 = t.binm[k];
    const f2 w = t.binw[k];
    const f2* melbar = pair_melbar(s);
    const f2 g0 = melbar[m0], g1 = melbar[m0 + 1];
    return f2{w.x * g0.x + w.y * g1.x, w.x * g0.y + w.y * g1.y};
}
template <int MODE>
DM_HD void pair_pack(int j, const PairConsts& c, const PairBinTab& t, PairSmem s, const PairX& x) {
    c2* lo = s.a + sw4(j);
    c2* hi = s.a + sw4((kH - j) & (kH - 1));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = j + 64 * i;
        if (k == 0) {
            f2 g0, gh, gq;
            if (MODE == kModePhaseWav) {  // P[] already holds the magnitude cotangent -(ref - |X|) (written by the caller)
                const f2* P = pair_energy(s);
                g0 = P[0];
                gh = P[kH];
                gq = P[kH / 2];
            } else {
                g0 = pair_bin_cotangent(0, t, s);
                gh = pair_bin_cotangent(kH, t, s);
                gq = pair_bin_cotangent(kH / 2, t, s);
            }
            // only the real part of the DC / Nyquist cotangent acts; Y[256] = Xbar / 2, Z[256] = 2 conj(Y)
            const float y0a = pair_xbar<MODE>(x.lo[0][0], g0.x).x, yha = pair_xbar<MODE>(x.hi[0][0], gh.x).x;
            const float y0b = pair_xbar<MODE>(x.lo[0][1], g0.y).x, yhb = pair_xbar<MODE>(x.hi[0][1], gh.y).x;
            s.a[0] = c2{y0a + yha, y0a - yha, y0b + yhb, y0b - yhb};
            const cf qa = pair_xbar<MODE>(x.q[0], gq.x), qb = pair_xbar<MODE>(x.q[1], gq.y);
            s.a[kH / 2] = c2{qa.x, -qa.y, qb.x, -qb.y};
        } else {
            f2 gk, gc;
            if (MODE == kModePhaseWav) {
                const f2* P = pair_energy(s);
                gk = P[k];
                gc = P[kH - k];
            } else {
                gk = pair_bin_cotangent(k, t, s);
                gc = pair_bin_cotangent(kH - k, t, s);
            }
            cf z[2][2];
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                cf yk = pair_xbar<MODE>(x.lo[i][f], f ? gk.y : gk.x);
                cf yc = pair_xbar<MODE>(x.hi[i][f], f ? gc.y : gc.x);
                yk = cf{0.5f * yk.x, 0.5f * yk.y};
                yc = cf{0.5f * yc.x, 0.5f * yc.y};
                irfft_pack_pair(yk, yc, c.wu[i], z[f][0], z[f][1]);
            }
            lo[64 * i] = c2{z[0][0].x, z[0][0].y, z[1][0].x, z[1][0].y};
            c2* pc = (j == 0) ? (s.a + sw4(kH - k)) : (hi - 64 * i);
            *pc = c2{z[0][1].x, z[0][1].y, z[1][1].x, z[1][1].y};
        }
    }
}

// ---- inverse STFT (torch.istft, istft.cu): the one-sided spectrum X of the owned bins (x, same ownership as the forward
// unpack) -> Hermitian-packed Z cells (-> a), scaled so that the unnormalised inverse passes return 1024 * irfft(X):
// interior bins enter as they are, only the real parts of DC / Nyquist act (as in irfft), Z[256] = 2 conj(X[256]).
DM_HD void pair_pack_spectrum(int j, const PairConsts& c, PairSmem s, const PairX& x) {
    c2* lo = s.a + sw4(j);
    c2* hi = s.a + sw4((kH - j) & (kH - 1));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = j + 64 * i;
        if (k == 0) {
            const float y0a = x.lo[0][0].x, yha = x.hi[0][0].x, y0b = x.lo[0][1].x, yhb = x.hi[0][1].x;
            s.a[0] = c2{y0a + yha, y0a - yha, y0b + yhb, y0b - yhb};
            s.a[kH / 2] = c2{2.f * x.q[0].x, -2.f * x.q[0].y, 2.f * x.q[1].x, -2.f * x.q[1].y};
        } else {
            cf z[2][2];
#pragma unroll
            for (int f = 0; f < 2; ++f) irfft_pack_pair(x.lo[i][f], x.hi[i][f], c.wu[i], z[f][0], z[f][1]);
            lo[64 * i] = c2{z[0][0].x, z[0][0].y, z[1][0].x, z[1][0].y};
            c2* pc = (j == 0) ? (s.a + sw4(kH - k)) : (hi - 64 * i);
            *pc = c2{z[0][1].x, z[0][1].y, z[1][1].x, z[1][1].y};
        }
    }
}
// Spectrum cells staged pair-major, spec[k] = (X_A[k], X_B[k]) for k = 0..512, into the owned-bin registers.
DM_HD void pair_load_spectrum(int j, const c2* __restrict__ spec, PairX& x) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = j + 64 * i;
        const c2 a = spec[k], b = spec[kH - k];  // k = 0: bins 0 and 512
        x.lo[i][0] = cf{a.ax, a.ay};
        x.lo[i][1] = cf{a.bx, a.by};
        x.hi[i][0] = cf{b.ax, b.ay};
        x.hi[i][1] = cf{b.bx, b.by};
    }
    const c2 q = spec[kH / 2];
    x.q[0] = cf{q.ax, q.ay};
    x.q[1] = cf{q.bx, q.by};
}
// rectangular-window overlap-add (torch.istft with window=None)
DM_HD void pair_ola_add_rect(int j, const cf (&v)[8], f2* acc2) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int h = j + 64 * q;
        f2 a = acc2[h];
        a.x += v[q].x;
        a.y += v[q].y;
        acc2[h] = a;
    }
}

// ---- inverse FFT of both frames: a -> b -> a -> registers, then windowed overlap-add from the registers ----
DM_HD void pair_inv_pass1(int j, PairSmem s) {
    cf va[8], vb[8];
    load8_c2(s.a, j, va, vb);
    dft8<+1>(va);
    dft8<+1>(vb);
    store8_c2<1>(s.b, j, va, vb);
}
DM_HD void pair_inv_pass2(int j, const PairConsts& c, PairSmem s) { pair_pass<8, +1>(j, c.w8, s.b, s.a); }
// last inverse pass, outputs kept in registers: va[q] / vb[q] = (unwindowed) frame-gradient samples 2h, 2h+1 of frame
// A / B for h = j + 64 q
DM_HD void pair_inv_pass3(int j, const PairConsts& c, PairSmem s, cf (&va)[8], cf (&vb)[8]) {
    load8_c2(s.a, j, va, vb);
    twiddle8<+1>(va, c.w64);
    twiddle8<+1>(vb, c.w64);
    dft8<+1>(va);
    dft8<+1>(vb);
}
// overlap-add of one frame straight from the registers: acc2[h] += v[q] * window (sample pairs, h = j + 64 q).
// The 64 threads of a group touch disjoint addresses; groups (and the two frames of a group) are serialised by the caller.
DM_HD void pair_ola_add(int j, const float* window, const cf (&v)[8], f2* acc2) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int h = j + 64 * q;
        const f2 w = reinterpret_cast<const f2*>(window)[h];
        f2 a = acc2[h];
        a.x = fmaf(v[q].x, w.x, a.x);
        a.y = fmaf(v[q].y, w.y, a.y);
        acc2[h] = a;
    }
}

}  // namespace dm
