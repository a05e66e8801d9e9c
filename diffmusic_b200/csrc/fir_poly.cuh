// Polyphase evaluation of torchaudio's sinc resampling FIR (integer decimation: new_freq/gcd == 1) and of its
// adjoint, as host/device per-thread bodies (validated on the CPU by tests/cpu_emul, see fft_core.cuh).
//
//   forward  y[j]    = sum_k  xz[orig*j + k] * w[k]                     xz = x zero-padded by `width` on the left
//   adjoint  xbar[i] = sum_j  ybar[j] * w[i + width - orig*j]           
//
// Splitting k = kappa + orig*m makes both sums sliding windows in m: a thread that owns R = 4 consecutive outputs of
// one phase reuses every staged sample and every weight four times (2 shared-memory loads per 4 FMAs instead of 8).
#pragma once
#include "fft_core.cuh"

namespace dm {

constexpr int kFirR = 4;  // outputs per thread
// staged-input index padding: thread t of the forward kernel reads xs[orig * 4 t + ...], a stride of 4*orig words;
// one extra word every 32 spreads that over all banks (same idea as swz() in fft_core.cuh)
DM_HD int fir_pad(int i) { return i + (i >> 5); }
DM_HDC int fir_padded_len(int n) { return n + (n >> 5) + 1; }

// ORIG / TAPS > 0 are compile-time specialisations (the reference's scale 2: orig 2, 28 taps; scale 10: orig 10, 132
// taps): the window loops unroll completely, the rotating weight registers and the tail selects disappear.
// ORIG = TAPS = 0 is the generic run-time version.

// ---- forward: outputs j0 .. j0+3 (block-relative), xs[n] = xz[orig*j_first + n] staged by the caller ----
// Needs xs[fir_pad(orig*(j0 + mm) + kappa)] for mm < M + 3, i.e. logical indices up to orig*(j0 + 3) + taps - 1.
template <int ORIG = 0, int TAPS = 0>
DM_HD void fir_fwd4(const float* xs, const float* w, int taps_rt, int orig_rt, int j0, float (&acc)[kFirR]) {
    const int orig = ORIG > 0 ? ORIG : orig_rt, taps = TAPS > 0 ? TAPS : taps_rt;
#pragma unroll
    for (int c = 0; c < kFirR; ++c) acc[c] = 0.f;
    if (ORIG > 0) {
#pragma unroll
        for (int kappa = 0; kappa < (ORIG > 0 ? ORIG : 1); ++kappa) {
            constexpr int O = ORIG > 0 ? ORIG : 1, TP = TAPS > 0 ? TAPS : 1;
            const int M = (TP - kappa + O - 1) / O;
            const int x0 = O * j0 + kappa;
#pragma unroll
            for (int mm = 0; mm < (TP + O - 1) / O + kFirR - 1; ++mm) {
                if (mm < M + kFirR - 1) {
                    const float xv = xs[fir_pad(x0 + O * mm)];
#pragma unroll
                    for (int c = 0; c < kFirR; ++c) {
                        const int m = mm - c;  // tap index of output c fed by this sample
                        if (m >= 0 && m < M) acc[c] = fmaf(xv, w[kappa + O * m], acc[c]);
                    }
                }
            }
        }
        return;
    }
    for (int kappa = 0; kappa < orig; ++kappa) {
        const int M = (taps - kappa + orig - 1) / orig;  // taps of this phase
        float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;    // weights m = mm, mm-1, mm-2, mm-3
        const int x0 = orig * j0 + kappa;
        for (int mm = 0; mm < M + kFirR - 1; ++mm) {
            w3 = w2;
            w2 = w1;
            w1 = w0;
            w0 = (mm < M) ? w[kappa + orig * mm] : 0.f;
            const float xv = xs[fir_pad(x0 + orig * mm)];
            acc[0] = fmaf(xv, w0, acc[0]);
            acc[1] = fmaf(xv, w1, acc[1]);
            acc[2] = fmaf(xv, w2, acc[2]);
            acc[3] = fmaf(xv, w3, acc[3]);
        }
    }
}

// ---- adjoint: outputs t0, t0+orig, t0+2*orig, t0+3*orig (chunk-relative input positions of one phase) ----
// ys[n] = (folded, scaled) ybar[j_base + n] staged by the caller with zeros outside [0, Ly);
// A = (i0 + width - taps + 1) - j_base*orig in (-orig, 0].
template <int ORIG = 0, int TAPS = 0>
DM_HD void fir_adj4(const float* ys, const float* w, int taps_rt, int orig_rt, int A, int t0, float (&acc)[kFirR]) {
    const int orig = ORIG > 0 ? ORIG : orig_rt, taps = TAPS > 0 ? TAPS : taps_rt;
    const int hi0 = A + taps - 1 + t0;  // (i + width) - j_base*orig for the first output, >= 0
    const int jb0 = hi0 / orig;
    const int kappa = hi0 - jb0 * orig;
    const int M = (taps - kappa + orig - 1) / orig;
#pragma unroll
    for (int c = 0; c < kFirR; ++c) acc[c] = 0.f;
    const float* yp = ys + jb0 + kFirR - 1;
    if (ORIG > 0) {
        // kappa is a run-time phase, but M only takes the values MMAX or MMAX - 1: run MMAX taps with a zero weight
        // for the (possibly) missing last one
        constexpr int O = ORIG > 0 ? ORIG : 1, TP = TAPS > 0 ? TAPS : 1;
        constexpr int MMAX = (TP + O - 1) / O;
        float wk[MMAX];
#pragma unroll
        for (int m = 0; m < MMAX; ++m) wk[m] = (m < M) ? w[kappa + O * m] : 0.f;
#pragma unroll
        for (int mm = 0; mm < MMAX + kFirR - 1; ++mm) {
            const float yv = yp[-mm];
#pragma unroll
            for (int c = 0; c < kFirR; ++c) {
                const int m = mm - (kFirR - 1 - c);  // output c pairs sample jb0 + 3 - mm with tap mm - (3 - c)
                if (m >= 0 && m < MMAX) acc[c] = fmaf(yv, wk[m], acc[c]);
            }
        }
        return;
    }
    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;  // weights m = mm, mm-1, mm-2, mm-3 (outputs c = 3, 2, 1, 0)
    for (int mm = 0; mm < M + kFirR - 1; ++mm) {
        w3 = w2;
        w2 = w1;
        w1 = w0;
        w0 = (mm < M) ? w[kappa + orig * mm] : 0.f;
        const float yv = yp[-mm];
        acc[3] = fmaf(yv, w0, acc[3]);
        acc[2] = fmaf(yv, w1, acc[2]);
        acc[1] = fmaf(yv, w2, acc[1]);
        acc[0] = fmaf(yv, w3, acc[0]);
    }
}

// ---- register-window bodies for the reference's scale-2 filter (orig 2, 28 taps, width 13; operator.py:180 with
//      run.py:188): a thread owns 8 consecutive outputs and keeps its whole input window in registers, so the FIR runs
//      without shared memory and the global traffic is 128-bit loads / stores only. ----
constexpr int kFir2Taps = 28, kFir2Width = 13, kFir2Out = 8;
constexpr int kFir2FwdWin = 48;  // x[2 j0 - 16 .. 2 j0 + 31]: 12 aligned float4 (j0 % 8 == 0)
constexpr int kFir2AdjWin = 20;  // ybar[i0/2 - 8 .. i0/2 + 11]: 5 aligned float4 (i0 % 8 == 0)
// y[j0 + c] = sum_k xz[2 (j0 + c) + k - 13] h[k] ; win[n] = xz[2 j0 - 16 + n]
DM_HD void fir2_fwd8(const float (&win)[kFir2FwdWin], const float (&h)[kFir2Taps], float (&y)[kFir2Out]) {
#pragma unroll
    for (int c = 0; c < kFir2Out; ++c) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kFir2Taps; ++k) a = fmaf(win[3 + 2 * c + k], h[k], a);
        y[c] = a;
    }
}
// xbar[i0 + 2 u]     = sum_t ybar[m - 7 + t] h[27 - 2 t]   (m = i0/2 + u, u < 4, t < 14)
// xbar[i0 + 2 u + 1] = sum_t ybar[m - 6 + t] h[26 - 2 t] ; win[n] = ybar[i0/2 - 8 + n]
DM_HD void fir2_adj8(const float (&win)[kFir2AdjWin], const float (&h)[kFir2Taps], float (&x)[kFir2Out]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float e = 0.f, o = 0.f;
#pragma unroll
        for (int t = 0; t < 14; ++t) {
            e = fmaf(win[u + 1 + t], h[27 - 2 * t], e);
            o = fmaf(win[u + 2 + t], h[26 - 2 * t], o);
        }
        x[2 * u] = e;
        x[2 * u + 1] = o;
    }
}

// ---- persistent forward kernel (resample2_fwd_stream_kernel): chunk geometry and shared-memory cell maps.
// A chunk = 1024 outputs = 8 per thread of a 128-thread CTA; its input span x[2 j0 - 16 .. 2 j0 + 2064) is staged as 520
// cells of 4 floats; logical cell c sits at rs2_cell(c) (its position inside the 128-byte row of 8 cells xor-ed with
// row & 3), which makes both the coalesced staging writes (8 consecutive cells per quarter-warp) and the window reads
// (thread t reads cells 4 t .. 4 t + 11: lane stride 64 B) bank-conflict free (audited on the host).
constexpr int kRs2ChunkOut = 1024;
constexpr int kRs2Threads = kRs2ChunkOut / kFir2Out;      // 128
constexpr int kRs2Cells = (2 * kRs2ChunkOut + 32) / 4;    // 520
constexpr int kRs2BufFloats = ((kRs2Cells + 7) / 8) * 8 * 4;
DM_HDC int rs2_cell(int c) { return (c & ~7) | ((c & 7) ^ ((c >> 3) & 3)); }
// physical cell of window cell q (0..11) of thread t = rs2_cell(4 t + q), in the cheap form the kernel evaluates:
// cells 4 u .. 4 u + 3 of u = t + q / 4 share a half row, so only the low two bits are permuted
DM_HDC int rs2_win_cell(int t, int q) { return 4 * (t + (q >> 2)) + ((q & 3) ^ (((t + (q >> 2)) >> 1) & 3)); }

// signed ceil-division (orig > 0)
DM_HD long long ceil_div_ll(long long a, long long b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }

}  // namespace dm
