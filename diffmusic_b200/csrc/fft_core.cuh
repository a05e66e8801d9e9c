// Radix-8 Stockham FFT building blocks shared by the STFT guidance kernel (N = 512 complex, one 1024-point real
// frame) and the overlap-save RIR convolution kernel (N = 4096 complex, one 8192-point real block).
//
// Everything here is `__host__ __device__` and written as *phases*: a phase is executed by every thread of an FFT
// group with a barrier between phases.  On the GPU the barrier is `bar.sync`; tests/cpu_emul runs the very same
// phase functions on the host, looping over thread ids, so the index arithmetic is validated without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#define DM_HDC __host__ __device__ constexpr
#else
#define DM_HD inline
#define DM_HDC constexpr
#endif

namespace dm {

struct cf {
    float x, y;
};
struct alignas(8) f2 {  // 64-bit pair for vector shared-memory accesses
    float x, y;
};

DM_HD cf cmul(cf a, cf b) { return cf{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
DM_HD cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
DM_HD cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }
DM_HD cf cconj(cf a) { return cf{a.x, -a.y}; }
// multiply by sign*i  (sign = +1: i*a, sign = -1: -i*a)
template <int SIGN>
DM_HD cf cmul_i(cf a) {
    return SIGN > 0 ? cf{-a.y, a.x} : cf{a.y, -a.x};
}

// Shared-memory index swizzle for the split re/im float arrays: XOR-ing the low five address bits with bits [3,8) of
// the index makes EVERY access pattern of the radix-8 Stockham passes bank-conflict free -- the unit-stride loads
// j + r*T, the stride-8 scatter of the first pass, the 8-block scatter of the second and the unit-stride scatters of the
// later ones (exhaustively checked for N = 512 and N = 4096: 96 / 1024 wavefronts = the minimum, versus 160 / 1792 for
// the usual pad-one-word-every-8 layout).  No padding words are needed.
DM_HD int swz(int i) { return i ^ ((i >> 3) & 31); }
DM_HDC int swz_len(int n) { return n; }

// 8-point DFT in registers: out[q] = sum_r v[r] * exp(SIGN * 2*pi*i * r*q / 8)
template <int SIGN>
DM_HD void dft8(cf v[8]) {
    const float h = 0.70710678118654752440f;
    cf a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    cf a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
    cf a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
    cf a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
    // odd branch twiddles w^1, w^2, w^3 with w = exp(SIGN*i*pi/4)
    a5 = cf{h * (a5.x - SIGN * a5.y), h * (a5.y + SIGN * a5.x)};
    a6 = cmul_i<SIGN>(a6);
    a7 = cf{h * (-a7.x - SIGN * a7.y), h * (-a7.y + SIGN * a7.x)};
    cf b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = cmul_i<SIGN>(csub(a1, a3));
    cf b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = cmul_i<SIGN>(csub(a5, a7));
    v[0] = cadd(b0, b1);
    v[4] = csub(b0, b1);
    v[2] = cadd(b2, b3);
    v[6] = csub(b2, b3);
    v[1] = cadd(b4, b5);
    v[5] = csub(b4, b5);
    v[3] = cadd(b6, b7);
    v[7] = csub(b6, b7);
}

// One Stockham pass for thread j of N/8.  `load(i)` returns input element i (natural order of the previous pass),
// `store(i, c)` writes output element i.  tw[m] = exp(-2*pi*i*m/N) (forward table; conjugated for SIGN=+1).
template <int N, int NS, int SIGN, class Load, class Store>
DM_HD void stockham_pass(int j, const cf* __restrict__ tw, Load load, Store store) {
    constexpr int T = N / 8;
    const int k = j % NS;
    cf v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = load(j + r * T);
    if (NS > 1) {
        constexpr int TWS = N / (NS * 8);
#pragma unroll
        for (int r = 1; r < 8; ++r) {
            cf w = tw[r * k * TWS];
            if (SIGN > 0) w.y = -w.y;
            v[r] = cmul(v[r], w);
        }
    }
    dft8<SIGN>(v);
    const int j0 = (j / NS) * NS * 8 + k;
#pragma unroll
    for (int q = 0; q < 8; ++q) store(j0 + q * NS, v[q]);
}

// Same pass with the thread's seven twiddles held in registers.  Thread j of a group always owns the same k, so a
// kernel that transforms many frames loads w[] once (thread_twiddles) instead of 7 conflicted shared-memory reads per
// pass and frame.  w[r-1] is the FORWARD twiddle exp(-2 pi i r k / (8 NS)); conjugated here for SIGN = +1.
template <int N, int NS>
DM_HD void thread_twiddles(int j, const cf* __restrict__ tw, cf (&w)[7]) {
    constexpr int TWS = N / (NS * 8);
    const int k = j % NS;
#pragma unroll
    for (int r = 1; r < 8; ++r) w[r - 1] = tw[r * k * TWS];
}
// Swizzled split-array I/O of one radix-8 butterfly: inputs j + r*T, outputs j0 + q*NS.
// The swizzled addresses have closed forms with ONE variable term per thread and compile-time constants per r / q,
// because the index pieces occupy disjoint bit fields (so + is ^ and the mask (i >> 3) & 31 splits the same way):
//   loads   j + r*T        T = 64 : (swz(j) ^ ((r & 3) << 3)) + 64 r          T = 512 : swz(j) + 512 r
//   NS = 1  8 j + q               : ((8 j) ^ (j & 31)) ^ q
//   NS = 8  64 a + k + 8 q        : (64 a ^ k ^ ((a & 3) << 3)) ^ (9 q)          a = j >> 3, k = j & 7
//   NS = 64 512 c + l + 64 q      : (((512 c + l) ^ (l >> 3)) ^ ((q & 3) << 3)) + 64 q     c = j >> 6, l = j & 63
//   NS = 512 j + 512 q            : swz(j) + 512 q
// (checked against swz() itself for every j, r, q in tests/cpu_emul; one LOP3/IADD per access instead of five.)
template <int N>
DM_HD int ld_addr(int b0, int r) {
    static_assert(N == 512 || N == 4096, "closed forms derived for N = 512 and N = 4096");
    return N == 512 ? ((b0 ^ ((r & 3) << 3)) + 64 * r) : (b0 + 512 * r);
}
template <int N>
DM_HD void load8_swz(const float* __restrict__ re, const float* __restrict__ im, int j, cf (&v)[8]) {
    const int b0 = swz(j);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int a = ld_addr<N>(b0, r);
        v[r] = cf{re[a], im[a]};
    }
}
template <int N, int NS>
DM_HD int st_base(int j) {
    if (NS == 1) return (8 * j) ^ (j & 31);
    if (NS == 8) return (64 * (j >> 3)) ^ (j & 7) ^ (((j >> 3) & 3) << 3);
    if (NS == 64) return ((512 * (j >> 6) + (j & 63)) ^ ((j & 63) >> 3));
    return swz(j);  // NS == 512
}
template <int N, int NS>
DM_HD int st_addr(int b, int q) {
    static_assert(NS == 1 || NS == 8 || NS == 64 || (NS == 512 && N == 4096), "unsupported pass");
    if (NS == 1) return b ^ q;
    if (NS == 8) return b ^ (9 * q);
    if (NS == 64) return (b ^ ((q & 3) << 3)) + 64 * q;
    return b + 512 * q;
}
template <int N, int NS>
DM_HD void store8_swz(float* __restrict__ re, float* __restrict__ im, int j, const cf (&v)[8]) {
    const int b = st_base<N, NS>(j);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int a = st_addr<N, NS>(b, q);
        re[a] = v[q].x;
        im[a] = v[q].y;
    }
}
// Seven twiddles of thread j generated from the first one by complex multiplication (w^2 = w*w, w^3 = w^2*w,
// w^4 = (w^2)^2, w^5 = w^4*w, w^6 = (w^3)^2, w^7 = w^4*w^3): one table read per pass and thread instead of seven strided
// ones.  Used by the 4096-point RIR transform, where each CTA does a single block and register-resident twiddles would
// not be reused.  Error: <= 3 extra roundings on |w| = 1, far below the 1e-4 parity bound.
template <int N, int NS, int SIGN>
DM_HD void twiddle8_rec(int j, const cf* __restrict__ tw, cf (&v)[8]) {
    constexpr int TWS = N / (NS * 8);
    cf w1 = tw[(j % NS) * TWS];
    if (SIGN > 0) w1.y = -w1.y;
    const cf w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
    v[1] = cmul(v[1], w1);
    v[2] = cmul(v[2], w2);
    v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], cmul(w4, w1));
    v[6] = cmul(v[6], cmul(w3, w3));
    v[7] = cmul(v[7], cmul(w4, w3));
}
// smem -> smem pass with recurrence twiddles (the 4096-point transform)
template <int N, int NS, int SIGN>
DM_HD void stockham_pass_rec_swz(int j, const cf* __restrict__ tw, const float* in_re, const float* in_im,
                                 float* out_re, float* out_im) {
    cf v[8];
    load8_swz<N>(in_re, in_im, j, v);
    twiddle8_rec<N, NS, SIGN>(j, tw, v);
    dft8<SIGN>(v);
    store8_swz<N, NS>(out_re, out_im, j, v);
}
template <int SIGN>
DM_HD void twiddle8(cf (&v)[8], const cf (&w)[7]) {
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        cf t = w[r - 1];
        if (SIGN > 0) t.y = -t.y;
        v[r] = cmul(v[r], t);
    }
}
// padded -> padded pass with register twiddles (NS > 1)
template <int N, int NS, int SIGN>
DM_HD void stockham_pass_swz(int j, const cf (&w)[7], const float* in_re, const float* in_im, float* out_re,
                             float* out_im) {
    cf v[8];
    load8_swz<N>(in_re, in_im, j, v);
    twiddle8<SIGN>(v, w);
    dft8<SIGN>(v);
    store8_swz<N, NS>(out_re, out_im, j, v);
}

// Swizzled split-array accessors (used by the phases that touch arbitrary indices: unpack / pack / first / last pass).
struct SwzLoad {
    const float* re;
    const float* im;
    DM_HD cf operator()(int i) const {
        int p = swz(i);
        return cf{re[p], im[p]};
    }
};
struct SwzStore {
    float* re;
    float* im;
    DM_HD void operator()(int i, cf c) const {
        int p = swz(i);
        re[p] = c.x;
        im[p] = c.y;
    }
};

// Real-FFT unpacking for a 2H-point real signal packed as H complex (z[n] = x[2n] + i x[2n+1]); Z = FFT_H(z).
// For the pair (k, H-k), 1 <= k < H/2:  X[k] = E + T, X[H-k] = conj(E - T) with E = (Z[k]+conj(Z[H-k]))/2,
// T = W^k * (-i/2) * (Z[k]-conj(Z[H-k])), W = exp(-2*pi*i/(2H)).
DM_HD void rfft_unpack_pair(cf zk, cf zc, cf w, cf& xk, cf& xc) {
    cf e = cf{0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y)};
    cf d = cf{0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y)};  // (Z[k] - conj(Z[H-k]))/2
    cf o = cf{d.y, -d.x};                                    // -i * d
    cf t = cmul(w, o);
    xk = cadd(e, t);
    xc = cconj(csub(e, t));
}
// Adjoint / inverse packing: given Y[k], Y[H-k] of a Hermitian half spectrum, Z[k] = A + S, Z[H-k] = conj(A - S)
// with A = Y[k] + conj(Y[H-k]), S = i * conj(W^k) * (Y[k] - conj(Y[H-k])).  IFFT_H(Z) (unnormalised) then holds
// x[2n] in re and x[2n+1] in im, where x[n] = Y0 + (-1)^n YH + 2 Re sum_{0<k<H} Y[k] e^{+2 pi i k n/(2H)}.
DM_HD void irfft_pack_pair(cf yk, cf yc, cf w, cf& zk, cf& zc) {
    cf a = cf{yk.x + yc.x, yk.y - yc.y};
    cf b = cf{yk.x - yc.x, yk.y + yc.y};
    cf vb = cmul(cconj(w), b);
    cf s = cf{-vb.y, vb.x};  // i * vb
    zk = cadd(a, s);
    zc = cconj(csub(a, s));
}

}  // namespace dm
