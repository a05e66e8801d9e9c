// Warp-per-frame-pair STFT guidance pipeline: ONE warp carries TWO 1024-sample real frames (A, B) through
//   window -> FFT -> |X|^2 (or |X|) -> sparse mel -> dB / clamp -> residual -> VJP -> inverse FFT -> window
// as ONE 1024-point COMPLEX transform of z = a + i b, factored 32 x 32:
//
//   pass 1  lane t holds z[t + 32 r], r = 0..31 (64 registers): in-register DFT-32 over r, times W1024^(t k1)
//   exchange (the only shared-memory transpose of the direction: 8 KB written, 8 KB read, all accesses conflict-free)
//   pass 2  lane k1 holds the 32 values of column k1: in-register DFT-32 over t  ->  Z[k1 + 32 k2]
//   split   A[k] = (Z[k] + conj Z[1024-k]) / 2, B[k] = (Z[k] - conj Z[1024-k]) / 2i : Z[1024-k] lives in lane 32-k1,
//           register 31-k2, so 32 shuffles with the partner lane replace the real-FFT unpack pass (and its twiddles);
//           lane k1 then owns the bins k = k1 + 32 i, i = 0..15, of BOTH frames (lane 0 also bin 512).
// The backward direction mirrors it: per-bin cotangents -> Q[k], Q[1024-k] (partner shuffle) -> DFT-32 -> twiddle ->
// exchange -> DFT-32 -> Re / Im are the two frame gradients, which are windowed and left in the warp's buffer for the
// CTA's gathered overlap-add.
//
// Compared with the 64-thread frame-pair pipeline (stft_pair.cuh: three radix-8 passes each way plus unpack / pack)
// this halves the shared-memory wavefronts per frame pair, has no named barriers (only __syncwarp) and 20 % fewer
// instructions.  The factor 1/2 of the split is folded into the staged window (exact: a power of two).
//
// Arithmetic restated from torchaudio (functional.spectrogram / MelScale / amplitude_to_DB) as used by the reference's
// diffmusic/inverse_problem/operator.py:24-36,123-124,153-154,162-171.  Everything is __host__ __device__ phase code:
// tests/cpu_emul runs it lane by lane on the host (shuffles = a phase boundary).
#pragma once
#include "stft_pair.cuh"

namespace dm {

struct alignas(16) f4 {
    float x, y, z, w;
};

constexpr int kWarpRow = 34;                        // cf per exchange row: 32 + 2 pad -> 272 B, LDS.128 conflict-free
constexpr int kWarpBufFloats = 32 * kWarpRow * 2;   // 2176 floats = 8704 B per warp
// While it is not an exchange buffer the warp buffer holds (float offsets):
constexpr int kWarpPOff = 0;        // f2 P[513]       per-bin energies (frame A, frame B)
constexpr int kWarpMelbarOff = 1040;  // f2 melbar[72]   mel cotangent (entries 64.. stay zero)
constexpr int kWarpGOff = 0;        // float G[2][1024] windowed frame gradients of frame A / B

// register position of output k of dft32 and its inverse map
DM_HDC int perm32(int p) { return (p >> 3) + 4 * (p & 7); }
DM_HDC int pinv32(int k) { return 8 * (k & 3) + (k >> 2); }

// cos(2 pi e / 32), e = 0..8
DM_HDC float cos32_q(int e) {
    return e == 0   ? 1.0f
           : e == 1 ? 0.98078528040323044913f
           : e == 2 ? 0.92387953251128675613f
           : e == 3 ? 0.83146961230254523708f
           : e == 4 ? 0.70710678118654752440f
           : e == 5 ? 0.55557023301960222474f
           : e == 6 ? 0.38268343236508977173f
           : e == 7 ? 0.19509032201612826785f
                    : 0.0f;
}
DM_HDC float cos32(int e) {
    e &= 31;
    return e <= 8 ? cos32_q(e) : e <= 16 ? -cos32_q(16 - e) : e <= 24 ? -cos32_q(e - 16) : cos32_q(32 - e);
}
DM_HDC float sin32(int e) { return cos32(e - 8); }

// a * exp(SIGN * 2 pi i E / 32) with the trivial and the 45-degree cases spelled out
template <int E, int SIGN>
DM_HD cf mul_w32(cf a) {
    constexpr int e = E & 31;
    constexpr float h = 0.70710678118654752440f;
    if (e == 0) return a;
    if (e == 8) return cmul_i<SIGN>(a);
    if (e == 16) return cf{-a.x, -a.y};
    if (e == 24) return cmul_i<-SIGN>(a);
    if (e == 4) return cf{h * (a.x - SIGN * a.y), h * (a.y + SIGN * a.x)};
    if (e == 12) return cf{h * (-a.x - SIGN * a.y), h * (-a.y + SIGN * a.x)};
    if (e == 20) return cf{h * (-a.x + SIGN * a.y), h * (-a.y - SIGN * a.x)};
    if (e == 28) return cf{h * (a.x + SIGN * a.y), h * (a.y - SIGN * a.x)};
    constexpr float c = cos32(e), s = SIGN * sin32(e);
    return cf{a.x * c - a.y * s, a.x * s + a.y * c};
}

template <int SIGN>
DM_HD void dft4(cf& a0, cf& a1, cf& a2, cf& a3) {
    const cf s0 = cadd(a0, a2), d0 = csub(a0, a2), s1 = cadd(a1, a3), d1 = cmul_i<SIGN>(csub(a1, a3));
    a0 = cadd(s0, s1);
    a2 = csub(s0, s1);
    a1 = cadd(d0, d1);
    a3 = csub(d0, d1);
}

template <int K1, int SIGN>
DM_HD void dft32_twiddle_row(cf (&v)[32]) {
    v[1 + 8 * K1] = mul_w32<1 * K1, SIGN>(v[1 + 8 * K1]);
    v[2 + 8 * K1] = mul_w32<2 * K1, SIGN>(v[2 + 8 * K1]);
    v[3 + 8 * K1] = mul_w32<3 * K1, SIGN>(v[3 + 8 * K1]);
    v[4 + 8 * K1] = mul_w32<4 * K1, SIGN>(v[4 + 8 * K1]);
    v[5 + 8 * K1] = mul_w32<5 * K1, SIGN>(v[5 + 8 * K1]);
    v[6 + 8 * K1] = mul_w32<6 * K1, SIGN>(v[6 + 8 * K1]);
    v[7 + 8 * K1] = mul_w32<7 * K1, SIGN>(v[7 + 8 * K1]);
}

// In-register 32-point DFT, V[k] = sum_r v[r] exp(SIGN 2 pi i r k / 32), as 4 x 8: r = r0 + 8 r1, k = k1 + 4 k0.
// Input natural order; output k sits in register pinv32(k) (register p holds output perm32(p)).
template <int SIGN>
DM_HD void dft32(cf (&v)[32]) {
#pragma unroll
    for (int r0 = 0; r0 < 8; ++r0) dft4<SIGN>(v[r0], v[r0 + 8], v[r0 + 16], v[r0 + 24]);  // -> u[r0][k1] at v[r0 + 8 k1]
    dft32_twiddle_row<1, SIGN>(v);
    dft32_twiddle_row<2, SIGN>(v);
    dft32_twiddle_row<3, SIGN>(v);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft8<SIGN>(v + 8 * k1);  // over r0 -> V[k1 + 4 k0] at v[8 k1 + k0]
}

// ---- exchange: multiply output k of dft32 by (W1024^lane)^k (conjugated for SIGN = +1) and store it to row k, column
// lane; then lane j reads row j.  tw4[m * 32 + lane] = (w^(2m), w^(2m+1)) with w = W1024^lane = exp(-2 pi i lane / 1024).
template <int SIGN>
DM_HD void warp_twiddle_store(int lane, cf (&v)[32], const f4* __restrict__ tw4, cf* __restrict__ xbuf) {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const f4 t = tw4[m * 32 + lane];
        cf w0 = cf{t.x, t.y}, w1 = cf{t.z, t.w};
        if (SIGN > 0) {
            w0.y = -w0.y;
            w1.y = -w1.y;
        }
        const int p0 = pinv32(2 * m), p1 = pinv32(2 * m + 1);
        if (m > 0) v[p0] = cmul(v[p0], w0);
        v[p1] = cmul(v[p1], w1);
    }
#pragma unroll
    for (int p = 0; p < 32; ++p) xbuf[perm32(p) * kWarpRow + lane] = v[p];
}
DM_HD void warp_xchg_load(int lane, const cf* __restrict__ xbuf, cf (&v)[32]) {
    const f4* row = reinterpret_cast<const f4*>(xbuf + lane * kWarpRow);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f4 q = row[j];
        v[2 * j] = cf{q.x, q.y};
        v[2 * j + 1] = cf{q.z, q.w};
    }
}

// ---- forward pass 1: z[n] = win[n]/2 * (a[n] + i b[n]), n = lane + 32 r.  win2[n] = (win[n], win[n + 512]) / 2.
DM_HD void warp_load_frames(int lane, const float* __restrict__ fa, const float* __restrict__ fb,
                            const f2* __restrict__ win2, cf (&v)[32]) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int n = lane + 32 * r;
        const f2 w = win2[n];
        v[r] = cf{fa[n] * w.x, fb[n] * w.x};
        v[r + 16] = cf{fa[n + 512] * w.y, fb[n + 512] * w.y};
    }
}

// Spectrum of the owned bins: slot i = bin lane + 32 i (i < 16); slot 16 = bin 512 (lane 0 only).
struct WarpX {
    cf a[17], b[17];
};

// ---- split, part 1: what this lane hands to its partner lane (32 - lane) & 31 at step i: Z[lane + 32 (31 - i)];
// lane 0 is its own partner and needs Z[32 ((32 - i) & 31)] back instead.
DM_HD void warp_split_send(int lane, const cf (&v)[32], cf (&snd)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const cf g = v[pinv32(31 - i)], z = v[pinv32((32 - i) & 31)];
        snd[i] = lane == 0 ? z : g;
    }
}
// part 2: rcv[i] = Z[1024 - k], k = lane + 32 i  ->  A[k], B[k]  (already halved through the window)
DM_HD void warp_split_recv(int lane, const cf (&v)[32], const cf (&rcv)[16], WarpX& x) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const cf z = v[pinv32(i)], p = rcv[i];
        x.a[i] = cf{z.x + p.x, z.y - p.y};
        x.b[i] = cf{z.y + p.y, p.x - z.x};
    }
    const cf z = v[pinv32(16)];  // bin 512 is its own mirror (lane 0)
    x.a[16] = cf{z.x + z.x, 0.f};
    x.b[16] = cf{z.y + z.y, 0.f};
}

// ---- per-bin energies -> P (shared); optional Gaussian noise on the magnitude is added by the caller in between
template <int MODE>
DM_HD void warp_energies(const WarpX& x, f2 (&e)[17]) {
#pragma unroll
    for (int i = 0; i < 17; ++i) e[i] = f2{pair_bin_energy<MODE>(x.a[i]), pair_bin_energy<MODE>(x.b[i])};
}
DM_HD void warp_store_energies(int lane, const f2 (&e)[17], f2* __restrict__ P) {
#pragma unroll
    for (int i = 0; i < 16; ++i) P[lane + 32 * i] = e[i];
    if (lane == 0) P[kH] = e[16];
}

// ---- mel projection: lane j sums band j, then band 63 - j (3 + 41 ... 12 + 12 bins: <= 44 per lane)
struct WarpMelConsts {
    int k0a, na, k0b, nb;
};
DM_HD void load_warp_mel_consts(int lane, const StftTables& t, WarpMelConsts& c) {
    c.k0a = t.mel_kstart[lane];
    c.na = t.mel_klen[lane];
    c.k0b = t.mel_kstart[63 - lane];
    c.nb = t.mel_klen[63 - lane];
}
DM_HD void warp_mel_project(int lane, const WarpMelConsts& c, const float* __restrict__ melw_t,
                            const f2* __restrict__ P, f2& lo, f2& hi) {
    lo = pair_mel_segment(melw_t, P, lane, c.k0a, 0, c.na);
    hi = pair_mel_segment(melw_t, P, 63 - lane, c.k0b, 0, c.nb);
}

// ---- backward: cotangent of the owned bins -> Q (natural order in v) and what the partner lane needs
// g[i] = d loss / d energy (or magnitude) of bin lane + 32 i for (frame A, frame B)
template <int MODE>
DM_HD float warp_bin_scale(cf x, float g) {
    if (MODE == kModeMelDb) return 2.f * g;  // d|X|^2 = 2 X
    const float e = x.x * x.x + x.y * x.y;
    return e > 0.f ? g * fast_rsqrt(e) : 0.f;  // d|X| = X / |X|, 0 at X = 0
}
DM_HD void warp_bin_cotangents(int lane, const PairBinTab& t, const f2* __restrict__ melbar, f2 (&g)[17]) {
#pragma unroll
    for (int i = 0; i < 17; ++i) {
        const int k = (i < 16) ? lane + 32 * i : kH;
        const int m0 = t.binm[k];
        const f2 w = t.binw[k];
        const f2 g0 = melbar[m0], g1 = melbar[m0 + 1];
        g[i] = f2{w.x * g0.x + w.y * g1.x, w.x * g0.y + w.y * g1.y};
    }
}
// Q[k] = Xbar_A[k] + i Xbar_B[k] -> v[i];  Q[1024 - k] = conj Xbar_A[k] + i conj Xbar_B[k] -> snd[i] (for the partner).
// (The 1/2 of Re(.) = (. + conj .)/2 rides on the halved synthesis window.)
template <int MODE>
DM_HD void warp_pack_send(int lane, const WarpX& x, const f2 (&g)[17], cf (&v)[32], cf (&snd)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float sa = warp_bin_scale<MODE>(x.a[i], g[i].x), sb = warp_bin_scale<MODE>(x.b[i], g[i].y);
        const float ar = sa * x.a[i].x, ai = sa * x.a[i].y, br = sb * x.b[i].x, bi = sb * x.b[i].y;
        v[i] = cf{ar - bi, ai + br};
        snd[i] = cf{ar + bi, br - ai};
    }
    if (lane == 0) {  // DC: both halves land on Q[0]; only the real parts of the cotangent act
        const float sa = warp_bin_scale<MODE>(x.a[0], g[0].x), sb = warp_bin_scale<MODE>(x.b[0], g[0].y);
        v[0] = cf{2.f * sa * x.a[0].x, 2.f * sb * x.b[0].x};
    }
}
// rcv[i] = partner's snd[i] = Q[lane + 32 (31 - i)]; lane 0 received its own: Q[32 (32 - i)], and owns Q[512]
template <int MODE>
DM_HD void warp_pack_recv(int lane, const WarpX& x, const f2 (&g)[17], const cf (&rcv)[16], cf (&v)[32]) {
    const float sa = warp_bin_scale<MODE>(x.a[16], g[16].x), sb = warp_bin_scale<MODE>(x.b[16], g[16].y);
    const cf q512 = cf{2.f * sa * x.a[16].x, 2.f * sb * x.b[16].x};
    v[16] = lane == 0 ? q512 : rcv[15];
#pragma unroll
    for (int k2 = 17; k2 < 32; ++k2) v[k2] = lane == 0 ? rcv[32 - k2] : rcv[31 - k2];
}

// ---- last inverse pass output -> windowed frame gradients in the warp buffer: G[0][n] (frame A), G[1][n] (frame B)
DM_HD void warp_store_gradients(int lane, const cf (&v)[32], const f2* __restrict__ win2, float* __restrict__ G) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int n = lane + 32 * r;
        const f2 w = win2[n];
        const cf lo = v[pinv32(r)], hi = v[pinv32(r + 16)];
        G[n] = lo.x * w.x;
        G[kNfft + n] = lo.y * w.x;
        G[n + 512] = hi.x * w.y;
        G[kNfft + n + 512] = hi.y * w.y;
    }
}

// W1024^e, e = 0..1023, from the quarter table w1024[0..256] (exact symmetries)
DM_HD cf w1024_any(const cf* __restrict__ w1024, int e) {
    e &= 1023;
    const int q = e >> 8, r = e & 255;
    const cf t = w1024[r];
    return q == 0 ? t : q == 1 ? cf{t.y, -t.x} : q == 2 ? cf{-t.x, -t.y} : cf{-t.y, t.x};
}

}  // namespace dm
