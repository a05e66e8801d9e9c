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
constexpr int kWarpPOff = 0;          // f2 P[514]       per-bin energies (frame A, frame B); P[513] = 0 (pair padding)
constexpr int kWarpMelbarOff = 1040;  // f2 melbar[72]   mel cotangent (entries 64.. stay zero)
constexpr int kWarpGStride = kWarpBufFloats / 2;  // float G[2][1088]: windowed frame gradients, frame f of the TILE at
                                                  // wbuf_all + f * kWarpGStride (uniform stride across warps)

// Host-built table image (diffmusic_b200/tables.py warp_image): float offsets of its sections
struct WarpImage {
    int win2, tw4, melp, lanek, binw, binm, total;
};
DM_HD WarpImage warp_image_layout(int na, int nb) {
    WarpImage l;
    l.win2 = 0;
    l.tw4 = l.win2 + kNfft;
    l.melp = l.tw4 + 16 * 32 * 4;
    l.lanek = l.melp + (na + nb) * 64;
    l.binw = l.lanek + 128;  // int32 [4][32]: pa0, pb0, ma, mb
    l.binm = l.binw + 2 * 514;
    l.total = l.binm + 132;
    return l;
}

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
    f2* x2 = reinterpret_cast<f2*>(xbuf);  // 64-bit stores (cf itself is only 4-byte aligned)
#pragma unroll
    for (int p = 0; p < 32; ++p) x2[perm32(p) * kWarpRow + lane] = f2{v[p].x, v[p].y};
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
// ssa / ssb: this lane's share of the two frames' windowed energies (see warp_balance)
DM_HD void warp_load_frames(int lane, const float* __restrict__ fa, const float* __restrict__ fb,
                            const f2* __restrict__ win2, cf (&v)[32], float& ssa, float& ssb) {
    ssa = ssb = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int n = lane + 32 * r;
        const f2 w = win2[n];
        v[r] = cf{fa[n] * w.x, fb[n] * w.x};
        v[r + 16] = cf{fa[n + 512] * w.y, fb[n + 512] * w.y};
        ssa = fmaf(v[r].x, v[r].x, fmaf(v[r + 16].x, v[r + 16].x, ssa));
        ssb = fmaf(v[r].y, v[r].y, fmaf(v[r + 16].y, v[r + 16].y, ssb));
    }
}
// The joint transform of z = a + i b carries rounding noise of ~1e-7 max(|A|, |B|) into BOTH spectra, so a quiet frame
// next to a loud one (an inpainting mask edge, an onset) would come out far less accurate than from a transform of its
// own -- and the dB floor's 1 / mel derivative amplifies exactly those frames.  Frame B is therefore scaled by the power
// of two s that brings its energy next to frame A's (exact), and its spectrum is scaled back by 1 / s after the split;
// an all-zero frame (masked stretch, digital silence) gets an exactly zero spectrum, as torch.stft returns.
// ssa / ssb: the frames' windowed energies summed over the warp.
DM_HD int f32_exponent(float x) {
#if defined(__CUDA_ARCH__)
    return (__float_as_int(x) >> 23) & 0xff;
#else
    union { float f; int i; } u;
    u.f = x;
    return (u.i >> 23) & 0xff;
#endif
}
DM_HD float f32_pow2(int k) {  // 2^k, -126 <= k <= 127
#if defined(__CUDA_ARCH__)
    return __int_as_float((127 + k) << 23);
#else
    union { float f; int i; } u;
    u.i = (127 + k) << 23;
    return u.f;
#endif
}
DM_HD void warp_balance(float ssa, float ssb, float& s, float& inv_s, bool& zero_a, bool& zero_b) {
    zero_a = ssa == 0.f;
    zero_b = ssb == 0.f;
    int k = 0;
    if (!zero_a && !zero_b) k = (f32_exponent(ssa) - f32_exponent(ssb)) >> 1;  // ~ log2 sqrt(ssa / ssb)
    k = k < -40 ? -40 : (k > 40 ? 40 : k);
    s = f32_pow2(k);
    inv_s = f32_pow2(-k);
}
DM_HD void warp_scale_b(cf (&v)[32], float s) {
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r].y *= s;
}

// Spectrum of the owned bins: slot i = bin lane + 32 i (i < 16); slot 16 = bin 512 (lane 0 only).
struct WarpX {
    cf a[17], b[17];
};

// ---- split, one owned bin at a time (keeps the register footprint at the 64 values of the transform): at step i this
// lane hands Z[lane + 32 (31 - i)] to its partner lane (32 - lane) & 31 -- lane 0 is its own partner and needs
// Z[32 ((32 - i) & 31)] back instead -- and receives p = Z[1024 - k] for its bin k = lane + 32 i:
//   A[k] = Z[k] + conj p,  B[k] = -i (Z[k] - conj p)        (already halved through the window)
DM_HD cf warp_split_send_i(int lane, const cf (&v)[32], int i) {
    const cf g = v[pinv32(31 - i)], z = v[pinv32((32 - i) & 31)];
    return lane == 0 ? z : g;
}
DM_HD void warp_split_recv_i(const cf (&v)[32], int i, cf p, cf& a, cf& b) {
    const cf z = v[pinv32(i)];
    a = cf{z.x + p.x, z.y - p.y};
    b = cf{z.y + p.y, p.x - z.x};
}
// bin 512 is its own mirror (lane 0)
DM_HD void warp_split_nyquist(const cf (&v)[32], cf& a, cf& b) {
    const cf z = v[pinv32(16)];
    a = cf{z.x + z.x, 0.f};
    b = cf{z.y + z.y, 0.f};
}

// ---- per-bin energies (-> P in shared memory, natural bin order, (frame A, frame B) per entry)
template <int MODE>
DM_HD f2 warp_bin_energies(cf a, cf b) {
    return f2{pair_bin_energy<MODE>(a), pair_bin_energy<MODE>(b)};
}
DM_HD int warp_bin_of(int lane, int i) { return i < 16 ? lane + 32 * i : kH; }

// ---- mel projection: lane l sums its short band ma[l] (na bin-pair rows from pair pa0[l]), then its long band mb[l]
// (nb rows from pb0[l]); every lane runs the same na + nb iterations (weights are zero outside the band) with one 128-bit
// load of (P_A, P_B) of two bins and one 64-bit load of the two weights per iteration.  The host picks the band -> lane
// assignment and the window starts so that the 128-bit loads are bank-conflict free (tables.py warp_image).
// melp[row * 32 + lane] = weights of bins 2 (p0 + i), 2 (p0 + i) + 1.
DM_HD void warp_mel_project(int lane, int na, int nb, int pa0, int pb0, const f2* __restrict__ melp,
                            const f2* __restrict__ P, f2& lo, f2& hi) {
    const f4* Pa = reinterpret_cast<const f4*>(P) + pa0;
    const f4* Pb = reinterpret_cast<const f4*>(P) + pb0;
    const f2* wa = melp + lane;
    const f2* wb = melp + na * 32 + lane;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 4
    for (int i = 0; i < na; ++i) {
        const f2 w = wa[i * 32];
        const f4 v = Pa[i];
        a0 = fmaf(w.x, v.x, a0);
        a1 = fmaf(w.x, v.y, a1);
        a0 = fmaf(w.y, v.z, a0);
        a1 = fmaf(w.y, v.w, a1);
    }
#pragma unroll 4
    for (int i = 0; i < nb; ++i) {
        const f2 w = wb[i * 32];
        const f4 v = Pb[i];
        b0 = fmaf(w.x, v.x, b0);
        b1 = fmaf(w.x, v.y, b1);
        b0 = fmaf(w.y, v.z, b0);
        b1 = fmaf(w.y, v.w, b1);
    }
    lo = f2{a0, a1};
    hi = f2{b0, b1};
}

// ---- backward, one owned bin at a time: spectrum cotangent Xbar = scale * X of both frames ->
//   Q[k]        = Xbar_A[k] + i Xbar_B[k]              (stays in this lane, register k2 = i of the inverse pass 1)
//   Q[1024 - k] = conj Xbar_A[k] + i conj Xbar_B[k]    (for the partner lane's register k2 = 31 - i)
// (the 1/2 of Re(.) = (. + conj .)/2 rides on the halved synthesis window)
template <int MODE>
DM_HD float warp_bin_scale(cf x, float g) {
    if (MODE == kModeMelDb) return 2.f * g;  // d|X|^2 = 2 X
    const float e = x.x * x.x + x.y * x.y;
    return e > 0.f ? g * fast_rsqrt(e) : 0.f;  // d|X| = X / |X|, 0 at X = 0
}
template <int MODE>
DM_HD cf warp_xbar(cf x, float g) {
    const float s = warp_bin_scale<MODE>(x, g);
    return cf{s * x.x, s * x.y};
}
// energy cotangent of bin k for (frame A, frame B) from the mel cotangent (<= 2 bands touch a bin)
DM_HD f2 warp_bin_cotangent(int k, const PairBinTab& t, const f2* __restrict__ melbar) {
    const int m0 = t.binm[k];
    const f2 w = t.binw[k];
    const f2 g0 = melbar[m0], g1 = melbar[m0 + 1];
    return f2{w.x * g0.x + w.y * g1.x, w.x * g0.y + w.y * g1.y};
}
DM_HD void warp_q_pair(cf ya, cf yb, cf& q, cf& qm) {
    q = cf{ya.x - yb.y, ya.y + yb.x};
    qm = cf{ya.x + yb.y, yb.x - ya.y};
}
// DC and Nyquist: both halves land on the same Q entry and only the real parts of the cotangent act
DM_HD cf warp_q_real(cf ya, cf yb) { return cf{2.f * ya.x, 2.f * yb.x}; }

// ---- last inverse pass output -> windowed frame gradients in the warp buffer: G[n] (frame A), G[kWarpGStride + n] (B)
DM_HD void warp_store_gradients(int lane, const cf (&v)[32], const f2* __restrict__ win2, float* __restrict__ G) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int n = lane + 32 * r;
        const f2 w = win2[n];
        const cf lo = v[pinv32(r)], hi = v[pinv32(r + 16)];
        G[n] = lo.x * w.x;
        G[kWarpGStride + n] = lo.y * w.x;
        G[n + 512] = hi.x * w.y;
        G[kWarpGStride + n + 512] = hi.y * w.y;
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
