// Fused STFT / mel guidance kernel for sm_100a, warp-per-frame-pair engine (stft_warp.cuh).
//
// A CTA of 8 warps walks over tiles of <= 16 consecutive frames of one clip (persistent: grid = 2 CTAs per SM, the tile
// list strided over the grid).  Its constant tables arrive as ONE cp.async.bulk of a host-built shared-memory image
// (tables.py warp_image), overlapped with the first tile's signal.  Per tile:
//   * the tile's signal span is staged in shared memory by one cp.async.bulk (interior fp32 tiles; edge tiles, 16-bit
//     waveforms and the inpainting mask go sample by sample).  The span is dead as soon as every warp has loaded its two
//     frames into registers, so the NEXT tile's span is fetched behind the current tile's transforms;
//   * warp w owns frames 2w, 2w+1 and runs the whole chain window -> FFT -> energies -> sparse mel -> dB / clamp ->
//     residual -> VJP -> inverse FFT out of registers and its private 8.5 KB buffer; warps only meet at CTA barriers
//     around the gather (no named-barrier ring, no ordered chain);
//   * gathered overlap-add: every warp leaves its two windowed frame gradients in its buffer, then all 256 threads sum,
//     per 4 output samples, the <= 7 frames that cover them in ascending frame order (bit-reproducible) and add the
//     result to the padded cotangent in HBM (tiles overlap by < 1 frame -> <= 2 commutative adds per address).
// Same inputs / outputs / arithmetic as stft_pair_kernel (stft_guidance.cu), which stays as the engine for hops that
// are not a multiple of 4 and tiles of more than 16 frames, and as the A/B reference of the tests.
#include "fir_poly.cuh"
#include "stft_params.cuh"
#include "stft_warp.cuh"

namespace dm {

constexpr int kWarpsPerCta = 8;  // warps (= frame pairs) per CTA.  (7 warps with 144 registers for 14-frame tiles was
                                 // measured: registers are allocated per four warps, so only ONE such CTA fits an SM: 44 vs 37 us)

struct WarpSmemLayout {
    int sig, img, red, wbuf, total;  // float offsets
};
__host__ __device__ inline WarpSmemLayout warp_smem_layout(int nf, int hop, int img_floats, int warps) {
    WarpSmemLayout l;
    const int span = ((nf - 1) * hop + kNfft + 3) & ~3;
    l.sig = 0;
    l.img = l.sig + span;
    l.red = l.img + img_floats;
    l.wbuf = l.red + 8;
    l.total = l.wbuf + warps * kWarpBufFloats;
    return l;
}
size_t stft_warp_smem_bytes(int nf, int hop, int img_floats) {
    return (size_t)warp_smem_layout(nf, hop, img_floats, kWarpsPerCta).total * sizeof(float);
}

__device__ __forceinline__ cf shfl_cf(cf v, int src) {
    return cf{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}

struct TileGeom {
    int b, tile, nfr, span;
    long long f0, base;
    const void* yb;
    const float* span_src;
    bool interior;
};
__device__ __forceinline__ TileGeom tile_geom(const StftParams& p, int item, int hop) {
    TileGeom g;
    g.b = item / p.ntiles;
    g.tile = item - g.b * p.ntiles;
    g.f0 = (long long)g.tile * p.nf;
    g.nfr = (int)min((long long)p.nf, p.T - g.f0);
    g.span = (g.nfr - 1) * hop + kNfft;
    g.base = g.f0 * hop;  // first padded-signal index of the tile
    g.yb = p.y ? wave_row(p.y, p.y_io, (long long)g.b * p.y_bstride) : nullptr;
    g.span_src = static_cast<const float*>(g.yb) + (g.base - kNfft / 2);
    // fp32 interior tiles without a mask: one TMA bulk copy; everything else is converted / mirrored / masked per sample
    g.interior = p.fir_x == nullptr && p.y_io == DM_IO_F32 && p.mask == nullptr && g.base >= kNfft / 2 &&
                 g.base - kNfft / 2 + g.span <= p.Ly && (reinterpret_cast<uintptr_t>(g.span_src) & 15) == 0;
    return g;
}
// stage the tile's signal span: asynchronously (thread 0 issues the bulk copy) or by all threads
__device__ __forceinline__ void stage_async(const TileGeom& g, float* sig, uint64_t* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bulk_load_span(sig, g.span_src, (uint32_t)g.span * 4u, bar);
}
template <int THREADS>
__device__ __forceinline__ void stage_generic(const StftParams& p, const TileGeom& g, float* sig, int tid) {
    // four samples per thread and round, all loads issued before the first store (edge tiles are few, but a CTA that
    // walks through them sample by sample becomes the straggler of the grid)
    for (int i0 = tid; i0 < g.span; i0 += 4 * THREADS) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * THREADS;
            v[u] = 0.f;
            if (i < g.span) {
                const long long j = reflect_src(g.base + i, p.Ly);
                v[u] = ld_wave(g.yb, p.y_io, j);
                if (p.mask) v[u] *= __ldg(p.mask + j);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * THREADS;
            if (i < g.span) sig[i] = v[u];
        }
    }
}


// ---- fused scale-2 super-resolution chain (dm_stft_guidance_fir2): the tile's signal span is the sinc-resampled
// waveform y[j] = sum_k xz[2 j + k - 13] h[k] (torchaudio Resample 2 -> 1, 28 taps: operator.py:180,203-205), computed
// from x straight into the span buffer with the register-window body of resample2_fwd_reg_kernel (same arithmetic, so
// the span is bit-identical to the resampled signal the unfused chain reads back from HBM).  The span is covered in
// windows of 8 outputs: the first tile of a CTA by all threads; every following tile while the current one is being
// transformed -- the warp that owns no frame pair takes all windows but 32 per transforming warp, which those add as ONE
// round after their inverse transform (a single warp would need 13 dependent load round trips per tile and become the
// critical path: measured 44 us against 33.5 us).  Positions in the reflect padding are mirrored inside the span afterwards.
__device__ __forceinline__ float fir2_at(const float* __restrict__ xb, long long L, const float (&h)[kFir2Taps],
                                         long long j) {  // h: the by-value taps of the kernel parameters
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < kFir2Taps; ++k) {
        const long long n = 2 * j + k - kFir2Width;
        a = fmaf((n >= 0 && n < L) ? __ldg(xb + n) : 0.f, h[k], a);
    }
    return a;
}
struct Fir2Span {
    const float* xb;
    long long jb, jlo, jhi;  // signal index of span position 0 (a multiple of 32); the in-range part [jlo, jhi) of the span
    int nwin;                // windows of 8 outputs covering [jlo, jhi)
    bool mirrored;           // the span reaches into the reflect padding
};
__device__ __forceinline__ Fir2Span fir2_span(const StftParams& p, const TileGeom& g) {
    Fir2Span f;
    f.xb = p.fir_x + (long long)g.b * p.fir_x_bstride;
    f.jb = g.base - kNfft / 2;
    f.jlo = f.jb < 0 ? 0 : f.jb;  // a multiple of 8
    f.jhi = min(p.Ly, f.jb + g.span);
    f.nwin = (int)((f.jhi - f.jlo + kFir2Out - 1) / kFir2Out);
    f.mirrored = f.jb < 0 || f.jb + g.span > p.Ly;
    return f;
}
// windows wi = first, first + stride, ... < last of the span (8 outputs each, all loads of a window issued up front)
__device__ __forceinline__ void fir2_windows(const StftParams& p, const Fir2Span& f, float* __restrict__ sig, int first,
                                             int last, int stride) {
    const long long L = p.fir_L;
    const float(&h)[kFir2Taps] = p.fir_h;  // uniform / constant-bank operands of the FFMAs
    for (int wi = first; wi < last; wi += stride) {
        const long long j0 = f.jlo + (long long)wi * kFir2Out;
        const long long x0 = 2 * j0 - 16;
        float win[kFir2FwdWin], out[kFir2Out];
        if (x0 >= 0 && x0 + kFir2FwdWin <= L) {
            const float4* src = reinterpret_cast<const float4*>(f.xb + x0);
#pragma unroll
            for (int q = 0; q < kFir2FwdWin / 4; ++q) {
                const float4 v = __ldg(src + q);
                win[4 * q] = v.x, win[4 * q + 1] = v.y, win[4 * q + 2] = v.z, win[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int n = 0; n < kFir2FwdWin; ++n) {
                const long long i = x0 + n;
                win[n] = (i >= 0 && i < L) ? __ldg(f.xb + i) : 0.f;
            }
        }
        fir2_fwd8(win, h, out);
        float* dst = sig + (j0 - f.jb);
        if (j0 + kFir2Out <= f.jhi) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(out[0], out[1], out[2], out[3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(out[4], out[5], out[6], out[7]);
        } else {
#pragma unroll
            for (int c = 0; c < kFir2Out; ++c)
                if (j0 + c < f.jhi) dst[c] = out[c];
        }
    }
}
// positions of the span inside the reflect padding, after every in-range sample is in place (a barrier in between)
__device__ __forceinline__ void fir2_mirror(const StftParams& p, const TileGeom& g, const Fir2Span& f,
                                            float* __restrict__ sig, int t, int nt) {
    for (int i = t; i < g.span; i += nt) {
        const long long j = f.jb + i;
        if (j >= 0 && j < p.Ly) continue;
        const long long jr = reflect_src(g.base + i, p.Ly);  // mirrored source, inside the signal
        sig[i] = (jr >= f.jlo && jr < f.jhi) ? sig[jr - f.jb] : fir2_at(f.xb, p.fir_L, p.fir_h, jr);
    }
}

// HOP: compile-time hop (160 in every shipped configuration) or 0 = read it from the parameters
// WARPS: warps per CTA (every warp owns one frame pair of the tile, so tiles hold <= 2 * WARPS frames)
// FIR: fused scale-2 resampling (see fir2_stage_tile): tiles hold <= 2 * (WARPS - 1) frames and the last warp, which owns
//      no frame pair, computes the NEXT tile's span behind the transforms of the current one
template <int MODE, int HOP, int WARPS, bool FIR>
__global__ void __launch_bounds__(32 * WARPS, 2) stft_warp_kernel(const StftParams p, int total_tiles) {
    constexpr int kWarpCtaThreads = 32 * WARPS;
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar_sig, bar_img;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int hop = HOP ? HOP : p.hop;
    const bool has_ref = p.ref != nullptr, want_grad = p.ypbar != nullptr;
    const int na = p.tab.warp_na, nb = p.tab.warp_nb;
    const WarpSmemLayout lay = warp_smem_layout(p.nf, hop, p.tab.warp_image_floats, WARPS);
    const WarpImage il = warp_image_layout(na, nb);
    float* sig = smem + lay.sig;
    float* img = smem + lay.img;
    const f2* win2 = reinterpret_cast<const f2*>(img + il.win2);
    const f4* tw4 = reinterpret_cast<const f4*>(img + il.tw4);
    const f2* melp = reinterpret_cast<const f2*>(img + il.melp);
    const int* lanek = reinterpret_cast<const int*>(img + il.lanek);
    const PairBinTab bins{reinterpret_cast<const f2*>(img + il.binw),
                          reinterpret_cast<const unsigned char*>(img + il.binm)};
    float* red = smem + lay.red;
    float* wbuf_all = smem + lay.wbuf;
    float* wbuf = wbuf_all + w * kWarpBufFloats;
    cf* xbuf = reinterpret_cast<cf*>(wbuf);
    f2* P = reinterpret_cast<f2*>(wbuf + kWarpPOff);
    f2* melbar = reinterpret_cast<f2*>(wbuf + kWarpMelbarOff);

    // ---- prologue: the table image and the first tile's signal are fetched concurrently ----
    TileGeom g = tile_geom(p, blockIdx.x, hop);
    pdl_trigger();  // the adjoint kernel of the chain may be scheduled as soon as CTAs of this one retire
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&bar_sig)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&bar_img)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulk_load_span(img, p.tab.warp_image, (uint32_t)p.tab.warp_image_floats * 4u, &bar_img);  // constant tables
    }
    pdl_wait();  // the signal (and the zeroed cotangent buffer) of the previous kernel from here on
    if (tid == 0 && g.interior) stage_async(g, sig, &bar_sig);
    if (FIR) {
        const Fir2Span f = fir2_span(p, g);
        fir2_windows(p, f, sig, tid, f.nwin, kWarpCtaThreads);
        if (f.mirrored) {  // uniform per tile
            __syncthreads();
            fir2_mirror(p, g, f, sig, tid, kWarpCtaThreads);
        }
    } else if (!g.interior) {
        stage_generic<kWarpCtaThreads>(p, g, sig, tid);
    }
    __syncthreads();  // barrier initialisation and a generically staged span are visible
    mbar_wait_parity(&bar_img, 0);
    const int pa0 = lanek[lane], pb0 = lanek[32 + lane], ma = lanek[64 + lane], mb = lanek[96 + lane];

    uint32_t sig_parity = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        if (g.interior) {
            mbar_wait_parity(&bar_sig, sig_parity);
            sig_parity ^= 1u;
        }
        const int b = g.b, nfr = g.nfr, f0 = (int)g.f0;

        // ---- transform phase: warp w owns frames 2w, 2w + 1 of the tile ----
        float lsum = 0.f;
        const int fa = 2 * w;
        const bool producer = FIR && w == WARPS - 1;  // owns no frames (the host keeps tiles <= 2 (WARPS - 1) frames)
        const bool active = fa < nfr && !producer;
        const bool active_b = fa + 1 < nfr;  // otherwise frame B recomputes frame A and its results are dropped
        const int fb = active_b ? fa + 1 : fa;
        const int ta = f0 + fa, tb = f0 + fb;
        cf v[32];
        float ssa = 0.f, ssb = 0.f;
        if (active) warp_load_frames(lane, sig + fa * hop, sig + fb * hop, win2, v, ssa, ssb);
        __syncthreads();  // every warp holds its frames in registers: the span is dead
        const int next = item + gridDim.x;
        if (!FIR && next < total_tiles && tid == 0) {
            const TileGeom gn = tile_geom(p, next, hop);
            if (gn.interior) stage_async(gn, sig, &bar_sig);  // lands behind this tile's transforms
        }
        if (active) {
            WarpX x;  // forward spectrum of the owned bins; in PhaseWav mode overwritten by its cotangent right away
            const int partner = (32 - lane) & 31;
            // balance the two frames' magnitudes for the joint transform (warp_balance)
            float bal_s, bal_inv;
            bool zero_a, zero_b;
            warp_balance(warp_sum(ssa), warp_sum(ssb), bal_s, bal_inv, zero_a, zero_b);
            warp_scale_b(v, bal_s);
            // forward: pass 1 -> exchange -> pass 2
            dft32<-1>(v);
            warp_twiddle_store<-1>(lane, v, tw4, xbuf);
            __syncwarp();
            warp_xchg_load(lane, xbuf, v);
            dft32<-1>(v);
            if (MODE != kModePhaseWav) {
                __syncwarp();  // every lane has read its exchange row before P overwrites the buffer
                if (lane < 8) melbar[64 + lane] = f2{0.f, 0.f};
                if (lane == 1) P[kH + 1] = f2{0.f, 0.f};  // touched by the bin-pair loads with a zero weight
            }
            // split into the two frames' spectra, one owned bin per step; energies leave for P (mel modes) or meet the
            // reference magnitude right here (PhaseWav)
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                if (i < 16) {
                    const cf rcv = shfl_cf(warp_split_send_i(lane, v, i), partner);
                    warp_split_recv_i(v, i, rcv, x.a[i], x.b[i]);
                } else {
                    warp_split_nyquist(v, x.a[16], x.b[16]);
                }
                x.b[i] = cf{x.b[i].x * bal_inv, x.b[i].y * bal_inv};
                if (zero_a) x.a[i] = cf{0.f, 0.f};
                if (zero_b) x.b[i] = cf{0.f, 0.f};
                const bool mine = i < 16 || lane == 0;
                const int k = warp_bin_of(lane, i);
                f2 e = warp_bin_energies<MODE>(x.a[i], x.b[i]);
                if (MODE != kModeMelDb && p.noise != nullptr && mine) {  // GaussianNoise on |STFT| (operator.py:171)
                    const float* nz = p.noise + ((long long)b * kBins + k) * p.T;
                    e.x += p.sigma * __ldg(nz + ta);
                    e.y += p.sigma * __ldg(nz + tb);
                }
                if (MODE == kModePhaseWav) {
                    f2 gk = f2{0.f, 0.f};
                    if (mine) {
                        const long long row = ((long long)b * kBins + k) * p.T;
                        if (p.out) {
                            p.out[row + ta] = e.x;
                            if (active_b) p.out[row + tb] = e.y;
                        }
                        if (has_ref) {
                            const float* rr = p.ref + (long long)b * p.ref_bstride + (long long)k * p.T;
                            const float da = __ldg(rr + ta) - e.x, db = __ldg(rr + tb) - e.y;
                            lsum = fmaf(da, da, lsum);
                            if (active_b) lsum = fmaf(db, db, lsum);
                            gk = f2{-da, -db};
                        }
                    }
                    x.a[i] = warp_xbar<MODE>(x.a[i], gk.x);
                    x.b[i] = warp_xbar<MODE>(x.b[i], gk.y);
                } else if (mine) {
                    P[k] = e;
                }
            }
            if (MODE != kModePhaseWav) {
                float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
                if (has_ref) {  // issued before the mel loop, consumed after it
                    const float* rb = p.ref + (long long)b * p.ref_bstride;
                    const float* rlo = rb + (long long)ma * p.T;
                    const float* rhi = rb + (long long)mb * p.T;
                    r0 = __ldg(rlo + ta), r1 = __ldg(rlo + tb), r2 = __ldg(rhi + ta), r3 = __ldg(rhi + tb);
                }
                __syncwarp();
                f2 lo, hi;
                warp_mel_project(lane, na, nb, pa0, pb0, melp, P, lo, hi);
                float v0, d0, v1, d1, v2, d2, v3, d3;
                mel_value<MODE>(lo.x, p.clamp != 0, v0, d0);
                mel_value<MODE>(lo.y, p.clamp != 0, v1, d1);
                mel_value<MODE>(hi.x, p.clamp != 0, v2, d2);
                mel_value<MODE>(hi.y, p.clamp != 0, v3, d3);
                if (has_ref) {
                    r0 -= v0, r1 -= v1, r2 -= v2, r3 -= v3;
                    melbar[ma] = f2{-r0 * d0, -r1 * d1};
                    melbar[mb] = f2{-r2 * d2, -r3 * d3};
                    lsum = fmaf(r0, r0, lsum);
                    lsum = fmaf(r2, r2, lsum);
                    if (active_b) {
                        lsum = fmaf(r1, r1, lsum);
                        lsum = fmaf(r3, r3, lsum);
                    }
                }
                if (p.out) {
                    float* ob = p.out + (long long)b * kMels * p.T;
                    ob[(long long)ma * p.T + ta] = v0;
                    ob[(long long)mb * p.T + ta] = v2;
                    if (active_b) {
                        ob[(long long)ma * p.T + tb] = v1;
                        ob[(long long)mb * p.T + tb] = v3;
                    }
                }
                if (want_grad) __syncwarp();  // melbar complete
            }
            if (want_grad) {
                // backward, one owned bin per step: Xbar -> Q[k] (register i) and Q[1024 - k] (partner's register 31 - i;
                // lane 0, its own partner, needs it in register 32 - i: one step later)
                cf prev = cf{0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    cf ya = x.a[i], yb = x.b[i];
                    if (MODE != kModePhaseWav) {
                        const f2 gk = warp_bin_cotangent(lane + 32 * i, bins, melbar);
                        ya = warp_xbar<MODE>(ya, gk.x);
                        yb = warp_xbar<MODE>(yb, gk.y);
                    }
                    cf q, qm;
                    warp_q_pair(ya, yb, q, qm);
                    if (i == 0 && lane == 0) q = warp_q_real(ya, yb);
                    v[i] = q;
                    const cf rcv = shfl_cf(qm, partner);
                    if (i > 0) v[32 - i] = lane == 0 ? rcv : prev;
                    prev = rcv;
                }
                {
                    cf ya = x.a[16], yb = x.b[16];
                    if (MODE != kModePhaseWav) {
                        const f2 gk = warp_bin_cotangent(kH, bins, melbar);
                        ya = warp_xbar<MODE>(ya, gk.x);
                        yb = warp_xbar<MODE>(yb, gk.y);
                    }
                    v[16] = lane == 0 ? warp_q_real(ya, yb) : prev;
                }
                dft32<+1>(v);
                __syncwarp();  // P / melbar are dead in every lane
                warp_twiddle_store<+1>(lane, v, tw4, xbuf);
                __syncwarp();
                warp_xchg_load(lane, xbuf, v);
                dft32<+1>(v);
                __syncwarp();  // exchange rows consumed before G overwrites them
                warp_store_gradients(lane, v, win2, wbuf);
            }
        }
        if (FIR && next < total_tiles) {  // this warp's share of the NEXT tile's span (the current one is dead)
            const Fir2Span f = fir2_span(p, tile_geom(p, next, hop));
            if (producer) fir2_windows(p, f, sig, 32 * (WARPS - 1) + lane, f.nwin, 32);
            else fir2_windows(p, f, sig, 32 * w + lane, min(f.nwin, 32 * w + 32), 32);
        }
        if (p.partial) {
            lsum = warp_sum(lsum);
            if (lane == 0) red[w] = lsum;
        }
        __syncthreads();
        if (FIR && next < total_tiles) {
            const TileGeom gn = tile_geom(p, next, hop);
            const Fir2Span f = fir2_span(p, gn);
            if (f.mirrored) fir2_mirror(p, gn, f, sig, tid, kWarpCtaThreads);  // visible after the barrier below
        }

        // ---- gathered overlap-add of the tile's frame gradients, straight to HBM ----
        if (want_grad) {
            float* gout = p.ypbar + (long long)b * (p.Ly + kNfft) + g.base;
            const bool vec = (reinterpret_cast<uintptr_t>(gout) & 15) == 0;
            // frame f of the tile sits at wbuf_all + f * kWarpGStride; sample x of the span is its sample x - f * hop
            const int fstep = kWarpGStride - hop;
            for (int q = tid; q < g.span / 4; q += kWarpCtaThreads) {
                const int xq = 4 * q;
                const int f_hi = min(nfr - 1, xq / hop);
                const int f_lo = xq > kNfft - 4 ? (xq - (kNfft - 4) + hop - 1) / hop : 0;
                const float* src = wbuf_all + xq + f_lo * fstep;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                const int n = f_hi - f_lo;
                if (HOP >= 160) {  // at most 7 frames cover a sample: predicated, fully unrolled
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        if (j <= n) {
                            const float4 t = *reinterpret_cast<const float4*>(src + j * fstep);
                            s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
                        }
                    }
                } else {
                    for (int j = 0; j <= n; ++j) {
                        const float4 t = *reinterpret_cast<const float4*>(src + j * fstep);
                        s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
                    }
                }
                if (vec) {
                    atomicAdd(reinterpret_cast<float4*>(gout + xq), s);
                } else {
                    atomicAdd(gout + xq, s.x);
                    atomicAdd(gout + xq + 1, s.y);
                    atomicAdd(gout + xq + 2, s.z);
                    atomicAdd(gout + xq + 3, s.w);
                }
            }
        }
        if (p.partial && tid == 0) {
            float t = 0.f;
            for (int i = 0; i < WARPS; ++i) t += red[i];  // idle warps left 0
            p.partial[(long long)b * p.ntiles + g.tile] = t;
        }
        if (next < total_tiles) {
            g = tile_geom(p, next, hop);
            if (!FIR && !g.interior) stage_generic<kWarpCtaThreads>(p, g, sig, tid);  // the span has been dead since the barrier above
        }
        __syncthreads();  // warp buffers and `red` are free, a generically staged span is visible
    }
}

int launch_stft_warp(const StftParams& p, int mode, cudaStream_t st) {
    if (p.tab.warp_image == nullptr || p.tab.warp_image_floats <= 0 || (p.tab.warp_image_floats & 3) != 0 ||
        (reinterpret_cast<uintptr_t>(p.tab.warp_image) & 15) != 0 || p.tab.warp_na < 1 || p.tab.warp_nb < 1 ||
        warp_image_layout(p.tab.warp_na, p.tab.warp_nb).total != p.tab.warp_image_floats)
        return fail(DM_ERR_INVALID, "%s: dm_stft_tables.warp_image is missing or malformed", __func__);
    const size_t smem = stft_warp_smem_bytes(p.nf, p.hop, p.tab.warp_image_floats);
    if (smem > 113 * 1024)
        return fail(DM_ERR_UNSUPPORTED, "%s: tile needs %zu B of shared memory (2 CTAs per SM need <= 113 KB)", __func__,
                    smem);
    const long long total = (long long)p.B * p.ntiles;
    if (total > 0x7fffffffLL) return fail(DM_ERR_INVALID, "%s: too many tiles", __func__);
    const int grid = (int)min(total, (long long)2 * num_sms());
#define DM_LAUNCH_WARP(M, H, W)                                                                   \
    do {                                                                                          \
        DM_SMEM_ONCE((stft_warp_kernel<M, H, W, false>), smem);                                   \
        DM_CARVEOUT_ONCE((stft_warp_kernel<M, H, W, false>));                                     \
        launch_pdl(stft_warp_kernel<M, H, W, false>, dim3(grid), dim3(32 * W), smem, st,          \
                   g_tuning[DM_TUNE_PDL] != 0, p, (int)total);                                    \
    } while (0)
#define DM_LAUNCH_WARP_HOP(M)                                        \
    do {                                                             \
        if (p.hop == 160) DM_LAUNCH_WARP(M, 160, kWarpsPerCta);      \
        else DM_LAUNCH_WARP(M, 0, kWarpsPerCta);                     \
    } while (0)
    if (p.fir_x != nullptr) {  // fused scale-2 resampling: mel-dB guidance with hop 160 only (the shipped chain)
        if (mode != DM_STFT_MEL_DB || p.hop != 160 || p.nf > 2 * (kWarpsPerCta - 1) || p.ref == nullptr ||
            p.mask != nullptr || p.out != nullptr)
            return fail(DM_ERR_UNSUPPORTED, "%s: the fused resampling chain needs mel-dB guidance, hop 160, <= %d frames "
                        "per tile", __func__, 2 * (kWarpsPerCta - 1));
        DM_SMEM_ONCE((stft_warp_kernel<kModeMelDb, 160, kWarpsPerCta, true>), smem);
        DM_CARVEOUT_ONCE((stft_warp_kernel<kModeMelDb, 160, kWarpsPerCta, true>));
        launch_pdl(stft_warp_kernel<kModeMelDb, 160, kWarpsPerCta, true>, dim3(grid), dim3(32 * kWarpsPerCta), smem, st,
                   g_tuning[DM_TUNE_PDL] != 0, p, (int)total);
    } else if (mode == DM_STFT_MEL_DB) DM_LAUNCH_WARP_HOP(kModeMelDb);
    else if (mode == DM_STFT_PHASE_MEL) DM_LAUNCH_WARP_HOP(kModePhaseMel);
    else DM_LAUNCH_WARP_HOP(kModePhaseWav);
#undef DM_LAUNCH_WARP_HOP
#undef DM_LAUNCH_WARP
    DM_LAUNCHED();
    return DM_OK;
}

}  // namespace dm
