// Fused STFT / mel guidance kernel for sm_100a, warp-per-frame-pair engine (stft_warp.cuh).
//
// A CTA of 8 warps walks over tiles of <= 16 consecutive frames of one clip (persistent: grid = 2 CTAs per SM, tile
// list strided over the grid; the tables are staged once per CTA).  Per tile:
//   * the tile's signal span is staged in shared memory (one cp.async.bulk for interior fp32 tiles; reflect padding,
//     16-bit waveforms and the inpainting mask sample by sample);
//   * warp w owns frames 2w, 2w+1 and runs the whole chain window -> FFT -> energies -> sparse mel -> dB / clamp ->
//     residual -> VJP -> inverse FFT out of registers and its private 8.5 KB buffer; warps only meet at the CTA
//     barrier that ends the tile's transform phase (no named-barrier ring, no ordered chain);
//   * gathered overlap-add: every warp leaves its two windowed frame gradients in its buffer, then all 256 threads sum,
//     per 4 output samples, the <= 7 frames that cover them in ascending frame order (bit-reproducible) and add the
//     result to the padded cotangent in HBM (tiles overlap by < 1 frame -> <= 2 commutative adds per address).
// Same inputs / outputs / arithmetic as stft_pair_kernel (stft_guidance.cu), which stays as the engine for hops that
// are not a multiple of 4 and tiles of more than 16 frames, and as the A/B reference of the tests.
#include "stft_params.cuh"
#include "stft_warp.cuh"

namespace dm {

constexpr int kWarpCtaThreads = 256;
constexpr int kWarpsPerCta = kWarpCtaThreads / 32;
constexpr int kWarpMaxFrames = 2 * kWarpsPerCta;

struct WarpSmemLayout {
    int sig, win2, tw4, melw, binw, binm, red, wbuf, total;  // float offsets
};
__host__ __device__ inline WarpSmemLayout warp_smem_layout(int nf, int hop, int mel_wstride) {
    WarpSmemLayout l;
    const int span = ((nf - 1) * hop + kNfft + 3) & ~3;
    l.sig = 0;
    l.win2 = l.sig + span;
    l.tw4 = l.win2 + kNfft;
    l.melw = l.tw4 + 16 * 32 * 4;
    l.binw = l.melw + mel_wstride * kMels;
    l.binm = l.binw + 2 * 514;
    l.red = l.binm + 132;
    l.wbuf = l.red + 8;
    l.total = l.wbuf + kWarpsPerCta * kWarpBufFloats;
    return l;
}
size_t stft_warp_smem_bytes(int nf, int hop, int mel_wstride) {
    return (size_t)warp_smem_layout(nf, hop, mel_wstride).total * sizeof(float);
}

__device__ __forceinline__ cf shfl_cf(cf v, int src) {
    return cf{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}

template <int MODE>
__global__ void __launch_bounds__(kWarpCtaThreads, 2) stft_warp_kernel(const StftParams p, int total_tiles) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t stage_bar;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const bool has_ref = p.ref != nullptr, want_grad = p.ypbar != nullptr;
    const WarpSmemLayout lay = warp_smem_layout(p.nf, p.hop, p.tab.mel_wstride);
    float* sig = smem + lay.sig;
    f2* win2 = reinterpret_cast<f2*>(smem + lay.win2);
    f4* tw4 = reinterpret_cast<f4*>(smem + lay.tw4);
    float* melw_t = smem + lay.melw;
    f2* binw = reinterpret_cast<f2*>(smem + lay.binw);
    unsigned char* binm = reinterpret_cast<unsigned char*>(smem + lay.binm);
    float* red = smem + lay.red;
    float* wbuf_all = smem + lay.wbuf;
    float* wbuf = wbuf_all + w * kWarpBufFloats;
    cf* xbuf = reinterpret_cast<cf*>(wbuf);
    f2* P = reinterpret_cast<f2*>(wbuf + kWarpPOff);
    f2* melbar = reinterpret_cast<f2*>(wbuf + kWarpMelbarOff);
    const PairBinTab bins{binw, binm};

    // ---- once per CTA: window (halved, paired), exchange twiddles, filterbank by band and by bin ----
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&stage_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < kH; i += kWarpCtaThreads)
        win2[i] = f2{0.5f * __ldg(p.tab.window + i), 0.5f * __ldg(p.tab.window + i + kH)};
    for (int i = tid; i < 16 * 32; i += kWarpCtaThreads) {
        const int m = i >> 5, l = i & 31;
        const cf a = w1024_any(p.tab.w1024, l * (2 * m)), b = w1024_any(p.tab.w1024, l * (2 * m + 1));
        tw4[i] = f4{a.x, a.y, b.x, b.y};
    }
    {
        const float4* src = reinterpret_cast<const float4*>(p.tab.mel_w);
        float4* dst = reinterpret_cast<float4*>(melw_t);
        for (int i = tid; i < p.tab.mel_wstride * kMels / 4; i += kWarpCtaThreads) dst[i] = __ldg(src + i);
    }
    for (int k = tid; k < kBins; k += kWarpCtaThreads) {
        binw[k] = f2{__ldg(p.tab.bin_w0 + k), __ldg(p.tab.bin_w1 + k)};
        binm[k] = (unsigned char)__ldg(p.tab.bin_m0 + k);
    }
    WarpMelConsts mc;
    load_warp_mel_consts(lane, p.tab, mc);
    __syncthreads();

    uint32_t bar_parity = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int b = item / p.ntiles, tile = item - b * p.ntiles;
        const long long f0 = (long long)tile * p.nf;
        const int nfr = (int)min((long long)p.nf, p.T - f0);
        const int span = (nfr - 1) * p.hop + kNfft;
        const long long base = f0 * p.hop;  // first padded-signal index of the tile

        // ---- stage the signal span ----
        const void* yb = wave_row(p.y, p.y_io, (long long)b * p.y_bstride);
        const float* span_src = static_cast<const float*>(yb) + (base - kNfft / 2);
        const bool interior = p.y_io == DM_IO_F32 && base >= kNfft / 2 && base - kNfft / 2 + span <= p.Ly &&
                              (reinterpret_cast<uintptr_t>(span_src) & 15) == 0;
        if (interior) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_load_span(sig, span_src, (uint32_t)span * 4u, &stage_bar);
            }
        } else {
            for (int i = tid; i < span; i += kWarpCtaThreads) {
                const long long j = reflect_src(base + i, p.Ly);
                float v = ld_wave(yb, p.y_io, j);
                if (p.mask) v *= __ldg(p.mask + j);
                sig[i] = v;
            }
        }
        __syncthreads();
        if (interior) {
            mbar_wait_parity(&stage_bar, bar_parity);
            bar_parity ^= 1u;
            if (p.mask) {
                const float* mk = p.mask + (base - kNfft / 2);
                for (int i = tid; i < span; i += kWarpCtaThreads) sig[i] *= __ldg(mk + i);
                __syncthreads();
            }
        }

        // ---- transform phase: warp w owns frames 2w, 2w + 1 of the tile ----
        float lsum = 0.f;
        const int fa = 2 * w;
        if (fa < nfr) {
            const bool active_b = fa + 1 < nfr;  // otherwise frame B recomputes frame A and its results are dropped
            const int fb = active_b ? fa + 1 : fa;
            const long long ta = f0 + fa, tb = f0 + fb;
            cf v[32];
            WarpX x;
            f2 g[17];
            // forward: pass 1 -> exchange -> pass 2 -> split
            warp_load_frames(lane, sig + fa * p.hop, sig + fb * p.hop, win2, v);
            dft32<-1>(v);
            warp_twiddle_store<-1>(lane, v, tw4, xbuf);
            __syncwarp();
            warp_xchg_load(lane, xbuf, v);
            dft32<-1>(v);
            {
                cf snd[16], rcv[16];
                warp_split_send(lane, v, snd);
                const int partner = (32 - lane) & 31;
#pragma unroll
                for (int i = 0; i < 16; ++i) rcv[i] = shfl_cf(snd[i], partner);
                warp_split_recv(lane, v, rcv, x);
            }
            f2 e[17];
            warp_energies<MODE>(x, e);
            if (MODE != kModeMelDb && p.noise != nullptr) {  // GaussianNoise on the magnitude (operator.py:171)
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    const int k = (i < 16) ? lane + 32 * i : kH;
                    if (i < 16 || lane == 0) {
                        const float* nz = p.noise + ((long long)b * kBins + k) * p.T;
                        e[i].x += p.sigma * __ldg(nz + ta);
                        e[i].y += p.sigma * __ldg(nz + tb);
                    }
                }
            }
            if (MODE == kModePhaseWav) {
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    const int k = (i < 16) ? lane + 32 * i : kH;
                    g[i] = f2{0.f, 0.f};
                    if (i < 16 || lane == 0) {
                        const long long row = ((long long)b * kBins + k) * p.T;
                        if (p.out) {
                            p.out[row + ta] = e[i].x;
                            if (active_b) p.out[row + tb] = e[i].y;
                        }
                        if (has_ref) {
                            const float* rr = p.ref + (long long)b * p.ref_bstride + (long long)k * p.T;
                            const float da = __ldg(rr + ta) - e[i].x, db = __ldg(rr + tb) - e[i].y;
                            lsum = fmaf(da, da, lsum);
                            if (active_b) lsum = fmaf(db, db, lsum);
                            g[i] = f2{-da, -db};
                        }
                    }
                }
            } else {
                __syncwarp();  // every lane has read its exchange row before P overwrites the buffer
                warp_store_energies(lane, e, P);
                if (lane < 8) melbar[64 + lane] = f2{0.f, 0.f};
                float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
                if (has_ref) {  // issued before the mel loop, consumed after it
                    const float* rb = p.ref + (long long)b * p.ref_bstride;
                    const float* rlo = rb + (long long)lane * p.T;
                    const float* rhi = rb + (long long)(63 - lane) * p.T;
                    r0 = __ldg(rlo + ta), r1 = __ldg(rlo + tb), r2 = __ldg(rhi + ta), r3 = __ldg(rhi + tb);
                }
                __syncwarp();
                f2 lo, hi;
                warp_mel_project(lane, mc, melw_t, P, lo, hi);
                float v0, d0, v1, d1, v2, d2, v3, d3;
                mel_value<MODE>(lo.x, p.clamp != 0, v0, d0);
                mel_value<MODE>(lo.y, p.clamp != 0, v1, d1);
                mel_value<MODE>(hi.x, p.clamp != 0, v2, d2);
                mel_value<MODE>(hi.y, p.clamp != 0, v3, d3);
                if (has_ref) {
                    r0 -= v0, r1 -= v1, r2 -= v2, r3 -= v3;
                    melbar[lane] = f2{-r0 * d0, -r1 * d1};
                    melbar[63 - lane] = f2{-r2 * d2, -r3 * d3};
                    lsum = fmaf(r0, r0, lsum);
                    lsum = fmaf(r2, r2, lsum);
                    if (active_b) {
                        lsum = fmaf(r1, r1, lsum);
                        lsum = fmaf(r3, r3, lsum);
                    }
                }
                if (p.out) {
                    float* ob = p.out + (long long)b * kMels * p.T;
                    ob[(long long)lane * p.T + ta] = v0;
                    ob[(long long)(63 - lane) * p.T + ta] = v2;
                    if (active_b) {
                        ob[(long long)lane * p.T + tb] = v1;
                        ob[(long long)(63 - lane) * p.T + tb] = v3;
                    }
                }
                if (want_grad) {
                    __syncwarp();
                    warp_bin_cotangents(lane, bins, melbar, g);
                }
            }
            if (want_grad) {
                // backward: Q -> pass 1 -> exchange -> pass 2 -> windowed frame gradients in the warp buffer
                {
                    cf snd[16], rcv[16];
                    warp_pack_send<MODE>(lane, x, g, v, snd);
                    const int partner = (32 - lane) & 31;
#pragma unroll
                    for (int i = 0; i < 16; ++i) rcv[i] = shfl_cf(snd[i], partner);
                    warp_pack_recv<MODE>(lane, x, g, rcv, v);
                }
                dft32<+1>(v);
                __syncwarp();  // P / melbar are dead in every lane
                warp_twiddle_store<+1>(lane, v, tw4, xbuf);
                __syncwarp();
                warp_xchg_load(lane, xbuf, v);
                dft32<+1>(v);
                __syncwarp();  // exchange rows consumed before G overwrites them
                warp_store_gradients(lane, v, win2, wbuf + kWarpGOff);
            }
        }
        if (p.partial) {
            lsum = warp_sum(lsum);
            if (lane == 0) red[w] = lsum;
        }
        __syncthreads();

        // ---- gathered overlap-add of the tile's frame gradients, straight to HBM ----
        if (want_grad) {
            float* gb = p.ypbar + (long long)b * (p.Ly + kNfft) + base;
            const bool vec = (reinterpret_cast<uintptr_t>(gb) & 15) == 0;
            for (int q = tid; q < span / 4; q += kWarpCtaThreads) {
                const int xq = 4 * q;
                const int f_hi = min(nfr - 1, xq / p.hop);
                const int f_lo = xq > kNfft - 4 ? (xq - (kNfft - 4) + p.hop - 1) / p.hop : 0;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int f = f_lo; f <= f_hi; ++f) {
                    const float4 t = *reinterpret_cast<const float4*>(wbuf_all + (f >> 1) * kWarpBufFloats + kWarpGOff +
                                                                      (f & 1) * kNfft + (xq - f * p.hop));
                    s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
                }
                if (vec) {
                    atomicAdd(reinterpret_cast<float4*>(gb + xq), s);
                } else {
                    atomicAdd(gb + xq, s.x);
                    atomicAdd(gb + xq + 1, s.y);
                    atomicAdd(gb + xq + 2, s.z);
                    atomicAdd(gb + xq + 3, s.w);
                }
            }
        }
        if (p.partial && tid == 0) {
            float t = 0.f;
            for (int i = 0; i < kWarpsPerCta; ++i) t += red[i];  // idle warps left 0
            p.partial[(long long)b * p.ntiles + tile] = t;
        }
        __syncthreads();  // warp buffers, the signal span and `red` are free for the next tile
    }
}

int launch_stft_warp(const StftParams& p, int mode, cudaStream_t st) {
    const size_t smem = stft_warp_smem_bytes(p.nf, p.hop, p.tab.mel_wstride);
    if (smem > 113 * 1024)
        return fail(DM_ERR_UNSUPPORTED, "%s: tile needs %zu B of shared memory (2 CTAs per SM need <= 113 KB)", __func__,
                    smem);
    const long long total = (long long)p.B * p.ntiles;
    if (total > 0x7fffffffLL) return fail(DM_ERR_INVALID, "%s: too many tiles", __func__);
    const int grid = (int)min(total, (long long)2 * num_sms());
#define DM_LAUNCH_WARP(M)                                                                         \
    do {                                                                                          \
        DM_SMEM_ONCE(stft_warp_kernel<M>, smem);                                                  \
        DM_CARVEOUT_ONCE(stft_warp_kernel<M>);                                                    \
        stft_warp_kernel<M><<<grid, kWarpCtaThreads, smem, st>>>(p, (int)total);                  \
    } while (0)
    if (mode == DM_STFT_MEL_DB) DM_LAUNCH_WARP(kModeMelDb);
    else if (mode == DM_STFT_PHASE_MEL) DM_LAUNCH_WARP(kModePhaseMel);
    else DM_LAUNCH_WARP(kModePhaseWav);
#undef DM_LAUNCH_WARP
    DM_LAUNCHED();
    return DM_OK;
}

}  // namespace dm
