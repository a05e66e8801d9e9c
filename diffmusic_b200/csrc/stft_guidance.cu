// Fused STFT / mel guidance kernel for sm_100a.
//
// One CTA (256 threads = 4 frame groups of 64) owns a tile of consecutive frames of one clip:
//   * the signal span of the tile is staged once in shared memory (reflect padding, optional inpainting mask);
//   * each group runs the per-frame pipeline of stft_frame.cuh entirely out of shared memory / registers:
//     FFT -> |X|^2 -> sparse mel -> dB/clamp -> residual -> VJP -> inverse FFT;
//   * frame gradients are windowed and overlap-added in shared memory by the whole CTA, in a fixed order;
//   * the tile's padded-signal cotangent is added to HBM once (tiles overlap by < 1 frame, 2 commutative adds
//     per address -> bit-reproducible), the tile's sum of squared residuals goes to a per-tile slot.
// HBM traffic is the algorithmic minimum: signal in, reference mel in, cotangent out.  The spectrum (4.1 MB/clip in
// the reference's autograd graph) never leaves the SM.
#include "dm_common.cuh"
#include "stft_frame.cuh"
#include "stft_pair.cuh"
#include "stft_params.cuh"

namespace dm {

constexpr int kCtaThreads = 256;
constexpr int kGroups = kCtaThreads / kGroupThreads;
constexpr int kTileLd = kMels + 1;  // padded row of the ref/out tile
__host__ __device__ constexpr int tile_floats(int nf) { return (nf * kTileLd + 3) & ~3; }

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kGroupThreads)); }

template <int MODE>
__global__ void __launch_bounds__(kCtaThreads, 3) stft_guidance_kernel(const StftParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, g = tid / kGroupThreads, gt = tid % kGroupThreads;
    const int b = blockIdx.y, tile = blockIdx.x;
    const long long f0 = (long long)tile * p.nf;
    const int nfr = (int)min((long long)p.nf, p.T - f0);
    const int span = (nfr - 1) * p.hop + kNfft;
    const int span_alloc = (p.nf - 1) * p.hop + kNfft;
    const long long base = f0 * p.hop;  // first padded-signal index of the tile
    const bool has_ref = p.ref != nullptr, want_grad = p.ypbar != nullptr;

    // ---- shared memory carve-up ----
    float* sig = smem;
    float* acc = sig + span_alloc;
    float* tilebuf = acc + span_alloc;                    // [nf][kTileLd] ref (guidance) or out (transform)
    float* win = tilebuf + tile_floats(p.nf);             // [1024]
    cf* w1024 = reinterpret_cast<cf*>(win + kNfft);       // [257] (+3 pad)
    float* melw_t = reinterpret_cast<float*>(w1024 + 260);  // [mel_wstride][64] transposed banded filterbank
    float* grp = melw_t + p.tab.mel_wstride * kMels;
    float* red = grp + kGroups * kFrameSmemFloats;        // [8]
    FrameSmem s;
    {
        float* q = grp + g * kFrameSmemFloats;
        s.a_re = q; q += swz_len(kH);
        s.a_im = q; q += swz_len(kH);
        s.b_re = q; q += swz_len(kH);
        s.b_im = q; q += swz_len(kH);
        s.melbar = q; q += 72;
        s.aux = q;
    }
    StftTables tab = p.tab;
    tab.window = win;
    tab.w1024 = w1024;
    ThreadConsts tc;  // this thread's twiddles and mel band: identical for every frame it will process
    load_thread_consts(gt, p.tab, tc);
    const bool aligned8 = (p.hop & 1) == 0;

    // ---- stage the signal span, tables and the reference tile ----
    // Interior tiles (no reflection) take their contiguous span with ONE TMA bulk copy (cp.async.bulk, completion on an
    // mbarrier) issued by thread 0 while all threads fill the tables; edge tiles mirror sample by sample.
    const float* yb = static_cast<const float*>(p.y) + (long long)b * p.y_bstride;  // this kernel: fp32 only
    __shared__ __align__(8) uint64_t stage_bar;
    const float* span_src = yb + (base - kNfft / 2);
    const bool interior = base >= kNfft / 2 && base - kNfft / 2 + span <= p.Ly &&
                          (reinterpret_cast<uintptr_t>(span_src) & 15) == 0 && (span & 3) == 0;
    if (interior) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&stage_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load_span(sig, span_src, (uint32_t)span * 4u, &stage_bar);
        }
        for (int i = tid; i < span; i += kCtaThreads) acc[i] = 0.f;
    } else {
        for (int i = tid; i < span; i += kCtaThreads) {
            long long j = reflect_src(base + i, p.Ly);
            float v = __ldg(yb + j);
            if (p.mask) v *= __ldg(p.mask + j);
            sig[i] = v;
            acc[i] = 0.f;
        }
    }
    for (int i = tid; i < kNfft; i += kCtaThreads) win[i] = __ldg(p.tab.window + i);
    for (int i = tid; i < p.tab.mel_wstride * kMels; i += kCtaThreads) melw_t[i] = __ldg(p.tab.mel_w + i);
    for (int i = tid; i < 257; i += kCtaThreads) w1024[i] = p.tab.w1024[i];
    if (gt < 8) s.melbar[64 + gt] = 0.f;
    if (MODE != kModePhaseWav && has_ref) {
        const float* rb = p.ref + (long long)b * p.ref_bstride;
        for (int i = tid; i < kMels * nfr; i += kCtaThreads) {
            int m = i / nfr, f = i - m * nfr;
            tilebuf[f * kTileLd + m] = __ldg(rb + (long long)m * p.T + f0 + f);
        }
    }
    __syncthreads();  // tables / tile visible, and the mbarrier initialisation precedes every wait
    if (interior) {
        mbar_wait_parity(&stage_bar, 0);  // bulk copy landed (async proxy writes are visible after the wait)
        if (p.mask) {                     // inpainting A(x) = x * mask on the staged span
            const float* mk = p.mask + (base - kNfft / 2);
            for (int i = tid; i < span; i += kCtaThreads) sig[i] *= __ldg(mk + i);
            __syncthreads();
        }
    }

    float lsum = 0.f;
    const int rounds = (nfr + kGroups - 1) / kGroups;
    for (int r = 0; r < rounds; ++r) {
        const int f = r * kGroups + g;
        const bool active = f < nfr;
        if (active) {
            const float* frame = sig + f * p.hop;
            if (aligned8) fwd_pass1<true>(gt, frame, win, s);
            else fwd_pass1<false>(gt, frame, win, s);
            group_sync(g);
            fwd_pass2(gt, tc, s);
            group_sync(g);
            fwd_pass3(gt, tc, s);
            group_sync(g);
            fwd_unpack<MODE>(gt, w1024, s);
            group_sync(g);
            if (MODE != kModeMelDb && p.noise != nullptr) {  // GaussianNoise on the magnitude (operator.py:171)
                const long long t = f0 + f;
                for (int k = gt; k < kBins; k += kGroupThreads)
                    p_at(s, k) += p.sigma * __ldg(p.noise + ((long long)b * kBins + k) * p.T + t);
                group_sync(g);
            }
            if (MODE == kModePhaseWav) {
                const long long t = f0 + f;
                for (int k = gt; k < kBins; k += kGroupThreads) {
                    float mag = p_at(s, k);
                    if (p.out) p.out[((long long)b * kBins + k) * p.T + t] = mag;
                    if (has_ref) {
                        float d = __ldg(p.ref + (long long)b * p.ref_bstride + (long long)k * p.T + t) - mag;
                        lsum = fmaf(d, d, lsum);
                        p_at(s, k) = -d;
                    }
                }
            } else {
                float v;
                lsum += mel_residual<MODE>(gt, tc, melw_t, s, p.clamp != 0, has_ref,
                                           has_ref ? tilebuf[f * kTileLd + gt] : 0.f, &v);
                if (p.out) tilebuf[f * kTileLd + gt] = v;
            }
            if (!want_grad) {
                group_sync(g);  // P lives in a_re: everyone must be done reading it before the next frame's pass 1
            } else {
                group_sync(g);
                bwd_pack<MODE>(gt, tab, s);
                group_sync(g);
                inv_pass1(gt, s);
                group_sync(g);
                inv_pass2(gt, tc, s);
                group_sync(g);
                inv_pass3(gt, tc, s);
            }
        }
        if (want_grad) {
            __syncthreads();
            // overlap-add the (up to) 4 frame gradients of this round frame by frame (fixed order -> deterministic);
            // all 256 threads take one complex output (= two samples) per iteration, a barrier separates frames
            // because consecutive frames overlap by 1024 - hop samples
            const int fr0 = r * kGroups;
            const int nact = min(kGroups, nfr - fr0);
            for (int q = 0; q < nact; ++q) {
                const float* gq = grp + q * kFrameSmemFloats;
                float* ap = acc + (fr0 + q) * p.hop;
                for (int h = tid; h < kH; h += kCtaThreads) {
                    const int pp = swz(h);
                    const float re = gq[pp], im = gq[swz_len(kH) + pp];
                    if (aligned8) {
                        f2* a2 = reinterpret_cast<f2*>(ap) + h;
                        const f2 w2 = reinterpret_cast<const f2*>(win)[h];
                        f2 v = *a2;
                        v.x = fmaf(re, w2.x, v.x);
                        v.y = fmaf(im, w2.y, v.y);
                        *a2 = v;
                    } else {
                        ap[2 * h] = fmaf(re, win[2 * h], ap[2 * h]);
                        ap[2 * h + 1] = fmaf(im, win[2 * h + 1], ap[2 * h + 1]);
                    }
                }
                __syncthreads();
            }
        }
    }

    // ---- tile epilogue ----
    if (want_grad) {
        float* gb = p.ypbar + (long long)b * (p.Ly + kNfft) + base;
        for (int i = tid; i < span; i += kCtaThreads) atomicAdd(gb + i, acc[i]);
    }
    if (p.out && MODE != kModePhaseWav) {
        __syncthreads();
        float* ob = p.out + (long long)b * kMels * p.T;
        for (int i = tid; i < kMels * nfr; i += kCtaThreads) {
            int m = i / nfr, f = i - m * nfr;
            ob[(long long)m * p.T + f0 + f] = tilebuf[f * kTileLd + m];
        }
    }
    if (p.partial) {
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) red[tid >> 5] = lsum;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < kCtaThreads / 32; ++i) t += red[i];
            p.partial[(long long)b * p.ntiles + tile] = t;
        }
    }
}

// ======================================================================================================================
// Frame-pair kernel (stft_pair.cuh): 4 groups x 2 frames = 8 frames per round, 128-bit shared-memory FFT traffic, the
// spectrum of the owned bins in registers, and the overlap-add done straight from the registers of the last inverse
// pass.  Groups never meet at a CTA-wide barrier inside the frame loop: the only coupling is the ORDER of the
// overlap-add into the tile accumulator (frames in ascending order -> bit-reproducible), enforced by a ring of named
// barriers (group g arrives on barrier 8+g when its two frames are added; group g+1 waits on it before adding).
// Used whenever the hop is even (every shipped configuration: hop 160); odd hops keep the frame-at-a-time kernel.
constexpr int kPairFrames = 2 * kGroups;  // frames per round

__device__ __forceinline__ void chain_wait(int pred) {
    asm volatile("bar.sync %0, %1;" ::"r"(8 + pred), "r"(2 * kGroupThreads) : "memory");
}
__device__ __forceinline__ void chain_arrive(int g) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(8 + g), "r"(2 * kGroupThreads) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kCtaThreads, 2) stft_pair_kernel(const StftParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, g = tid / kGroupThreads, gt = tid % kGroupThreads;
    const int b = blockIdx.y, tile = blockIdx.x;
    const long long f0 = (long long)tile * p.nf;
    const int nfr = (int)min((long long)p.nf, p.T - f0);
    const int span = (nfr - 1) * p.hop + kNfft;
    const int span_alloc = ((p.nf - 1) * p.hop + kNfft + 3) & ~3;
    const long long base = f0 * p.hop;  // first padded-signal index of the tile
    const bool has_ref = p.ref != nullptr, want_grad = p.ypbar != nullptr;
    const int hop2 = p.hop >> 1;

    // ---- shared memory carve-up ----
    float* sig = smem;
    float* acc = sig + span_alloc;
    float* tilebuf = acc + span_alloc;                 // [nf][kTileLd] ref (guidance) or out (transform)
    float* win = tilebuf + tile_floats(p.nf);          // [1024]
    f2* binw = reinterpret_cast<f2*>(win + kNfft);     // [513] (+1 pad)
    float* grp = reinterpret_cast<float*>(binw + 514); // [kGroups][kPairSmemFloats]
    float* red = grp + kGroups * kPairSmemFloats;      // [8]
    unsigned char* binm = reinterpret_cast<unsigned char*>(red + 8);  // [513] (+3 pad)
    float* melw_t = red + 8 + 132;                     // [mel_wstride][64] banded filterbank, transposed (16-B aligned)
    PairSmem s;
    {
        float* q = grp + g * kPairSmemFloats;
        s.a = reinterpret_cast<c2*>(q);
        s.b = reinterpret_cast<c2*>(q + 4 * kH);
    }
    const PairBinTab bins{binw, binm};
    PairConsts pc;
    load_pair_consts(gt, p.tab, pc);

    // ---- stage the signal span, the window, the filterbank (by band and by bin) and the reference tile ----
    const void* yb = wave_row(p.y, p.y_io, (long long)b * p.y_bstride);
    __shared__ __align__(8) uint64_t stage_bar;
    const float* span_src = static_cast<const float*>(yb) + (base - kNfft / 2);
    // fp32 interior tiles: one TMA bulk copy; 16-bit waveforms and edge tiles: converted / mirrored sample by sample
    const bool interior = p.y_io == DM_IO_F32 && base >= kNfft / 2 && base - kNfft / 2 + span <= p.Ly &&
                          (reinterpret_cast<uintptr_t>(span_src) & 15) == 0 && (span & 3) == 0;
    if (interior) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&stage_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load_span(sig, span_src, (uint32_t)span * 4u, &stage_bar);
        }
    } else {
        for (int i = tid; i < span; i += kCtaThreads) {
            long long j = reflect_src(base + i, p.Ly);
            float v = ld_wave(yb, p.y_io, j);
            if (p.mask) v *= __ldg(p.mask + j);
            sig[i] = v;
        }
    }
    if (want_grad) {
        float4* a4 = reinterpret_cast<float4*>(acc);
        for (int i = tid; i < span_alloc / 4; i += kCtaThreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    reinterpret_cast<float4*>(win)[tid] = __ldg(reinterpret_cast<const float4*>(p.tab.window) + tid);
    {
        const float4* src = reinterpret_cast<const float4*>(p.tab.mel_w);
        float4* dst = reinterpret_cast<float4*>(melw_t);
        for (int i = tid; i < p.tab.mel_wstride * kMels / 4; i += kCtaThreads) dst[i] = __ldg(src + i);
    }
    for (int k = tid; k < kBins; k += kCtaThreads) {
        binw[k] = f2{__ldg(p.tab.bin_w0 + k), __ldg(p.tab.bin_w1 + k)};
        binm[k] = (unsigned char)__ldg(p.tab.bin_m0 + k);
    }
    if (MODE != kModePhaseWav && has_ref) {  // warp w takes bands w, w + 8, ...; lanes run along the frames (coalesced)
        const int lane = tid & 31, w = tid >> 5;
        const float* rb = p.ref + (long long)b * p.ref_bstride + f0 + (long long)w * p.T + lane;
        float* tb = tilebuf + lane * kTileLd + w;
        const long long rstep = 8 * p.T;
        if (lane < nfr)
#pragma unroll
            for (int i = 0; i < kMels / 8; ++i) tb[8 * i] = __ldg(rb + i * rstep);
        static_assert(kCtaThreads / 32 == 8, "ref-tile staging assumes 8 warps");
    }
    __syncthreads();
    if (interior) {
        mbar_wait_parity(&stage_bar, 0);
        if (p.mask) {
            const float* mk = p.mask + (base - kNfft / 2);
            for (int i = tid; i < span; i += kCtaThreads) sig[i] *= __ldg(mk + i);
            __syncthreads();
        }
    }

    float lsum = 0.f;
    const int rounds = (nfr + kPairFrames - 1) / kPairFrames;
    f2* acc2 = reinterpret_cast<f2*>(acc);
    for (int r = 0; r < rounds; ++r) {
        const int fa = r * kPairFrames + 2 * g;
        const bool active = fa < nfr;          // frame A exists
        const bool active_b = fa + 1 < nfr;    // frame B exists (otherwise B recomputes A and its results are dropped)
        const int fb = active_b ? fa + 1 : fa;
        if (active) {
            PairX x;
            pair_fwd_pass1(gt, sig + fa * p.hop, sig + fb * p.hop, win, s);
            group_sync(g);
            pair_fwd_pass2(gt, pc, s);
            group_sync(g);
            pair_fwd_pass3(gt, pc, s);
            group_sync(g);
            pair_unpack<MODE>(gt, pc, s, x);
            group_sync(g);
            f2* P = pair_energy(s);
            const long long ta = f0 + fa, tb = f0 + fb;
            if (MODE != kModeMelDb && p.noise != nullptr) {  // GaussianNoise on the magnitude (operator.py:171)
                for (int k = gt; k < kBins; k += kGroupThreads) {
                    const float* nz = p.noise + ((long long)b * kBins + k) * p.T;
                    f2 v = P[k];
                    v.x += p.sigma * __ldg(nz + ta);
                    v.y += p.sigma * __ldg(nz + tb);
                    P[k] = v;
                }
                group_sync(g);
            }
            if (MODE == kModePhaseWav) {
                for (int k = gt; k < kBins; k += kGroupThreads) {
                    const f2 mag = P[k];
                    const long long row = ((long long)b * kBins + k) * p.T;
                    if (p.out) {
                        p.out[row + ta] = mag.x;
                        if (active_b) p.out[row + tb] = mag.y;
                    }
                    if (has_ref) {
                        const float* rr = p.ref + (long long)b * p.ref_bstride + (long long)k * p.T;
                        const float da = __ldg(rr + ta) - mag.x, db = __ldg(rr + tb) - mag.y;
                        lsum = fmaf(da, da, lsum);
                        if (active_b) lsum = fmaf(db, db, lsum);
                        P[k] = f2{-da, -db};
                    }
                }
            } else {
                f2 mel = pair_mel_project(gt, pc, melw_t, s);
                group_sync(g);
                mel = pair_mel_combine(gt, mel, s);
                float va, da_, vb, db_;
                mel_value<MODE>(mel.x, p.clamp != 0, va, da_);
                mel_value<MODE>(mel.y, p.clamp != 0, vb, db_);
                if (has_ref) {
                    const float ra = tilebuf[fa * kTileLd + gt] - va, rb2 = tilebuf[fb * kTileLd + gt] - vb;
                    pair_melbar(s)[gt] = f2{-ra * da_, -rb2 * db_};
                    lsum = fmaf(ra, ra, lsum);
                    if (active_b) lsum = fmaf(rb2, rb2, lsum);
                }
                if (p.out) {
                    tilebuf[fa * kTileLd + gt] = va;
                    if (active_b) tilebuf[fb * kTileLd + gt] = vb;
                }
            }
            group_sync(g);  // P / melbar complete (and, without a gradient, consumed before the next round reuses b)
            if (want_grad) {
                pair_pack<MODE>(gt, pc, bins, s, x);
                group_sync(g);
                pair_inv_pass1(gt, s);
                group_sync(g);
                pair_inv_pass2(gt, pc, s);
                group_sync(g);
            }
        }
        if (want_grad) {
            // ordered overlap-add: ... -> (round r-1, group 3) -> (r, 0) -> (r, 1) -> (r, 2) -> (r, 3) -> (r+1, 0) ...
            // Groups without frames in the last round still pass the baton.
            cf va[8], vb[8];
            if (active) pair_inv_pass3(gt, pc, s, va, vb);
            if (r > 0 || g > 0) chain_wait((g + kGroups - 1) % kGroups);
            if (active) {
                pair_ola_add(gt, win, va, acc2 + fa * hop2);
                group_sync(g);  // frame B overlaps frame A; also: every thread has read `a` before the next pass 1 writes it
                if (active_b) pair_ola_add(gt, win, vb, acc2 + fb * hop2);
            }
            if (r + 1 < rounds || g + 1 < kGroups) chain_arrive(g);
        }
    }

    // ---- tile epilogue ----
    if (want_grad) {
        __syncthreads();
        float* gb = p.ypbar + (long long)b * (p.Ly + kNfft) + base;
        for (int i = tid; i < span; i += kCtaThreads) atomicAdd(gb + i, acc[i]);
    }
    if (p.out && MODE != kModePhaseWav) {
        __syncthreads();
        float* ob = p.out + (long long)b * kMels * p.T + f0;
        for (int m = tid >> 5; m < kMels; m += kCtaThreads / 32)
            for (int f = tid & 31; f < nfr; f += 32) ob[(long long)m * p.T + f] = tilebuf[f * kTileLd + m];
    }
    if (p.partial) {
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) red[tid >> 5] = lsum;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < kCtaThreads / 32; ++i) t += red[i];
            p.partial[(long long)b * p.ntiles + tile] = t;
        }
    }
}

static size_t stft_pair_smem_bytes(int nf, int hop, int mel_wstride) {
    size_t span = ((size_t)(nf - 1) * hop + kNfft + 3) & ~(size_t)3;
    size_t fl = 2 * span + (size_t)tile_floats(nf) + kNfft + 2 * 514 + (size_t)kGroups * kPairSmemFloats + 8 + 132 +
                (size_t)mel_wstride * kMels;
    return fl * sizeof(float);
}

// mel projection of an already materialised magnitude (PhaseRetrievalOperator.transform, operator.py:153-154):
// out[b, m, t] = clamp(sum_k fb[k, m] * mag[b, k, t], +-80) with the banded filterbank; coalesced along t.
__global__ void __launch_bounds__(128) mel_project_kernel(const float* __restrict__ mag, long long T, StftTables tab,
                                                          int clamp, float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
    const int m = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    const int k0 = tab.mel_kstart[m], n = tab.mel_klen[m];
    const float* src = mag + ((long long)b * kBins + k0) * T + t;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc = fmaf(__ldg(tab.mel_w + i * kMels + m), src[(long long)i * T], acc);
    if (clamp) acc = clamp_nan(acc, -80.f, 80.f);
    out[((long long)b * kMels + m) * T + t] = acc;
}

static size_t stft_smem_bytes(int nf, int hop, int mel_wstride) {
    size_t span = (size_t)(nf - 1) * hop + kNfft;
    size_t fl = 2 * span + (size_t)tile_floats(nf) + kNfft + 2 * 260 + (size_t)mel_wstride * kMels +
                (size_t)kGroups * kFrameSmemFloats + 8;
    return fl * sizeof(float);
}

}  // namespace dm

using namespace dm;

static int g_stft_engine = DM_STFT_ENGINE_AUTO;

extern "C" int dm_stft_set_engine(int engine) {
    DM_REQUIRE(engine == DM_STFT_ENGINE_AUTO || engine == DM_STFT_ENGINE_FRAME || engine == DM_STFT_ENGINE_PAIR);
    g_stft_engine = engine;
    return DM_OK;
}

extern "C" int dm_stft_num_tiles(long long Ly, int hop, int frames_per_tile) {
    if (Ly <= 0 || hop <= 0 || frames_per_tile <= 0) return DM_ERR_INVALID;
    long long T = 1 + Ly / hop;
    return (int)((T + frames_per_tile - 1) / frames_per_tile);
}

extern "C" int dm_stft_guidance(const dm_stft_tables* tab, int mode, int clamp, int hop, const float* y,
                                long long y_bstride, long long Ly, const float* mask, int B, const float* ref,
                                long long ref_bstride, const float* noise, float sigma, float* out, float* ypbar,
                                float* partial, int frames_per_tile, dm_stream_t stream) {
    return dm_stft_guidance_io(tab, mode, clamp, hop, y, DM_IO_F32, y_bstride, Ly, mask, B, ref, ref_bstride, noise,
                               sigma, out, ypbar, partial, frames_per_tile, stream);
}

extern "C" int dm_stft_guidance_io(const dm_stft_tables* tab, int mode, int clamp, int hop, const void* y, int y_dtype,
                                   long long y_bstride, long long Ly, const float* mask, int B, const float* ref,
                                   long long ref_bstride, const float* noise, float sigma, float* out, float* ypbar,
                                   float* partial, int frames_per_tile, dm_stream_t stream) {
    DM_REQUIRE(tab != nullptr && y != nullptr && io_dtype_ok(y_dtype));
    DM_REQUIRE(y_dtype == DM_IO_F32 || (hop & 1) == 0);  // 16-bit waveforms: frame-pair kernel only
    DM_REQUIRE(mode >= 0 && mode <= 2);
    DM_REQUIRE(B > 0 && Ly > kNfft / 2);  // reflect padding needs pad < length (torch.stft raises otherwise)
    DM_REQUIRE(hop > 0 && hop <= kNfft);
    DM_REQUIRE(frames_per_tile >= 1 && frames_per_tile <= 64);
    DM_REQUIRE((ref != nullptr) != (out != nullptr));  // guidance mode xor transform mode
    DM_REQUIRE(ypbar == nullptr || ref != nullptr);
    DM_REQUIRE(ref == nullptr || partial != nullptr);
    DM_REQUIRE(noise == nullptr || mode != DM_STFT_MEL_DB);
    StftParams p;
    p.tab = StftTables{tab->window, reinterpret_cast<const cf*>(tab->tw512), reinterpret_cast<const cf*>(tab->w1024),
                       tab->mel_kstart, tab->mel_klen, tab->mel_w, tab->mel_wstride, tab->bin_m0, tab->bin_w0,
                       tab->bin_w1, tab->warp_image, tab->warp_image_floats, tab->warp_na, tab->warp_nb};
    p.clamp = clamp;
    p.hop = hop;
    p.B = B;
    p.nf = frames_per_tile;
    p.Ly = Ly;
    p.T = 1 + Ly / hop;
    p.ntiles = (int)((p.T + p.nf - 1) / p.nf);
    p.y_bstride = y_bstride;
    p.ref_bstride = ref_bstride;
    p.y = y;
    p.y_io = y_dtype;
    p.mask = mask;
    p.ref = ref;
    p.noise = noise;
    p.sigma = sigma;
    p.out = out;
    p.ypbar = ypbar;
    p.partial = partial;
    DM_REQUIRE(tab->mel_wstride >= 1 && tab->mel_wstride <= 64);
    dim3 grid(p.ntiles, B), block(kCtaThreads);
    cudaStream_t st = as_stream(stream);
    // warp-per-frame-pair engine (stft_warp.cu): hops that keep the gathered overlap-add 16-byte aligned, one round per tile
    if ((hop & 3) == 0 && p.nf <= 16 && g_stft_engine == DM_STFT_ENGINE_AUTO && tab->warp_image != nullptr)
        return launch_stft_warp(p, mode, st);
    if ((hop & 1) == 0 && (g_stft_engine != DM_STFT_ENGINE_FRAME || y_dtype != DM_IO_F32)) {  // frame-pair kernel
        size_t smem2 = stft_pair_smem_bytes(p.nf, hop, tab->mel_wstride);
        if (smem2 > 227 * 1024)
            return fail(DM_ERR_UNSUPPORTED, "%s: tile needs %zu B of shared memory", __func__, smem2);
#define DM_LAUNCH_PAIR(M)                                       \
    do {                                                        \
        DM_SMEM_ONCE(stft_pair_kernel<M>, smem2);               \
        DM_CARVEOUT_ONCE(stft_pair_kernel<M>);                  \
        stft_pair_kernel<M><<<grid, block, smem2, st>>>(p);     \
    } while (0)
        if (mode == DM_STFT_MEL_DB) DM_LAUNCH_PAIR(kModeMelDb);
        else if (mode == DM_STFT_PHASE_MEL) DM_LAUNCH_PAIR(kModePhaseMel);
        else DM_LAUNCH_PAIR(kModePhaseWav);
#undef DM_LAUNCH_PAIR
        DM_LAUNCHED();
        return DM_OK;
    }
    size_t smem = stft_smem_bytes(p.nf, hop, tab->mel_wstride);
    if (smem > 227 * 1024) return fail(DM_ERR_UNSUPPORTED, "%s: tile needs %zu B of shared memory", __func__, smem);
#define DM_LAUNCH_STFT(M)                                          \
    do {                                                           \
        DM_SMEM_ONCE(stft_guidance_kernel<M>, smem);               \
        DM_CARVEOUT_ONCE(stft_guidance_kernel<M>);                 \
        stft_guidance_kernel<M><<<grid, block, smem, st>>>(p);     \
    } while (0)
    if (mode == DM_STFT_MEL_DB) DM_LAUNCH_STFT(kModeMelDb);
    else if (mode == DM_STFT_PHASE_MEL) DM_LAUNCH_STFT(kModePhaseMel);
    else DM_LAUNCH_STFT(kModePhaseWav);
#undef DM_LAUNCH_STFT
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_stft_guidance_fir2(const dm_stft_tables* tab, int clamp, int hop, const float* x,
                                     long long x_bstride, long long L, const float* taps, int B, const float* ref,
                                     long long ref_bstride, float* ypbar, float* partial, int frames_per_tile,
                                     dm_stream_t stream) {
    DM_REQUIRE(tab != nullptr && x != nullptr && taps != nullptr && ref != nullptr && partial != nullptr);
    DM_REQUIRE(tab->warp_image != nullptr && hop == 160);
    DM_REQUIRE(B > 0 && L > kNfft);  // Ly = ceil(L / 2) > 512: reflect padding needs pad < length
    DM_REQUIRE(frames_per_tile >= 1 && frames_per_tile <= 14);
    DM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_bstride & 3) == 0);
    StftParams p;
    p.tab = StftTables{tab->window, reinterpret_cast<const cf*>(tab->tw512), reinterpret_cast<const cf*>(tab->w1024),
                       tab->mel_kstart, tab->mel_klen, tab->mel_w, tab->mel_wstride, tab->bin_m0, tab->bin_w0,
                       tab->bin_w1, tab->warp_image, tab->warp_image_floats, tab->warp_na, tab->warp_nb};
    p.clamp = clamp;
    p.hop = hop;
    p.B = B;
    p.nf = frames_per_tile;
    p.Ly = (L + 1) / 2;
    p.T = 1 + p.Ly / hop;
    p.ntiles = (int)((p.T + p.nf - 1) / p.nf);
    p.y_bstride = 0;
    p.ref_bstride = ref_bstride;
    p.y = nullptr;
    p.y_io = DM_IO_F32;
    p.mask = nullptr;
    p.ref = ref;
    p.noise = nullptr;
    p.sigma = 0.f;
    p.out = nullptr;
    p.ypbar = ypbar;
    p.partial = partial;
    p.fir_x = x;
    p.fir_x_bstride = x_bstride;
    p.fir_L = L;
    for (int k = 0; k < 28; ++k) p.fir_h[k] = taps[k];  // HOST array
    return launch_stft_warp(p, DM_STFT_MEL_DB, as_stream(stream));
}

extern "C" int dm_mel_project(const dm_stft_tables* tab, const float* mag, int B, long long T, int clamp, float* out,
                              dm_stream_t stream) {
    DM_REQUIRE(tab && mag && out && B > 0 && T > 0);
    StftTables t{tab->window, reinterpret_cast<const cf*>(tab->tw512), reinterpret_cast<const cf*>(tab->w1024),
                 tab->mel_kstart, tab->mel_klen, tab->mel_w, tab->mel_wstride, tab->bin_m0, tab->bin_w0, tab->bin_w1};
    mel_project_kernel<<<dim3((unsigned)((T + 127) / 128), kMels, B), 128, 0, as_stream(stream)>>>(mag, T, t, clamp,
                                                                                                 out);
    DM_LAUNCHED();
    return DM_OK;
}
