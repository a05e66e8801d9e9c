// Fused scheduler updates on the latent.
//
// DDIM / DPS / MPGD are pure elementwise passes: 128-bit vectorised, one read of each input, one write of each output.
// DSG / DiffMusic need per-clip norms: one thread-block CLUSTER (8 CTAs) per clip, partial sums exchanged through
// distributed shared memory, the data-dependent slerp branch resolved on the device (no host sync) and the second
// sweep served from L2.  Multiplications / additions are kept un-fused (__fmul_rn/__fadd_rn) where the reference's
// eager torch ops round after every step, so results track the reference to the last bit or two.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "dm_common.cuh"

namespace cg = cooperative_groups;

namespace dm {

constexpr int kThreads = 256;
constexpr int kCluster = 8;

template <int W>
__device__ __forceinline__ void ldv(const float* __restrict__ p, long long v, float (&o)[W]) {
    if (W == 4) {
        float4 t = reinterpret_cast<const float4*>(p)[v];
        o[0] = t.x;
        o[1 % W] = t.y;
        o[2 % W] = t.z;
        o[3 % W] = t.w;
    } else {
        o[0] = p[v];
    }
}
template <int W>
__device__ __forceinline__ void stv(float* __restrict__ p, long long v, const float (&o)[W]) {
    if (W == 4) {
        reinterpret_cast<float4*>(p)[v] = make_float4(o[0], o[1 % W], o[2 % W], o[3 % W]);
    } else {
        p[v] = o[0];
    }
}

// Latent-typed tensors (what the pipeline hands in and gets back: sample, model_output, prev_sample,
// pred_original_sample) may be fp32, fp16 or bf16 (the reference pipelines run in fp16, run.py:218); everything the
// step keeps to itself (x0 for the guidance, the gradient, the noise) and all arithmetic stay fp32.
template <int IO, int W>
__device__ __forceinline__ void ldio(const void* __restrict__ p, long long v, float (&o)[W]) {
    if (IO == DM_IO_F32) {
        ldv<W>(static_cast<const float*>(p), v, o);
    } else if (W == 4) {
        const uint2 raw = reinterpret_cast<const uint2*>(p)[v];
        if (IO == DM_IO_F16) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            o[0] = a.x; o[1 % W] = a.y; o[2 % W] = b.x; o[3 % W] = b.y;
        } else {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
            o[0] = a.x; o[1 % W] = a.y; o[2 % W] = b.x; o[3 % W] = b.y;
        }
    } else {
        o[0] = IO == DM_IO_F16 ? __half2float(static_cast<const __half*>(p)[v])
                               : __bfloat162float(static_cast<const __nv_bfloat16*>(p)[v]);
    }
}
template <int IO, int W>
__device__ __forceinline__ void stio(void* __restrict__ p, long long v, const float (&o)[W]) {
    if (IO == DM_IO_F32) {
        stv<W>(static_cast<float*>(p), v, o);
    } else if (W == 4) {
        uint2 raw;
        if (IO == DM_IO_F16) {
            *reinterpret_cast<__half2*>(&raw.x) = __floats2half2_rn(o[0], o[1 % W]);
            *reinterpret_cast<__half2*>(&raw.y) = __floats2half2_rn(o[2 % W], o[3 % W]);
        } else {
            *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(o[0], o[1 % W]);
            *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(o[2 % W], o[3 % W]);
        }
        reinterpret_cast<uint2*>(p)[v] = raw;
    } else if (IO == DM_IO_F16) {
        static_cast<__half*>(p)[v] = __float2half_rn(o[0]);
    } else {
        static_cast<__nv_bfloat16*>(p)[v] = __float2bfloat16_rn(o[0]);
    }
}

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }

// ------------------------------------------------------------------------------------------------ elementwise
struct EwParams {
    const void* x;    // latent-typed
    const float* x0;
    const void* eps;  // latent-typed
    const void* g0;   // latent-typed: dLoss/d(leaf) as autograd returns it, leaf = leaf_scale * x0 (see x0_leaf)
    const float* z;
    void* prev;       // latent-typed
    float* x0_out;    // fp32 x0 (kX0) / guided x0 (kMpgd) kept by the step
    void* x0_pub;     // optional latent-typed copy of x0_out for the caller (pred_original_sample)
    void* x0_leaf;    // kX0, optional latent-typed: leaf_scale * x0 = the VAE decoder input 1/scaling_factor * x0
                      // (scheduling_dps.py:195-197), so torch needs no scaling kernel (nor its backward)
    float leaf_scale; // kX0: the factor above; updates: dLoss/dx0 = leaf_scale * g0 (chain rule of that scaling)
    const float* losses;  // optional per-clip losses (n_losses of them) and where to put their 2-norm: the 0-d `loss` of
    int n_losses;         // the reference (torch.linalg.norm over the whole batch, scheduling_dps.py:211) without a
    float* loss_total;    // reduction kernel of its own
    float sqrt_a, sqrt_b, sqrt_p, dir_coef, std, rate, clip_range;
    int clip;
    const float* coef;  // optional device-resident [sqrt_a, sqrt_b, sqrt_p, dir_coef, std, r]: overrides the by-value
                        // scalars so a captured CUDA graph can be replayed for every timestep
};

// batch loss = sqrt(sum_b loss_b^2) by one thread (B <= a few hundred); NaN / Inf propagate
__device__ __forceinline__ void write_loss_total(const float* __restrict__ losses, int n, float* __restrict__ out) {
    float s = 0.f;
    for (int b = 0; b < n; ++b) s = fmaf(losses[b], losses[b], s);
    *out = n == 1 ? losses[0] : sqrtf(s);
}

__device__ __forceinline__ void load_coef(const float* __restrict__ c, float& sqrt_a, float& sqrt_b, float& sqrt_p,
                                          float& dir_coef, float& std) {
    if (c != nullptr) {
        sqrt_a = __ldg(c + 0);
        sqrt_b = __ldg(c + 1);
        sqrt_p = __ldg(c + 2);
        dir_coef = __ldg(c + 3);
        std = __ldg(c + 4);
    }
}

enum EwKind { kX0 = 0, kDdim = 1, kDps = 2, kMpgd = 3 };

// one vector of the elementwise step, split into its load and its compute / store half so that the kernel can keep the
// loads of TWO grid-stride iterations in flight per thread (these kernels are pure HBM streams: bytes in flight per SM
// decide the achieved bandwidth)
template <int KIND, int W, int IO>
struct EwVec {
    float x[W], a[W], g[W], z[W];
    bool has_z;
    __device__ __forceinline__ void load(const EwParams& p, long long v) {
        ldio<IO, W>(p.x, v, x);
        if (KIND == kX0) {
            ldio<IO, W>(p.eps, v, a);
            return;
        }
        ldv<W>(p.x0, v, a);
        if (KIND != kDdim) ldio<IO, W>(p.g0, v, g);
        has_z = (KIND != kDdim) && p.z != nullptr;
        if (has_z) ldv<W>(p.z, v, z);
    }
    __device__ __forceinline__ void finish(const EwParams& p, long long v) {
        float o[W], o2[W];
        if (KIND == kX0) {
#pragma unroll
            for (int i = 0; i < W; ++i) {
                float t = dvd(sub(x[i], mul(p.sqrt_b, a[i])), p.sqrt_a);
                if (p.clip) t = t < -p.clip_range ? -p.clip_range : (t > p.clip_range ? p.clip_range : t);  // NaN-preserving
                o[i] = t;
            }
            stv<W>(p.x0_out, v, o);
            if (IO != DM_IO_F32 && p.x0_pub != nullptr) stio<IO, W>(p.x0_pub, v, o);
            if (p.x0_leaf != nullptr) {
#pragma unroll
                for (int i = 0; i < W; ++i) o2[i] = mul(p.leaf_scale, o[i]);
                stio<IO, W>(p.x0_leaf, v, o2);
            }
            return;
        }
        if (KIND != kDdim) {
#pragma unroll
            for (int i = 0; i < W; ++i) g[i] = mul(g[i], p.leaf_scale);  // dLoss/dx0 (autograd of leaf_scale * x0)
        }
#pragma unroll
        for (int i = 0; i < W; ++i) {
            float x0 = a[i];
            if (KIND == kMpgd) x0 = sub(x0, mul(p.rate, g[i]));                 // scheduling_mpgd.py:199-200
            float e = dvd(sub(x[i], mul(p.sqrt_a, x0)), p.sqrt_b);              // noise_pred
            float prev = add(mul(p.sqrt_p, x0), mul(p.dir_coef, e));
            if (has_z) prev = add(prev, mul(p.std, z[i]));
            if (KIND == kDps) prev = sub(prev, mul(p.rate, dvd(g[i], p.sqrt_a)));  // scheduling_dps.py:212-213
            o[i] = prev;
            o2[i] = x0;
        }
        stio<IO, W>(p.prev, v, o);
        if (KIND == kMpgd) {
            if (IO == DM_IO_F32 || p.x0_pub == nullptr) stv<W>(p.x0_out, v, o2);
            else stio<IO, W>(p.x0_pub, v, o2);
        }
    }
};

template <int KIND, int W, int IO>
__global__ void __launch_bounds__(kThreads) ew_update_kernel(EwParams p, long long nvec) {
    load_coef(p.coef, p.sqrt_a, p.sqrt_b, p.sqrt_p, p.dir_coef, p.std);
    if (KIND != kX0 && p.loss_total != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
        write_loss_total(p.losses, p.n_losses, p.loss_total);
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += 2 * stride) {
        const long long v2 = v + stride;
        EwVec<KIND, W, IO> e0, e1;
        e0.load(p, v);
        if (v2 < nvec) e1.load(p, v2);
        e0.finish(p, v);
        if (v2 < nvec) e1.finish(p, v2);
    }
}

static bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool aligned8(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
static bool io_ok(int io) { return io == DM_IO_F32 || io == DM_IO_F16 || io == DM_IO_BF16; }
static bool aligned_io(const void* p, int io) { return io == DM_IO_F32 ? aligned16(p) : aligned8(p); }

template <int KIND, int IO>
static void launch_ew_io(const EwParams& p, long long n, cudaStream_t st) {
    const bool vec = (n % 4 == 0) && aligned_io(p.x, IO) && aligned16(p.x0) && aligned_io(p.eps, IO) &&
                     aligned_io(p.g0, IO) && aligned16(p.z) && aligned_io(p.prev, IO) && aligned16(p.x0_out) &&
                     aligned_io(p.x0_pub, IO) && aligned_io(p.x0_leaf, IO);
    const long long nvec = vec ? n / 4 : n;
    const int nblk = (int)std::max<long long>(1, std::min<long long>((nvec + kThreads - 1) / kThreads,
                                                                      (long long)num_sms() * 8));
    if (vec)
        ew_update_kernel<KIND, 4, IO><<<nblk, kThreads, 0, st>>>(p, nvec);
    else
        ew_update_kernel<KIND, 1, IO><<<nblk, kThreads, 0, st>>>(p, nvec);
}
template <int KIND>
static int launch_ew(const EwParams& p, long long n, int io, cudaStream_t st) {
    if (io == DM_IO_F16) launch_ew_io<KIND, DM_IO_F16>(p, n, st);
    else if (io == DM_IO_BF16) launch_ew_io<KIND, DM_IO_BF16>(p, n, st);
    else launch_ew_io<KIND, DM_IO_F32>(p, n, st);
    return 0;
}

// ------------------------------------------------------------------------------------------------ per-clip norms
struct NormParams {
    const float* x0;
    const void* eps;  // latent-typed
    const void* g0;   // latent-typed, dLoss/d(leaf); dLoss/dx0 = leaf_scale * g0
    const float* z;
    void* prev;       // latent-typed
    long long n_clip;
    float sqrt_a, sqrt_p, dir_coef, std, rate, r, grad_scale, e, threshold;
    const float* coef;  // optional device-resident coefficients (see EwParams::coef)
    float leaf_scale;
    const float* losses;  // see EwParams
    int n_losses;
    float* loss_total;
};
// g0 as loaded (dLoss/d leaf) -> dLoss/dx0
template <int IO, int W>
__device__ __forceinline__ void ldg0(const NormParams& p, long long v, float (&g)[W]) {
    ldio<IO, W>(p.g0, v, g);
#pragma unroll
    for (int i = 0; i < W; ++i) g[i] = mul(g[i], p.leaf_scale);
}

// block-level sum of up to 3 doubles, result broadcast through smem slot `out[0..2]`
__device__ __forceinline__ void block_sum3(double a, double b, double c, double* wred, double* out) {
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        wred[warp * 3 + 0] = a;
        wred[warp * 3 + 1] = b;
        wred[warp * 3 + 2] = c;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += wred[w * 3 + threadIdx.x];
        out[threadIdx.x] = t;
    }
}

// sum the per-CTA slots of every CTA in the cluster through DSMEM, in rank order (deterministic)
__device__ __forceinline__ void cluster_sum3(cg::cluster_group& cluster, double* slot, double (&tot)[3]) {
    cluster.sync();  // all slots written
    tot[0] = tot[1] = tot[2] = 0.0;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) {
        const double* remote = cluster.map_shared_rank(slot, r);
        tot[0] += remote[0];
        tot[1] += remote[1];
        tot[2] += remote[2];
    }
}

enum NormKind { kDsg = 0, kDiffMusic = 1 };

// per-element pieces of B.4 / B.5, written once and used by all sweeps
struct NormScalars {
    float gn, zn, w0, w1, mix_scale;
    bool lin;
};
__device__ __forceinline__ float elem_g(const NormParams& p, float g0) { return dvd(mul(p.grad_scale, g0), p.sqrt_a); }
__device__ __forceinline__ float elem_mix(const NormParams& p, float gn, float g0, float z) {
    const float dstar = dvd(mul(-p.r, elem_g(p, g0)), add(gn, p.e));  // scheduling_dsg.py:214
    const float ds = mul(p.std, z);
    return add(ds, mul(p.rate, sub(dstar, ds)));                       // :221-222
}
template <int KIND>
__device__ __forceinline__ float elem_out(const NormParams& p, const NormScalars& c, float x0, float ep, float g0,
                                          float z) {
    const float mean = add(mul(p.sqrt_p, x0), mul(p.dir_coef, ep));
    if (KIND == kDsg) return add(mean, dvd(mul(p.r, elem_mix(p, c.gn, g0, z)), c.mix_scale));  // :224
    const float u = -mul(dvd(elem_g(p, g0), add(c.gn, p.e)), c.zn);      // -normalized_grad, scheduling_diffmusic.py:221
    const float m = c.lin ? add(z, mul(p.rate, sub(u, z))) : add(mul(c.w0, z), mul(c.w1, u));  // slerp :59-68
    return add(mean, mul(p.std, m));
}

// One 8-CTA cluster per clip.  CACHED: every thread keeps its <= kIt vectors of x0 / eps / g0 / z in registers, so the
// clip is read from memory exactly once (all loads issued up front, their latency overlaps the reductions); otherwise
// (clips longer than 8 * 256 * kIt * W elements) the later sweeps re-read from L2.
constexpr int kIt = 4;

template <int KIND, int W, bool CACHED, int IO>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads) norm_update_kernel(NormParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    if (p.coef != nullptr) {
        float unused_b;
        load_coef(p.coef, p.sqrt_a, unused_b, p.sqrt_p, p.dir_coef, p.std);
        p.r = __ldg(p.coef + 5);
    }
    __shared__ double wred[(kThreads / 32) * 3];
    __shared__ double slot1[3], slot2[3];
    if (p.loss_total != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
        write_loss_total(p.losses, p.n_losses, p.loss_total);
    const unsigned rank = cluster.block_rank();
    const long long clip = blockIdx.x / kCluster;
    const long long nv_clip = p.n_clip / W;
    const long long chunk = (nv_clip + kCluster - 1) / kCluster;
    const long long lo = rank * chunk, hi = min(nv_clip, lo + chunk);
    const long long off = clip * nv_clip;  // in vectors

    float rx0[CACHED ? kIt : 1][W], rep[CACHED ? kIt : 1][W], rg[CACHED ? kIt : 1][W], rz[CACHED ? kIt : 1][W];
    if (CACHED) {
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const long long v = lo + threadIdx.x + (long long)it * kThreads;
            if (v < hi) {
                ldg0<IO, W>(p, off + v, rg[it]);
                ldv<W>(p.z, off + v, rz[it]);
                ldv<W>(p.x0, off + v, rx0[it]);
                ldio<IO, W>(p.eps, off + v, rep[it]);
            } else {
#pragma unroll
                for (int i = 0; i < W; ++i) rg[it][i] = rz[it][i] = rx0[it][i] = rep[it][i] = 0.f;
            }
        }
    }

    // ---- sweep 1: |g|^2 (+ |z|^2 and <z, g> for DiffMusic) ----
    double s_gg = 0.0, s_zz = 0.0, s_gz = 0.0;
    auto acc1 = [&](const float (&g)[W], const float (&z)[W]) {
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const float gi = elem_g(p, g[i]);
            a = fmaf(gi, gi, a);
            if (KIND == kDiffMusic) {
                b = fmaf(z[i], z[i], b);
                c = fmaf(z[i], gi, c);
            }
        }
        s_gg += a;
        s_zz += b;
        s_gz += c;
    };
    if (CACHED) {
#pragma unroll
        for (int it = 0; it < kIt; ++it) acc1(rg[it], rz[it]);  // out-of-range slots hold zeros
    } else {
        for (long long v = lo + threadIdx.x; v < hi; v += kThreads) {
            float g[W], z[W] = {};
            ldg0<IO, W>(p, off + v, g);
            if (KIND == kDiffMusic) ldv<W>(p.z, off + v, z);
            acc1(g, z);
        }
    }
    block_sum3(s_gg, s_zz, s_gz, wred, slot1);
    double t1[3];
    cluster_sum3(cluster, slot1, t1);
    NormScalars c{};
    c.gn = sqrtf((float)t1[0]);

    if (KIND == kDsg) {
        // ---- sweep 2: |mix|^2 (scheduling_dsg.py:214-223) ----
        double s_mm = 0.0;
        auto acc2 = [&](const float (&g)[W], const float (&z)[W], bool live) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < W; ++i) {
                const float mix = elem_mix(p, c.gn, g[i], z[i]);
                a = fmaf(mix, mix, a);
            }
            if (live) s_mm += a;
        };
        if (CACHED) {
#pragma unroll
            for (int it = 0; it < kIt; ++it) acc2(rg[it], rz[it], lo + threadIdx.x + (long long)it * kThreads < hi);
        } else {
            for (long long v = lo + threadIdx.x; v < hi; v += kThreads) {
                float g[W], z[W];
                ldg0<IO, W>(p, off + v, g);
                ldv<W>(p.z, off + v, z);
                acc2(g, z, true);
            }
        }
        __syncthreads();  // wred reuse
        block_sum3(s_mm, 0.0, 0.0, wred, slot2);
        double t2[3];
        cluster_sum3(cluster, slot2, t2);
        c.mix_scale = add(sqrtf((float)t2[0]), p.e);
    } else {
        // ---- slerp weights (scheduling_diffmusic.py:59-68), per clip, on the device ----
        c.zn = sqrtf((float)t1[1]);
        // u = -g/(gn+e) * zn ; |u| = gn/(gn+e) * zn ; cos = <z,u>/(|z||u|) = -<z,g>/(gn zn)   (NaN when gn == 0, as ref)
        const float cs = (float)(-t1[2] / ((double)c.gn * (double)c.zn));
        c.lin = fabsf(cs) > p.threshold;  // data-dependent branch of slerp, resolved here instead of on the host
        if (!c.lin) {
            const float th = acosf(cs);
            const float sn = sinf(th);
            c.w0 = dvd(sinf(mul(1.f - p.rate, th)), sn);
            c.w1 = dvd(sinf(mul(p.rate, th)), sn);
        }
    }

    // ---- final sweep: write prev ----
    auto emit = [&](long long v, const float (&x0)[W], const float (&ep)[W], const float (&g)[W], const float (&z)[W]) {
        float o[W];
#pragma unroll
        for (int i = 0; i < W; ++i) o[i] = elem_out<KIND>(p, c, x0[i], ep[i], g[i], z[i]);
        stio<IO, W>(p.prev, off + v, o);
    };
    if (CACHED) {
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const long long v = lo + threadIdx.x + (long long)it * kThreads;
            if (v < hi) emit(v, rx0[it], rep[it], rg[it], rz[it]);
        }
    } else {
        for (long long v = lo + threadIdx.x; v < hi; v += kThreads) {
            float x0[W], ep[W], g[W], z[W];
            ldv<W>(p.x0, off + v, x0);
            ldio<IO, W>(p.eps, off + v, ep);
            ldg0<IO, W>(p, off + v, g);
            ldv<W>(p.z, off + v, z);
            emit(v, x0, ep, g, z);
        }
    }
    cluster.sync();  // keep every CTA's shared memory alive until all remote reads are done
}

template <int KIND, int IO>
static void launch_norm_io(const NormParams& p, int n_clips, cudaStream_t st) {
    const bool vec = (p.n_clip % 4 == 0) && aligned16(p.x0) && aligned_io(p.eps, IO) && aligned_io(p.g0, IO) &&
                     aligned16(p.z) && aligned_io(p.prev, IO);
    const int W = vec ? 4 : 1;
    const long long chunk = (p.n_clip / W + kCluster - 1) / kCluster;
    const bool cached = vec && chunk <= (long long)kIt * kThreads;  // the 10 s latent: 1000 vectors per CTA
    if (cached)
        norm_update_kernel<KIND, 4, true, IO><<<n_clips * kCluster, kThreads, 0, st>>>(p);
    else if (vec)
        norm_update_kernel<KIND, 4, false, IO><<<n_clips * kCluster, kThreads, 0, st>>>(p);
    else
        norm_update_kernel<KIND, 1, false, IO><<<n_clips * kCluster, kThreads, 0, st>>>(p);
}
template <int KIND>
static int launch_norm(const NormParams& p, int n_clips, int io, cudaStream_t st) {
    if (io == DM_IO_F16) launch_norm_io<KIND, DM_IO_F16>(p, n_clips, st);
    else if (io == DM_IO_BF16) launch_norm_io<KIND, DM_IO_BF16>(p, n_clips, st);
    else launch_norm_io<KIND, DM_IO_F32>(p, n_clips, st);
    return 0;
}

}  // namespace dm

using namespace dm;

extern "C" int dm_sched_x0_io(const void* x, const void* eps, float* x0, void* x0_pub, void* x0_leaf, float leaf_scale,
                              long long n, float sqrt_a, float sqrt_b, int clip, float clip_range, const float* coef,
                              int io_dtype, dm_stream_t stream) {
    DM_REQUIRE(x && eps && x0 && n > 0 && io_ok(io_dtype));
    EwParams p{};
    p.coef = coef;
    p.x = x;
    p.eps = eps;
    p.x0_out = x0;
    p.x0_pub = x0_pub;
    p.x0_leaf = x0_leaf;
    p.leaf_scale = leaf_scale;
    p.sqrt_a = sqrt_a;
    p.sqrt_b = sqrt_b;
    p.clip = clip;
    p.clip_range = clip_range;
    launch_ew<kX0>(p, n, io_dtype, as_stream(stream));
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_sched_x0(const float* x, const float* eps, float* x0, long long n, float sqrt_a, float sqrt_b,
                           int clip, float clip_range, const float* coef, dm_stream_t stream) {
    return dm_sched_x0_io(x, eps, x0, nullptr, nullptr, 1.f, n, sqrt_a, sqrt_b, clip, clip_range, coef, DM_IO_F32,
                          stream);
}

extern "C" int dm_sched_ddim_update_io(const void* x, const float* x0, void* prev, long long n, float sqrt_a,
                                       float sqrt_b, float sqrt_p, float sqrt_1mp, const float* coef, int io_dtype,
                                       dm_stream_t stream) {
    DM_REQUIRE(x && x0 && prev && n > 0 && io_ok(io_dtype));
    EwParams p{};
    p.coef = coef;
    p.x = x;
    p.x0 = x0;
    p.prev = prev;
    p.sqrt_a = sqrt_a;
    p.sqrt_b = sqrt_b;
    p.sqrt_p = sqrt_p;
    p.dir_coef = sqrt_1mp;
    launch_ew<kDdim>(p, n, io_dtype, as_stream(stream));
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_sched_ddim_update(const float* x, const float* x0, float* prev, long long n, float sqrt_a,
                                    float sqrt_b, float sqrt_p, float sqrt_1mp, const float* coef,
                                    dm_stream_t stream) {
    return dm_sched_ddim_update_io(x, x0, prev, n, sqrt_a, sqrt_b, sqrt_p, sqrt_1mp, coef, DM_IO_F32, stream);
}

extern "C" int dm_sched_dps_update_io(const void* x, const float* x0, const void* g0, float leaf_scale, const float* z,
                                      void* prev, long long n, float sqrt_a, float sqrt_b, float sqrt_p,
                                      float dir_coef, float std, float rate, const float* coef, int io_dtype,
                                      const float* losses, int n_losses, float* loss_total, dm_stream_t stream) {
    DM_REQUIRE(x && x0 && g0 && prev && n > 0 && io_ok(io_dtype));
    DM_REQUIRE(loss_total == nullptr || (losses != nullptr && n_losses > 0));
    EwParams p{};
    p.leaf_scale = leaf_scale;
    p.losses = losses;
    p.n_losses = n_losses;
    p.loss_total = loss_total;
    p.coef = coef;
    p.x = x;
    p.x0 = x0;
    p.g0 = g0;
    p.z = z;
    p.prev = prev;
    p.sqrt_a = sqrt_a;
    p.sqrt_b = sqrt_b;
    p.sqrt_p = sqrt_p;
    p.dir_coef = dir_coef;
    p.std = std;
    p.rate = rate;
    launch_ew<kDps>(p, n, io_dtype, as_stream(stream));
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_sched_dps_update(const float* x, const float* x0, const float* g0, const float* z, float* prev,
                                   long long n, float sqrt_a, float sqrt_b, float sqrt_p, float dir_coef, float std,
                                   float rate, const float* coef, dm_stream_t stream) {
    return dm_sched_dps_update_io(x, x0, g0, 1.f, z, prev, n, sqrt_a, sqrt_b, sqrt_p, dir_coef, std, rate, coef,
                                  DM_IO_F32, nullptr, 0, nullptr, stream);
}

extern "C" int dm_sched_mpgd_update_io(const void* x, const float* x0, const void* g0, float leaf_scale, const float* z,
                                       void* prev, void* x0_out, long long n, float sqrt_a, float sqrt_b, float sqrt_p,
                                       float dir_coef, float std, float rate, const float* coef, int io_dtype,
                                       const float* losses, int n_losses, float* loss_total, dm_stream_t stream) {
    DM_REQUIRE(x && x0 && g0 && prev && x0_out && n > 0 && io_ok(io_dtype));
    DM_REQUIRE(loss_total == nullptr || (losses != nullptr && n_losses > 0));
    EwParams p{};
    p.leaf_scale = leaf_scale;
    p.losses = losses;
    p.n_losses = n_losses;
    p.loss_total = loss_total;
    p.coef = coef;
    p.x = x;
    p.x0 = x0;
    p.g0 = g0;
    p.z = z;
    p.prev = prev;
    if (io_dtype == DM_IO_F32) p.x0_out = static_cast<float*>(x0_out);
    else p.x0_pub = x0_out;
    p.sqrt_a = sqrt_a;
    p.sqrt_b = sqrt_b;
    p.sqrt_p = sqrt_p;
    p.dir_coef = dir_coef;
    p.std = std;
    p.rate = rate;
    launch_ew<kMpgd>(p, n, io_dtype, as_stream(stream));
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_sched_mpgd_update(const float* x, const float* x0, const float* g0, const float* z, float* prev,
                                    float* x0_out, long long n, float sqrt_a, float sqrt_b, float sqrt_p,
                                    float dir_coef, float std, float rate, const float* coef,
                                    dm_stream_t stream) {
    return dm_sched_mpgd_update_io(x, x0, g0, 1.f, z, prev, x0_out, n, sqrt_a, sqrt_b, sqrt_p, dir_coef, std, rate,
                                   coef, DM_IO_F32, nullptr, 0, nullptr, stream);
}

extern "C" int dm_sched_dsg_update_io(const float* x0, const void* eps, const void* g0, float leaf_scale, const float* z,
                                      void* prev, int n_clips, long long n_clip, float sqrt_a, float sqrt_p,
                                      float dir_coef, float std, float rate, float r, float grad_scale, float e,
                                      const float* coef, int io_dtype, const float* losses, float* loss_total,
                                      dm_stream_t stream) {
    DM_REQUIRE(x0 && eps && g0 && z && prev && n_clips > 0 && n_clip > 0 && io_ok(io_dtype));
    DM_REQUIRE(loss_total == nullptr || losses != nullptr);
    NormParams p{x0, eps, g0, z, prev, n_clip, sqrt_a, sqrt_p, dir_coef, std, rate, r, grad_scale, e, 0.f, coef,
                 leaf_scale, losses, n_clips, loss_total};
    launch_norm<kDsg>(p, n_clips, io_dtype, as_stream(stream));
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_sched_dsg_update(const float* x0, const float* eps, const float* g0, const float* z, float* prev,
                                   int n_clips, long long n_clip, float sqrt_a, float sqrt_p, float dir_coef,
                                   float std, float rate, float r, float grad_scale, float e, const float* coef,
                                   dm_stream_t stream) {
    return dm_sched_dsg_update_io(x0, eps, g0, 1.f, z, prev, n_clips, n_clip, sqrt_a, sqrt_p, dir_coef, std, rate, r,
                                  grad_scale, e, coef, DM_IO_F32, nullptr, nullptr, stream);
}

extern "C" int dm_sched_diffmusic_update_io(const float* x0, const void* eps, const void* g0, float leaf_scale,
                                            const float* z, void* prev, int n_clips, long long n_clip, float sqrt_a,
                                            float sqrt_p, float dir_coef, float std, float rate, float grad_scale,
                                            float e, float threshold, const float* coef, int io_dtype,
                                            const float* losses, float* loss_total, dm_stream_t stream) {
    DM_REQUIRE(x0 && eps && g0 && z && prev && n_clips > 0 && n_clip > 0 && io_ok(io_dtype));
    DM_REQUIRE(loss_total == nullptr || losses != nullptr);
    NormParams p{x0, eps, g0, z, prev, n_clip, sqrt_a, sqrt_p, dir_coef, std, rate, 0.f, grad_scale, e, threshold,
                 coef, leaf_scale, losses, n_clips, loss_total};
    launch_norm<kDiffMusic>(p, n_clips, io_dtype, as_stream(stream));
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_sched_diffmusic_update(const float* x0, const float* eps, const float* g0, const float* z,
                                         float* prev, int n_clips, long long n_clip, float sqrt_a, float sqrt_p,
                                         float dir_coef, float std, float rate, float grad_scale, float e,
                                         float threshold, const float* coef, dm_stream_t stream) {
    return dm_sched_diffmusic_update_io(x0, eps, g0, 1.f, z, prev, n_clips, n_clip, sqrt_a, sqrt_p, dir_coef, std, rate,
                                        grad_scale, e, threshold, coef, DM_IO_F32, nullptr, nullptr, stream);
}
