// Shared host-side helpers for the C ABI: error string, launch accounting, argument checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/dm_abi.h"

namespace dm {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;
extern int g_tuning[DM_TUNE_COUNT];  // dm_set_tuning

template <typename... A>
inline int fail(int code, const char* fmt, A... a) {
    snprintf(g_err, sizeof(g_err), fmt, a...);
    return code;
}

#define DM_REQUIRE(cond)                                                                                   \
    do {                                                                                                   \
        if (!(cond)) return dm::fail(DM_ERR_INVALID, "%s: requirement failed: %s", __func__, #cond);       \
    } while (0)

// call after every kernel launch: counts it and turns a launch error into a return code
#define DM_LAUNCHED()                                                                                      \
    do {                                                                                                   \
        dm::g_launches.fetch_add(1, std::memory_order_relaxed);                                            \
        cudaError_t e__ = cudaGetLastError();                                                              \
        if (e__ != cudaSuccess)                                                                            \
            return dm::fail(DM_ERR_CUDA, "%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
    } while (0)

#define DM_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return dm::fail(DM_ERR_CUDA, "%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));   \
    } while (0)

// Raise the dynamic shared-memory limit of a kernel.  The limit only ever grows, and the call is skipped once it is
// large enough, so steady-state launches (and CUDA-graph capture) make no non-stream runtime calls.
#define DM_SMEM_ONCE(kernel, bytes)                                                                        \
    do {                                                                                                   \
        static size_t have__ = 0;                                                                          \
        if ((size_t)(bytes) > have__) {                                                                    \
            DM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            have__ = (size_t)(bytes);                                                                      \
        }                                                                                                  \
    } while (0)

// ask for the largest shared-memory carve-out once (kernels that want 3 x ~74 KB CTAs resident per SM)
#define DM_CARVEOUT_ONCE(kernel)                                                                           \
    do {                                                                                                   \
        static bool done__ = false;                                                                        \
        if (!done__) {                                                                                     \
            DM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,           \
                                         cudaSharedmemCarveoutMaxShared));                                 \
            done__ = true;                                                                                 \
        }                                                                                                  \
    } while (0)

// Launch with the programmatic-stream-serialization attribute (when `pdl`): the kernel may be scheduled while the
// previous kernel of the stream is still running (after all its CTAs have executed griddepcontrol.launch_dependents, or
// have exited); everything it reads from / writes in common with that kernel comes after its own griddepcontrol.wait
// (pdl_wait()), which returns once the previous kernel has completed and flushed.  Both instructions are no-ops in a
// kernel launched without the attribute / without a dependent.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline bool io_dtype_ok(int io) { return io == DM_IO_F32 || io == DM_IO_F16 || io == DM_IO_BF16; }
inline cudaStream_t as_stream(dm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ---- device helpers ----
#if defined(__CUDACC__)
}  // namespace dm
#include <cuda_bf16.h>
#include <cuda_fp16.h>
namespace dm {
// Waveform-typed tensors (the vocoder output handed to the operators, and dLoss/dwav handed back to autograd) may be
// fp32, fp16 or bf16 (DM_IO_*): converted on load / store, all arithmetic in fp32.
__device__ __forceinline__ float ld_wave(const void* __restrict__ p, int io, long long i) {
    if (io == DM_IO_F16) return __half2float(static_cast<const __half*>(p)[i]);
    if (io == DM_IO_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
    return static_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_wave(void* __restrict__ p, int io, long long i, float v) {
    if (io == DM_IO_F16) static_cast<__half*>(p)[i] = __float2half_rn(v);
    else if (io == DM_IO_BF16) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else static_cast<float*>(p)[i] = v;
}
__device__ __forceinline__ const void* wave_row(const void* p, int io, long long elems) {
    return static_cast<const char*>(p) + elems * (io == DM_IO_F32 ? 4 : 2);
}
__device__ __forceinline__ void* wave_row(void* p, int io, long long elems) {
    return static_cast<char*>(p) + elems * (io == DM_IO_F32 ? 4 : 2);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Deterministic per-clip scale from the per-tile partial sums of squared residuals: every CTA of an adjoint kernel
// recomputes it (<= a few hundred floats, L2 resident) in a fixed order.  loss = sqrt(sum); scale = 1/loss, 0 at 0
// (torch.linalg.norm backward is masked at 0); NaN / Inf propagate.
__device__ __forceinline__ float clip_loss(const float* __restrict__ partial, int ntiles, float* smem_scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        double acc = 0.0;
        for (int i = lane; i < ntiles; i += 32) acc += (double)partial[i];
        acc = warp_sum(acc);
        if (lane == 0) smem_scratch[0] = sqrtf((float)acc);
    }
    __syncthreads();
    return smem_scratch[0];
}
__device__ __forceinline__ float inv_loss(float loss) { return loss == 0.f ? 0.f : 1.0f / loss; }
#endif

}  // namespace dm
