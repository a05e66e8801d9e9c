// Launch parameters and TMA / mbarrier helpers shared by the STFT guidance kernels (stft_guidance.cu, stft_warp.cu).
#pragma once
#include "dm_common.cuh"
#include "stft_frame.cuh"

namespace dm {

struct StftParams {
    StftTables tab;
    int clamp, hop, B, nf, ntiles;
    long long Ly, T, y_bstride, ref_bstride;
    const void* y;  // waveform-typed (y_io)
    int y_io;
    const float* mask;
    const float* ref;
    const float* noise;
    float sigma;
    float* out;
    float* ypbar;
    float* partial;
    // fused scale-2 super-resolution chain (dm_stft_guidance_fir2, warp engine only): the signal of the tiles is
    // y = sinc-resample(x) computed inside the kernel from fir_x (B, fir_L) fp32 rows; `y` is NULL then
    const float* fir_x = nullptr;
    long long fir_x_bstride = 0, fir_L = 0;
    float fir_h[28] = {};  // the filter taps BY VALUE: FFMA reads them straight from the constant bank, no registers
};

// ---- TMA bulk copy (cp.async.bulk) of an interior tile's contiguous signal span into shared memory ----
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load_span(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_addr_u32(bar))
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_addr_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

// defined in stft_warp.cu: the warp-per-frame-pair kernel (even hops, <= 16 frames per tile)
int launch_stft_warp(const StftParams& p, int mode, cudaStream_t st);
size_t stft_warp_smem_bytes(int nf, int hop, int image_floats);

}  // namespace dm
