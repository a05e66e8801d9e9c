// Per-clip Gaussian noise for a whole batch in ONE launch, bit-identical to the reference's per-clip draws.
//
// The reference draws the step noise with `randn_tensor(shape, generator=[g_0 .. g_{B-1}])` (diffmusic/torch_utils.py:31-76):
// one `torch.randn((1, C, H, W), generator=g_b)` per clip, concatenated -- B tiny kernels plus a concat per step, which at
// B = 16 costs more device time than the fused scheduler update itself.  torch's CUDA normal kernel is curand's Philox4x32-10
// + Box-Muller (`curand_normal4`) on a fixed thread/index mapping (ATen/native/cuda/DistributionTemplates.h,
// `distribution_elementwise_grid_stride_kernel`):
//     grid = min(#SM * (max threads per SM / 256), ceil(n / 256)) blocks of 256 threads, thread idx = subsequence idx,
//     state = curand_init(seed, idx, offset); each loop trip draws 4 normals r[0..3] for elements idx + ii * grid * 256
// and the generator's offset then advances by ((n - 1) / (256 * grid * 4) + 1) * 4.  This kernel runs exactly that mapping
// for every clip (blockIdx.y = clip) with the clip's own (seed, offset), using the same curand device functions, so the
// values are the ones torch would have produced; the host advances each generator's offset by the same amount.
#include <curand_kernel.h>

#include "dm_common.cuh"

namespace dm {

struct RngClipsParams {
    unsigned long long seed[DM_RNG_MAX_CLIPS];
    unsigned long long offset[DM_RNG_MAX_CLIPS];
};

__global__ void __launch_bounds__(256) randn_clips_kernel(const __grid_constant__ RngClipsParams p, long long n,
                                                          int round_dtype, float* __restrict__ out) {
    const int clip = blockIdx.y;
    const unsigned idx = blockIdx.x * 256u + threadIdx.x;
    curandStatePhilox4_32_10_t state;
    curand_init(p.seed[clip], idx, p.offset[clip], &state);
    const long long stride = 256LL * gridDim.x;
    const long long rounded = ((n - 1) / (stride * 4) + 1) * stride * 4;
    float* o = out + (long long)clip * n;
    for (long long li0 = idx; li0 < rounded; li0 += stride * 4) {
        const float4 r = curand_normal4(&state);
        const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const long long li = li0 + stride * ii;
            if (li < n) {
                float x = v[ii];  // torch: static_cast<scalar_t>(rand * std + mean) with std = 1, mean = 0
                if (round_dtype == DM_IO_F16) x = __half2float(__float2half_rn(x));
                else if (round_dtype == DM_IO_BF16) x = __bfloat162float(__float2bfloat16_rn(x));
                o[li] = x;
            }
        }
    }
}

static int rng_grid(long long n) {
    static long long cap = 0;  // #SM * (max threads per SM / 256): torch's grid cap (one device model per process)
    if (cap == 0) {
        int dev = 0, sms = 0, tpsm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
        cap = (long long)sms * (tpsm / 256);
    }
    const long long g = (n + 255) / 256;
    return (int)(g < cap ? g : cap);
}

}  // namespace dm

using namespace dm;

extern "C" long long dm_randn_offset_increment(long long n) {
    if (n <= 0) return 0;
    const long long grid = rng_grid(n);
    return ((n - 1) / (256 * grid * 4) + 1) * 4;
}

extern "C" int dm_randn_clips(const unsigned long long* seeds, const unsigned long long* offsets, int n_clips,
                              long long n_per_clip, int round_dtype, float* out, dm_stream_t stream) {
    DM_REQUIRE(seeds && offsets && out && n_per_clip > 0);
    DM_REQUIRE(n_clips >= 1 && n_clips <= DM_RNG_MAX_CLIPS && io_dtype_ok(round_dtype));
    RngClipsParams p;
    for (int i = 0; i < n_clips; ++i) {
        p.seed[i] = seeds[i];
        p.offset[i] = offsets[i];
    }
    randn_clips_kernel<<<dim3(rng_grid(n_per_clip), n_clips), 256, 0, as_stream(stream)>>>(p, n_per_clip, round_dtype,
                                                                                          out);
    DM_LAUNCHED();
    return DM_OK;
}
