// Dereverberation operator and its VJP by overlap-save FFT (see rir_block.cuh).
// K = 5000 taps (run.py:208-210) costs 1.6 GFLOP per clip as a direct correlation; as 8192-point FFT blocks it is
// ~26 MFLOP per clip and the kernel is bound by shared-memory traffic of the Stockham passes, not by math.
#include "dm_common.cuh"
#include "rir_block.cuh"

namespace dm {

__device__ __forceinline__ float fold_at_rir(const float* __restrict__ yp, long long j, long long Ly) {
    float v = yp[512 + j];
    if (j >= 1 && j <= 512) v += yp[512 - j];
    if (j >= Ly - 513 && j <= Ly - 2) v += yp[512 + 2 * (Ly - 1) - j];
    return v;
}

struct SrcPlain {  // waveform-typed input (fp32 / fp16 / bf16)
    const void* x;
    int io;
    long long off, L;
    __device__ __forceinline__ float operator()(int n) const {
        long long i = off + n;
        return (i >= 0 && i < L) ? ld_wave(x, io, i) : 0.f;
    }
};
struct SrcFolded {  // scaled, reflect-folded cotangent
    const float* yb;
    long long off, Ly;
    int pad;
    float scale;
    __device__ __forceinline__ float operator()(int n) const {
        long long i = off + n;
        if (i < 0 || i >= Ly) return 0.f;
        return (pad ? fold_at_rir(yb, i, Ly) : yb[i]) * scale;
    }
};

__device__ __forceinline__ RirSmem carve(float* smem) {
    RirSmem s;
    s.a_re = smem;
    s.a_im = s.a_re + swz_len(kRirH);
    s.b_re = s.a_im + swz_len(kRirH);
    s.b_im = s.b_re + swz_len(kRirH);
    return s;
}

__global__ void __launch_bounds__(kRirThreads) rir_spectrum_kernel(const float* __restrict__ ir, int K,
                                                                   const cf* __restrict__ tw,
                                                                   const cf* __restrict__ w8192,
                                                                   cf* __restrict__ spec) {
    extern __shared__ __align__(16) float smem[];
    RirSmem s = carve(smem);
    SrcPlain src{ir, DM_IO_F32, 0, K};
    RirStore st{nullptr, 0, 0, 0.f};
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        rir_block_phase<false>(ph, threadIdx.x, tw, w8192, nullptr, s, src, st);
        __syncthreads();
    }
    rir_unpack_spectrum(threadIdx.x, SwzLoad{s.b_re, s.b_im}, w8192, spec);
}

__global__ void __launch_bounds__(kRirThreads, 2) rir_correlate_kernel(const void* __restrict__ x, int x_io,
                                                                    long long x_bstride, RirGeom g,
                                                                    const cf* __restrict__ spec,
                                                                    const cf* __restrict__ tw,
                                                                    const cf* __restrict__ w8192,
                                                                    float* __restrict__ y) {
    extern __shared__ __align__(16) float smem[];
    RirSmem s = carve(smem);
    const int b = blockIdx.y;
    pdl_trigger();  // the STFT kernel of the chain may stage its tables while this one runs
    const long long i0 = (long long)blockIdx.x * g.valid;
    SrcPlain src{wave_row(x, x_io, (long long)b * x_bstride), x_io, i0 - g.pad, g.L};
    RirStore st{y + (long long)b * g.nout + i0, 0, (int)min((long long)g.valid, g.nout - i0), 1.0f / kRirN};
#pragma unroll
    for (int ph = 0; ph < kRirPhases; ++ph) {
        rir_block_phase<true>(ph, threadIdx.x, tw, w8192, spec, s, src, st);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kRirThreads, 2) rir_adjoint_kernel(const float* __restrict__ ybar, int pad,
                                                                  RirGeom g, const float* __restrict__ partial,
                                                                  int ntiles, const cf* __restrict__ spec,
                                                                  const cf* __restrict__ tw,
                                                                  const cf* __restrict__ w8192,
                                                                  void* __restrict__ dwav, int dw_io,
                                                                  long long dwav_bstride,
                                                                  float* __restrict__ loss) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float scratch[2];
    RirSmem s = carve(smem);
    const int b = blockIdx.y;
    pdl_wait();  // cotangent and partial sums of the STFT kernel
    const float l = clip_loss(partial + (long long)b * ntiles, ntiles, scratch);
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) loss[b] = l;
    const long long j0 = (long long)blockIdx.x * g.valid;
    SrcFolded src{ybar + (long long)b * (g.nout + 2 * pad), j0 + g.pad - (g.K - 1), g.nout, pad, inv_loss(l)};
    RirStore st{static_cast<float*>(wave_row(dwav, dw_io, (long long)b * dwav_bstride + j0)), g.K - 1,
                (int)min((long long)g.valid, g.L - j0), 1.0f / kRirN, dw_io};
#pragma unroll
    for (int ph = 0; ph < kRirPhases; ++ph) {
        rir_block_phase<false>(ph, threadIdx.x, tw, w8192, spec, s, src, st);
        __syncthreads();
    }
}

constexpr size_t kRirSmemBytes = (size_t)kRirSmemFloats * sizeof(float);

}  // namespace dm

using namespace dm;

extern "C" int dm_rir_spectrum(const float* ir, int K, const float* tw4096, const float* w8192, float* spec,
                               dm_stream_t stream) {
    DM_REQUIRE(ir && tw4096 && w8192 && spec);
    DM_REQUIRE(K >= 1 && K <= DM_RIR_MAX_TAPS);
    DM_SMEM_ONCE(rir_spectrum_kernel, kRirSmemBytes);
    rir_spectrum_kernel<<<1, kRirThreads, kRirSmemBytes, as_stream(stream)>>>(
        ir, K, reinterpret_cast<const cf*>(tw4096), reinterpret_cast<const cf*>(w8192), reinterpret_cast<cf*>(spec));
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_rir_correlate(const float* x, long long x_bstride, long long L, int B, const float* spec, int K,
                                const float* tw4096, const float* w8192, float* y, long long Ly,
                                dm_stream_t stream) {
    return dm_rir_correlate_io(x, DM_IO_F32, x_bstride, L, B, spec, K, tw4096, w8192, y, Ly, stream);
}
extern "C" int dm_rir_correlate_io(const void* x, int x_dtype, long long x_bstride, long long L, int B,
                                   const float* spec, int K, const float* tw4096, const float* w8192, float* y,
                                   long long Ly, dm_stream_t stream) {
    DM_REQUIRE(x && spec && tw4096 && w8192 && y && L > 0 && B > 0 && io_dtype_ok(x_dtype));
    DM_REQUIRE(K >= 1 && K <= DM_RIR_MAX_TAPS);
    RirGeom g = rir_geom(L, K);
    DM_REQUIRE(Ly == g.nout);
    const int nblk = (int)((g.nout + g.valid - 1) / g.valid);
    DM_SMEM_ONCE(rir_correlate_kernel, kRirSmemBytes);
    rir_correlate_kernel<<<dim3(nblk, B), kRirThreads, kRirSmemBytes, as_stream(stream)>>>(
        x, x_dtype, x_bstride, g, reinterpret_cast<const cf*>(spec), reinterpret_cast<const cf*>(tw4096),
        reinterpret_cast<const cf*>(w8192), y);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_rir_adjoint(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                              const float* spec, int K, const float* tw4096, const float* w8192, float* dwav,
                              long long dwav_bstride, long long L, float* loss, dm_stream_t stream) {
    return dm_rir_adjoint_io(ybar, pad, Ly, B, partial, ntiles, spec, K, tw4096, w8192, dwav, DM_IO_F32, dwav_bstride, L,
                             loss, stream);
}
extern "C" int dm_rir_adjoint_io(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                                 const float* spec, int K, const float* tw4096, const float* w8192, void* dwav,
                                 int dwav_dtype, long long dwav_bstride, long long L, float* loss, dm_stream_t stream) {
    DM_REQUIRE(ybar && partial && spec && tw4096 && w8192 && dwav && L > 0 && B > 0 && ntiles > 0);
    DM_REQUIRE(io_dtype_ok(dwav_dtype));
    DM_REQUIRE(K >= 1 && K <= DM_RIR_MAX_TAPS);
    RirGeom g = rir_geom(L, K);
    DM_REQUIRE(Ly == g.nout);
    DM_REQUIRE(pad == 0 || (pad == 512 && Ly > 512));
    const int nblk = (int)((L + g.valid - 1) / g.valid);
    DM_SMEM_ONCE(rir_adjoint_kernel, kRirSmemBytes);
    launch_pdl(rir_adjoint_kernel, dim3(nblk, B), dim3(kRirThreads), kRirSmemBytes, as_stream(stream),
               g_tuning[DM_TUNE_PDL] != 0, ybar, pad, g, partial, ntiles, reinterpret_cast<const cf*>(spec),
               reinterpret_cast<const cf*>(tw4096), reinterpret_cast<const cf*>(w8192), dwav, dwav_dtype, dwav_bstride,
               loss);
    DM_LAUNCHED();
    return DM_OK;
}
