// FAD embedding statistics: raw moments (n, sum x, sum x x^T) of an (N, d) fp16 block, accumulated in float64.
// fadtk/fad.py:41-47 (np.mean / np.cov) and fadtk/utils.py:13-46 (Chan merge) reduce to these three sums; summing them
// over ranks with ONE all-reduce and finalising once is algebraically the Chan merge of all files.
//
// fp16 x fp16 products are exact in fp32; each CTA accumulates a 64x64 tile of X^T X over a slab of rows in fp32 and
// adds it to the float64 accumulator, so the only rounding is the fp32 running sum inside one slab.
#include <cuda_fp16.h>

#include <algorithm>

#include "dm_common.cuh"

namespace dm {

constexpr int kFadTile = 64;
constexpr int kFadSlab = 32;
constexpr int kFadThreads = 256;

__global__ void __launch_bounds__(kFadThreads) fad_xtx_kernel(const __half* __restrict__ X, long long N, int d,
                                                              long long rows_per_cta, double* __restrict__ sxx) {
    const int ti = blockIdx.x, tj = blockIdx.y;
    if (tj < ti) return;  // X^T X is symmetric: compute the upper tiles, mirror on write
    __shared__ float As[kFadSlab][kFadTile + 4];
    __shared__ float Bs[kFadSlab][kFadTile + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long r_begin = (long long)blockIdx.z * rows_per_cta;
    const long long r_end = min(N, r_begin + rows_per_cta);
    float acc[4][4] = {};
    for (long long r0 = r_begin; r0 < r_end; r0 += kFadSlab) {
#pragma unroll
        for (int q = 0; q < (kFadSlab * kFadTile) / kFadThreads; ++q) {
            const int idx = tid + kFadThreads * q;
            const int row = idx / kFadTile, col = idx % kFadTile;
            const long long r = r0 + row;
            const int ca = ti * kFadTile + col, cb = tj * kFadTile + col;
            As[row][col] = (r < r_end && ca < d) ? __half2float(X[r * d + ca]) : 0.f;
            Bs[row][col] = (r < r_end && cb < d) ? __half2float(X[r * d + cb]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < kFadSlab; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = As[kk][ty * 4 + i];
                b[i] = Bs[kk][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gi = ti * kFadTile + ty * 4 + i, gj = tj * kFadTile + tx * 4 + j;
            if (gi < d && gj < d) {
                atomicAdd(&sxx[(long long)gi * d + gj], (double)acc[i][j]);
                if (ti != tj) atomicAdd(&sxx[(long long)gj * d + gi], (double)acc[i][j]);
            }
        }
}

__global__ void __launch_bounds__(kFadThreads) fad_colsum_kernel(const __half* __restrict__ X, long long N, int d,
                                                                 long long rows_per_cta, double* __restrict__ acc) {
    const int c = blockIdx.x * kFadThreads + threadIdx.x;
    const long long r_begin = (long long)blockIdx.y * rows_per_cta;
    const long long r_end = min(N, r_begin + rows_per_cta);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(&acc[0], (double)N);
    if (c >= d) return;
    double s = 0.0;
    float part = 0.f;
    int cnt = 0;
    for (long long r = r_begin; r < r_end; ++r) {
        part += __half2float(X[r * d + c]);
        if (++cnt == 64) {  // short fp32 runs, float64 across runs
            s += part;
            part = 0.f;
            cnt = 0;
        }
    }
    s += part;
    atomicAdd(&acc[1 + c], s);
}

// Column sums at HBM speed for d % 8 == 0: thread (p, g) owns 8 consecutive columns (one 128-bit load per row) and every
// P-th row of the block's slab; fp32 partials over 32 rows, float64 across, one float64 atomic per column at the end.
__global__ void __launch_bounds__(1024) fad_colsum_vec_kernel(const __half* __restrict__ X, long long N, int d, int G,
                                                              int P, long long rows_per_cta,
                                                              double* __restrict__ acc) {
    const int g = threadIdx.x % G, pr = threadIdx.x / G;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&acc[0], (double)N);
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(N, r_begin + rows_per_cta);
    double s[8];
    float part[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.0, part[i] = 0.f;
    int cnt = 0;
    for (long long r = r_begin + pr; r < r_end; r += P) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + r * d) + g);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            part[2 * i] += f.x;
            part[2 * i + 1] += f.y;
        }
        if (++cnt == 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] += part[i], part[i] = 0.f;
            cnt = 0;
        }
    }
    // block-level reduction over the P row groups in shared memory, then ONE float64 atomic per column per CTA
    extern __shared__ double red[];  // [P][d]
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(size_t)pr * d + g * 8 + i] = s[i] + (double)part[i];
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        double t = 0.0;
        for (int q = 0; q < P; ++q) t += red[(size_t)q * d + c];
        atomicAdd(&acc[1 + c], t);
    }
}

__global__ void __launch_bounds__(kFadThreads) fad_finalize_kernel(const double* __restrict__ acc, int d,
                                                                   double* __restrict__ mu,
                                                                   double* __restrict__ cov) {
    const double n = acc[0];
    const long long idx = (long long)blockIdx.x * kFadThreads + threadIdx.x;
    if (idx >= (long long)d * d) return;
    const int i = (int)(idx / d), j = (int)(idx % d);
    const double mi = acc[1 + i] / n, mj = acc[1 + j] / n;
    if (j == 0) mu[i] = mi;
    // fadtk/utils.py:42-46: cov = S/(n-1), zeros when n < 2 ; S = sum xx^T - n mu mu^T
    cov[idx] = (n < 2.0) ? 0.0 : (acc[1 + d + idx] - n * mi * mj) / (n - 1.0);
}

}  // namespace dm

namespace dm {
int fad_xtx_tc(const void* x_f16, long long N, int d, double* sxx, double* sx, double* n_out, bool pair_mma,
               cudaStream_t st);  // fad_tc.cu
}

using namespace dm;

extern "C" int dm_fad_moments(const void* x_f16, long long N, int d, double* acc, dm_stream_t stream) {
    return dm_fad_moments_ex(x_f16, N, d, acc, DM_FAD_AUTO, stream);
}

extern "C" int dm_fad_moments_ex(const void* x_f16, long long N, int d, double* acc, int engine,
                                 dm_stream_t stream) {
    DM_REQUIRE(x_f16 && acc && N > 0 && d > 0);
    DM_REQUIRE(engine == DM_FAD_AUTO || engine == DM_FAD_SIMT || engine == DM_FAD_TCGEN05 ||
               engine == DM_FAD_TCGEN05_PAIR);
    const __half* X = reinterpret_cast<const __half*>(x_f16);
    bool xtx_done = false;
    if (engine != DM_FAD_SIMT) {
        // the tensor-core kernel also produces the column sums and the count (one extra N = 16 MMA per k-step on the
        // diagonal tiles), so X is read exactly once; the cta_group::2 engine leaves them to the column-sum kernel
        // Measured (tools/fad_bench.py, N = 511 k): fused sums win up to 6 column blocks (d = 512: 0.345 -> 0.273 ms,
        // d = 768: 0.587 -> 0.525 ms); at d = 1024 the 8 diagonal CTAs become the stragglers (0.897 -> 0.931 ms) and the
        // separate HBM-speed column-sum kernel stays.
        const bool pair = engine == DM_FAD_TCGEN05_PAIR;
        const bool fuse_sums = !pair && d <= 6 * 128;
        int rc = fad_xtx_tc(x_f16, N, d, acc + 1 + d, fuse_sums ? acc + 1 : nullptr, fuse_sums ? acc : nullptr, pair,
                            as_stream(stream));
        if (rc == DM_OK && fuse_sums) return DM_OK;
        if (rc == DM_OK) xtx_done = true;
        else if (engine == DM_FAD_TCGEN05 || engine == DM_FAD_TCGEN05_PAIR || rc != DM_ERR_UNSUPPORTED) return rc;
    }
    const int nt = (d + kFadTile - 1) / kFadTile;
    // enough row slabs to fill the machine, each long enough to amortise the float64 atomics
    long long want = std::max<long long>(1, (2LL * num_sms()) / std::max(1, nt * (nt + 1) / 2));
    long long rows = std::max<long long>(256, (N + want - 1) / want);
    rows = (rows + kFadSlab - 1) / kFadSlab * kFadSlab;
    const int nz = (int)((N + rows - 1) / rows);
    if (!xtx_done) {
        fad_xtx_kernel<<<dim3(nt, nt, nz), kFadThreads, 0, as_stream(stream)>>>(X, N, d, rows, acc + 1 + d);
        DM_LAUNCHED();
    }
    if (d % 8 == 0 && d / 8 <= 1024 && (reinterpret_cast<uintptr_t>(x_f16) & 15) == 0) {
        const int G = d / 8, P = 1024 / G;  // G * P <= 1024 threads, every thread owns (row group, 8 columns)
        const long long want_ctas = 2LL * num_sms();
        const long long rows = std::max<long long>(8LL * P, (N + want_ctas - 1) / want_ctas);
        const size_t smem = (size_t)P * d * sizeof(double);
        DM_SMEM_ONCE(fad_colsum_vec_kernel, smem);
        fad_colsum_vec_kernel<<<(unsigned)((N + rows - 1) / rows), G * P, smem, as_stream(stream)>>>(X, N, d, G, P,
                                                                                                   rows, acc);
        DM_LAUNCHED();
    } else {
        const long long crow = std::max<long long>(64, (N + 63) / 64);
        fad_colsum_kernel<<<dim3((d + kFadThreads - 1) / kFadThreads, (unsigned)((N + crow - 1) / crow)), kFadThreads,
                            0, as_stream(stream)>>>(X, N, d, crow, acc);
        DM_LAUNCHED();
    }
    return DM_OK;
}

extern "C" int dm_fad_finalize(const double* acc, int d, double* mu, double* cov, dm_stream_t stream) {
    DM_REQUIRE(acc && mu && cov && d > 0);
    const long long n = (long long)d * d;
    fad_finalize_kernel<<<(unsigned)((n + kFadThreads - 1) / kFadThreads), kFadThreads, 0, as_stream(stream)>>>(
        acc, d, mu, cov);
    DM_LAUNCHED();
    return DM_OK;
}
