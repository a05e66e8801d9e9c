// Frechet distance between two Gaussians on the GPU, float64 (fadtk/fad.py:50-119 `calc_frechet_distance`), and the row
// gather of the FAD-inf bootstrap (fadtk/fad.py:303-350 `score_inf`: embeds[np.random.choice(N, n)]).
//
//     d^2 = |mu1 - mu2|^2 + tr C1 + tr C2 - 2 tr sqrt(C1 C2)
//
// The reference takes scipy's eigendecomposition of the non-symmetric product C1 C2 and sums sqrt of the (complex)
// eigenvalues.  Here the same spectrum is obtained from symmetric problems only, with a hand-written one-sided Jacobi
// (Hestenes) solver that needs nothing but row rotations -- every step is d/2 independent row pairs, one CTA each:
//
//   1. W = C1; rotate ROWS of W until they are mutually orthogonal.  For symmetric PSD C1 = Q L Q^T this gives
//      W = L Q^T (up to order / sign): row norms are the eigenvalues, and  F = diag(L)^(-1/2) W  satisfies F^T F = C1.
//   2. M = F C2 F^T (two float64 GEMMs): symmetric PSD with eig(M) = eig(C2 F^T F) = eig(C2 C1) = eig(C1 C2).
//   3. rotate the rows of M the same way: row norms = eig(M);  tr sqrt(C1 C2) = sum sqrt(norms).
//
// Pairing: the circle method of a round-robin tournament -- d - 1 rounds (d even; an odd d plays a dummy) cover every
// pair once per sweep.  One kernel launch per round; all launches of `max_sweeps` sweeps are enqueued up front and the
// CTAs of a launch return immediately once the previous sweep's largest normalised inner product is below `tol`
// (three rotating slots in device memory: written / read / cleared), so nothing synchronises with the host.
#include <cuda_fp16.h>

#include <cstdlib>

#include "dm_common.cuh"

namespace dm {

constexpr int kJacThreads = 128;

__device__ __forceinline__ void round_robin_pair(int n, int r, int k, int& i, int& j) {
    if (k == 0) {
        i = n - 1;
        j = r;
    } else {
        i = (r + k) % (n - 1);
        j = (r - k + (n - 1)) % (n - 1);
    }
}

// state[0..2]: largest |<wi,wj>| / (|wi||wj|) over the significant pairs of sweep s in slot s % 3 (bits of a
// non-negative double, so integer max orders them); state[3]: sweeps actually executed; state[4]: bits of |W|_F^2;
// state[5]: converged flag; state[6]: sweep counter of the graph-replayed path.
__global__ void __launch_bounds__(kJacThreads) jacobi_round_kernel(double* __restrict__ W, int d, int n_even, int round,
                                                                   int sweep, double tol,
                                                                   unsigned long long* __restrict__ state) {
    if (state[5] != 0ull) return;  // converged in an earlier sweep (sticky)
    if (sweep < 0) sweep = (int)state[6];  // graph replay: the sweep index lives on the device (jacobi_sweep_end_kernel)
    if (sweep > 0 && __longlong_as_double((long long)state[(sweep + 2) % 3]) <= tol) {
        if (threadIdx.x == 0) state[5] = 1ull;  // every CTA of this launch takes the same decision from the same slot
        return;
    }
    if (round == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        state[(sweep + 1) % 3] = 0ull;
        state[3] = (unsigned long long)(sweep + 1);
    }
    int i, j;
    round_robin_pair(n_even, round, blockIdx.x, i, j);
    if (i >= d || j >= d) return;
    extern __shared__ double rows[];  // [2][d]
    __shared__ double red[3][kJacThreads / 32];
    __shared__ double rot[2];
    double* wi = W + (size_t)i * d;
    double* wj = W + (size_t)j * d;
    double a = 0.0, b = 0.0, g = 0.0;
    for (int c = threadIdx.x; c < d; c += kJacThreads) {
        const double x = wi[c], y = wj[c];
        rows[c] = x;
        rows[d + c] = y;
        a = fma(x, x, a);
        b = fma(y, y, b);
        g = fma(x, y, g);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    g = warp_sum(g);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = a;
        red[1][warp] = b;
        red[2][warp] = g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = g = 0.0;
        for (int w = 0; w < kJacThreads / 32; ++w) {
            a += red[0][w];
            b += red[1][w];
            g += red[2][w];
        }
        const double den = sqrt(a * b);
        const double off = den > 0.0 ? fabs(g) / den : 0.0;
        // rows at rounding-noise level (rank-deficient covariances) keep being rotated but do not hold up convergence
        const double floor2 = 1e-28 * __longlong_as_double((long long)state[4]);
        if (fmin(a, b) > floor2) atomicMax(&state[sweep % 3], (unsigned long long)__double_as_longlong(off));
        double c = 1.0, s = 0.0;
        if (off >= 1e-16) {
            const double zeta = (b - a) / (2.0 * g);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            c = 1.0 / sqrt(1.0 + t * t);
            s = c * t;
        }
        rot[0] = c;
        rot[1] = s;
    }
    __syncthreads();
    const double c = rot[0], s = rot[1];
    if (s == 0.0) return;
    for (int k = threadIdx.x; k < d; k += kJacThreads) {
        const double x = rows[k], y = rows[d + k];
        wi[k] = c * x - s * y;
        wj[k] = s * x + c * y;
    }
}

__global__ void __launch_bounds__(256) frob2_kernel(const double* __restrict__ W, long long n,
                                                    unsigned long long* __restrict__ state) {
    __shared__ double red[8];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) acc = fma(W[i], W[i], acc);  // one CTA: fixed order, d <= 2048
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        state[0] = state[1] = state[2] = state[3] = state[5] = state[6] = 0ull;
        state[4] = (unsigned long long)__double_as_longlong(t);
    }
}

// eig[k] = |row k|; optionally scale the row by eig^(-1/2) (the factor F of step 1)
__global__ void __launch_bounds__(kJacThreads) row_norm_kernel(double* __restrict__ W, int d, double* __restrict__ eig,
                                                               int make_factor) {
    __shared__ double red[kJacThreads / 32];
    __shared__ double nrm;
    double* w = W + (size_t)blockIdx.x * d;
    double a = 0.0;
    for (int c = threadIdx.x; c < d; c += kJacThreads) a = fma(w[c], w[c], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kJacThreads / 32; ++k) t += red[k];
        nrm = sqrt(t);
        eig[blockIdx.x] = nrm;
    }
    __syncthreads();
    if (make_factor) {
        const double sc = nrm > 0.0 ? 1.0 / sqrt(nrm) : 0.0;
        for (int c = threadIdx.x; c < d; c += kJacThreads) w[c] *= sc;
    }
}

// C (M x N) = A (M x K) * B, row-major; TB: B is given as (N x K) and used transposed.  64 x 64 x 16 tiles,
// 4 x 4 outputs per thread.
template <bool TB>
__global__ void __launch_bounds__(256) dgemm_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                    double* __restrict__ C, int M, int N, int K) {
    __shared__ double As[16][65];
    __shared__ double Bs[16][65];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = tid + 256 * it;
            {
                const int m = idx >> 4, k = idx & 15;
                As[k][m] = (m0 + m < M && k0 + k < K) ? A[(size_t)(m0 + m) * K + k0 + k] : 0.0;
            }
            if (TB) {
                const int n = idx >> 4, k = idx & 15;
                Bs[k][n] = (n0 + n < N && k0 + k < K) ? B[(size_t)(n0 + n) * K + k0 + k] : 0.0;
            } else {
                const int k = idx >> 6, n = idx & 63;
                Bs[k][n] = (n0 + n < N && k0 + k < K) ? B[(size_t)(k0 + k) * N + n0 + n] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = As[k][ty * 4 + i];
                b[i] = Bs[k][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) C[(size_t)m * N + n] = acc[i][j];
        }
}

// out[0] = |mu1 - mu2|^2 + tr C1 + tr C2 - 2 sum sqrt(eig);  out[1] = sum sqrt(eig);  out[2], out[3] = sweeps used
__global__ void __launch_bounds__(256) frechet_finish_kernel(const double* __restrict__ mu1,
                                                             const double* __restrict__ cov1,
                                                             const double* __restrict__ mu2,
                                                             const double* __restrict__ cov2, int d,
                                                             const double* __restrict__ eig,
                                                             const unsigned long long* __restrict__ st1,
                                                             const unsigned long long* __restrict__ st2,
                                                             double* __restrict__ out) {
    __shared__ double red[2][8];
    double a = 0.0, t = 0.0;
    for (int k = threadIdx.x; k < d; k += 256) {
        const double df = mu1[k] - mu2[k];
        a += df * df + cov1[(size_t)k * d + k] + cov2[(size_t)k * d + k];
        t += sqrt(fmax(eig[k], 0.0));
    }
    a = warp_sum(a);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = a;
        red[1][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = t = 0.0;
        for (int w = 0; w < 8; ++w) {
            a += red[0][w];
            t += red[1][w];
        }
        out[0] = a - 2.0 * t;
        out[1] = t;
        out[2] = (double)st1[3];
        out[3] = (double)st2[3];
    }
}

__global__ void __launch_bounds__(128) gather_rows_kernel(const uint4* __restrict__ x, long long N, int d8,
                                                          const long long* __restrict__ idx, long long n,
                                                          uint4* __restrict__ out) {
    // one CTA per output row group: rows blockIdx.x, + gridDim.x, ...; 16-byte chunks along the row
    for (long long r = blockIdx.x; r < n; r += gridDim.x) {
        long long s = idx[r];
        s = s < 0 ? 0 : (s >= N ? N - 1 : s);
        const uint4* src = x + s * d8;
        uint4* dst = out + r * d8;
        for (int c = threadIdx.x; c < d8; c += 128) dst[c] = __ldg(src + c);
    }
}
__global__ void __launch_bounds__(128) gather_rows_kernel_h(const __half* __restrict__ x, long long N, int d,
                                                            const long long* __restrict__ idx, long long n,
                                                            __half* __restrict__ out) {
    for (long long r = blockIdx.x; r < n; r += gridDim.x) {
        long long s = idx[r];
        s = s < 0 ? 0 : (s >= N ? N - 1 : s);
        for (int c = threadIdx.x; c < d; c += 128) out[r * d + c] = x[s * d + c];
    }
}

__global__ void jacobi_sweep_end_kernel(unsigned long long* __restrict__ state) { state[6] += 1ull; }

// DM_JACOBI_GRAPH=0 keeps the plain launch loop
static bool jacobi_use_graph() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DM_JACOBI_GRAPH");
        v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
    }
    return v != 0;
}
// capture happens on a private stream (the caller's may be the legacy default stream, which cannot be captured)
static cudaStream_t jacobi_capture_stream() {
    static thread_local cudaStream_t cap[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (cap[dev] == nullptr && cudaStreamCreateWithFlags(&cap[dev], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        cap[dev] = nullptr;
    }
    return cap[dev];
}

// One sweep = n_even - 1 dependent rounds of ~2 us kernels: launched one by one the host (and the launch path) bound the
// solve.  The rounds of ONE sweep are captured into a CUDA graph once per solve (the sweep index is read from
// state[6] on the device, so every sweep replays the same graph) and the graph is launched max_sweeps times.
static int jacobi_rows(double* W, int d, int max_sweeps, double tol, unsigned long long* state, cudaStream_t st) {
    frob2_kernel<<<1, 256, 0, st>>>(W, (long long)d * d, state);
    DM_LAUNCHED();
    const int n_even = d + (d & 1);
    const size_t smem = 2 * (size_t)d * sizeof(double);
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    DM_CUDA(cudaStreamIsCapturing(st, &capturing));
    cudaStream_t cap = nullptr;
    if (jacobi_use_graph() && capturing == cudaStreamCaptureStatusNone && n_even - 1 >= 16)
        cap = jacobi_capture_stream();
    if (cap != nullptr) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        DM_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        for (int r = 0; r < n_even - 1; ++r)
            jacobi_round_kernel<<<n_even / 2, kJacThreads, smem, cap>>>(W, d, n_even, r, -1, tol, state);
        jacobi_sweep_end_kernel<<<1, 1, 0, cap>>>(state);
        cudaError_t e = cudaStreamEndCapture(cap, &graph);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph != nullptr) cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(DM_ERR_CUDA, "%s: capturing a Jacobi sweep failed: %s", __func__, cudaGetErrorString(e));
        }
        for (int s = 0; s < max_sweeps && e == cudaSuccess; ++s) e = cudaGraphLaunch(exec, st);
        cudaGraphExecDestroy(exec);  // released once the launches in flight have completed
        if (e != cudaSuccess)
            return fail(DM_ERR_CUDA, "%s: launching a Jacobi sweep failed: %s", __func__, cudaGetErrorString(e));
        g_launches.fetch_add((unsigned long long)max_sweeps * (unsigned long long)n_even, std::memory_order_relaxed);
        return DM_OK;
    }
    for (int s = 0; s < max_sweeps; ++s)
        for (int r = 0; r < n_even - 1; ++r) {
            jacobi_round_kernel<<<n_even / 2, kJacThreads, smem, st>>>(W, d, n_even, r, s, tol, state);
            DM_LAUNCHED();
        }
    return DM_OK;
}

}  // namespace dm

using namespace dm;

extern "C" long long dm_frechet_workspace_doubles(int d) { return 3LL * d * d + 2LL * d + 16; }

extern "C" int dm_sym_eig_jacobi(double* W, int d, int max_sweeps, double tol, double* eig, unsigned long long* state,
                                 dm_stream_t stream) {
    DM_REQUIRE(W && eig && state && d >= 1 && d <= 2048 && max_sweeps >= 1 && max_sweeps <= 64 && tol > 0);
    cudaStream_t st = as_stream(stream);
    if (d > 1) {
        int rc = jacobi_rows(W, d, max_sweeps, tol, state, st);
        if (rc != DM_OK) return rc;
    } else {
        DM_CUDA(cudaMemsetAsync(state, 0, 5 * sizeof(unsigned long long), st));
    }
    row_norm_kernel<<<d, kJacThreads, 0, st>>>(W, d, eig, 0);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_frechet_distance(const double* mu1, const double* cov1, const double* mu2, const double* cov2, int d,
                                   int max_sweeps, double tol, double* work, double* out, dm_stream_t stream) {
    DM_REQUIRE(mu1 && cov1 && mu2 && cov2 && work && out && d >= 1 && d <= 2048);
    DM_REQUIRE(max_sweeps >= 1 && max_sweeps <= 64 && tol > 0);
    cudaStream_t st = as_stream(stream);
    const size_t dd = (size_t)d * d;
    double* W = work;            // C1 -> L Q^T -> F
    double* T = work + dd;       // F C2
    double* Mm = work + 2 * dd;  // F C2 F^T
    double* eig = work + 3 * dd; // [2][d]
    unsigned long long* st1 = reinterpret_cast<unsigned long long*>(work + 3 * dd + 2 * (size_t)d);
    unsigned long long* st2 = st1 + 8;
    DM_CUDA(cudaMemcpyAsync(W, cov1, dd * sizeof(double), cudaMemcpyDeviceToDevice, st));
    DM_CUDA(cudaMemsetAsync(st1, 0, 16 * sizeof(unsigned long long), st));
    if (d > 1) {
        int rc = jacobi_rows(W, d, max_sweeps, tol, st1, st);
        if (rc != DM_OK) return rc;
    }
    row_norm_kernel<<<d, kJacThreads, 0, st>>>(W, d, eig, 1);
    DM_LAUNCHED();
    const dim3 grid((d + 63) / 64, (d + 63) / 64);
    dgemm_kernel<false><<<grid, 256, 0, st>>>(W, cov2, T, d, d, d);
    DM_LAUNCHED();
    dgemm_kernel<true><<<grid, 256, 0, st>>>(T, W, Mm, d, d, d);
    DM_LAUNCHED();
    if (d > 1) {
        int rc = jacobi_rows(Mm, d, max_sweeps, tol, st2, st);
        if (rc != DM_OK) return rc;
    }
    row_norm_kernel<<<d, kJacThreads, 0, st>>>(Mm, d, eig + d, 0);
    DM_LAUNCHED();
    frechet_finish_kernel<<<1, 256, 0, st>>>(mu1, cov1, mu2, cov2, d, eig + d, st1, st2, out);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_fad_gather_rows(const void* x_f16, long long N, int d, const long long* idx, long long n,
                                  void* out_f16, dm_stream_t stream) {
    DM_REQUIRE(x_f16 && idx && out_f16 && N > 0 && d > 0 && n > 0);
    const int blocks = (int)(n < (long long)num_sms() * 16 ? n : (long long)num_sms() * 16);
    const bool vec = d % 8 == 0 && (reinterpret_cast<uintptr_t>(x_f16) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out_f16) & 15) == 0;
    if (vec)
        gather_rows_kernel<<<blocks, 128, 0, as_stream(stream)>>>(static_cast<const uint4*>(x_f16), N, d / 8, idx, n,
                                                                  static_cast<uint4*>(out_f16));
    else
        gather_rows_kernel_h<<<blocks, 128, 0, as_stream(stream)>>>(static_cast<const __half*>(x_f16), N, d, idx, n,
                                                                    static_cast<__half*>(out_f16));
    DM_LAUNCHED();
    return DM_OK;
}
