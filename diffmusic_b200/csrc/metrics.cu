// Evaluation metrics next to the guidance path, on the same STFT building blocks:
//
//   * LogSpectralDistance.score (diffmusic/metrics/lsd.py:17-40): |STFT| of a background and an eval clip (n_fft 1024,
//     Hann window, centred), log10(. + eps), squared difference, sqrt of the mean over the 513 bins per frame.
//     The frame-pair pipeline of stft_pair.cuh carries frame t of the BACKGROUND clip as frame A and frame t of the EVAL
//     clip as frame B through one FFT: both magnitudes of a bin land side by side in shared memory and the per-frame
//     distance is reduced in place -- no spectrogram is ever written (the reference materialises two of them per clip).
//   * MeanSquaredError.score (diffmusic/metrics/mse.py:9-29): per-clip mean of squared differences.
//   Both sanitise their inputs like the reference's np.nan_to_num(nan=0, posinf=1, neginf=-1).
#include "dm_common.cuh"
#include "stft_pair.cuh"

namespace dm {

constexpr int kLsdThreads = 256;
constexpr int kLsdGroups = kLsdThreads / kGroupThreads;

__device__ __forceinline__ float nan_to_num(float v) {
    if (v != v) return 0.f;
    if (isinf(v)) return v > 0.f ? 1.f : -1.f;
    return v;
}
__device__ __forceinline__ void lsd_group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kGroupThreads));
}

struct LsdParams {
    StftTables tab;
    const float* ref;
    const float* est;
    long long ref_bstride, est_bstride, L, T;
    int hop, nf, pad_reflect, sanitize_ref;
    float eps;
    float* out;  // (B, T) per-frame distance
};

__global__ void __launch_bounds__(kLsdThreads, 2) lsd_pair_kernel(const LsdParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, g = tid / kGroupThreads, gt = tid % kGroupThreads;
    const int b = blockIdx.y, tile = blockIdx.x;
    const long long f0 = (long long)tile * p.nf;
    const int nfr = (int)min((long long)p.nf, p.T - f0);
    const int span = (nfr - 1) * p.hop + kNfft;
    const int span_alloc = ((p.nf - 1) * p.hop + kNfft + 3) & ~3;
    const long long base = f0 * p.hop;  // first padded-signal index of the tile

    float* sa = smem;                 // background span
    float* sb = sa + span_alloc;      // eval span
    float* win = sb + span_alloc;     // [1024]
    float* grp = win + kNfft;         // [kLsdGroups][kPairSmemFloats]
    float* red = grp + kLsdGroups * kPairSmemFloats;  // [kLsdGroups][2]
    PairSmem s;
    s.a = reinterpret_cast<c2*>(grp + g * kPairSmemFloats);
    s.b = reinterpret_cast<c2*>(grp + g * kPairSmemFloats + 4 * kH);
    PairConsts pc;
    load_pair_consts(gt, p.tab, pc);

    const float* rb = p.ref + (long long)b * p.ref_bstride;
    const float* eb = p.est + (long long)b * p.est_bstride;
    for (int i = tid; i < span; i += kLsdThreads) {
        float va = 0.f, vb = 0.f;
        long long j = base + i - kNfft / 2;
        bool inside = j >= 0 && j < p.L;
        if (!inside && p.pad_reflect) {
            j = reflect_src(base + i, p.L);
            inside = true;
        }
        if (inside) {
            va = __ldg(rb + j);
            vb = nan_to_num(__ldg(eb + j));
            if (p.sanitize_ref) va = nan_to_num(va);
        }
        sa[i] = va;
        sb[i] = vb;
    }
    reinterpret_cast<float4*>(win)[tid] = __ldg(reinterpret_cast<const float4*>(p.tab.window) + tid);
    __syncthreads();

    for (int f = g; f < nfr; f += kLsdGroups) {
        PairX x;
        pair_fwd_pass1(gt, sa + f * p.hop, sb + f * p.hop, win, s);
        lsd_group_sync(g);
        pair_fwd_pass2(gt, pc, s);
        lsd_group_sync(g);
        pair_fwd_pass3(gt, pc, s);
        lsd_group_sync(g);
        pair_unpack<kModePhaseWav>(gt, pc, s, x);  // P[k] = (|X_ref[k]|, |X_est[k]|)
        lsd_group_sync(g);
        const f2* P = pair_energy(s);
        float acc = 0.f;
        for (int k = gt; k < kBins; k += kGroupThreads) {
            const f2 m = P[k];
            const float d = log10f(m.x + p.eps) - log10f(m.y + p.eps);
            acc = fmaf(d, d, acc);
        }
        acc = warp_sum(acc);
        if ((gt & 31) == 0) red[2 * g + (gt >> 5)] = acc;
        lsd_group_sync(g);
        if (gt == 0) p.out[(long long)b * p.T + f0 + f] = sqrtf((red[2 * g] + red[2 * g + 1]) * (1.0f / kBins));
        // the next frame's pass 1 only writes `a`; P (in b) and red are rewritten after further group barriers
    }
}

constexpr int kMseThreads = 256;
constexpr int kMseChunk = 16384;

__global__ void __launch_bounds__(kMseThreads) mse_partial_kernel(const float* __restrict__ ref, long long ref_bstride,
                                                                  const float* __restrict__ est, long long est_bstride,
                                                                  long long n, int nchunks,
                                                                  double* __restrict__ partial) {
    __shared__ double red[kMseThreads / 32];
    const int b = blockIdx.y;
    const long long i0 = (long long)blockIdx.x * kMseChunk;
    const long long i1 = min(n, i0 + kMseChunk);
    const float* r = ref + (long long)b * ref_bstride;
    const float* e = est + (long long)b * est_bstride;
    float acc = 0.f;
    for (long long i = i0 + threadIdx.x; i < i1; i += kMseThreads) {
        const float d = nan_to_num(__ldg(r + i)) - nan_to_num(__ldg(e + i));
        acc = fmaf(d, d, acc);
    }
    double a = warp_sum((double)acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kMseThreads / 32; ++w) t += red[w];
        partial[(long long)b * nchunks + blockIdx.x] = t;
    }
}
__global__ void mse_finish_kernel(const double* __restrict__ partial, int nchunks, long long n, int B,
                                  float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double t = 0.0;
    for (int c = 0; c < nchunks; ++c) t += partial[(long long)b * nchunks + c];
    out[b] = (float)(t / (double)n);
}

static size_t lsd_smem_bytes(int nf, int hop) {
    size_t span = ((size_t)(nf - 1) * hop + kNfft + 3) & ~(size_t)3;
    return (2 * span + kNfft + (size_t)kLsdGroups * kPairSmemFloats + 2 * kLsdGroups) * sizeof(float);
}

}  // namespace dm

using namespace dm;

extern "C" int dm_lsd_frames(const dm_stft_tables* tab, const float* ref, long long ref_bstride, const float* est,
                             long long est_bstride, long long L, int B, int hop, int pad_reflect, int sanitize_ref,
                             float eps, float* out, dm_stream_t stream) {
    DM_REQUIRE(tab && ref && est && out && B > 0 && L > 0);
    DM_REQUIRE(hop > 0 && hop <= kNfft && (hop & 1) == 0);
    DM_REQUIRE(!pad_reflect || L > kNfft / 2);
    LsdParams p;
    p.tab = StftTables{tab->window, reinterpret_cast<const cf*>(tab->tw512), reinterpret_cast<const cf*>(tab->w1024),
                       tab->mel_kstart, tab->mel_klen, tab->mel_w, tab->mel_wstride, tab->bin_m0, tab->bin_w0,
                       tab->bin_w1};
    p.ref = ref;
    p.est = est;
    p.ref_bstride = ref_bstride;
    p.est_bstride = est_bstride;
    p.L = L;
    p.T = 1 + L / hop;
    p.hop = hop;
    p.nf = hop <= 256 ? 16 : 8;  // frames per tile: ~30 KB of staged signal either way
    p.pad_reflect = pad_reflect;
    p.sanitize_ref = sanitize_ref;
    p.eps = eps;
    p.out = out;
    const size_t smem = lsd_smem_bytes(p.nf, hop);
    DM_SMEM_ONCE(lsd_pair_kernel, smem);
    const dim3 grid((unsigned)((p.T + p.nf - 1) / p.nf), B);
    lsd_pair_kernel<<<grid, kLsdThreads, smem, as_stream(stream)>>>(p);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" long long dm_mse_num_chunks(long long n) { return n <= 0 ? 0 : (n + kMseChunk - 1) / kMseChunk; }

extern "C" int dm_mse(const float* ref, long long ref_bstride, const float* est, long long est_bstride, long long n,
                      int B, double* partial, float* out, dm_stream_t stream) {
    DM_REQUIRE(ref && est && partial && out && n > 0 && B > 0);
    const int nchunks = (int)dm_mse_num_chunks(n);
    mse_partial_kernel<<<dim3(nchunks, B), kMseThreads, 0, as_stream(stream)>>>(ref, ref_bstride, est, est_bstride, n,
                                                                               nchunks, partial);
    DM_LAUNCHED();
    mse_finish_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(partial, nchunks, n, B, out);
    DM_LAUNCHED();
    return DM_OK;
}
