// mel_spectrogram_to_waveform_with_phase (diffmusic/pipelines/pipeline_musicldm.py:263-301, plpeline_audioldm2.py:681;
// SURVEY.md 8f rank 3): mel -> torchaudio InverseMelScale -> times exp(i * original_phase) -> torch.istft -> clip / pad.
//
//   * InverseMelScale (driver "gels", filterbank 513 x 64 of full column rank) is the minimum-norm least-squares solution
//     followed by relu: a FIXED 513 x 64 matrix W = fb (fb^T fb)^-1 applied to every frame.  The host computes W once in
//     fp64 (tables.py); here it is a per-tile product out of L2 with the mel tile broadcast from shared memory.  The
//     (B, 513, T) linear spectrogram and the complex spectrogram the reference materialises are never written.
//   * torch.istft(n_fft 1024, hop, win 1024, window=None -> rectangular, center=True): irfft of every frame, overlap-add,
//     division by the window envelope (= the number of frames covering a sample), trimmed by 512 on both sides; the
//     frame-pair inverse FFT of stft_pair.cuh carries two frames per 64-thread group, the overlap-add runs in shared
//     memory per tile and tiles are joined with atomics (<= 2 tiles meet in a sample, so the sum is order-independent).
#include "dm_common.cuh"
#include "stft_pair.cuh"

namespace dm {

constexpr int kIstftThreads = 256;
constexpr int kIstftGroups = kIstftThreads / kGroupThreads;
constexpr int kIstftFrames = 8;                  // frames per tile = one round of 4 groups x 2 frames
constexpr int kSpecLd = 516;                     // cells per staged frame pair (513 bins, padded)

__device__ __forceinline__ void istft_group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kGroupThreads));
}

struct IstftParams {
    StftTables tab;
    const float* winv_t;  // (64, 513): W^T
    const float* mel;
    long long mel_bstride, mel_mstride, mel_tstride;
    const float* phase;   // (513, T) per clip (phase_bstride) or shared (0)
    long long phase_bstride, T, ola_len;
    int hop;
    float* ola;           // (B, ola_len) zeroed
};

// phases of bin k for the frames of the tile (rows of the (513, T) phase are 32-byte runs per thread: issued early, the
// latency hides behind the band loop)
__device__ __forceinline__ void istft_load_phase(const float* __restrict__ phb, long long T, long long f0, int nfr, int k,
                                                 float (&ph)[kIstftFrames]) {
    const float* src = phb + (long long)k * T + f0;
#pragma unroll
    for (int f = 0; f < kIstftFrames; ++f) ph[f] = f < nfr ? __ldg(src + f) : 0.f;
}
// bin k of the tile: relu, times exp(i phase), stored as (frame 2q, frame 2q + 1) cells; frames past the end have a
// zero magnitude (their mel tile is zero)
__device__ __forceinline__ void istft_emit(c2* spec, int k, const float (&lin)[kIstftFrames],
                                           const float (&ph)[kIstftFrames]) {
    float xr[kIstftFrames], xi[kIstftFrames];
#pragma unroll
    for (int f = 0; f < kIstftFrames; ++f) {
        const float mag = lin[f] < 0.f ? 0.f : lin[f];  // torch.relu: NaN stays NaN
        float sn, cs;
        sincosf(ph[f], &sn, &cs);
        xr[f] = mag * cs;
        xi[f] = mag * sn;
    }
#pragma unroll
    for (int q = 0; q < kIstftFrames / 2; ++q)
        spec[q * kSpecLd + k] = c2{xr[2 * q], xi[2 * q], xr[2 * q + 1], xi[2 * q + 1]};
}

__global__ void __launch_bounds__(kIstftThreads, 2) istft_pair_kernel(const IstftParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, g = tid / kGroupThreads, gt = tid % kGroupThreads;
    const int b = blockIdx.y, tile = blockIdx.x;
    const long long f0 = (long long)tile * kIstftFrames;
    const int nfr = (int)min((long long)kIstftFrames, p.T - f0);
    const int span = (nfr - 1) * p.hop + kNfft;
    const int span_alloc = ((kIstftFrames - 1) * p.hop + kNfft + 3) & ~3;
    const int hop2 = p.hop >> 1;

    c2* spec = reinterpret_cast<c2*>(smem);                        // [kIstftFrames / 2][kSpecLd]
    float* grp = smem + (kIstftFrames / 2) * kSpecLd * 4;          // [kIstftGroups][kPairSmemFloats]
    float* acc = grp + kIstftGroups * kPairSmemFloats;             // [span_alloc]
    float* mel_s = acc + span_alloc;                               // [64][kIstftFrames]
    PairSmem s;
    s.a = reinterpret_cast<c2*>(grp + g * kPairSmemFloats);
    s.b = reinterpret_cast<c2*>(grp + g * kPairSmemFloats + 4 * kH);
    PairConsts pc;
    load_pair_consts(gt, p.tab, pc);

    // ---- phases of the two bins this thread will emit: in flight while the mel tile is staged ----
    const float* phb = p.phase + (long long)b * p.phase_bstride;
    float ph0[kIstftFrames], ph1[kIstftFrames];
    istft_load_phase(phb, p.T, f0, nfr, tid, ph0);
    istft_load_phase(phb, p.T, f0, nfr, tid + kIstftThreads, ph1);

    // ---- stage the mel tile (frames past the end read as zero) and clear the overlap-add span ----
    const float* mb = p.mel + (long long)b * p.mel_bstride;
    for (int i = tid; i < kMels * kIstftFrames; i += kIstftThreads) {
        const int m = p.mel_mstride == 1 ? i % kMels : i / kIstftFrames;  // run along the contiguous axis
        const int f = p.mel_mstride == 1 ? i / kMels : i % kIstftFrames;
        mel_s[m * kIstftFrames + f] = f < nfr ? __ldg(mb + m * p.mel_mstride + (f0 + f) * p.mel_tstride) : 0.f;
    }
    for (int i = tid; i < span_alloc / 4; i += kIstftThreads)
        reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // ---- linear magnitude of every bin of the tile: relu(W mel), times exp(i phase) -> pair-major spectrum cells ----
    // thread t carries bins t and t + 256 through one pass over the 64 bands (16 accumulators per W / mel load pair);
    // the odd bin 512 is split over warp 0 (lane = 8 * band quarter + frame) and reduced with two shuffles.
    {
        float l0[kIstftFrames], l1[kIstftFrames];
#pragma unroll
        for (int f = 0; f < kIstftFrames; ++f) l0[f] = l1[f] = 0.f;
        const float* w = p.winv_t + tid;
#pragma unroll 4
        for (int m = 0; m < kMels; ++m) {
            const float w0 = __ldg(w + m * kBins), w1 = __ldg(w + m * kBins + kIstftThreads);
            const float4 a = reinterpret_cast<const float4*>(mel_s + m * kIstftFrames)[0];
            const float4 c = reinterpret_cast<const float4*>(mel_s + m * kIstftFrames)[1];
            const float mv[kIstftFrames] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int f = 0; f < kIstftFrames; ++f) {
                l0[f] = fmaf(w0, mv[f], l0[f]);
                l1[f] = fmaf(w1, mv[f], l1[f]);
            }
        }
        static_assert(kIstftFrames == 8 && kBins == 2 * kIstftThreads + 1, "bin ownership of the tile product");
        istft_emit(spec, tid, l0, ph0);
        istft_emit(spec, tid + kIstftThreads, l1, ph1);
    }
    if (tid < 32) {
        const int f = tid & 7, part = tid >> 3;
        float lin = 0.f;
#pragma unroll 4
        for (int m = 16 * part; m < 16 * part + 16; ++m)
            lin = fmaf(__ldg(p.winv_t + m * kBins + (kBins - 1)), mel_s[m * kIstftFrames + f], lin);
        lin += __shfl_xor_sync(0xffffffffu, lin, 8);
        lin += __shfl_xor_sync(0xffffffffu, lin, 16);
        if (part == 0) {
            const float mag = lin < 0.f ? 0.f : lin;
            float sn = 0.f, cs = 0.f;
            if (f < nfr) sincosf(__ldg(phb + (long long)(kBins - 1) * p.T + f0 + f), &sn, &cs);
            float* cell = reinterpret_cast<float*>(spec + (f >> 1) * kSpecLd + (kBins - 1)) + 2 * (f & 1);
            cell[0] = mag * cs;
            cell[1] = mag * sn;
        }
    }
    __syncthreads();

    // ---- inverse FFT of the group's frame pair, then the overlap-add group by group ----
    const int fa = 2 * g;
    const bool active = fa < nfr, active_b = fa + 1 < nfr;  // a missing frame B has a zero spectrum and is dropped
    cf va[8], vb[8];
    if (active) {
        PairX x;
        pair_load_spectrum(gt, spec + g * kSpecLd, x);
        pair_pack_spectrum(gt, pc, s, x);
        istft_group_sync(g);
        pair_inv_pass1(gt, s);
        istft_group_sync(g);
        pair_inv_pass2(gt, pc, s);
        istft_group_sync(g);
        pair_inv_pass3(gt, pc, s, va, vb);
    }
    f2* acc2 = reinterpret_cast<f2*>(acc);
    for (int turn = 0; turn < kIstftGroups; ++turn) {
        if (turn == g && active) {
            pair_ola_add_rect(gt, va, acc2 + fa * hop2);
            istft_group_sync(g);  // frame B overlaps frame A
            if (active_b) pair_ola_add_rect(gt, vb, acc2 + (fa + 1) * hop2);
        }
        __syncthreads();
    }
    float* ob = p.ola + (long long)b * p.ola_len + f0 * p.hop;
    for (int i = tid; i < span; i += kIstftThreads) atomicAdd(ob + i, acc[i]);
}

// out[b, j] = ola[b, j + 512] / 1024 / (frames covering padded sample j + 512) for j < hop (T - 1), zero padding after.
// Four consecutive samples per thread (128-bit loads / stores when the rows allow it), 32-bit index arithmetic.
// ola * (1/1024) is exact; the division by the frame count is the one rounding, as in torch's y / window_envelop.
__device__ __forceinline__ float istft_sample(float v, unsigned i, unsigned hop, unsigned t_last) {
    const unsigned t_hi = min(t_last, i / hop);
    const unsigned t_lo = i >= (unsigned)kNfft ? (i - kNfft) / hop + 1u : 0u;
    return v * (1.0f / kNfft) / (float)(t_hi - t_lo + 1u);
}
__global__ void __launch_bounds__(256) istft_finish_kernel(const float* __restrict__ ola, long long ola_len, long long T,
                                                           int hop, float* __restrict__ out, long long out_len) {
    const int b = blockIdx.y;
    const unsigned n_valid = (unsigned)((long long)hop * (T - 1)), t_last = (unsigned)(T - 1), h = (unsigned)hop;
    const float* src = ola + (long long)b * ola_len + kNfft / 2;
    float* dst = out + (long long)b * out_len;
    const bool vec = ((ola_len | out_len) & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(ola) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    for (long long j0 = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); j0 < out_len;
         j0 += 4LL * gridDim.x * blockDim.x) {
        const unsigned j = (unsigned)j0;
        if (vec && j + 4 <= n_valid && j0 + 4 <= out_len) {
            float4 v = *reinterpret_cast<const float4*>(src + j);
            v.x = istft_sample(v.x, j + 512u, h, t_last);
            v.y = istft_sample(v.y, j + 513u, h, t_last);
            v.z = istft_sample(v.z, j + 514u, h, t_last);
            v.w = istft_sample(v.w, j + 515u, h, t_last);
            *reinterpret_cast<float4*>(dst + j) = v;
        } else {
            for (unsigned q = 0; q < 4 && j0 + q < out_len; ++q)
                dst[j + q] = (j + q < n_valid) ? istft_sample(src[j + q], j + q + 512u, h, t_last) : 0.f;
        }
    }
}

// ---- companion forward: waveform_to_spectrogram (diffmusic/utils.py:11-20; run.py:305 takes the phase of the degraded
// clip from it): torch.stft(1024, hop, 1024, window = tab.window (the reference passes none: rectangular), centred with
// reflection) -> |X| and angle(X), both (B, 513, T).  Frames t and t + 1 of a clip ride through one pair FFT; every
// thread emits the magnitude and atan2 of the bins it owns straight from its registers.
constexpr int kSpecThreads = 256;
constexpr int kSpecGroups = kSpecThreads / kGroupThreads;

struct SpectrogramParams {
    StftTables tab;
    const float* wav;
    long long wav_bstride, L, T;
    int hop, nf;
    float* mag;    // (B, 513, T) or null
    float* phase;  // (B, 513, T) or null
};

__device__ __forceinline__ void spec_emit(const SpectrogramParams& p, long long row, long long ta, bool has_b, cf xa,
                                          cf xb) {
    if (p.mag) {
        p.mag[row + ta] = pair_bin_energy<kModePhaseWav>(xa);
        if (has_b) p.mag[row + ta + 1] = pair_bin_energy<kModePhaseWav>(xb);
    }
    if (p.phase) {
        p.phase[row + ta] = atan2f(xa.y, xa.x);
        if (has_b) p.phase[row + ta + 1] = atan2f(xb.y, xb.x);
    }
}

__global__ void __launch_bounds__(kSpecThreads, 2) spectrogram_pair_kernel(const SpectrogramParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, g = tid / kGroupThreads, gt = tid % kGroupThreads;
    const int b = blockIdx.y, tile = blockIdx.x;
    const long long f0 = (long long)tile * p.nf;
    const int nfr = (int)min((long long)p.nf, p.T - f0);
    const int span = (nfr - 1) * p.hop + kNfft;
    const int span_alloc = ((p.nf - 1) * p.hop + kNfft + 3) & ~3;
    const long long base = f0 * p.hop;

    float* sig = smem;
    float* win = sig + span_alloc;
    float* grp = win + kNfft;
    PairSmem s;
    s.a = reinterpret_cast<c2*>(grp + g * kPairSmemFloats);
    s.b = reinterpret_cast<c2*>(grp + g * kPairSmemFloats + 4 * kH);
    PairConsts pc;
    load_pair_consts(gt, p.tab, pc);

    const float* wb = p.wav + (long long)b * p.wav_bstride;
    for (int i = tid; i < span; i += kSpecThreads) sig[i] = __ldg(wb + reflect_src(base + i, p.L));
    reinterpret_cast<float4*>(win)[tid] = __ldg(reinterpret_cast<const float4*>(p.tab.window) + tid);
    __syncthreads();

    for (int fa = 2 * g; fa < nfr; fa += 2 * kSpecGroups) {
        const bool has_b = fa + 1 < nfr;
        const int fb = has_b ? fa + 1 : fa;
        PairX x;
        pair_fwd_pass1(gt, sig + fa * p.hop, sig + fb * p.hop, win, s);
        istft_group_sync(g);
        pair_fwd_pass2(gt, pc, s);
        istft_group_sync(g);
        pair_fwd_pass3(gt, pc, s);
        istft_group_sync(g);
        pair_unpack<kModePhaseWav>(gt, pc, s, x);
        const long long ta = f0 + fa;
        const long long rows = (long long)b * kBins;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = gt + 64 * i;  // k = 0: lo = bin 0, hi = bin 512
            spec_emit(p, (rows + k) * p.T, ta, has_b, x.lo[i][0], x.lo[i][1]);
            spec_emit(p, (rows + kH - k) * p.T, ta, has_b, x.hi[i][0], x.hi[i][1]);
        }
        if (gt == 0) spec_emit(p, (rows + kH / 2) * p.T, ta, has_b, x.q[0], x.q[1]);
        istft_group_sync(g);  // the unpack's reads of `a` (and writes of P in `b`) are over before the next pass 1
    }
}

static size_t spectrogram_smem_bytes(int nf, int hop) {
    size_t span = ((size_t)(nf - 1) * hop + kNfft + 3) & ~(size_t)3;
    return (span + kNfft + (size_t)kSpecGroups * kPairSmemFloats) * sizeof(float);
}

static size_t istft_smem_bytes(int hop) {
    size_t span = ((size_t)(kIstftFrames - 1) * hop + kNfft + 3) & ~(size_t)3;
    return ((size_t)(kIstftFrames / 2) * kSpecLd * 4 + (size_t)kIstftGroups * kPairSmemFloats + span +
            (size_t)kMels * kIstftFrames) * sizeof(float);
}

}  // namespace dm

using namespace dm;

extern "C" long long dm_istft_workspace_floats(int B, long long T, int hop) {
    if (B <= 0 || T <= 0 || hop <= 0) return 0;
    return (long long)B * ((long long)hop * (T - 1) + kNfft);
}

extern "C" int dm_istft_mel_phase(const dm_stft_tables* tab, const float* winv_t, const float* mel,
                                  long long mel_bstride, long long mel_mstride, long long mel_tstride,
                                  const float* phase, long long phase_bstride, int B, long long T, int hop, float* ola,
                                  float* out, long long out_len, dm_stream_t stream) {
    DM_REQUIRE(tab && winv_t && mel && phase && ola && out && B > 0 && T > 0 && out_len > 0);
    DM_REQUIRE(hop > 0 && hop <= kNfft && (hop & 1) == 0);
    IstftParams p;
    p.tab = StftTables{tab->window, reinterpret_cast<const cf*>(tab->tw512), reinterpret_cast<const cf*>(tab->w1024),
                       tab->mel_kstart, tab->mel_klen, tab->mel_w, tab->mel_wstride, tab->bin_m0, tab->bin_w0,
                       tab->bin_w1};
    p.winv_t = winv_t;
    p.mel = mel;
    p.mel_bstride = mel_bstride;
    p.mel_mstride = mel_mstride;
    p.mel_tstride = mel_tstride;
    p.phase = phase;
    p.phase_bstride = phase_bstride;
    p.T = T;
    p.hop = hop;
    p.ola_len = (long long)hop * (T - 1) + kNfft;
    p.ola = ola;
    DM_CUDA(cudaMemsetAsync(ola, 0, sizeof(float) * (size_t)B * (size_t)p.ola_len, as_stream(stream)));
    const size_t smem = istft_smem_bytes(hop);
    if (smem > 227 * 1024) return fail(DM_ERR_UNSUPPORTED, "%s: tile needs %zu B of shared memory", __func__, smem);
    DM_SMEM_ONCE(istft_pair_kernel, smem);
    const dim3 grid((unsigned)((T + kIstftFrames - 1) / kIstftFrames), B);
    istft_pair_kernel<<<grid, kIstftThreads, smem, as_stream(stream)>>>(p);
    DM_LAUNCHED();
    DM_REQUIRE(p.ola_len < (1LL << 31) && out_len < (1LL << 31));
    const unsigned fx = (unsigned)min((out_len + 1023) / 1024, 1184LL);
    istft_finish_kernel<<<dim3(fx, B), 256, 0, as_stream(stream)>>>(ola, p.ola_len, T, hop, out, out_len);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_stft_spectrogram(const dm_stft_tables* tab, const float* wav, long long wav_bstride, long long L, int B,
                                   int hop, float* mag, float* phase, dm_stream_t stream) {
    DM_REQUIRE(tab && wav && (mag || phase) && B > 0);
    DM_REQUIRE(L > kNfft / 2);  // reflect padding needs pad < length (torch.stft raises otherwise)
    DM_REQUIRE(hop > 0 && hop <= kNfft && (hop & 1) == 0);
    SpectrogramParams p;
    p.tab = StftTables{tab->window, reinterpret_cast<const cf*>(tab->tw512), reinterpret_cast<const cf*>(tab->w1024),
                       tab->mel_kstart, tab->mel_klen, tab->mel_w, tab->mel_wstride, tab->bin_m0, tab->bin_w0,
                       tab->bin_w1};
    p.wav = wav;
    p.wav_bstride = wav_bstride;
    p.L = L;
    p.T = 1 + L / hop;
    p.hop = hop;
    p.nf = hop <= 256 ? 16 : 8;
    p.mag = mag;
    p.phase = phase;
    const size_t smem = spectrogram_smem_bytes(p.nf, hop);
    DM_SMEM_ONCE(spectrogram_pair_kernel, smem);
    const dim3 grid((unsigned)((p.T + p.nf - 1) / p.nf), B);
    spectrogram_pair_kernel<<<grid, kSpecThreads, smem, as_stream(stream)>>>(p);
    DM_LAUNCHED();
    return DM_OK;
}
