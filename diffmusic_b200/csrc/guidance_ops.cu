// Bandwidth-bound pieces of the guidance path: wav-space residual, reflect-fold + mask adjoint, sinc resampling
// forward / adjoint, noise add.  All coalesced, per-clip reductions through per-chunk partial slots (deterministic).
#include <algorithm>

#include "dm_common.cuh"
#include "fir_poly.cuh"

namespace dm {

constexpr int kEwThreads = 256;

// cotangent of the un-padded signal at j from a reflect-padded cotangent (pad = 512 each side)
__device__ __forceinline__ float fold_at(const float* __restrict__ yp, long long j, long long Ly) {
    float v = yp[512 + j];
    if (j >= 1 && j <= 512) v += yp[512 - j];
    if (j >= Ly - 513 && j <= Ly - 2) v += yp[512 + 2 * (Ly - 1) - j];
    return v;
}
__device__ __forceinline__ float ybar_at(const float* __restrict__ yb, int pad, long long j, long long Ly) {
    return pad ? fold_at(yb, j, Ly) : yb[j];
}

// ---------------------------------------------------------------------------------------------------- residual
__global__ void __launch_bounds__(kEwThreads) residual_wav_kernel(const void* __restrict__ y, int y_io,
                                                                  long long y_bstride,
                                                                  long long n, const float* __restrict__ mask,
                                                                  const float* __restrict__ meas,
                                                                  long long meas_bstride, float* __restrict__ ybar,
                                                                  float* __restrict__ partial, int ntiles) {
    __shared__ float red[kEwThreads / 32];
    const int b = blockIdx.y;
    const long long lo = (long long)blockIdx.x * DM_RESID_CHUNK;
    const long long hi = min(n, lo + DM_RESID_CHUNK);
    const void* yb = wave_row(y, y_io, (long long)b * y_bstride);
    const float* mb = meas + (long long)b * meas_bstride;
    float* ob = ybar + (long long)b * n;
    float s = 0.f;
    // full chunks of aligned rows: four vectors of 4 samples per thread (128-bit fp32 / 64-bit fp16, bf16 loads), all
    // loads issued before the first use; the same summation order for every waveform dtype
    const bool fast = hi - lo == DM_RESID_CHUNK && ((n | y_bstride | meas_bstride) & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & (y_io == DM_IO_F32 ? 15 : 7)) == 0 &&
                      ((reinterpret_cast<uintptr_t>(meas) | reinterpret_cast<uintptr_t>(ybar) |
                        reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
    if (fast) {
        constexpr int kV = DM_RESID_CHUNK / 4 / kEwThreads;  // 4
        const float4* m4 = reinterpret_cast<const float4*>(mb + lo);
        const float4* k4 = mask ? reinterpret_cast<const float4*>(mask + lo) : nullptr;
        float4* o4 = reinterpret_cast<float4*>(ob + lo);
        float4 v[kV], m[kV], k[kV];
        if (y_io == DM_IO_F32) {
            const float4* y4 = reinterpret_cast<const float4*>(static_cast<const float*>(yb) + lo);
#pragma unroll
            for (int u = 0; u < kV; ++u) v[u] = y4[threadIdx.x + u * kEwThreads];
        } else {
            const uint2* y2 = reinterpret_cast<const uint2*>(static_cast<const unsigned short*>(yb) + lo);
#pragma unroll
            for (int u = 0; u < kV; ++u) {
                const uint2 raw = y2[threadIdx.x + u * kEwThreads];
                float2 a, c;
                if (y_io == DM_IO_F16) {
                    a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                    c = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                } else {
                    a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                    c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                }
                v[u] = make_float4(a.x, a.y, c.x, c.y);
            }
        }
#pragma unroll
        for (int u = 0; u < kV; ++u) {
            m[u] = m4[threadIdx.x + u * kEwThreads];
            k[u] = k4 ? __ldg(k4 + threadIdx.x + u * kEwThreads) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int u = 0; u < kV; ++u) {
            float4 d;
            if (k4) v[u].x *= k[u].x, v[u].y *= k[u].y, v[u].z *= k[u].z, v[u].w *= k[u].w;
            d.x = m[u].x - v[u].x, d.y = m[u].y - v[u].y, d.z = m[u].z - v[u].z, d.w = m[u].w - v[u].w;
            o4[threadIdx.x + u * kEwThreads] = make_float4(-d.x, -d.y, -d.z, -d.w);
            s = fmaf(d.x, d.x, s), s = fmaf(d.y, d.y, s), s = fmaf(d.z, d.z, s), s = fmaf(d.w, d.w, s);
        }
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += kEwThreads) {
            float v = ld_wave(yb, y_io, i);
            if (mask) v *= __ldg(mask + i);
            float d = mb[i] - v;
            ob[i] = -d;
            s = fmaf(d, d, s);
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kEwThreads / 32; ++i) t += red[i];
        partial[(long long)b * ntiles + blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------------- fold adjoint
__global__ void __launch_bounds__(kEwThreads) fold_adjoint_kernel(const float* __restrict__ ybar, int pad,
                                                                  long long Ly, const float* __restrict__ mask,
                                                                  const float* __restrict__ partial, int ntiles,
                                                                  void* __restrict__ dwav, int dwav_io,
                                                                  long long dwav_bstride,
                                                                  float* __restrict__ loss) {
    __shared__ float scratch[2];
    const int b = blockIdx.y;
    pdl_wait();  // cotangent and partial sums of the previous kernel of the chain
    const float* yb = ybar + (long long)b * (Ly + 2 * pad);
    void* ob = wave_row(dwav, dwav_io, (long long)b * dwav_bstride);
    const long long lo = (long long)blockIdx.x * (kEwThreads * 8);
    const long long hi = min(Ly, lo + kEwThreads * 8);
    // interior chunks (no reflected samples fold into them) of 16-byte aligned fp32 rows: 128-bit loads and stores, the
    // loads issued BEFORE the per-clip scale is reduced from the partial sums, so their latency hides behind it
    const bool fast = dwav != nullptr && dwav_io == DM_IO_F32 && hi - lo == kEwThreads * 8 &&
                      (pad == 0 || (lo >= 513 && hi <= Ly - 513)) && ((Ly + 2 * pad) & 3) == 0 &&
                      (dwav_bstride & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(ybar) | reinterpret_cast<uintptr_t>(dwav) |
                        reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (fast) {
        const float4* src = reinterpret_cast<const float4*>(yb + pad + lo);
        v0 = src[threadIdx.x];
        v1 = src[threadIdx.x + kEwThreads];
        if (mask) {
            const float4 m0 = __ldg(reinterpret_cast<const float4*>(mask + lo) + threadIdx.x);
            const float4 m1 = __ldg(reinterpret_cast<const float4*>(mask + lo) + threadIdx.x + kEwThreads);
            v0.x *= m0.x, v0.y *= m0.y, v0.z *= m0.z, v0.w *= m0.w;
            v1.x *= m1.x, v1.y *= m1.y, v1.z *= m1.z, v1.w *= m1.w;
        }
    }
    const float l = clip_loss(partial + (long long)b * ntiles, ntiles, scratch);
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) loss[b] = l;
    if (dwav == nullptr) return;  // loss only
    const float sc = inv_loss(l);
    if (fast) {
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(ob) + lo);
        dst[threadIdx.x] = make_float4(v0.x * sc, v0.y * sc, v0.z * sc, v0.w * sc);
        dst[threadIdx.x + kEwThreads] = make_float4(v1.x * sc, v1.y * sc, v1.z * sc, v1.w * sc);
        return;
    }
    for (long long j = lo + threadIdx.x; j < hi; j += kEwThreads) {
        float v = ybar_at(yb, pad, j, Ly) * sc;
        if (mask) v *= __ldg(mask + j);
        st_wave(ob, dwav_io, j, v);
    }
}

// ---------------------------------------------------------------------------------------------------- resampling
// Both directions stage the input span of a CTA's output chunk in shared memory once (coalesced, with the reflect
// fold and the 1/loss scale applied on the way in for the adjoint) and then run the short FIR out of shared memory
// with 32-bit index arithmetic.
constexpr int kRsChunk = 2048;  // outputs per CTA

__global__ void __launch_bounds__(kEwThreads) resample_fwd_kernel(const void* __restrict__ x, int x_io,
                                                                  long long x_bstride,
                                                                  long long L, const float* __restrict__ kernel,
                                                                  int n_new, int taps, int orig, int width,
                                                                  float* __restrict__ y, long long Ly, int span) {
    extern __shared__ float sm[];
    float* kw = sm;                   // [n_new * taps]
    float* xs = sm + n_new * taps;    // [span] input samples orig*j_lo - width ...
    const int b = blockIdx.y;
    const long long o0 = (long long)blockIdx.x * kRsChunk;
    const int no = (int)min((long long)kRsChunk, Ly - o0);
    const long long j_lo = o0 / n_new;
    const long long x_lo = (long long)orig * j_lo - width;
    for (int i = threadIdx.x; i < n_new * taps; i += kEwThreads) kw[i] = __ldg(kernel + i);
    const void* xb = wave_row(x, x_io, (long long)b * x_bstride);
    for (int i = threadIdx.x; i < span; i += kEwThreads) {
        long long g = x_lo + i;
        xs[i] = (g >= 0 && g < L) ? ld_wave(xb, x_io, g) : 0.f;
    }
    __syncthreads();
    const int ph0 = (int)(o0 - j_lo * n_new);
    for (int t = threadIdx.x; t < no; t += kEwThreads) {
        const int q = ph0 + t;
        const int jr = q / n_new, ph = q - jr * n_new;  // block index relative to j_lo, phase
        const float* w = kw + ph * taps;
        const float* xv = xs + jr * orig;
        float acc = 0.f;
#pragma unroll 4
        for (int k = 0; k < taps; ++k) acc = fmaf(xv[k], w[k], acc);
        y[(long long)b * Ly + o0 + t] = acc;
    }
}

__global__ void __launch_bounds__(kEwThreads) resample_adjoint_kernel(
    const float* __restrict__ ybar, int pad, long long Ly, const float* __restrict__ partial, int ntiles,
    const float* __restrict__ kernel, int n_new, int taps, int orig, int width, void* __restrict__ dwav, int dw_io,
    long long dwav_bstride, long long L, float* __restrict__ loss, int span) {
    extern __shared__ float sm[];
    __shared__ float scratch[2];
    float* kw = sm;                 // [n_new * taps]
    float* ys = sm + n_new * taps;  // [span] folded, scaled cotangent of the resampled signal
    const int b = blockIdx.y;
    const float l = clip_loss(partial + (long long)b * ntiles, ntiles, scratch);
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) loss[b] = l;
    const float sc = inv_loss(l);
    const long long i0 = (long long)blockIdx.x * kRsChunk;
    const int ni = (int)min((long long)kRsChunk, L - i0);
    // blocks j with 0 <= i + width - orig*j < taps for some i of the chunk
    const long long num = i0 + width - taps + 1;
    const long long j_lo = num <= 0 ? 0 : (num + orig - 1) / orig;
    const long long o_lo = j_lo * n_new;
    for (int i = threadIdx.x; i < n_new * taps; i += kEwThreads) kw[i] = __ldg(kernel + i);
    const float* yb = ybar + (long long)b * (Ly + 2 * pad);
    for (int i = threadIdx.x; i < span; i += kEwThreads) {
        long long o = o_lo + i;
        ys[i] = (o < Ly) ? ybar_at(yb, pad, o, Ly) * sc : 0.f;
    }
    __syncthreads();
    // 32-bit arithmetic relative to the chunk: i + width - taps + 1 = j_lo*orig + A + t
    const int A = (int)(num - j_lo * orig);
    for (int t = threadIdx.x; t < ni; t += kEwThreads) {
        const int lo = A + t;                              // first tap position, relative
        const int ja = lo <= 0 ? 0 : (lo + orig - 1) / orig;
        const int hi = lo + taps - 1;                      // i + width, relative; >= 0 because taps >= orig
        const int jb = hi / orig;
        float acc = 0.f;
        for (int j = ja; j <= jb; ++j) {
            const int k = hi - orig * j;
            const float* yv = ys + j * n_new;
            for (int ph = 0; ph < n_new; ++ph) acc = fmaf(yv[ph], kw[ph * taps + k], acc);
        }
        st_wave(dwav, dw_io, (long long)b * dwav_bstride + i0 + t, acc);
    }
}

// ---- integer decimation (n_new == 1: scale 2 and 10 of run.py / operator.py): polyphase, 4 outputs per thread ----
template <int ORIG, int TAPS>
__global__ void __launch_bounds__(kEwThreads) resample_fwd_poly_kernel(const void* __restrict__ x, int x_io,
                                                                       long long x_bstride, long long L,
                                                                       const float* __restrict__ kernel, int taps,
                                                                       int orig, int width, float* __restrict__ y,
                                                                       long long Ly, int span) {
    extern __shared__ float sm[];
    float* kw = sm;            // [taps]
    float* xs = kw + taps;                     // [fir_padded_len(span)] xz[orig*o0 ...], bank-padded
    float* outs = xs + fir_padded_len(span);   // [kRsChunk]
    const int b = blockIdx.y;
    const long long o0 = (long long)blockIdx.x * kRsChunk;
    const int no = (int)min((long long)kRsChunk, Ly - o0);
    const long long x_lo = (long long)orig * o0 - width;
    for (int i = threadIdx.x; i < taps; i += kEwThreads) kw[i] = __ldg(kernel + i);
    const void* xb = wave_row(x, x_io, (long long)b * x_bstride);
    // valid staged range in 32-bit block-relative indices: 0 <= x_lo + i < L
    const int i_lo = (int)max(0LL, -x_lo), i_hi = (int)min((long long)span, L - x_lo);
    for (int i = threadIdx.x; i < span; i += kEwThreads)
        xs[fir_pad(i)] = (i >= i_lo && i < i_hi) ? ld_wave(xb, x_io, x_lo + i) : 0.f;
    __syncthreads();
    for (int j0 = threadIdx.x * kFirR; j0 < no; j0 += kEwThreads * kFirR) {
        float acc[kFirR];
        fir_fwd4<ORIG, TAPS>(xs, kw, taps, orig, j0, acc);
#pragma unroll
        for (int c = 0; c < kFirR; ++c) outs[j0 + c] = acc[c];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < no; t += kEwThreads) y[(long long)b * Ly + o0 + t] = outs[t];
}

template <int ORIG, int TAPS>
__global__ void __launch_bounds__(kEwThreads) resample_adjoint_poly_kernel(
    const float* __restrict__ ybar, int pad, long long Ly, const float* __restrict__ partial, int ntiles,
    const float* __restrict__ kernel, int taps, int orig, int width, void* __restrict__ dwav, int dw_io,
    long long dwav_bstride, long long L, float* __restrict__ loss, int span, int chunk) {
    extern __shared__ float sm[];
    __shared__ float scratch[2];
    float* kw = sm;              // [taps]
    float* ys = kw + taps + 1;   // [span], ys[-1] = 0 (the specialised body may touch it with a zero weight)
    float* outs = ys + span;     // [chunk]
    const int b = blockIdx.y;
    const float l = clip_loss(partial + (long long)b * ntiles, ntiles, scratch);
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) loss[b] = l;
    const float sc = inv_loss(l);
    const long long i0 = (long long)blockIdx.x * chunk;
    const int ni = (int)min((long long)chunk, L - i0);
    const long long num = i0 + width - taps + 1;
    const long long j_base = ceil_div_ll(num, orig);  // may be negative: the staged window then starts with zeros
    const int A = (int)(num - j_base * orig);
    for (int i = threadIdx.x; i < taps; i += kEwThreads) kw[i] = __ldg(kernel + i);
    const float* yb = ybar + (long long)b * (Ly + 2 * pad);
    if (threadIdx.x == 0) ys[-1] = 0.f;
    for (int i = threadIdx.x; i < span; i += kEwThreads) {
        long long o = j_base + i;
        ys[i] = (o >= 0 && o < Ly) ? ybar_at(yb, pad, o, Ly) * sc : 0.f;
    }
    __syncthreads();
    const int items = chunk / kFirR;  // (phase, unit) pairs; chunk is a multiple of orig * kFirR
    for (int wi = threadIdx.x; wi < items; wi += kEwThreads) {
        const int phi = wi % orig, u = wi / orig;
        const int t0 = phi + orig * kFirR * u;
        float acc[kFirR];
        fir_adj4<ORIG, TAPS>(ys, kw, taps, orig, A, t0, acc);
#pragma unroll
        for (int c = 0; c < kFirR; ++c) outs[t0 + orig * c] = acc[c];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ni; t += kEwThreads)
        st_wave(dwav, dw_io, (long long)b * dwav_bstride + i0 + t, outs[t]);
}

// ---- scale 2 (orig 2, 28 taps, width 13): register-window kernels, no shared-memory staging ----
// Each thread produces 8 consecutive outputs from a window of aligned float4 loads (all issued before anything else, so
// a CTA has its whole input in flight at once); threads whose window crosses a row end or the reflect-fold region take
// a guarded scalar path.
__device__ __forceinline__ void load_taps28(const float* __restrict__ kernel, float (&h)[kFir2Taps]) {
#pragma unroll
    for (int q = 0; q < kFir2Taps / 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(kernel) + q);
        h[4 * q] = t.x;
        h[4 * q + 1] = t.y;
        h[4 * q + 2] = t.z;
        h[4 * q + 3] = t.w;
    }
}

// 8 consecutive 16-bit values (one 128-bit load) to fp32
template <int IO>
__device__ __forceinline__ void unpack8(const uint4 raw, float* o) {
    const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f;
        if (IO == DM_IO_F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        o[2 * i] = f.x;
        o[2 * i + 1] = f.y;
    }
}
template <int IO>
__device__ __forceinline__ uint4 pack8(const float* v) {
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (IO == DM_IO_F16) *reinterpret_cast<__half2*>(&w[i]) = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        else *reinterpret_cast<__nv_bfloat162*>(&w[i]) = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int IO>
__global__ void __launch_bounds__(kEwThreads) resample2_fwd_reg_kernel(const void* __restrict__ x, long long x_bstride,
                                                                       long long L, const float* __restrict__ kernel,
                                                                       float* __restrict__ y, long long Ly) {
    const int b = blockIdx.y;
    const long long j0 = ((long long)blockIdx.x * kEwThreads + threadIdx.x) * kFir2Out;
    if (j0 >= Ly) return;
    const void* xb = wave_row(x, IO, (long long)b * x_bstride);
    const long long x0 = 2 * j0 - 16;
    float win[kFir2FwdWin];
    if (x0 >= 0 && x0 + kFir2FwdWin <= L) {
        if (IO == DM_IO_F32) {
            const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(xb) + x0);
#pragma unroll
            for (int q = 0; q < kFir2FwdWin / 4; ++q) {
                const float4 t = src[q];
                win[4 * q] = t.x;
                win[4 * q + 1] = t.y;
                win[4 * q + 2] = t.z;
                win[4 * q + 3] = t.w;
            }
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const unsigned short*>(xb) + x0);
#pragma unroll
            for (int q = 0; q < kFir2FwdWin / 8; ++q) unpack8<IO>(src[q], win + 8 * q);
        }
    } else {
#pragma unroll
        for (int n = 0; n < kFir2FwdWin; ++n) {
            const long long g = x0 + n;
            win[n] = (g >= 0 && g < L) ? ld_wave(xb, IO, g) : 0.f;
        }
    }
    float h[kFir2Taps], out[kFir2Out];
    load_taps28(kernel, h);
    fir2_fwd8(win, h, out);
    float* yb = y + (long long)b * Ly + j0;
    if (j0 + kFir2Out <= Ly) {
        reinterpret_cast<float4*>(yb)[0] = make_float4(out[0], out[1], out[2], out[3]);
        reinterpret_cast<float4*>(yb)[1] = make_float4(out[4], out[5], out[6], out[7]);
    } else {
#pragma unroll
        for (int c = 0; c < kFir2Out; ++c)
            if (j0 + c < Ly) yb[c] = out[c];
    }
}

template <int IO>
__global__ void __launch_bounds__(kEwThreads) resample2_adjoint_reg_kernel(
    const float* __restrict__ ybar, int pad, long long Ly, const float* __restrict__ partial, int ntiles,
    const float* __restrict__ kernel, void* __restrict__ dwav, long long dwav_bstride, long long L,
    float* __restrict__ loss) {
    __shared__ float scratch[2];
    const int b = blockIdx.y;
    const long long i0 = ((long long)blockIdx.x * kEwThreads + threadIdx.x) * kFir2Out;
    const float* yb = ybar + (long long)b * (Ly + 2 * pad);
    const long long m0 = i0 / 2 - 8;  // first cotangent sample of the window
    float win[kFir2AdjWin];
    if (i0 < L) {
        // no fold terms for 513 <= j <= Ly - 514 (pad = 512); without padding any in-range window is plain
        const bool plain = pad ? (m0 >= 513 && m0 + kFir2AdjWin <= Ly - 513) : (m0 >= 0 && m0 + kFir2AdjWin <= Ly);
        if (plain) {
            const float4* src = reinterpret_cast<const float4*>(yb + pad + m0);
#pragma unroll
            for (int q = 0; q < kFir2AdjWin / 4; ++q) {
                const float4 t = src[q];
                win[4 * q] = t.x;
                win[4 * q + 1] = t.y;
                win[4 * q + 2] = t.z;
                win[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int n = 0; n < kFir2AdjWin; ++n) {
                const long long j = m0 + n;
                win[n] = (j >= 0 && j < Ly) ? ybar_at(yb, pad, j, Ly) : 0.f;
            }
        }
    }
    // the per-clip scale is only needed at the very end: the window loads above are already in flight
    const float l = clip_loss(partial + (long long)b * ntiles, ntiles, scratch);
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss) loss[b] = l;
    if (i0 >= L) return;
    const float sc = inv_loss(l);
    float h[kFir2Taps], out[kFir2Out];
    load_taps28(kernel, h);
    fir2_adj8(win, h, out);
    void* ob = wave_row(dwav, IO, (long long)b * dwav_bstride + i0);
#pragma unroll
    for (int c = 0; c < kFir2Out; ++c) out[c] *= sc;
    if (i0 + kFir2Out <= L) {
        if (IO == DM_IO_F32) {
            reinterpret_cast<float4*>(ob)[0] = make_float4(out[0], out[1], out[2], out[3]);
            reinterpret_cast<float4*>(ob)[1] = make_float4(out[4], out[5], out[6], out[7]);
        } else {
            *reinterpret_cast<uint4*>(ob) = pack8<IO>(out);
        }
    } else {
#pragma unroll
        for (int c = 0; c < kFir2Out; ++c)
            if (i0 + c < L) st_wave(ob, IO, c, out[c]);
    }
}

// ---- persistent ("stream") variants for the shipped shapes.  The kernels above spend one CTA on 2048 samples: at 16
// clips that is a thousand CTAs living one or two dependent memory latencies each (the adjoints first reduce the clip's
// loss), and the forward window loads (lane stride 64 B) cost 16 L1 wavefronts per 128-bit load.  Here a fixed grid
// walks the work: every CTA owns a contiguous range of windows, reduces the scales of the (one or two) clips it touches
// once, keeps two windows in flight per thread and stores a full 32-byte sector per lane (st.global.v8); the forward
// stages its input span with cp.async into a swizzled shared-memory buffer (double buffered) from which the same
// window reads are bank-conflict free.  Arithmetic is unchanged (bit-identical).  [Measured and dropped: the same
// treatment of fold_adjoint_kernel (slower: it is already a plain two-vector stream).]
constexpr int kStreamMaxClips = 1024;

// s_scale[b - b_lo] = 1 / loss_b (0 at 0) for the clips b_lo..b_hi a CTA works on, one warp per clip, loss_b exactly as
// clip_loss() computes it; the CTA that owns a clip's first item publishes its loss
__device__ __forceinline__ void clip_scales_smem(const float* __restrict__ partial, int ntiles, int b_lo, int b_hi,
                                                 int b_pub_lo, float* __restrict__ loss, float* __restrict__ s_scale) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = b_lo + warp; b <= b_hi; b += kEwThreads / 32) {
        const float* pb = partial + (long long)b * ntiles;
        double acc = 0.0;
        for (int i = lane; i < ntiles; i += 32) acc += (double)pb[i];
        acc = warp_sum(acc);
        if (lane == 0) {
            const float l = sqrtf((float)acc);
            s_scale[b - b_lo] = inv_loss(l);
            if (b >= b_pub_lo && loss) loss[b] = l;
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void st_global_256(float* p, const float (&o)[8]) {  // one full 32-byte sector per lane
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                 "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                 : "memory");
}

struct Rs2AdjWin {
    float win[kFir2AdjWin];
    int b, i0;
};
template <int IO>
__device__ __forceinline__ void rs2_adj_load(Rs2AdjWin& w, int item, int nwin_clip, const float* __restrict__ ybar,
                                             int pad, long long Ly) {
    w.b = item / nwin_clip;
    w.i0 = (item - w.b * nwin_clip) * kFir2Out;
    const float* yb = ybar + (long long)w.b * (Ly + 2 * pad);
    const long long m0 = w.i0 / 2 - 8;
    const bool plain = pad ? (m0 >= 513 && m0 + kFir2AdjWin <= Ly - 513) : (m0 >= 0 && m0 + kFir2AdjWin <= Ly);
    if (plain) {
        const float4* src = reinterpret_cast<const float4*>(yb + pad + m0);
#pragma unroll
        for (int q = 0; q < kFir2AdjWin / 4; ++q) {
            const float4 t = src[q];
            w.win[4 * q] = t.x, w.win[4 * q + 1] = t.y, w.win[4 * q + 2] = t.z, w.win[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int n = 0; n < kFir2AdjWin; ++n) {
            const long long j = m0 + n;
            w.win[n] = (j >= 0 && j < Ly) ? ybar_at(yb, pad, j, Ly) : 0.f;
        }
    }
}
template <int IO>
__device__ __forceinline__ void rs2_adj_finish(const Rs2AdjWin& w, const float (&h)[kFir2Taps],
                                               const float* __restrict__ s_scale, int b_lo, void* __restrict__ dwav,
                                               long long dwav_bstride, long long L) {
    float out[kFir2Out];
    fir2_adj8(w.win, h, out);
    const float sc = s_scale[w.b - b_lo];
    void* ob = wave_row(dwav, IO, (long long)w.b * dwav_bstride + w.i0);
#pragma unroll
    for (int c = 0; c < kFir2Out; ++c) out[c] *= sc;
    if (w.i0 + kFir2Out <= L) {
        if (IO == DM_IO_F32) {
            if ((reinterpret_cast<uintptr_t>(ob) & 31) == 0) {
                st_global_256(static_cast<float*>(ob), out);
            } else {
                reinterpret_cast<float4*>(ob)[0] = make_float4(out[0], out[1], out[2], out[3]);
                reinterpret_cast<float4*>(ob)[1] = make_float4(out[4], out[5], out[6], out[7]);
            }
        } else {
            *reinterpret_cast<uint4*>(ob) = pack8<IO>(out);
        }
    } else {
#pragma unroll
        for (int c = 0; c < kFir2Out; ++c)
            if (w.i0 + c < L) st_wave(ob, IO, c, out[c]);
    }
}
template <int IO>
__global__ void __launch_bounds__(kEwThreads) resample2_adjoint_stream_kernel(
    const float* __restrict__ ybar, int pad, long long Ly, int B, const float* __restrict__ partial, int ntiles,
    const float* __restrict__ kernel, void* __restrict__ dwav, long long dwav_bstride, long long L,
    float* __restrict__ loss) {
    __shared__ float s_scale[kStreamMaxClips];
    const int nwin_clip = (int)((L + kFir2Out - 1) / kFir2Out), total = nwin_clip * B;  // host: total < 2^30
    const int per_cta = ((total + gridDim.x - 1) / gridDim.x + kEwThreads - 1) / kEwThreads * kEwThreads;  // as the host
    const int start = blockIdx.x * per_cta, end = min(total, start + per_cta);
    if (start >= end) return;
    const int b_lo = start / nwin_clip, b_hi = (end - 1) / nwin_clip;
    int item = start + threadIdx.x;
    Rs2AdjWin w0, w1;
    float h[kFir2Taps];
    load_taps28(kernel, h);
    pdl_wait();  // cotangent and partial sums of the STFT kernel from here on
    if (item < end) rs2_adj_load<IO>(w0, item, nwin_clip, ybar, pad, Ly);  // in flight behind the scale reduction
    clip_scales_smem(partial, ntiles, b_lo, b_hi, (start + nwin_clip - 1) / nwin_clip, loss, s_scale);
    for (; item < end; item += 2 * kEwThreads) {
        const int item1 = item + kEwThreads;
        if (item1 < end) rs2_adj_load<IO>(w1, item1, nwin_clip, ybar, pad, Ly);
        rs2_adj_finish<IO>(w0, h, s_scale, b_lo, dwav, dwav_bstride, L);
        const int item2 = item1 + kEwThreads;
        if (item2 < end) rs2_adj_load<IO>(w0, item2, nwin_clip, ybar, pad, Ly);
        if (item1 < end) rs2_adj_finish<IO>(w1, h, s_scale, b_lo, dwav, dwav_bstride, L);
    }
}

// forward: cp.async staging of the chunk's input span, swizzled so that the window reads (lane stride 64 B) are
// conflict-free: 16-byte cell c sits at (c & ~7) | ((c & 7) ^ ((c >> 3) & 3)).  [Per quarter-warp the cells 4 i + q of 8
// consecutive lanes fall into two columns a, a ^ 4 of four consecutive 128-byte rows each; xor-ing the column with
// row & 3 spreads each set over a ^ {0..3} resp. a ^ 4 ^ {0..3}: eight different bank groups.]
// (chunk geometry and the cell maps rs2_cell / rs2_win_cell: fir_poly.cuh, shared with the host emulation)
__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__global__ void __launch_bounds__(kRs2Threads) resample2_fwd_stream_kernel(const float* __restrict__ x,
                                                                           long long x_bstride, long long L, int B,
                                                                           const float* __restrict__ kernel,
                                                                           float* __restrict__ y, long long Ly,
                                                                           float4* __restrict__ fill,
                                                                           long long fill_n4) {
    __shared__ __align__(16) float buf[2][kRs2BufFloats];
    const int t = threadIdx.x;
    pdl_trigger();  // the STFT kernel of the chain may stage its tables while this one runs
    // optional: zero a side buffer on the way (the cotangent buffer the STFT kernel accumulates into next: saves the
    // fill launch of the chain).  Stores only, issued before anything else.
    for (long long i = (long long)blockIdx.x * kRs2Threads + t; i < fill_n4; i += (long long)gridDim.x * kRs2Threads)
        fill[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int chunks_clip = (int)((Ly + kRs2ChunkOut - 1) / kRs2ChunkOut);
    const int total = chunks_clip * B;  // the host keeps it below 2^31
    // chunk-independent addressing, computed once: where this thread's staged cells go, where its window cells are
    constexpr int kStageIt = (kRs2Cells + kRs2Threads - 1) / kRs2Threads;  // 5, the last one for 8 threads only
    int stage_off[kStageIt];
#pragma unroll
    for (int k = 0; k < kStageIt; ++k) stage_off[k] = 4 * rs2_cell(t + k * kRs2Threads);

    auto stage = [&](int item, float* dst) {
        const int b = item / chunks_clip;
        const long long j0c = (long long)(item - b * chunks_clip) * kRs2ChunkOut;
        const float* xb = x + (long long)b * x_bstride;
        const long long x0 = 2 * j0c - 16;  // a multiple of 4 floats
        if (x0 >= 0 && x0 + 4 * kRs2Cells <= L) {  // interior chunk (uniform): plain 16-byte copies
            const float* src = xb + x0 + 4 * t;
#pragma unroll
            for (int k = 0; k < kStageIt; ++k)
                if (k + 1 < kStageIt || t + k * kRs2Threads < kRs2Cells) cp_async16(dst + stage_off[k], src + 4 * k * kRs2Threads);
        } else {
#pragma unroll
            for (int k = 0; k < kStageIt; ++k) {
                const int c = t + k * kRs2Threads;
                if (c >= kRs2Cells) break;
                const long long g = x0 + 4 * c;
                int nbytes = 0;
                if (g >= 0 && g < L) nbytes = (int)min((long long)16, (L - g) * 4);
                cp_async16_zfill(dst + stage_off[k], nbytes ? xb + g : xb, nbytes);  // zero-fills outside [0, L)
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float h[kFir2Taps];
    load_taps28(kernel, h);
    int item = blockIdx.x;
    if (item < total) stage(item, buf[0]);
    int cur = 0;
    for (; item < total; item += gridDim.x, cur ^= 1) {
        const int next = item + gridDim.x;
        if (next < total) {
            stage(next, buf[cur ^ 1]);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int b = item / chunks_clip;
        const long long j0 = (long long)(item - b * chunks_clip) * kRs2ChunkOut + t * kFir2Out;
        if (j0 < Ly) {
            float win[kFir2FwdWin], out[kFir2Out];
            const float4* cells = reinterpret_cast<const float4*>(buf[cur]);
#pragma unroll
            for (int q = 0; q < kFir2FwdWin / 4; ++q) {
                const float4 v = cells[rs2_win_cell(t, q)];
                win[4 * q] = v.x, win[4 * q + 1] = v.y, win[4 * q + 2] = v.z, win[4 * q + 3] = v.w;
            }
            fir2_fwd8(win, h, out);
            float* yb = y + (long long)b * Ly + j0;
            if (j0 + kFir2Out <= Ly) {
                if ((reinterpret_cast<uintptr_t>(yb) & 31) == 0) {
                    st_global_256(yb, out);
                } else {
                    reinterpret_cast<float4*>(yb)[0] = make_float4(out[0], out[1], out[2], out[3]);
                    reinterpret_cast<float4*>(yb)[1] = make_float4(out[4], out[5], out[6], out[7]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < kFir2Out; ++c)
                    if (j0 + c < Ly) yb[c] = out[c];
            }
        }
        __syncthreads();  // the buffer is free for the chunk after next
    }
}

// ---- pure streaming kernels: 128-bit accesses (two float4 in flight per thread and iteration) with a scalar tail;
// the scalar path also takes rows / pointers that are not 16-byte aligned
__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(kEwThreads) mask_apply_kernel(const float* __restrict__ x, long long x_bstride,
                                                                long long L, const float* __restrict__ mask,
                                                                float* __restrict__ y) {
    const int b = blockIdx.y;
    const float* xb = x + (long long)b * x_bstride;
    float* yb = y + (long long)b * L;
    const long long stride = (long long)gridDim.x * kEwThreads;
    const long long t0 = (long long)blockIdx.x * kEwThreads + threadIdx.x;
    long long done = 0;
    if (aligned16_dev(xb) && aligned16_dev(yb) && aligned16_dev(mask)) {
        const long long nv = L / 4;
        const float4* x4 = reinterpret_cast<const float4*>(xb);
        const float4* m4 = reinterpret_cast<const float4*>(mask);
        float4* y4 = reinterpret_cast<float4*>(yb);
        for (long long i = t0; i < nv; i += 2 * stride) {
            const long long i2 = i + stride;
            const float4 a = x4[i], m = __ldg(m4 + i);
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f), n = c;
            if (i2 < nv) c = x4[i2], n = __ldg(m4 + i2);
            y4[i] = make_float4(a.x * m.x, a.y * m.y, a.z * m.z, a.w * m.w);
            if (i2 < nv) y4[i2] = make_float4(c.x * n.x, c.y * n.y, c.z * n.z, c.w * n.w);
        }
        done = nv * 4;
    }
    for (long long i = done + t0; i < L; i += stride) yb[i] = xb[i] * __ldg(mask + i);
}

__global__ void __launch_bounds__(kEwThreads) add_scaled_kernel(float* __restrict__ y,
                                                                const float* __restrict__ noise, float sigma,
                                                                long long n) {
    const long long stride = (long long)gridDim.x * kEwThreads;
    const long long t0 = (long long)blockIdx.x * kEwThreads + threadIdx.x;
    long long done = 0;
    if (aligned16_dev(y) && aligned16_dev(noise)) {
        const long long nv = n / 4;
        float4* y4 = reinterpret_cast<float4*>(y);
        const float4* z4 = reinterpret_cast<const float4*>(noise);
        for (long long i = t0; i < nv; i += 2 * stride) {
            const long long i2 = i + stride;
            float4 a = y4[i];
            const float4 z = z4[i];
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f), w = c;
            if (i2 < nv) c = y4[i2], w = z4[i2];
            // y + noise * sigma with the product rounded first, like the reference's two torch ops (noise.py:17-18)
            a.x = __fadd_rn(a.x, __fmul_rn(z.x, sigma)), a.y = __fadd_rn(a.y, __fmul_rn(z.y, sigma));
            a.z = __fadd_rn(a.z, __fmul_rn(z.z, sigma)), a.w = __fadd_rn(a.w, __fmul_rn(z.w, sigma));
            y4[i] = a;
            if (i2 < nv) {
                c.x = __fadd_rn(c.x, __fmul_rn(w.x, sigma)), c.y = __fadd_rn(c.y, __fmul_rn(w.y, sigma));
                c.z = __fadd_rn(c.z, __fmul_rn(w.z, sigma)), c.w = __fadd_rn(c.w, __fmul_rn(w.w, sigma));
                y4[i2] = c;
            }
        }
        done = nv * 4;
    }
    for (long long i = done + t0; i < n; i += stride) y[i] = __fadd_rn(y[i], __fmul_rn(noise[i], sigma));
}

}  // namespace dm

using namespace dm;

extern "C" int dm_residual_wav_io(const void* y, int y_dtype, long long y_bstride, long long n, int B,
                                  const float* mask, const float* meas, long long meas_bstride, float* ybar,
                                  float* partial, dm_stream_t stream) {
    DM_REQUIRE(y && meas && ybar && partial && n > 0 && B > 0 && io_dtype_ok(y_dtype));
    const int ntiles = (int)((n + DM_RESID_CHUNK - 1) / DM_RESID_CHUNK);
    residual_wav_kernel<<<dim3(ntiles, B), kEwThreads, 0, as_stream(stream)>>>(y, y_dtype, y_bstride, n, mask, meas,
                                                                               meas_bstride, ybar, partial, ntiles);
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_residual_wav(const float* y, long long y_bstride, long long n, int B, const float* mask,
                               const float* meas, long long meas_bstride, float* ybar, float* partial,
                               dm_stream_t stream) {
    return dm_residual_wav_io(y, DM_IO_F32, y_bstride, n, B, mask, meas, meas_bstride, ybar, partial, stream);
}

extern "C" int dm_fold_adjoint(const float* ybar, int pad, long long Ly, int B, const float* mask,
                               const float* partial, int ntiles, float* dwav, long long dwav_bstride, float* loss,
                               dm_stream_t stream) {
    return dm_fold_adjoint_io(ybar, pad, Ly, B, mask, partial, ntiles, dwav, DM_IO_F32, dwav_bstride, loss, stream);
}
extern "C" int dm_fold_adjoint_io(const float* ybar, int pad, long long Ly, int B, const float* mask,
                                  const float* partial, int ntiles, void* dwav, int dwav_dtype, long long dwav_bstride,
                                  float* loss, dm_stream_t stream) {
    DM_REQUIRE(partial && Ly > 0 && B > 0 && ntiles > 0 && io_dtype_ok(dwav_dtype));
    DM_REQUIRE((dwav == nullptr && loss != nullptr) || ybar != nullptr);
    DM_REQUIRE(pad == 0 || (pad == 512 && Ly > 512));
    const int nblk = dwav ? (int)((Ly + kEwThreads * 8 - 1) / (kEwThreads * 8)) : 1;
    launch_pdl(fold_adjoint_kernel, dim3(nblk, B), dim3(kEwThreads), 0, as_stream(stream), g_tuning[DM_TUNE_PDL] != 0,
               ybar, pad, Ly, mask, partial, ntiles, dwav, dwav_dtype, dwav_bstride, loss);
    DM_LAUNCHED();
    return DM_OK;
}

template <int IO>
static void launch_rs2_fwd(const void* x, long long x_bstride, long long L, int B, const float* kernel, float* y,
                           long long Ly, cudaStream_t st, float* fill = nullptr, long long fill_n = 0) {
    const long long chunks = ((Ly + kRs2ChunkOut - 1) / kRs2ChunkOut) * B;
    if (IO == DM_IO_F32 && g_tuning[DM_TUNE_STREAM_KERNELS] && chunks < 0x7fffffffLL) {
        const int grid = (int)std::min<long long>(chunks, (long long)num_sms() * 12);
        const bool fuse_fill = fill != nullptr && (fill_n & 3) == 0 && (reinterpret_cast<uintptr_t>(fill) & 15) == 0;
        if (fill != nullptr && !fuse_fill) cudaMemsetAsync(fill, 0, (size_t)fill_n * sizeof(float), st);
        resample2_fwd_stream_kernel<<<grid, kRs2Threads, 0, st>>>(static_cast<const float*>(x), x_bstride, L, B, kernel,
                                                                  y, Ly, reinterpret_cast<float4*>(fill),
                                                                  fuse_fill ? fill_n / 4 : 0);
        return;
    }
    if (fill != nullptr) cudaMemsetAsync(fill, 0, (size_t)fill_n * sizeof(float), st);
    const long long nthr = (Ly + kFir2Out - 1) / kFir2Out;
    resample2_fwd_reg_kernel<IO><<<dim3((unsigned)((nthr + kEwThreads - 1) / kEwThreads), B), kEwThreads, 0, st>>>(
        x, x_bstride, L, kernel, y, Ly);
}
template <int IO>
static void launch_rs2_adj(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                           const float* kernel, void* dwav, long long dwav_bstride, long long L, float* loss,
                           cudaStream_t st) {
    const long long nwin = ((L + kFir2Out - 1) / kFir2Out) * B;
    // The persistent kernel wins once the work exceeds about two resident waves of the plain one (measured under ncu,
    // 10 s clips: 128 clips 33 us against 46 us; 16 clips 11.9 us against 10.5 us) -- below that the plain kernel stays.
    if (g_tuning[DM_TUNE_STREAM_KERNELS] && B <= kStreamMaxClips && nwin < (1LL << 30) &&
        (nwin > (long long)num_sms() * 4096 || g_tuning[DM_TUNE_STREAM_KERNELS] == 2)) {
        // one resident wave (4 CTAs of 64 registers per SM), every CTA a whole number of 256-window rounds
        long long per_cta = (nwin + (long long)num_sms() * 4 - 1) / ((long long)num_sms() * 4);
        per_cta = std::max<long long>(2, (per_cta + kEwThreads - 1) / kEwThreads) * kEwThreads;
        const int grid = (int)((nwin + per_cta - 1) / per_cta);
        launch_pdl(resample2_adjoint_stream_kernel<IO>, dim3(grid), dim3(kEwThreads), 0, st, g_tuning[DM_TUNE_PDL] != 0,
                   ybar, pad, Ly, B, partial, ntiles, kernel, dwav, dwav_bstride, L, loss);
        return;
    }
    const long long nthr = (L + kFir2Out - 1) / kFir2Out;
    resample2_adjoint_reg_kernel<IO><<<dim3((unsigned)((nthr + kEwThreads - 1) / kEwThreads), B), kEwThreads, 0, st>>>(
        ybar, pad, Ly, partial, ntiles, kernel, dwav, dwav_bstride, L, loss);
}
static bool rs2_filter(int n_new, int orig, int taps, int width, const float* kernel) {
    return n_new == 1 && orig == 2 && taps == kFir2Taps && width == kFir2Width &&
           (reinterpret_cast<uintptr_t>(kernel) & 15) == 0;
}

// ---- resampling entry points: one typed core per direction (the waveform side is fp32 / fp16 / bf16) ----
template <int IO>
static bool rs2_fwd_ok(const void* x, long long x_bstride, const float* y, long long Ly) {
    const long long q = IO == DM_IO_F32 ? 4 : 8;  // row stride in elements that keeps rows 16-byte aligned
    return x_bstride % q == 0 && Ly % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(y) & 15) == 0;
}
static int resample_fwd_core(const void* x, int x_io, long long x_bstride, long long L, int B, const float* kernel,
                             int n_new, int taps, int orig, int width, float* y, long long Ly, cudaStream_t st,
                             float* fill = nullptr, long long fill_n = 0) {
    DM_REQUIRE(x && kernel && y && L > 0 && B > 0 && Ly > 0 && io_dtype_ok(x_io));
    DM_REQUIRE(n_new >= 1 && taps >= 1 && orig >= 1 && width >= 0);
    DM_REQUIRE(fill == nullptr || fill_n > 0);
    if (rs2_filter(n_new, orig, taps, width, kernel)) {  // scale 2: register-window kernel
        if (x_io == DM_IO_F32 && rs2_fwd_ok<DM_IO_F32>(x, x_bstride, y, Ly)) {
            launch_rs2_fwd<DM_IO_F32>(x, x_bstride, L, B, kernel, y, Ly, st, fill, fill_n);
            DM_LAUNCHED();
            return DM_OK;
        }
    }
    if (fill != nullptr) DM_CUDA(cudaMemsetAsync(fill, 0, (size_t)fill_n * sizeof(float), st));  // every other path
    if (rs2_filter(n_new, orig, taps, width, kernel)) {
        if (x_io != DM_IO_F32 && rs2_fwd_ok<DM_IO_F16>(x, x_bstride, y, Ly)) {
            if (x_io == DM_IO_F16) launch_rs2_fwd<DM_IO_F16>(x, x_bstride, L, B, kernel, y, Ly, st);
            else launch_rs2_fwd<DM_IO_BF16>(x, x_bstride, L, B, kernel, y, Ly, st);
            DM_LAUNCHED();
            return DM_OK;
        }
    }
    if (n_new == 1) {  // integer decimation: polyphase kernel out of shared memory
        const int span1 = orig * (kRsChunk + kFirR) + taps;
        const size_t smem1 = ((size_t)taps + fir_padded_len(span1) + kRsChunk) * sizeof(float);
        if (smem1 <= 200 * 1024) {
            const dim3 grid((unsigned)((Ly + kRsChunk - 1) / kRsChunk), B);
#define DM_RS_FWD(O, T)                                                                                     \
    do {                                                                                                    \
        DM_SMEM_ONCE((resample_fwd_poly_kernel<O, T>), smem1);                                              \
        resample_fwd_poly_kernel<O, T><<<grid, kEwThreads, smem1, st>>>(x, x_io, x_bstride, L, kernel, taps, orig, \
                                                                        width, y, Ly, span1);               \
    } while (0)
            if (orig == 2 && taps == 28) DM_RS_FWD(2, 28);          // scale 2 (run.py:188), unaligned rows
            else if (orig == 10 && taps == 132) DM_RS_FWD(10, 132);  // scale 10 (operator.py:179 default)
            else DM_RS_FWD(0, 0);
#undef DM_RS_FWD
            DM_LAUNCHED();
            return DM_OK;
        }
    }
    // input span of one chunk: blocks j_lo .. j_lo + ceil((chunk + n_new - 1)/n_new), each orig apart, plus the taps
    const int span = ((kRsChunk + n_new - 1) / n_new + 1) * orig + taps;
    const size_t smem = ((size_t)n_new * taps + span) * sizeof(float);
    if (smem > 200 * 1024) return fail(DM_ERR_UNSUPPORTED, "%s: resampling ratio %d/%d needs %zu B of shared memory",
                                       __func__, n_new, orig, smem);
    DM_SMEM_ONCE(resample_fwd_kernel, smem);
    const int nblk = (int)((Ly + kRsChunk - 1) / kRsChunk);
    resample_fwd_kernel<<<dim3(nblk, B), kEwThreads, smem, st>>>(x, x_io, x_bstride, L, kernel, n_new, taps, orig,
                                                                 width, y, Ly, span);
    DM_LAUNCHED();
    return DM_OK;
}

static int resample_adjoint_core(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                                 const float* kernel, int n_new, int taps, int orig, int width, void* dwav, int dw_io,
                                 long long dwav_bstride, long long L, float* loss, cudaStream_t st) {
    DM_REQUIRE(ybar && partial && kernel && dwav && L > 0 && B > 0 && Ly > 0 && ntiles > 0 && io_dtype_ok(dw_io));
    DM_REQUIRE(pad == 0 || (pad == 512 && Ly > 512));
    if (rs2_filter(n_new, orig, taps, width, kernel) && dwav_bstride % (dw_io == DM_IO_F32 ? 4 : 8) == 0 &&
        (Ly + 2 * pad) % 4 == 0 && (reinterpret_cast<uintptr_t>(ybar) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dwav) & 15) == 0) {
        if (dw_io == DM_IO_F32)
            launch_rs2_adj<DM_IO_F32>(ybar, pad, Ly, B, partial, ntiles, kernel, dwav, dwav_bstride, L, loss, st);
        else if (dw_io == DM_IO_F16)
            launch_rs2_adj<DM_IO_F16>(ybar, pad, Ly, B, partial, ntiles, kernel, dwav, dwav_bstride, L, loss, st);
        else
            launch_rs2_adj<DM_IO_BF16>(ybar, pad, Ly, B, partial, ntiles, kernel, dwav, dwav_bstride, L, loss, st);
        DM_LAUNCHED();
        return DM_OK;
    }
    if (n_new == 1 && taps >= orig) {  // integer decimation: polyphase kernel
        const int unit = orig * kFirR;
        const int chunk = unit * ((kRsChunk + unit - 1) / unit);
        const int span1 = (taps + chunk) / orig + kFirR + 2;
        const size_t smem1 = ((size_t)taps + 1 + span1 + chunk) * sizeof(float);
        const dim3 grid((unsigned)((L + chunk - 1) / chunk), B);
#define DM_RS_ADJ(O, T)                                                                                         \
    do {                                                                                                        \
        DM_SMEM_ONCE((resample_adjoint_poly_kernel<O, T>), smem1);                                              \
        resample_adjoint_poly_kernel<O, T><<<grid, kEwThreads, smem1, st>>>(ybar, pad, Ly, partial, ntiles, kernel, \
                                                                            taps, orig, width, dwav, dw_io,     \
                                                                            dwav_bstride, L, loss, span1, chunk); \
    } while (0)
        if (orig == 2 && taps == 28) DM_RS_ADJ(2, 28);
        else if (orig == 10 && taps == 132) DM_RS_ADJ(10, 132);
        else DM_RS_ADJ(0, 0);
#undef DM_RS_ADJ
        DM_LAUNCHED();
        return DM_OK;
    }
    // cotangent span of one chunk of inputs: (chunk + taps)/orig + 2 blocks of n_new samples
    const int span = ((kRsChunk + taps) / orig + 2) * n_new;
    const size_t smem = ((size_t)n_new * taps + span) * sizeof(float);
    if (smem > 200 * 1024) return fail(DM_ERR_UNSUPPORTED, "%s: resampling ratio %d/%d needs %zu B of shared memory",
                                       __func__, n_new, orig, smem);
    DM_SMEM_ONCE(resample_adjoint_kernel, smem);
    const int nblk = (int)((L + kRsChunk - 1) / kRsChunk);
    resample_adjoint_kernel<<<dim3(nblk, B), kEwThreads, smem, st>>>(ybar, pad, Ly, partial, ntiles, kernel, n_new, taps,
                                                                     orig, width, dwav, dw_io, dwav_bstride, L, loss,
                                                                     span);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_resample_fwd_io(const void* x, int x_dtype, long long x_bstride, long long L, int B,
                                  const float* kernel, int n_new, int taps, int orig, int width, float* y, long long Ly,
                                  dm_stream_t stream) {
    return resample_fwd_core(x, x_dtype, x_bstride, L, B, kernel, n_new, taps, orig, width, y, Ly, as_stream(stream));
}
extern "C" int dm_resample_fwd_fill_io(const void* x, int x_dtype, long long x_bstride, long long L, int B,
                                       const float* kernel, int n_new, int taps, int orig, int width, float* y,
                                       long long Ly, float* fill, long long fill_count, dm_stream_t stream) {
    return resample_fwd_core(x, x_dtype, x_bstride, L, B, kernel, n_new, taps, orig, width, y, Ly, as_stream(stream),
                             fill, fill_count);
}
extern "C" int dm_resample_fwd(const float* x, long long x_bstride, long long L, int B, const float* kernel,
                               int n_new, int taps, int orig, int width, float* y, long long Ly,
                               dm_stream_t stream) {
    return resample_fwd_core(x, DM_IO_F32, x_bstride, L, B, kernel, n_new, taps, orig, width, y, Ly,
                             as_stream(stream));
}
extern "C" int dm_resample_adjoint_io(const float* ybar, int pad, long long Ly, int B, const float* partial,
                                      int ntiles, const float* kernel, int n_new, int taps, int orig, int width,
                                      void* dwav, int dwav_dtype, long long dwav_bstride, long long L, float* loss,
                                      dm_stream_t stream) {
    return resample_adjoint_core(ybar, pad, Ly, B, partial, ntiles, kernel, n_new, taps, orig, width, dwav, dwav_dtype,
                                 dwav_bstride, L, loss, as_stream(stream));
}
extern "C" int dm_resample_adjoint(const float* ybar, int pad, long long Ly, int B, const float* partial, int ntiles,
                                   const float* kernel, int n_new, int taps, int orig, int width, float* dwav,
                                   long long dwav_bstride, long long L, float* loss, dm_stream_t stream) {
    return resample_adjoint_core(ybar, pad, Ly, B, partial, ntiles, kernel, n_new, taps, orig, width, dwav, DM_IO_F32,
                                 dwav_bstride, L, loss, as_stream(stream));
}

extern "C" int dm_mask_apply(const float* x, long long x_bstride, long long L, int B, const float* mask, float* y,
                             dm_stream_t stream) {
    DM_REQUIRE(x && mask && y && L > 0 && B > 0);
    const int nblk = (int)min((long long)num_sms() * 4, (L / 8 + kEwThreads) / kEwThreads);  // two float4 per thread
    mask_apply_kernel<<<dim3(nblk, B), kEwThreads, 0, as_stream(stream)>>>(x, x_bstride, L, mask, y);
    DM_LAUNCHED();
    return DM_OK;
}

__global__ void __launch_bounds__(kEwThreads) copy_f32_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                              long long n) {
    const long long stride = (long long)gridDim.x * kEwThreads;
    const long long t0 = (long long)blockIdx.x * kEwThreads + threadIdx.x;
    long long done = 0;
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const long long nv = n / 4;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (long long i = t0; i < nv; i += stride) d4[i] = s4[i];
        done = nv * 4;
    }
    for (long long i = done + t0; i < n; i += stride) dst[i] = src[i];
}

extern "C" int dm_copy_f32(float* dst, const float* src, long long n, dm_stream_t stream) {
    DM_REQUIRE(dst && src && n > 0);
    const int nblk = (int)min((long long)num_sms() * 2, (n / 4 + kEwThreads) / kEwThreads);
    copy_f32_kernel<<<nblk, kEwThreads, 0, as_stream(stream)>>>(dst, src, n);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_add_scaled(float* y, const float* noise, float sigma, long long n, dm_stream_t stream) {
    DM_REQUIRE(y && noise && n > 0);
    const int nblk = (int)min((long long)num_sms() * 8, (n / 8 + kEwThreads) / kEwThreads);
    add_scaled_kernel<<<nblk, kEwThreads, 0, as_stream(stream)>>>(y, noise, sigma, n);
    DM_LAUNCHED();
    return DM_OK;
}
