// The one exchange step of the path: summing the per-rank FAD moments  acc = [n | sum x | sum x x^T]
// (fadtk/utils.py:36-40, the Chan merge, as a sum of raw moments) over the GPUs of one node.
//
// All-reduce over NVLink peer memory, written here instead of calling a collective library:
//   * every rank keeps its accumulator (and the buffer the sum lands in) in memory its peers have mapped (CUDA IPC; the
//     host side only exchanges the handles once, diffmusic_b200/parallel.py PeerGroup);
//   * dm_fad_allreduce_push (the default, "two-shot"): ONE kernel per rank.  Block 0 raises this rank's READY flag in
//     every peer's flag pad (system-scope release) and every block waits until all peers are READY (acquire).  Rank r then
//     REDUCES the rows i = r, r + W, r + 2W, ... of the upper triangle of sum x x^T -- reading that row from every peer in
//     rank order, so all ranks end up with bit-identical sums -- and PUSHES the reduced row into every peer's sum buffer.
//     Per GPU 2 (W - 1) / W of the 1 + d + d (d + 1) / 2 doubles cross the links (4.1 MB at d = 768, W = 8) instead of
//     W - 1 times the triangle (16.6 MB) for the one-shot variant below.  The last block raises DONE in every peer's pad
//     (after a system fence: the pushed rows are visible); dm_fad_finalize_shared waits for every rank's DONE before it
//     reads the sum;
//   * dm_fad_allreduce_peers (one-shot, kept for A/B and tiny worlds): every rank reads the whole triangle of every peer;
//   * dm_fad_reset_shared (before the next accumulation) waits for every peer's DONE of the last exchange round before it
//     clears the accumulator, so no rank overwrites moments a slower peer is still reading.
// Flags are monotonically increasing round numbers: no flag is ever reset, nothing synchronises with the host.
//
// dm_fad_allreduce is the same exchange through ncclAllReduce for callers that own a communicator (resolved from the
// process at run time: the library does not link NCCL).
#include <dlfcn.h>

#include "dm_common.cuh"

namespace dm {

constexpr int kXchgThreads = 256;
constexpr int kXchgMaxWorld = 16;
// flag pad of one rank (unsigned): [0, 16) READY round of peer r, [16, 32) DONE round of peer r, [32] block counter
constexpr int kFlagReady = 0, kFlagDone = kXchgMaxWorld, kFlagCounter = 2 * kXchgMaxWorld;

struct PeerPtrs {
    const double* acc[kXchgMaxWorld];
    unsigned* flags[kXchgMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer data must come from the owner's memory every round (never from a line this SM cached in an earlier round)
__device__ __forceinline__ double ld_peer(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flags(const unsigned* pad, int world, unsigned round) {
    if ((int)threadIdx.x < world) {
        while (ld_acquire_sys(pad + threadIdx.x) < round) __nanosleep(64);
    }
    __syncthreads();
}

// clear the shared accumulator once every peer has finished reading exchange round `done_round` (0 = none yet)
__global__ void __launch_bounds__(kXchgThreads) fad_reset_shared_kernel(double* __restrict__ acc, long long n,
                                                                        const unsigned* __restrict__ my_flags,
                                                                        int world, unsigned done_round) {
    wait_flags(my_flags + kFlagDone, world, done_round);
    const long long stride = (long long)gridDim.x * kXchgThreads;
    for (long long i = (long long)blockIdx.x * kXchgThreads + threadIdx.x; i < n; i += stride) acc[i] = 0.0;
}

// out = sum over ranks of acc_r on [n | sx | upper triangle of sxx]; grid.x = 1 + d: block 0 takes n and sx, block 1 + i
// row i of the triangle (columns i .. d - 1: coalesced 8-byte reads from every peer)
__global__ void __launch_bounds__(kXchgThreads) fad_allreduce_peers_kernel(PeerPtrs p, int world, int rank, int d,
                                                                           unsigned round, double* __restrict__ out) {
    unsigned* mine = p.flags[rank];
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();  // this rank's moments (written by the kernels before this one) are visible system-wide
        st_release_sys(p.flags[threadIdx.x] + kFlagReady + rank, round);
    }
    wait_flags(mine + kFlagReady, world, round);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < 1 + d; i += kXchgThreads) {
            double s = 0.0;
            for (int r = 0; r < world; ++r) s += ld_peer(p.acc[r] + i);
            out[i] = s;
        }
    } else {
        const int i = blockIdx.x - 1;
        const long long row = 1 + d + (long long)i * d;
        for (int j = i + threadIdx.x; j < d; j += kXchgThreads) {
            double s = 0.0;
            for (int r = 0; r < world; ++r) s += ld_peer(p.acc[r] + row + j);
            out[row + j] = s;
        }
    }
    // last block to finish: every peer may overwrite its accumulator as far as this rank is concerned
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(mine + kFlagCounter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) mine[kFlagCounter] = 0u;
        if ((int)threadIdx.x < world) st_release_sys(p.flags[threadIdx.x] + kFlagDone + rank, round);
    }
}

// two-shot: reduce my rows of the triangle from all peers, push them into every peer's sum buffer
struct PushPtrs {
    double* sum[kXchgMaxWorld];
};
__global__ void __launch_bounds__(kXchgThreads) fad_allreduce_push_kernel(PeerPtrs p, PushPtrs q, int world, int rank,
                                                                          int d, unsigned round) {
    unsigned* mine = p.flags[rank];
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();  // this rank's moments (written by the kernels before this one) are visible system-wide
        st_release_sys(p.flags[threadIdx.x] + kFlagReady + rank, round);
    }
    wait_flags(mine + kFlagReady, world, round);
    if (blockIdx.x == 0) {  // n and sum x: d + 1 values, every rank sums them for itself
        double* out = q.sum[rank];
        for (int i = threadIdx.x; i < 1 + d; i += kXchgThreads) {
            double s = 0.0;
            for (int r = 0; r < world; ++r) s += ld_peer(p.acc[r] + i);
            out[i] = s;
        }
    } else {
        const int i = rank + world * (blockIdx.x - 1);  // rows are interleaved over the ranks: equal triangle shares
        if (i < d) {
            const long long row = 1 + d + (long long)i * d;
            for (int j = i + threadIdx.x; j < d; j += kXchgThreads) {
                double s = 0.0;
                for (int r = 0; r < world; ++r) s += ld_peer(p.acc[r] + row + j);
                for (int r = 0; r < world; ++r) q.sum[r][row + j] = s;
            }
        }
    }
    // last block: every pushed row is visible system-wide, then DONE (= "my rows have landed in your sum buffer" and
    // "I have finished reading your accumulator")
    __threadfence_system();  // every thread: its pushed values before the block counts itself done
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(mine + kFlagCounter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) mine[kFlagCounter] = 0u;
        if ((int)threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(p.flags[threadIdx.x] + kFlagDone + rank, round);
        }
    }
}

// mu, cov from an accumulator whose sxx holds (at least) the upper triangle
__global__ void __launch_bounds__(kXchgThreads) fad_finalize_sym_kernel(const double* __restrict__ acc, int d,
                                                                        double* __restrict__ mu,
                                                                        double* __restrict__ cov,
                                                                        const unsigned* __restrict__ my_flags,
                                                                        int world, unsigned round) {
    if (my_flags != nullptr) wait_flags(my_flags + kFlagDone, world, round);  // every rank's rows have landed in acc
    const double n = acc[0];
    const long long idx = (long long)blockIdx.x * kXchgThreads + threadIdx.x;
    if (idx >= (long long)d * d) return;
    const int i = (int)(idx / d), j = (int)(idx % d);
    const int lo = min(i, j), hi = max(i, j);
    const double mi = acc[1 + i] / n, mj = acc[1 + j] / n;
    if (j == 0) mu[i] = mi;
    // fadtk/utils.py:42-46: cov = S/(n-1), zeros when n < 2 ; S = sum xx^T - n mu mu^T
    cov[idx] = (n < 2.0) ? 0.0 : (acc[1 + d + (long long)lo * d + hi] - n * mi * mj) / (n - 1.0);
}

// triangle <-> packed vector (for the NCCL variant): packed = [n | sx | rows i: columns i .. d-1]
__global__ void __launch_bounds__(kXchgThreads) fad_pack_tri_kernel(const double* __restrict__ acc, int d,
                                                                    double* __restrict__ packed, int unpack,
                                                                    double* __restrict__ acc_out) {
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < 1 + d; i += kXchgThreads) {
            if (unpack) acc_out[i] = packed[i];
            else packed[i] = acc[i];
        }
        return;
    }
    const int i = blockIdx.x - 1;
    const long long row = 1 + d + (long long)i * d;
    const long long prow = 1 + d + (long long)i * d - (long long)i * (i - 1) / 2 - i;  // packed offset of (i, i) minus i
    for (int j = i + threadIdx.x; j < d; j += kXchgThreads) {
        if (unpack) acc_out[row + j] = packed[prow + j];
        else packed[prow + j] = acc[row + j];
    }
}

}  // namespace dm

using namespace dm;

extern "C" long long dm_fad_packed_doubles(int d) { return d > 0 ? 1LL + d + (long long)d * (d + 1) / 2 : 0; }
extern "C" int dm_fad_flag_words(void) { return 2 * kXchgMaxWorld + 8; }

extern "C" int dm_fad_reset_shared(double* acc, int d, const unsigned* my_flags, int world, unsigned done_round,
                                   dm_stream_t stream) {
    DM_REQUIRE(acc && my_flags && d > 0 && world >= 1 && world <= kXchgMaxWorld);
    const long long n = 1LL + d + (long long)d * d;
    const int grid = (int)std::min<long long>((n + kXchgThreads - 1) / kXchgThreads, 4LL * num_sms());
    fad_reset_shared_kernel<<<grid, kXchgThreads, 0, as_stream(stream)>>>(acc, n, my_flags, world, done_round);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_fad_allreduce_peers(const double* const* peer_acc, unsigned* const* peer_flags, int world, int rank,
                                      int d, unsigned round, double* out_acc, dm_stream_t stream) {
    DM_REQUIRE(peer_acc && peer_flags && out_acc && d > 0 && world >= 1 && world <= kXchgMaxWorld);
    DM_REQUIRE(rank >= 0 && rank < world && round >= 1);
    PeerPtrs p;
    for (int r = 0; r < kXchgMaxWorld; ++r) {
        p.acc[r] = r < world ? peer_acc[r] : nullptr;
        p.flags[r] = r < world ? peer_flags[r] : nullptr;
        DM_REQUIRE(r >= world || (p.acc[r] != nullptr && p.flags[r] != nullptr));
    }
    fad_allreduce_peers_kernel<<<1 + d, kXchgThreads, 0, as_stream(stream)>>>(p, world, rank, d, round, out_acc);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_fad_finalize_sym(const double* acc, int d, double* mu, double* cov, dm_stream_t stream) {
    DM_REQUIRE(acc && mu && cov && d > 0);
    const long long n = (long long)d * d;
    fad_finalize_sym_kernel<<<(unsigned)((n + kXchgThreads - 1) / kXchgThreads), kXchgThreads, 0, as_stream(stream)>>>(
        acc, d, mu, cov, nullptr, 0, 0u);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_fad_finalize_shared(const double* sum, int d, const unsigned* my_flags, int world, unsigned round,
                                      double* mu, double* cov, dm_stream_t stream) {
    DM_REQUIRE(sum && mu && cov && my_flags && d > 0 && world >= 1 && world <= kXchgMaxWorld && round >= 1);
    const long long n = (long long)d * d;
    fad_finalize_sym_kernel<<<(unsigned)((n + kXchgThreads - 1) / kXchgThreads), kXchgThreads, 0, as_stream(stream)>>>(
        sum, d, mu, cov, my_flags, world, round);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_fad_allreduce_push(const double* const* peer_acc, double* const* peer_sum,
                                     unsigned* const* peer_flags, int world, int rank, int d, unsigned round,
                                     dm_stream_t stream) {
    DM_REQUIRE(peer_acc && peer_sum && peer_flags && d > 0 && world >= 1 && world <= kXchgMaxWorld);
    DM_REQUIRE(rank >= 0 && rank < world && round >= 1);
    PeerPtrs p;
    PushPtrs q;
    for (int r = 0; r < kXchgMaxWorld; ++r) {
        p.acc[r] = r < world ? peer_acc[r] : nullptr;
        p.flags[r] = r < world ? peer_flags[r] : nullptr;
        q.sum[r] = r < world ? peer_sum[r] : nullptr;
        DM_REQUIRE(r >= world || (p.acc[r] != nullptr && p.flags[r] != nullptr && q.sum[r] != nullptr));
    }
    const int rows = (d - rank + world - 1) / world;  // rows rank, rank + W, ...
    fad_allreduce_push_kernel<<<1 + (rows > 0 ? rows : 0), kXchgThreads, 0, as_stream(stream)>>>(p, q, world, rank, d,
                                                                                                 round);
    DM_LAUNCHED();
    return DM_OK;
}

extern "C" int dm_fad_pack_tri(const double* acc, int d, double* packed, dm_stream_t stream) {
    DM_REQUIRE(acc && packed && d > 0);
    fad_pack_tri_kernel<<<1 + d, kXchgThreads, 0, as_stream(stream)>>>(acc, d, packed, 0, nullptr);
    DM_LAUNCHED();
    return DM_OK;
}
extern "C" int dm_fad_unpack_tri(const double* packed, int d, double* acc, dm_stream_t stream) {
    DM_REQUIRE(acc && packed && d > 0);
    fad_pack_tri_kernel<<<1 + d, kXchgThreads, 0, as_stream(stream)>>>(nullptr, d, const_cast<double*>(packed), 1, acc);
    DM_LAUNCHED();
    return DM_OK;
}

// ---- the same exchange through a caller-owned NCCL communicator -------------------------------------------------------
namespace {
using nccl_allreduce_fn = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
nccl_allreduce_fn resolve_nccl() {
    static nccl_allreduce_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");  // already loaded by the host framework?
        if (!sym) {
            for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
                if (void* h = dlopen(name, RTLD_NOW | RTLD_GLOBAL)) {
                    sym = dlsym(h, "ncclAllReduce");
                    if (sym) break;
                }
            }
        }
        fn = reinterpret_cast<nccl_allreduce_fn>(sym);
    }
    return fn;
}
}  // namespace

extern "C" int dm_fad_allreduce(void* nccl_comm, double* acc, int d, double* packed_work, dm_stream_t stream) {
    DM_REQUIRE(nccl_comm && acc && packed_work && d > 0);
    nccl_allreduce_fn fn = resolve_nccl();
    if (!fn) return fail(DM_ERR_UNSUPPORTED, "%s: ncclAllReduce is not available in this process", __func__);
    int rc = dm_fad_pack_tri(acc, d, packed_work, stream);
    if (rc != DM_OK) return rc;
    const int kNcclFloat64 = 8, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values of nccl.h
    const int nrc = fn(packed_work, packed_work, (size_t)dm_fad_packed_doubles(d), kNcclFloat64, kNcclSum, nccl_comm,
                       as_stream(stream));
    if (nrc != 0) return fail(DM_ERR_CUDA, "%s: ncclAllReduce failed with code %d", __func__, nrc);
    return dm_fad_unpack_tri(packed_work, d, acc, stream);
}

// Workspace sizes (bytes) the caller has to provide (the library never allocates):
//   DM_WS_STFT_COTANGENT : (B, Ly + 1024) fp32 padded cotangent of dm_stft_guidance           a = Ly, b = B
//   DM_WS_STFT_PARTIAL   : (B, ntiles) fp32 per-tile sums of squares                          a = Ly, b = B, c = frames/tile
//   DM_WS_FAD_ACC        : [n | sx | sxx] float64 accumulator                                 c = d
//   DM_WS_FAD_PACKED     : packed upper-triangle exchange buffer                              c = d
//   DM_WS_FAD_FLAGS      : flag pad of the peer exchange
extern "C" long long dm_workspace_bytes(int kind, long long a, long long b, int c) {
    switch (kind) {
        case DM_WS_STFT_COTANGENT: return a > 0 && b > 0 ? 4LL * b * (a + 1024) : -1;
        case DM_WS_STFT_PARTIAL: {
            if (a <= 0 || b <= 0 || c <= 0) return -1;
            const long long T = 1 + a / 160;
            return 4LL * b * ((T + c - 1) / c);
        }
        case DM_WS_FAD_ACC: return c > 0 ? 8LL * (1 + c + (long long)c * c) : -1;
        case DM_WS_FAD_PACKED: return c > 0 ? 8LL * dm_fad_packed_doubles(c) : -1;
        case DM_WS_FAD_FLAGS: return 4LL * dm_fad_flag_words();
        default: return -1;
    }
}
