// FAD second moments  S += X^T X  on the 5th-generation tensor cores (tcgen05), operands fed by TMA.
//
//   X : (N, d) fp16 row-major (what fadtk caches, model_loader.py:46-48).  fp16 x fp16 products are exact in fp32, the
//   TMEM accumulator is fp32 and is drained into float64 REGISTER accumulators every kFlush K-blocks (512 rows), so the
//   only rounding is the fp32 running sum inside one 512-row slab (measured: 1.8e-6 relative on sum x x^T, growing
//   linearly with the slab length -- the tensor core's fp32 accumulation truncates); np.cov (fadtk/fad.py:47) is float64.
//   Slab length and stage depth were tuned on the B200 (profiles/README.md): 64-row stages with 256-row slabs left the
//   tensor pipe at 39 % (MMA-issue / hand-off overhead per stage and per slab), 128-row stages with 512-row slabs
//   reach 55 %.
//
//   C tile (128 x 128) = A^T-tile * B-tile with A = X[:, i-block], B = X[:, j-block]: both operands are "MN-major"
//   (contiguous along the output dimension), which kind::f16 supports directly -- no transpose pass over X.
//   TMA loads 64-column x 64-row boxes with the 128-byte swizzle straight into the canonical UMMA layout:
//       atom = 64 (MN) x 8 (K) fp16 = 1024 B, K groups every 1024 B (SBO), MN halves every 8192 B (LBO).
//
//   Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected thread),
//   warps 2..9 = epilogue: TMEM -> registers (tcgen05.ld) -> float64 accumulators; two TMEM accumulator stages let the
//   drain of slab s overlap the MMAs of slab s+1.  Upper-triangular tiles only; the mirror is written on the way out.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "dm_common.cuh"

namespace dm {

constexpr int kTile = 128;          // output tile (M = N = 128)
constexpr int kBoxCols = 64;        // TMA box: 64 fp16 = 128 B (one swizzle row)
constexpr int kBlockK = 128;        // rows of X per pipeline stage
constexpr int kUmmaK = 16;          // K of one tcgen05.mma (fp16)
constexpr int kStages = 3;
constexpr int kAccStages = 2;
constexpr int kFlush = 4;           // K-blocks per TMEM slab (512 rows) before draining to float64
constexpr int kTcThreads = 320;
constexpr int kEpiWarps = 8;
constexpr uint32_t kBoxBytes = kBoxCols * kBlockK * 2;          // 8 KB
constexpr uint32_t kOperandBytes = 2 * kBoxBytes;               // 128 columns x 64 rows = 16 KB
constexpr uint32_t kStageBytes = 2 * kOperandBytes;             // A + B
constexpr int kSumCols = 16;                                    // N of the column-sum MMA (the smallest N for M = 128)
constexpr uint32_t kTmemCols = 512;                             // 2 x 128 accumulator columns + 2 x 16 column-sum columns
constexpr uint32_t kSumTmemCol = kAccStages * kTile;            // first column of the column-sum accumulators
constexpr uint32_t kOnesBytes = kBoxBytes;                      // a 64 x kBlockK tile of fp16 ones (any layout)
constexpr size_t kTcSmemBytes = (size_t)kStages * kStageBytes + kOnesBytes + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                                  uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%4, %5}], [%2], %3;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// UMMA shared-memory descriptor, MN-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((kBoxBytes >> 4) & 0x3FFF) << 16;  // LBO: next 64 MN elements = next TMA box
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;      // SBO: next 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both MN-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((kTile >> 3) << 17) |
                            ((kTile >> 4) << 24);
// the same with N = 16: A^T (128 x K) times a K x 16 block of ones = the column sums of the A block, 16 times over
constexpr uint32_t kIdescSum = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((kSumCols >> 3) << 17) |
                               ((kTile >> 4) << 24);

struct FadTcParams {
    int d, ntile;           // columns, tiles per side
    long long N, rows_per_cta;
    double* sxx;
    double* sx;             // optional: column sums (d) and the row count, produced by the diagonal tiles with one extra
    double* n_out;          // N = 16 MMA per k-step against a block of ones -- X is not read a second time for them
};

// kCluster == 2: the two CTAs of a cluster own tiles (ti, tj) and (ti, tj+1) of the same row range, i.e. they need the
// SAME A operand (column block ti).  Each CTA fetches one 64-column half of A and TMA-multicasts it into both CTAs'
// shared memory, so the L2 -> SM traffic per CTA drops from A + B to A/2 + B (the kernel is L2-fabric bound, see
// profiles/README.md).  A stage may be refilled only when BOTH consumers are done with it: the MMA warp's commit is
// multicast to both CTAs' `empty` barriers (count 2).
// kCluster == 1: the leftover tile of rows with an odd number of upper-triangular tiles, no multicast.
template <int kCluster>
__global__ void __launch_bounds__(kTcThreads, 1)
    fad_xtx_tc_kernel(const __grid_constant__ CUtensorMap tmap, const FadTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ones = smem + (size_t)kStages * kStageBytes;  // fp16 ones: the B operand of the column-sum MMA
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + kOnesBytes);
    uint64_t* full = bars;                       // [kStages]  TMA -> MMA
    uint64_t* empty = bars + kStages;            // [kStages]  MMA -> TMA
    uint64_t* tfull = bars + 2 * kStages;        // [kAccStages] MMA -> epilogue
    uint64_t* tempty = tfull + kAccStages;       // [kAccStages] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile assignment over the upper triangle (tj >= ti).  Row ti has n = ntile - ti tiles: n/2 pairs handled by
    // 2-CTA clusters, and when n is odd its last tile (ti, ntile-1) by a single CTA.
    const uint32_t crank = (kCluster == 2) ? cluster_ctarank() : 0;
    int ti = 0, tj;
    if (kCluster == 2) {
        int pair = blockIdx.x >> 1;
        while (pair >= (p.ntile - ti) / 2) {
            pair -= (p.ntile - ti) / 2;
            ++ti;
        }
        tj = ti + 2 * pair + (int)crank;
    } else {
        int single = blockIdx.x;
        for (;; ++ti) {
            if ((p.ntile - ti) & 1) {
                if (single == 0) break;
                --single;
            }
        }
        tj = p.ntile - 1;
    }
    const bool diag = (ti == tj);
    const bool sums = diag && p.sx != nullptr;  // every column block has exactly one diagonal tile
    constexpr uint16_t kAllCtas = (uint16_t)((1u << kCluster) - 1);
    for (uint32_t i = threadIdx.x; i < kOnesBytes / 4; i += kTcThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA's reads
    const long long r_begin = (long long)blockIdx.y * p.rows_per_cta;
    const long long r_end = min(p.N, r_begin + p.rows_per_cta);
    const int nkb = r_end > r_begin ? (int)((r_end - r_begin + kBlockK - 1) / kBlockK) : 0;
    const int nslab = (nkb + kFlush - 1) / kFlush;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kCluster);  // every CTA that received a multicast half must release the stage
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();  // peer barriers are initialised before any multicast / remote arrive
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
                uint8_t* a = smem + (size_t)s * kStageBytes;
                uint8_t* b = a + kOperandBytes;
                const int row = (int)(r_begin + (long long)kb * kBlockK);
                mbar_expect_tx(&full[s], diag ? kOperandBytes : 2 * kOperandBytes);
                if (kCluster == 2) {  // my half of the shared A operand goes to both CTAs
                    tma_load_2d_mcast(&tmap, &full[s], a + crank * kBoxBytes, ti * kTile + (int)crank * kBoxCols, row,
                                      kAllCtas);
                } else {
                    tma_load_2d(&tmap, &full[s], a, ti * kTile, row);
                    tma_load_2d(&tmap, &full[s], a + kBoxBytes, ti * kTile + kBoxCols, row);
                }
                if (!diag) {
                    tma_load_2d(&tmap, &full[s], b, tj * kTile, row);
                    tma_load_2d(&tmap, &full[s], b + kBoxBytes, tj * kTile + kBoxCols, row);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const int slab = kb / kFlush, as = slab % kAccStages;
                if (kb % kFlush == 0) {  // new slab: its TMEM stage must have been drained
                    mbar_wait(&tempty[as], ((slab / kAccStages) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(&full[s], (kb / kStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + (size_t)s * kStageBytes);
                const uint32_t b_addr = diag ? a_addr : a_addr + kOperandBytes;
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * kTile);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    const uint32_t koff = (uint32_t)k * kUmmaK * 128u;  // 16 K rows x 128 B
                    umma_f16(d_tmem, make_desc_mn_sw128(a_addr + koff), make_desc_mn_sw128(b_addr + koff), kIdesc,
                             (kb % kFlush != 0 || k != 0) ? 1u : 0u);
                }
                if (sums) {  // column sums of the A block: every element of the B operand is 1, so its layout is moot
                    const uint64_t ones_desc = make_desc_mn_sw128(smem_u32(ones));
                    const uint32_t s_tmem = tmem_base + kSumTmemCol + (uint32_t)(as * kSumCols);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_f16(s_tmem, make_desc_mn_sw128(a_addr + (uint32_t)k * kUmmaK * 128u), ones_desc, kIdescSum,
                                 (kb % kFlush != 0 || k != 0) ? 1u : 0u);
                }
                // smem stage free (in every CTA that multicasts into it) once these MMAs have read it
                if (kCluster == 2) umma_commit_mcast(&empty[s], kAllCtas);
                else umma_commit(&empty[s]);
                if (kb % kFlush == kFlush - 1 || kb == nkb - 1) umma_commit(&tfull[as]);
            }
        }
    } else {
        // ===================== epilogue: TMEM -> float64 registers =====================
        const int e = warp - 2;
        const int quarter = warp & 3;           // TMEM lanes a warp may touch: 32 * (warp % 4) ...
        const int half = e >> 2;                // which 64 of the 128 accumulator columns
        double acc[64];
        double sx_acc = 0.0;
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = 0.0;
        for (int slab = 0; slab < nslab; ++slab) {
            const int as = slab % kAccStages;
            mbar_wait(&tfull[as], (slab / kAccStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kTile + half * 64);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tmem_ld32(taddr + h * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[h * 32 + c] += (double)__uint_as_float(v[c]);
            }
            if (sums && half == 0) {  // lane = column of the A block; the 16 sum columns are identical, read the first
                uint32_t sv;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];"
                             : "=r"(sv)
                             : "r"(tmem_base + ((uint32_t)(quarter * 32) << 16) + kSumTmemCol + (uint32_t)(as * kSumCols))
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                sx_acc += (double)__uint_as_float(sv);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
        // ---- add the tile into the float64 accumulator (and its mirror) ----
        const int gi = ti * kTile + quarter * 32 + lane;
        if (sums && half == 0 && gi < p.d && nkb > 0) atomicAdd(&p.sx[gi], sx_acc);
        if (sums && ti == 0 && blockIdx.y == 0 && e == 0 && lane == 0) atomicAdd(p.n_out, (double)p.N);
        if (gi < p.d && nkb > 0) {
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const int gj = tj * kTile + half * 64 + c;
                if (gj < p.d) {
                    atomicAdd(&p.sxx[(long long)gi * p.d + gj], acc[c]);
                    if (!diag) atomicAdd(&p.sxx[(long long)gj * p.d + gi], acc[c]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();  // the peer may still arrive on / multicast into this CTA's shared memory
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ======================================================================================================================
// cta_group::2 variant: a CTA PAIR (two SMs of one TPC) owns a 256 x 128 output super-tile -- column blocks (2 tp, 2 tp + 1)
// of X against column block tj.  One tcgen05.mma.cta_group::2 (M = 256, N = 128) issued by the leader CTA drives the
// tensor cores of both SMs: every CTA stages ITS 128 columns of A (16 KB per 64-row k-block) but only HALF of B (64
// columns, 8 KB) -- the other half is read out of the peer's shared memory by the pair's MMA datapath.  Per SM and
// 2.1 MFLOP that is 24 KB of L2 -> SM traffic instead of 32 KB for two independent 128 x 128 tiles.  Accumulators: each
// CTA keeps its own 128 x 128 half in TMEM and drains it to float64 registers as above.
// MEASURED (profiles/README.md): correct, but not faster than the multicast-cluster kernel above -- 43 % tensor-pipe
// activity against 55 %: the super-tiles cover 24 block tiles where 21 are needed (the block below the diagonal of every
// diagonal super-tile is wasted, diagonal tiles cannot share A = B), the total L2 read volume is the same 24 KB per
// tile, and the two SMs of a pair stall together.  Kept as engine DM_FAD_TCGEN05_PAIR; DM_FAD_AUTO uses the kernel above.
//   barriers   leader: full[s] (both CTAs' TMA loads land their transaction bytes there), tempty[a] (16 epilogue warps
//              of both CTAs arrive);  every CTA: empty[s], tfull[a] (multicast tcgen05.commit of the leader).
constexpr uint32_t kPairBBytes = kBoxBytes;                                  // 64 columns of B per CTA
constexpr uint32_t kPairStageBytes = kOperandBytes + kPairBBytes;           // 24 KB per CTA and stage
constexpr int kPairStages = 4;
constexpr size_t kTc2SmemBytes = (size_t)kPairStages * kPairStageBytes + 1024 + 256;
// instruction descriptor: as kIdesc with M = 256
constexpr uint32_t kIdesc2 = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((kTile >> 3) << 17) |
                             ((256u >> 4) << 24);

__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose transaction bytes are accounted on a barrier given by its
// shared::cluster address (the leader's `full` barrier)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0,
                                                 int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
// bounded wait: a protocol error becomes a trap (-> a CUDA error at the next sync) instead of a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    for (unsigned spin = 0;; ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > 64) __nanosleep(64);
        if (spin > (1u << 24)) __trap();  // ~2 s: a protocol error, not a long wait
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    fad_xtx_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const FadTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kPairStages * kPairStageBytes);
    uint64_t* full = bars;                          // [kPairStages]  (leader's copy is the live one)
    uint64_t* empty = bars + kPairStages;           // [kPairStages]
    uint64_t* tfull = bars + 2 * kPairStages;       // [kAccStages]
    uint64_t* tempty = tfull + kAccStages;          // [kAccStages]  (leader's copy is the live one)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    // super-tile (tp, tj), tj >= 2 tp, enumerated row pair by row pair
    int tp = 0, rem = blockIdx.x >> 1;
    while (rem >= p.ntile - 2 * tp) {
        rem -= p.ntile - 2 * tp;
        ++tp;
    }
    const int tj = 2 * tp + rem;
    const int bi = 2 * tp + (int)crank;  // the column block of X this CTA's half of the accumulator belongs to
    const long long r_begin = (long long)blockIdx.y * p.rows_per_cta;
    const long long r_end = min(p.N, r_begin + p.rows_per_cta);
    const int nkb = r_end > r_begin ? (int)((r_end - r_begin + kBlockK - 1) / kBlockK) : 0;
    const int nslab = (nkb + kFlush - 1) / kFlush;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kPairStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 2 * kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // same warp in both CTAs: the allocation is a pair-wide operation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit / peer TMA signal
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kPairStages;
                mbar_wait_bounded(&empty[s], ((kb / kPairStages) & 1) ^ 1);
                uint8_t* a = smem + (size_t)s * kPairStageBytes;
                uint8_t* b = a + kOperandBytes;
                const int row = (int)(r_begin + (long long)kb * kBlockK);
                if (leader) mbar_expect_tx(&full[s], 2 * kPairStageBytes);  // both CTAs' bytes land on this barrier
                const uint32_t lead_full = mapa_rank(smem_u32(&full[s]), 0);
                tma_load_2d_pair(&tmap, lead_full, a, bi * kTile, row);
                tma_load_2d_pair(&tmap, lead_full, a + kBoxBytes, bi * kTile + kBoxCols, row);
                tma_load_2d_pair(&tmap, lead_full, b, tj * kTile + (int)crank * kBoxCols, row);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kPairStages;
                const int slab = kb / kFlush, as = slab % kAccStages;
                if (kb % kFlush == 0) {
                    mbar_wait_bounded(&tempty[as], ((slab / kAccStages) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait_bounded(&full[s], (kb / kPairStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + (size_t)s * kPairStageBytes);
                const uint32_t b_addr = a_addr + kOperandBytes;
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * kTile);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    const uint32_t koff = (uint32_t)k * kUmmaK * 128u;
                    umma_f16_pair(d_tmem, make_desc_mn_sw128(a_addr + koff), make_desc_mn_sw128(b_addr + koff), kIdesc2,
                                  (kb % kFlush != 0 || k != 0) ? 1u : 0u);
                }
                umma_commit_pair(&empty[s]);
                if (kb % kFlush == kFlush - 1 || kb == nkb - 1) umma_commit_pair(&tfull[as]);
            }
        }
    } else {
        // ===================== epilogue (both CTAs): TMEM -> float64 registers =====================
        const int e = warp - 2;
        const int quarter = warp & 3;
        const int half = e >> 2;
        const uint32_t lead_tempty0 = mapa_rank(smem_u32(&tempty[0]), 0);
        double acc[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = 0.0;
        for (int slab = 0; slab < nslab; ++slab) {
            const int as = slab % kAccStages;
            mbar_wait_bounded(&tfull[as], (slab / kAccStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kTile + half * 64);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tmem_ld32(taddr + h * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[h * 32 + c] += (double)__uint_as_float(v[c]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty0 + (uint32_t)(as * sizeof(uint64_t)));
        }
        // block (bi, tj): above the diagonal -> tile and mirror; on it -> tile; below it (bi = tj + 1) -> the mirror of
        // block (tj, bi), which the super-tile (tp, tj + 1) writes
        const int gi = bi * kTile + quarter * 32 + lane;
        if (bi <= tj && gi < p.d && nkb > 0) {
            const bool diag = bi == tj;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const int gj = tj * kTile + half * 64 + c;
                if (gj < p.d) {
                    atomicAdd(&p.sxx[(long long)gi * p.d + gj], acc[c]);
                    if (!diag) atomicAdd(&p.sxx[(long long)gj * p.d + gi], acc[c]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // the peer may still arrive on this CTA's barriers / read its shared memory through the MMA
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// returns DM_ERR_UNSUPPORTED when the shape cannot go through TMA (caller falls back to the SIMT kernel)
// sx / n_out (may be NULL): column sums and row count, fused into the diagonal tiles (not by the cta_group::2 engine)
int fad_xtx_tc(const void* x_f16, long long N, int d, double* sxx, double* sx, double* n_out, bool pair_mma,
               cudaStream_t st) {
    if (d % 8 != 0 || (reinterpret_cast<uintptr_t>(x_f16) & 15) != 0 || N > 0x7fffffffLL)
        return fail(DM_ERR_UNSUPPORTED, "%s: TMA needs d %% 8 == 0 and a 16-byte aligned base", __func__);
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) return fail(DM_ERR_CUDA, "%s: cuTensorMapEncodeTiled not available", __func__);
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
    const cuuint32_t box[2] = {kBoxCols, kBlockK};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x_f16), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DM_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", __func__, (int)r);
    FadTcParams p;
    p.d = d;
    p.ntile = (d + kTile - 1) / kTile;
    p.N = N;
    p.sxx = sxx;
    p.sx = pair_mma ? nullptr : sx;
    p.n_out = pair_mma ? nullptr : n_out;
    int npairs = 0, nsingles = 0;
    for (int ti = 0; ti < p.ntile; ++ti) {
        npairs += (p.ntile - ti) / 2;
        nsingles += (p.ntile - ti) & 1;
    }
    const long long slab_rows = (long long)kFlush * kBlockK;
    if (pair_mma) {  // cta_group::2: one CTA pair per 256 x 128 super-tile
        int nsuper = 0;
        for (int tp = 0; 2 * tp < p.ntile; ++tp) nsuper += p.ntile - 2 * tp;
        long long splits = std::max<long long>(1, num_sms() / (2 * nsuper));
        long long rows = (N + splits - 1) / splits;
        p.rows_per_cta = std::max<long long>(slab_rows, (rows + slab_rows - 1) / slab_rows * slab_rows);
        const int gy = (int)((N + p.rows_per_cta - 1) / p.rows_per_cta);
        DM_SMEM_ONCE(fad_xtx_tc2_kernel, kTc2SmemBytes);
        fad_xtx_tc2_kernel<<<dim3(2 * nsuper, gy), kTcThreads, kTc2SmemBytes, st>>>(tmap, p);
        DM_LAUNCHED();
        return DM_OK;
    }
    // one CTA per SM (192 KB of pipeline stages): each launch splits the rows so that its grid covers the machine about
    // once, in multiples of the flush slab so every CTA drains whole slabs
    auto rows_for = [&](int tiles) {
        long long splits = std::max<long long>(1, num_sms() / tiles);
        long long rows = (N + splits - 1) / splits;
        return std::max<long long>(slab_rows, (rows + slab_rows - 1) / slab_rows * slab_rows);
    };
    if (npairs > 0) {
        p.rows_per_cta = rows_for(2 * npairs);
        const int gy = (int)((N + p.rows_per_cta - 1) / p.rows_per_cta);
        DM_SMEM_ONCE(fad_xtx_tc_kernel<2>, kTcSmemBytes);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * npairs, gy);
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = kTcSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        DM_CUDA(cudaLaunchKernelEx(&cfg, fad_xtx_tc_kernel<2>, tmap, p));
        DM_LAUNCHED();
    }
    if (nsingles > 0) {
        p.rows_per_cta = rows_for(nsingles);
        const int gy = (int)((N + p.rows_per_cta - 1) / p.rows_per_cta);
        DM_SMEM_ONCE(fad_xtx_tc_kernel<1>, kTcSmemBytes);
        fad_xtx_tc_kernel<1><<<dim3(nsingles, gy), kTcThreads, kTcSmemBytes, st>>>(tmap, p);
        DM_LAUNCHED();
    }
    return DM_OK;
}

}  // namespace dm
