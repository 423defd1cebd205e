// Episode Gram blocks on the 5th-generation tensor cores (tcgen05 + TMEM), for NK <= 32 and NK + NQ <= 192.
//
//   gram[b, i, j] = <x_i, x_j>,  i over the task's NK support + NQ query rows, j over its NK support rows
//
// Same contract as gram.cu (reference: the per-step F.linear(x, W0_task) of fumi.py:161,178 in Gram form);
// this is the kernel that streams the sampled feature rows from HBM, so it is built to be HBM-bound.
// A first version with both operands in shared memory was shared-memory-bandwidth bound: with N = 32 every
// tcgen05.mma re-reads its 4 KB A tile for 32 columns of output, 120 KB of operand reads per 24 KB of new
// data.  So the big operand (all rows of the task) never touches shared memory:
//
//   warps 0-5 (loaders)    gather: cp.async (LDGSTS), 16 bytes per lane, 8 lanes per 128-byte row piece, the
//                          next 32 features of every row of the task per stage, four stages (96 KB) in
//                          flight, into a swizzled staging ring (chunk c of row r at c ^ (r & 7), so that both
//                          the coalesced writes and the row-per-thread reads are conflict-free).
//                          convert: thread r reads row r's 128 bytes back.  The raw fp32 bits ARE the tf32
//                          "hi" operand (the tensor core ignores the low 13 mantissa bits), lo = x - trunc(x)
//                          is computed in registers, and both go to TENSOR MEMORY with tcgen05.st (lane =
//                          row, column = feature: the A-operand layout of an M=128 MMA).  (Loading a row
//                          per thread straight from global memory was tried: 32 lines per load instruction,
//                          L1 tag-bound, 2x slower.)  Warp 0 owns the support rows and also writes them to
//                          shared memory (K-major SWIZZLE_128B, hi and lo planes, 8 KB per stage): B.
//   warp 6 (MMA issuer)    one thread: per stage, for both row tiles (rows 0-127, 128-191) and 4 k-steps
//                          (lo.hi, hi.lo, hi.hi) tcgen05.mma.kind::tf32 M128 N32 K8 with A from TMEM, into
//                          an accumulator that is restarted every 2 stages (the tensor core accumulates with
//                          truncation; see dense_tc.cu).  tcgen05.commit frees the stage / publishes it.
//   warps 8-11 (drain)     tcgen05.ld the finished accumulator and add it into fp32 registers (round to
//                          nearest); at the end of the task the NK-wide rows go to HBM.
// Measured (tools/gram_time.py, FUMI_GRAM_DBG ablations, 4096 tasks): 1.89 ms = 3.3 TB/s of gathered rows vs 2.15 ms
// for the warp-level kernel.  The bound is the MMA stream itself: a tcgen05.mma M128 K8 tf32 costs ~60-85
// cycles whatever N <= 32 or M is (the 4 KB A tile is fetched at ~64 B/clk), so 24 MMAs per 32-feature stage take
// ~2,000 cycles where HBM needs ~1,000; gather and TMEM stores are hidden behind it (skipping them changes
// nothing), alternating accumulators between consecutive MMAs is 35 % slower.  Shapes with N >= 128 per
// instruction are what tcgen05 wants; the next step here is K = 16 per instruction (fp16 hi/lo planes with a
// bank-wide power-of-two scale), which halves the MMA count.
// TMEM columns: [0,128) two accumulator buffers x (tile 0 | tile 1) x 32; [128,512) three A stages x
// (tile 0 hi | tile 0 lo | tile 1 hi | tile 1 lo) x 32.  One persistent CTA per SM walks tasks
// b = blockIdx.x, blockIdx.x + gridDim.x, ...; the stage ring runs across task boundaries.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/fumi_b200.h"
#include "common.cuh"

namespace {

constexpr int kRows = 192;                       // rows per task: tile 0 = rows 0-127, tile 1 = rows 128-191
template <bool F16> struct GramCfg {              // F16: fp16 (hi, lo) bank planes, 64 features per 128-byte stage row
    static constexpr int kFeat = F16 ? 64 : 32;                  // features per stage
    static constexpr int kRingSlots = F16 ? 4 : 6;               // staging ring of gathered row pieces
    static constexpr int kAhead = F16 ? 2 : 4;                   // stages of cp.async in flight (96 KB either way)
    static constexpr uint32_t kRingSlot = 192u * 128u * (F16 ? 2u : 1u);      // 24 KB per plane
    static constexpr int kDrain = F16 ? 1 : 2;                   // stages (64 features) per TMEM accumulator
    static constexpr uint32_t kSmem = 3u * 2u * 4096u + kRingSlots * kRingSlot + 1024 + 256;
};
constexpr int kBK = 32;                          // 32-bit words per stage row = 128 bytes
constexpr int kStagesG = 3;                      // A stages in TMEM / B stages in shared memory
constexpr uint32_t kBPlane = 32 * kBK * 4;       // 4 KB: 32 support rows x 128 B
constexpr uint32_t kBStage = 2 * kBPlane;        // hi + lo
constexpr int kThreadsG = 12 * 32;               // warps 0-5 loaders, 6 MMA, 7 idle, 8-11 drain
constexpr int kTmemCols = 512;
constexpr uint32_t kAccCols = 128, kAStageCols = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GWAIT_DONE;\n"
        "bra GWAIT_LOOP;\n"
        "GWAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// K-major SWIZZLE_128B descriptor: start>>4 | LBO=1 | SBO=1024>>4 | version 1 | layout 2 (see dense_tc.cu)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}
// A operand from tensor memory (lane = row, 32-bit column = k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

struct GramTcParams {
    const void* feats;        // fp32 bank, or the fp16 hi plane
    const void* feats_lo;     // fp16 lo plane (F16 only)
    const float* absmax;      // device scalar that fixed the fp16 plane scale (F16 only)
    int64_t D;
    const int64_t* sup_rows;
    const int64_t* qry_rows;
    int64_t B;
    int NK, NQ;
    float* gram;
    int dbg;      // diagnostics (FUMI_GRAM_DBG): 1 skip the MMAs, 2 skip the gather, 4 skip the TMEM stores
};

__host__ __device__ __forceinline__ int gram_f16_scale_exp(float amax) {      // = f16_scale_exp of dense_tc.cu
    if (!(amax > 0.f) || !(amax < 3.0e38f)) return 0;
    int e;
    frexpf(amax, &e);
    const int k = 14 - e;
    return k > 100 ? 100 : (k < -100 ? -100 : k);
}

template <bool F16>
__global__ void __launch_bounds__(kThreadsG, 1) gram_tc_kernel(const GramTcParams p) {
    using Cfg = GramCfg<F16>;
    constexpr int kRingSlots = Cfg::kRingSlots, kAhead = Cfg::kAhead, kDrain = Cfg::kDrain;
    constexpr uint32_t kRingSlot = Cfg::kRingSlot;
    constexpr int kEsz = F16 ? 2 : 4;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t ring = base + kStagesG * kBStage;            // staging ring after the B stages
    const uint32_t bars = ring + kRingSlots * kRingSlot;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * kStagesG, tmem_full_bar = bars + 16 * kStagesG;
    const uint32_t tmem_empty_bar = tmem_full_bar + 16, tmem_slot = tmem_empty_bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = p.NK + p.NQ;
    const bool two_tiles = rows > 128;
    const int n_loader_warps = two_tiles ? 6 : 4;
    const int k_stages = int(p.D / Cfg::kFeat);                            // stages per task
    const int64_t my_tasks = p.B > blockIdx.x ? (p.B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tasks * k_stages;                             // stages this CTA runs

    // rows [NK + NQ, ..) of the staging ring are never written: they stay zero for the whole kernel
    for (uint32_t i = threadIdx.x; i < kRingSlots * kRingSlot / 16; i += kThreadsG)
        reinterpret_cast<float4*>(base_ptr + kStagesG * kBStage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesG; ++s) { mbar_init(full_bar + 8 * s, n_loader_warps); mbar_init(empty_bar + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar + 8 * b, 1); mbar_init(tmem_empty_bar + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 6) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < n_loader_warps) {
        // ===================== loaders =====================
        const int t = threadIdx.x;                               // 0 .. 32 n_loader_warps - 1
        const int LT = n_loader_warps * 32;
        // gather mapping: piece j of this thread = 16-byte chunk c of row (t >> 3) + (LT / 8) j
        const int c = t & 7, rg0 = t >> 3, rstep = LT >> 3;
        // convert mapping: this thread owns row r = t
        const int r = t;
        const int tile = r >> 7;                                 // 0: rows 0-127, 1: rows 128-191
        // TMEM address of row r's A slot: lane = r % 128 (= 32 (warp % 4) + lane), column block by tile
        const uint32_t a_lane = uint32_t((r & 127) & ~31) << 16;
        const uint32_t a_col0 = kAccCols + uint32_t(tile) * 64u;
        const uint32_t sw = uint32_t(r & 7);                     // swizzle phase of row r
        int64_t rowoff[8];                                       // byte offset of (row, chunk c) in a bank plane
        const uint8_t* plane_hi = static_cast<const uint8_t*>(p.feats);
        const uint8_t* plane_lo = static_cast<const uint8_t*>(p.feats_lo);
        int64_t gi = 0, bi = blockIdx.x;                         // issue side: stage, its task, its k chunk
        int kci = 0;
        auto task_rows = [&](int64_t b) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int rg = rg0 + rstep * j;
                const int64_t row = rg < p.NK ? p.sup_rows[b * p.NK + rg] : (rg < rows ? p.qry_rows[b * p.NQ + (rg - p.NK)] : 0);
                rowoff[j] = row * p.D * kEsz + c * 16;
            }
        };
        auto issue = [&]() {                                     // cp.async of stage gi
            if (gi < total && !(p.dbg & 2)) {
                const uint32_t dst = ring + uint32_t(gi % kRingSlots) * kRingSlot;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rg = rg0 + rstep * j;
                    if (rg < rows) {
                        const uint32_t d = dst + uint32_t(rg) * 128u + uint32_t((c ^ (rg & 7)) << 4);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                                     ::"r"(d), "l"(plane_hi + rowoff[j] + int64_t(kci) * 128) : "memory");
                        if (F16)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                                         ::"r"(d + kRingSlot / 2), "l"(plane_lo + rowoff[j] + int64_t(kci) * 128) : "memory");
                    }
                }
                if (++kci == k_stages) {
                    kci = 0;
                    bi += gridDim.x;
                    if (gi + 1 < total) task_rows(bi);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            ++gi;
        };
        auto process = [&](const uint8_t* src, int64_t g) {      // stage g: staging ring -> TMEM (and B planes)
            const int s = int(g % kStagesG);
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(src + ((uint32_t(j) ^ sw) << 4));
            mbar_wait(empty_bar + 8 * s, (uint32_t(g / kStagesG) & 1) ^ 1);
            uint32_t u[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                u[4 * j] = __float_as_uint(v[j].x); u[4 * j + 1] = __float_as_uint(v[j].y);
                u[4 * j + 2] = __float_as_uint(v[j].z); u[4 * j + 3] = __float_as_uint(v[j].w);
            }
            const uint32_t ta = tmem_base + a_lane + a_col0 + uint32_t(s) * kAStageCols;
            if (!(p.dbg & 4)) tmem_st32(ta, u);                  // hi plane (fp32: the raw bits)
            uint8_t* bst = base_ptr + s * kBStage;
            if (warp == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(bst + r * 128 + ((uint32_t(j) ^ sw) << 4)) = v[j];
            }
            if (F16) {                                           // the lo plane was gathered too
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = *reinterpret_cast<const float4*>(src + kRingSlot / 2 + ((uint32_t(j) ^ sw) << 4));
                    u[4 * j] = __float_as_uint(v[j].x); u[4 * j + 1] = __float_as_uint(v[j].y);
                    u[4 * j + 2] = __float_as_uint(v[j].z); u[4 * j + 3] = __float_as_uint(v[j].w);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float x = __uint_as_float(u[i]);
                    u[i] = __float_as_uint(x - __uint_as_float(u[i] & 0xFFFFE000u));
                }
            }
            if (!(p.dbg & 4)) tmem_st32(ta + 32, u);             // lo plane
            if (warp == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(bst + kBPlane + r * 128 + ((uint32_t(j) ^ sw) << 4)) =
                        make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                                    __uint_as_float(u[4 * j + 3]));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar + 8 * s);
        };
        if (total > 0) {
            task_rows(bi);
#pragma unroll
            for (int a = 0; a < kAhead; ++a) issue();
            for (int64_t g = 0; g < total; ++g) {
                // slot (g + kAhead) % kRingSlots was last read in iteration g - 2; every loader thread has passed the
                // named barrier of iteration g - 1 since
                issue();
                asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");   // this thread's pieces of stage g
                asm volatile("bar.sync 1, %0;" ::"r"(LT) : "memory");               // ... and everyone else's
                process(base_ptr + kStagesG * kBStage + uint32_t(g % kRingSlots) * kRingSlot + r * 128, g);
            }
        }
    } else if (warp == 6) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (uint32_t(32 >> 3) << 17) |
                                   (uint32_t(128 >> 4) << 24);
            auto mma = [&](uint32_t td, uint32_t ta, uint64_t db, uint32_t accum) {
                if (F16) umma_f16_ts(td, ta, db, idesc, accum); else umma_tf32_ts(td, ta, db, idesc, accum);
            };
            for (int64_t g = 0; g < total; ++g) {
                const int s = int(g % kStagesG);
                const int64_t grp = g / kDrain;
                const int sub = int(g - grp * kDrain);
                const int tb = int(grp & 1);
                if (sub == 0) mbar_wait(tmem_empty_bar + 8 * tb, (uint32_t(grp >> 1) & 1) ^ 1);
                mbar_wait(full_bar + 8 * s, uint32_t(g / kStagesG) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t bhi = base + s * kBStage, blo = bhi + kBPlane;
                // tile by tile (alternating the two accumulators between consecutive MMAs measured 35 % slower)
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if ((t == 1 && !two_tiles) || (p.dbg & 1)) break;
                    const uint32_t td = tmem_base + uint32_t(tb * 64 + t * 32);
                    const uint32_t ahi = tmem_base + kAccCols + uint32_t(s) * kAStageCols + uint32_t(t) * 64u, alo = ahi + 32;
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {          // small cross products first
                        mma(td, alo + 8 * k, umma_desc_sw128(bhi + k * 32), (k | sub) != 0);
                        mma(td, ahi + 8 * k, umma_desc_sw128(blo + k * 32), 1);
                    }
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) mma(td, ahi + 8 * k, umma_desc_sw128(bhi + k * 32), 1);
                }
                umma_commit(empty_bar + 8 * s);                  // A stage (TMEM) and B stage (smem) are free once these retire
                if (sub == kDrain - 1 || g == total - 1) umma_commit(tmem_full_bar + 8 * tb);
            }
        }
    } else if (warp >= 8) {
        // ===================== drain + epilogue: warp q owns TMEM lanes [32q, 32q+32) =====================
        const int q = warp & 3;
        const int groups_per_task = k_stages / kDrain;
        const bool has1 = two_tiles && q < 2;                    // tile 1: rows 128 + lane of quarters 0, 1
        int64_t grp = 0;
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
            float acc0[32], acc1[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
            for (int it = 0; it < groups_per_task; ++it, ++grp) {
                const int tb = int(grp & 1);
                mbar_wait(tmem_full_bar + 8 * tb, uint32_t(grp >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(tb * 64);
                tmem_ld32(taddr, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc0[j] += __uint_as_float(v[j]);
                if (has1) {
                    tmem_ld32(taddr + 32, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc1[j] += __uint_as_float(v[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar + 8 * tb);
            }
            if (F16) {                                           // undo the plane scale (squared): exact powers of two
                const float inv = ldexpf(1.f, -gram_f16_scale_exp(*p.absmax));
#pragma unroll
                for (int j = 0; j < 32; ++j) { acc0[j] = (acc0[j] * inv) * inv; acc1[j] = (acc1[j] * inv) * inv; }
            }
            float* g = p.gram + b * int64_t(rows) * p.NK;
            const int r0 = q * 32 + lane;
            if (r0 < rows) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < p.NK) g[int64_t(r0) * p.NK + j] = acc0[j];
            }
            const int r1 = 128 + q * 32 + lane;
            if (has1 && r1 < rows) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < p.NK) g[int64_t(r1) * p.NK + j] = acc1[j];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 6) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

}  // namespace

namespace {
template <bool F16>
int launch_gram_tc(const GramTcParams& p, void* stream) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return fumi_cuda_fail(cudaGetLastError(), "fumi_gram (device query)");
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(GramCfg<F16>::kSmem));
        if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaFuncSetAttribute(gram_tc_kernel)");
        attr_done = true;
    }
    const unsigned grid = unsigned(p.B < sms ? p.B : sms);
    gram_tc_kernel<F16><<<grid, kThreadsG, GramCfg<F16>::kSmem, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fumi_cuda_fail(e, "gram_tc_kernel");
    return FUMI_OK;
}
int gram_dbg() {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("FUMI_GRAM_DBG"); dbg = e ? atoi(e) : 0; }
    return dbg;
}
}  // namespace

// Returns FUMI_OK when the launch was made, 1 when the shape is outside this kernel (caller falls back to the
// warp-level kernel of gram.cu), < 0 on error.
int fumi_gram_tc_launch(const float* feats, int64_t D, const int64_t* sup_rows, const int64_t* qry_rows, int64_t B,
                        int32_t NK, int32_t NQ, float* gram, void* stream) {
    if (NK > 32 || NK + NQ > kRows || D % 64 != 0) return 1;
    GramTcParams p{feats, nullptr, nullptr, D, sup_rows, qry_rows, B, NK, NQ, gram, gram_dbg()};
    return launch_gram_tc<false>(p, stream);
}

// fp16 (hi, lo) bank planes (fumi_split_f16): half the MMAs per feature (K = 16), no split arithmetic in the kernel
extern "C" int fumi_gram_f16(const void* feats_hi, const void* feats_lo, const float* absmax, int64_t num_rows, int64_t D,
                             const int64_t* sup_rows, const int64_t* qry_rows, int64_t B, int32_t NK, int32_t NQ,
                             float* gram, void* stream) {
    FUMI_CHECK_ARG(B >= 0 && NK >= 1 && NK <= 32 && NQ >= 0 && NK + NQ <= kRows && num_rows >= 1,
                   "fp16-plane Gram kernel: NK <= 32 and NK + NQ <= 192");
    FUMI_CHECK_ARG(D >= 64 && D % 64 == 0, "feature dim must be a multiple of 64");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(feats_hi && feats_lo && absmax && sup_rows && (qry_rows || NQ == 0) && gram, "null pointer");
    FUMI_CHECK_ARG((uintptr_t(feats_hi) | uintptr_t(feats_lo)) % 16 == 0, "planes must be 16-byte aligned");
    GramTcParams p{feats_hi, feats_lo, absmax, D, sup_rows, qry_rows, B, NK, NQ, gram, gram_dbg()};
    return launch_gram_tc<true>(p, stream);
}
