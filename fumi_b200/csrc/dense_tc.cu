// tcgen05 (5th-gen tensor core) dense contraction with fp32-level accuracy: "3xTF32".
//
//   C[M,N] (=|+=) act( A[M,K] . B[N,K]^T + bias[N] ),      A = A_hi + A_lo,  B = B_hi + B_lo
//   A.B^T ~= A_hi.B_hi^T + A_lo.B_hi^T + A_hi.B_lo^T       (dropped term ~2^-22 relative)
//
// where x_hi = tf32-rounded x (low 13 mantissa bits zero) and x_lo = x - x_hi are produced once by
// fumi_split_tf32 (static feature bank: at load time; weights: once per outer step).  This is the
// path of the real dense contractions of FuMI: the hypernetwork layers (fumi.py:70-107), the first
// image layer applied to the whole split bank once per outer step (fumi.py:215, hoisted; see
// DESIGN.md) and its weight gradient dW0 = d_proj^T X (fumi.py:192), which runs as the same
// "both operands K-major" kernel on a pre-transposed copy of the bank with split-K.
//
// Structure (one 128x256 output tile per CTA, 320 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2D, SWIZZLE_128B boxes [32 fp32 x rows] of the four
//              operand planes into a 2-stage shared-memory ring, mbarrier complete_tx
//   warp 1   : MMA issuer    -- one elected thread issues tcgen05.mma.kind::tf32 (M128 N256 K8): per 32-wide
//              k stage the 8 small cross products (lo.hi, hi.lo) first, then the 4 hi.hi products, into a
//              FRESH 256-column TMEM accumulator (two ping-pong buffers = all 512 TMEM columns);
//              tcgen05.commit releases the smem stage and publishes the TMEM buffer
//   warps 2-9: accumulate + epilogue -- every stage: tcgen05.ld 32x32b.x32 TMEM -> registers and add into
//              128 fp32 register accumulators per thread with round-to-nearest; at the end bias +
//              activation and float4 stores (split-K: the split's partial tile goes to a workspace that
//              splitk_reduce_kernel sums in split order -- no floating-point atomics, run-to-run reproducible).
// Why two levels: the tensor core adds into its accumulator with truncation, a bias that grows linearly
// with the number of accumulated MMAs (measured 5.6e-6 relative at K=768 in a single accumulator, 10x
// plain fp32); restarting the accumulator every stage keeps <= 4 significant truncations per partial sum.
// Out-of-range rows / k are zero-filled by TMA, so M, N, K need no padding (leading dims: multiple of 4).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/fumi_b200.h"
#include "common.cuh"

namespace {

// Stage rows are 64 bytes (SWIZZLE_64B): 16 tf32 / 32 fp16 k per stage, 48 KB per stage, FOUR stages.  Both bank-sized
// contractions are bound by operand arrival, not by MMA rate: with two 96 KB stages of 128-byte rows a CTA had one
// stage in flight and measured ~4.5k cycles per stage against 1.5k cycles of MMAs; three 48 KB stages in flight keep
// 1.5x the bytes outstanding at half the granularity.
constexpr int BM = 128, BN = 256, BK = 16;          // BK fp32 = 64 bytes = one swizzle row
constexpr int kRowBytes = BK * 4;
constexpr int kStages = 4;
constexpr int kThreadsTc = 320;
constexpr uint32_t kABytes = BM * kRowBytes;        // 8 KB per plane
constexpr uint32_t kBBytes = BN * kRowBytes;        // 16 KB per plane
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // 48 KB
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct TcParams {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;      // b_*: box of BN rows, or of BN / 2 rows when mcast
    int mcast;              // 1: CTA pairs (cluster 2x1x1 along M) share every B stage: each CTA loads half of it and
                            //    multicasts it into both CTAs' shared memory (half the L2 reads of the re-streamed operand)
    float* C;
    const float* bias;
    int64_t M, N, ldc;
    int m_tiles;            // M tiles (even when mcast); a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (persistent:
                            // the operand ring and the TMEM ping-pong run on across tiles, so prologue / epilogue overlap)
    int k_tiles;            // total BK tiles along K
    int k_tiles_per_split;
    int act;
    int atomic;             // 0: store (bias / act fused); 1: C += tile (single writer); 2: split-K partial tile -> parts
    float* parts;           // split-K workspace [splits][M][N] (dense), summed in split order by splitk_reduce_kernel
    int drain;              // k stages accumulated in one TMEM buffer before it is drained into registers
    const float* a_absmax;  // f16x3 only: device scalars max|A|, max|B| that fixed the power-of-two plane scales
    const float* b_absmax;
};

// f16x3 planes hold x * s with s = 2^(14 - e), 2^(e-1) <= max|x| < 2^e: the largest element lands in [2^13, 2^14),
// the fp16 "lo" plane of typical elements stays normal, and the scale is undone exactly in the epilogue.
__host__ __device__ __forceinline__ int f16_scale_exp(float amax) {
    if (!(amax > 0.f) || !(amax < 3.0e38f)) return 0;
    int e;
    frexpf(amax, &e);
    int k = 14 - e;
    return k > 100 ? 100 : (k < -100 ? -100 : k);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64): SWIZZLE_128B = 2, SWIZZLE_64B = 4
// SBO = bytes between 8-row groups = 8 x 64
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t((8 * kRowBytes) >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(4) << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask` (a stage is free when BOTH CTAs of a pair have read it)
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
// (not inlined: 128 expansions of tanhf / expf per thread made the kernel 300 KB of SASS, and the once-per-tile epilogue
// then ran out of the instruction cache -- 38 % of all stall samples were no_instruction)
__device__ __noinline__ float act_f(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return tanhf(v);
    if (act == 3) return 1.f / (1.f + expf(-v));
    return v;
}

// F16 = false: tf32 planes (fp32 containers, 32 k per 128-byte stage row, K = 8 per MMA);
// F16 = true : fp16 planes (64 k per 128-byte stage row, K = 16 per MMA at twice the rate, half the bytes).
// Stage geometry in bytes, descriptors and the k-step advance (32 B) are identical.
template <bool F16>
__global__ void __launch_bounds__(kThreadsTc, 1) gemm_x3_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;          // SWIZZLE_128B atoms: 1024 B aligned
    const uint32_t bars = base + kStages * kStageBytes;     // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
    const uint32_t full_bar = bars, empty_bar = bars + 8 * kStages, tmem_full_bar = bars + 16 * kStages;
    const uint32_t tmem_empty_bar = tmem_full_bar + 16;
    const uint32_t tmem_slot = tmem_empty_bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * BN;
    const int kt0 = blockIdx.z * p.k_tiles_per_split;
    const int kt1 = min(p.k_tiles, kt0 + p.k_tiles_per_split);
    const int nk = kt1 - kt0;

    const uint32_t crank = p.mcast ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, p.mcast ? 2 : 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar + 8 * b, 1); mbar_init(tmem_empty_bar + 8 * b, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {       // TMEM: two 256-column x 128-lane fp32 accumulators (all 512 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (p.mcast) cluster_sync_all();          // the peer's barriers are initialised before anything is multicast into them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && nk > 0) {
            int git = 0;                                   // stage counter across this CTA's tiles
            for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
            const int m0 = mt * BM;
            for (int it = 0; it < nk; ++it, ++git) {
                const int s = git % kStages;
                const uint32_t ph = (git / kStages) & 1;
                mbar_wait(empty_bar + 8 * s, ph ^ 1);
                const uint32_t st = base + s * kStageBytes;
                mbar_expect_tx(full_bar + 8 * s, kStageBytes);
                const int kc = (kt0 + it) * (F16 ? 2 * BK : BK);              // element coordinate of the stage
                tma_load_2d(st, &p.a_hi, full_bar + 8 * s, kc, m0);
                tma_load_2d(st + kABytes, &p.a_lo, full_bar + 8 * s, kc, m0);
                if (p.mcast) {                 // my half of the B rows, into both CTAs of the pair
                    const uint32_t off = crank * (kBBytes / 2);
                    const int nr = n0 + int(crank) * (BN / 2);
                    tma_load_2d_mcast(st + 2 * kABytes + off, &p.b_hi, full_bar + 8 * s, kc, nr, uint16_t(3));
                    tma_load_2d_mcast(st + 2 * kABytes + kBBytes + off, &p.b_lo, full_bar + 8 * s, kc, nr, uint16_t(3));
                } else {
                    tma_load_2d(st + 2 * kABytes, &p.b_hi, full_bar + 8 * s, kc, n0);
                    tma_load_2d(st + 2 * kABytes + kBBytes, &p.b_lo, full_bar + 8 * s, kc, n0);
                }
            }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && nk > 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, K-major both, N>>3, M>>4
            // (kind::f16: A = B = f16 is format 0)
            const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (uint32_t(BN >> 3) << 17) |
                                   (uint32_t(BM >> 4) << 24);
            auto mma = [&](uint32_t td, uint64_t da, uint64_t db, uint32_t accum) {
                if (F16) umma_f16(td, da, db, idesc, accum); else umma_tf32(td, da, db, idesc, accum);
            };
            int git = 0, gg = 0;                          // stage / TMEM-group counters across this CTA's tiles
            const int ngroups_t = (nk + p.drain - 1) / p.drain;
            for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, gg += ngroups_t)
            for (int it = 0; it < nk; ++it, ++git) {
                const int s = git % kStages;
                const uint32_t ph = (git / kStages) & 1;
                const int grp_t = it / p.drain, sub = it - grp_t * p.drain;
                const int grp = gg + grp_t;
                const int tb = grp & 1;                                  // TMEM ping-pong buffer of this group of stages
                if (sub == 0) mbar_wait(tmem_empty_bar + 8 * tb, ((grp >> 1) & 1) ^ 1);   // accumulate warps drained it
                mbar_wait(full_bar + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * kStageBytes;
                const uint32_t td = tmem_base + uint32_t(tb * BN);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {                       // small cross products first
                    const uint64_t ahi = umma_desc_sw128(st + k * 32), alo = umma_desc_sw128(st + kABytes + k * 32);
                    const uint64_t bhi = umma_desc_sw128(st + 2 * kABytes + k * 32);
                    const uint64_t blo = umma_desc_sw128(st + 2 * kABytes + kBBytes + k * 32);
                    mma(td, alo, bhi, (k | sub) != 0);
                    mma(td, ahi, blo, 1);
                }
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint64_t ahi = umma_desc_sw128(st + k * 32);
                    const uint64_t bhi = umma_desc_sw128(st + 2 * kABytes + k * 32);
                    mma(td, ahi, bhi, 1);
                }
                if (p.mcast) umma_commit_mcast(empty_bar + 8 * s, uint16_t(3));
                else umma_commit(empty_bar + 8 * s);   // frees the smem stage once these MMAs retire
                if (sub == p.drain - 1 || it == nk - 1) umma_commit(tmem_full_bar + 8 * tb);   // partial sums complete
            }
        }
    } else {
        // ===== accumulate + epilogue: warps 2..9; lane quarter = warp % 4, column half = (warp - 2) / 4 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int ngroups = (nk + p.drain - 1) / p.drain;
        int gg = 0;
        for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, gg += ngroups) {
        const int row = mt * BM + q * 32 + lane;
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.f;
        for (int git = gg; git < gg + ngroups; ++git) {
            const int tb = git & 1;
            mbar_wait(tmem_full_bar + 8 * tb, (git >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(tb * BN + half * 128 + c * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                    "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar + 8 * tb) : "memory");
        }
        if (F16) {                                    // undo the plane scales (exact powers of two)
            const float ia = ldexpf(1.f, -f16_scale_exp(p.a_absmax ? *p.a_absmax : 0.f));
            const float ib = ldexpf(1.f, -f16_scale_exp(p.b_absmax ? *p.b_absmax : 0.f));
#pragma unroll
            for (int j = 0; j < 128; ++j) acc[j] = (acc[j] * ia) * ib;
        }
        if ((nk > 0 || p.atomic == 2) && row < p.M) {       // (an empty split still owns -- and zeroes -- its partial tile)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col0 = n0 + half * 128 + c * 32;
                if (col0 >= p.N) break;
                float* crow = p.atomic == 2 ? p.parts + (int64_t(blockIdx.z) * p.M + row) * p.N + col0
                                            : p.C + int64_t(row) * p.ldc + col0;
                const int ncols = min(32, int(p.N) - col0);
                if (p.atomic == 2) {                  // deterministic split-K: every split owns its partial tile
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) crow[j] = acc[c * 32 + j];
                } else if (p.atomic == 1) {           // accumulate, one writer per element
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) crow[j] += acc[c * 32 + j];
                } else if (ncols == 32 && (p.ldc & 3) == 0 && p.bias == nullptr && p.act == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)       // the bank-sized contractions: plain stores
                        *reinterpret_cast<float4*>(crow + j) = make_float4(acc[c * 32 + j], acc[c * 32 + j + 1], acc[c * 32 + j + 2], acc[c * 32 + j + 3]);
                } else if (ncols == 32 && (p.ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o;
                        o.x = act_f(acc[c * 32 + j] + (p.bias ? p.bias[col0 + j] : 0.f), p.act);
                        o.y = act_f(acc[c * 32 + j + 1] + (p.bias ? p.bias[col0 + j + 1] : 0.f), p.act);
                        o.z = act_f(acc[c * 32 + j + 2] + (p.bias ? p.bias[col0 + j + 2] : 0.f), p.act);
                        o.w = act_f(acc[c * 32 + j + 3] + (p.bias ? p.bias[col0 + j + 3] : 0.f), p.act);
                        *reinterpret_cast<float4*>(crow + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) crow[j] = act_f(acc[c * 32 + j] + (p.bias ? p.bias[col0 + j] : 0.f), p.act);
                }
            }
        }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (p.mcast) cluster_sync_all();          // no CTA leaves while its peer may still arrive on its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

// C[m][n] (=|+=) sum over splits of parts[s][m][n], in split order (deterministic: no floating-point atomics)
__global__ void splitk_reduce_kernel(const float* __restrict__ parts, int splits, int64_t M, int64_t N, float* __restrict__ C,
                                     int64_t ldc, int accumulate) {
    const int64_t total = M * N;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        float v = 0.f;
        for (int sidx = 0; sidx < splits; ++sidx) v += parts[int64_t(sidx) * total + i];
        const int64_t m = i / N, n = i - m * N;
        float* dst = C + m * ldc + n;
        *dst = accumulate ? *dst + v : v;
    }
}

// x -> hi (tf32 round-to-nearest, stored as fp32 with 13 zero low mantissa bits) and lo = x - hi
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const float v = x[i];
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hf = __uint_as_float(h);
        hi[i] = hf;
        lo[i] = v - hf;
    }
}

// x[R,Ccols] -> hiT / loT [Ccols, ldt] (transposed, split), 32x32 tiles through shared memory
__global__ void transpose_split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hiT, float* __restrict__ loT,
                                            int64_t R, int64_t Ccols, int64_t ldt) {
    __shared__ float tile[32][33];
    const int64_t r0 = int64_t(blockIdx.x) * 32, c0 = int64_t(blockIdx.y) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < Ccols) ? x[r * Ccols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < Ccols && r < ldt) {
            const float v = r < R ? tile[threadIdx.x][i] : 0.f;
            uint32_t h;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
            const float hf = __uint_as_float(h);
            hiT[c * ldt + r] = hf;
            loT[c * ldt + r] = v - hf;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2D row-major [rows, cols] with leading dimension ld (elements); box = [64 bytes of columns x box_rows]
int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f16) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { fumi_set_error("cuTensorMapEncodeTiled is not available from the driver"); return FUMI_ERR_CUDA; }
    const int esz = f16 ? 2 : 4;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * esz};
    cuuint32_t box[2] = {cuuint32_t(kRowBytes / esz), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fumi_set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
        return FUMI_ERR_CUDA;
    }
    return FUMI_OK;
}

// max |x| over a buffer, as the bit pattern of a non-negative float (ordered like an unsigned integer)
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// x -> fp16 planes of x * 2^k (k from the device scalar max|x|): hi = fp16(xs), lo = fp16(xs - hi)
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n,
                                 const float* __restrict__ absmax) {
    const float sc = ldexpf(1.f, f16_scale_exp(*absmax));
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        split_f16(x[i] * sc, hi[i], lo[i]);
}

// x[R,Ccols] -> hiT / loT [Ccols, ldt] fp16 (transposed, scaled, split), 32x32 tiles through shared memory
__global__ void transpose_split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hiT, __half* __restrict__ loT,
                                           int64_t R, int64_t Ccols, int64_t ldt, const float* __restrict__ absmax) {
    __shared__ float tile[32][33];
    const float sc = ldexpf(1.f, f16_scale_exp(*absmax));
    const int64_t r0 = int64_t(blockIdx.x) * 32, c0 = int64_t(blockIdx.y) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < Ccols) ? x[r * Ccols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < Ccols && r < ldt) split_f16(r < R ? tile[threadIdx.x][i] * sc : 0.f, hiT[c * ldt + r], loT[c * ldt + r]);
    }
}

}  // namespace

extern "C" int fumi_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
    FUMI_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return FUMI_OK;
    FUMI_CHECK_ARG(x && hi && lo, "null pointer");
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
    FUMI_CHECK_LAUNCH("split_tf32_kernel");
    return FUMI_OK;
}

extern "C" int fumi_transpose_split_tf32(const float* x, float* hiT, float* loT, int64_t R, int64_t C, int64_t ldt,
                                         void* stream) {
    FUMI_CHECK_ARG(R >= 1 && C >= 1 && ldt >= R && (ldt & 3) == 0, "need ldt >= R and ldt % 4 == 0");
    FUMI_CHECK_ARG(x && hiT && loT, "null pointer");
    dim3 grid((unsigned)((ldt + 31) / 32), (unsigned)((C + 31) / 32));
    FUMI_CHECK_ARG(grid.y <= 65535, "too many columns");
    transpose_split_tf32_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, hiT, loT, R, C, ldt);
    FUMI_CHECK_LAUNCH("transpose_split_tf32_kernel");
    return FUMI_OK;
}

namespace {
int launch_gemm_x3(bool f16, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, const float* a_absmax,
                   const float* b_absmax, const float* bias, float* c, int64_t M, int64_t N, int64_t K, int64_t lda,
                   int64_t ldb, int64_t ldc, int32_t act, int32_t accumulate, int32_t split_k, void* stream) {
    FUMI_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "bad shape");
    FUMI_CHECK_ARG(a_hi && a_lo && b_hi && b_lo && c, "null pointer");
    const int al = f16 ? 8 : 4;            // elements per 16 bytes
    FUMI_CHECK_ARG(lda >= K && ldb >= K && ldc >= N && lda % al == 0 && ldb % al == 0,
                   "leading dimensions must cover K / N and be multiples of 16 bytes (TMA strides)");
    FUMI_CHECK_ARG(act >= 0 && act <= 3, "act must be 0..3");
    FUMI_CHECK_ARG((uintptr_t(a_hi) | uintptr_t(a_lo) | uintptr_t(b_hi) | uintptr_t(b_lo)) % 16 == 0,
                   "operand planes must be 16-byte aligned");
    TcParams p;
    std::memset(&p, 0, sizeof(p));
    int rc;
    if ((rc = make_map(&p.a_hi, a_hi, M, K, lda, BM, f16)) != FUMI_OK) return rc;
    if ((rc = make_map(&p.a_lo, a_lo, M, K, lda, BM, f16)) != FUMI_OK) return rc;
    // CTA pairs along M share the B stages (TMA multicast) whenever there are at least two M tiles
    static int mcast_on = -1;
    if (mcast_on < 0) { const char* e = getenv("FUMI_GEMM_MCAST"); mcast_on = (e && atoi(e) == 0) ? 0 : 1; }
    p.mcast = (mcast_on && M > BM) ? 1 : 0;
    if ((rc = make_map(&p.b_hi, b_hi, N, K, ldb, p.mcast ? BN / 2 : BN, f16)) != FUMI_OK) return rc;
    if ((rc = make_map(&p.b_lo, b_lo, N, K, ldb, p.mcast ? BN / 2 : BN, f16)) != FUMI_OK) return rc;
    p.C = c; p.bias = bias; p.M = M; p.N = N; p.ldc = ldc; p.act = act;
    p.a_absmax = a_absmax; p.b_absmax = b_absmax;
    {   // stages per TMEM drain: 1 = best accuracy (<= 4 truncating adds per partial sum), more = fewer drains
        static int drain = -1;
        if (drain < 0) {
            const char* e = getenv("FUMI_GEMM_DRAIN");
            drain = e ? atoi(e) : 2;
            if (drain < 1) drain = 1;
            if (drain > 8) drain = 8;
        }
        // `drain` counts 32-k tf32 slices; a stage is 16 tf32 k or 32 fp16 k
        p.drain = f16 ? 2 * ((drain + 1) / 2) : 2 * drain;
    }
    const int kstage = f16 ? 2 * BK : BK;
    p.k_tiles = int((K + kstage - 1) / kstage);
    const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int splits = split_k;
    if (splits <= 0 && (bias != nullptr || act != 0)) splits = 1;     // fused epilogue needs whole-K tiles
    if (splits <= 0) {                                   // auto: fill the SMs when there are few output tiles
        int sms = fumi_device_sm_count();
        if (sms <= 0) return sms;
        splits = tiles >= sms ? 1 : int(sms / tiles);    // one wave: tiles * splits <= #SMs (one CTA per SM)
        if (splits > p.k_tiles / 4) splits = p.k_tiles / 4 > 0 ? p.k_tiles / 4 : 1;
    }
    if (splits > p.k_tiles) splits = p.k_tiles;
    p.k_tiles_per_split = (p.k_tiles + splits - 1) / splits;
    splits = (p.k_tiles + p.k_tiles_per_split - 1) / p.k_tiles_per_split;
    p.atomic = splits > 1 ? 2 : (accumulate ? 1 : 0);
    if (p.atomic) FUMI_CHECK_ARG(bias == nullptr && act == 0, "bias / activation cannot be fused into a split-K or accumulating GEMM");
    if (p.atomic == 2) {                                 // workspace of the partial tiles (grown on demand, kept per process)
        static float* ws = nullptr;
        static size_t ws_bytes = 0;
        const size_t need = size_t(splits) * size_t(M) * size_t(N) * sizeof(float);
        if (need > ws_bytes) {
            if (ws) cudaFree(ws);
            ws = nullptr; ws_bytes = 0;
            cudaError_t e = cudaMalloc(&ws, need);
            if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaMalloc(split-K workspace)");
            ws_bytes = need;
        }
        p.parts = ws;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
        if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaFuncSetAttribute(gemm_x3)");
        attr_done = true;
    }
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)splits);
    FUMI_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "grid too large");
    if (p.mcast) grid.x = (grid.x + 1) & ~1u;  // whole pairs: an odd last M tile gets an idle partner (all its rows out of range)
    p.m_tiles = int(grid.x);
    {   // persistent over M: at most one CTA per SM; a CTA walks M tiles with its ring / TMEM ping-pong running on
        const int sms = fumi_device_sm_count();
        if (sms <= 0) return sms;
        unsigned budget = unsigned(sms) / (grid.y * grid.z);
        if (p.mcast) budget &= ~1u;
        if (budget >= (p.mcast ? 2u : 1u) && grid.x > budget) grid.x = budget;
    }
    if (p.mcast) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(kThreadsTc); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        cudaError_t e = f16 ? cudaLaunchKernelEx(&cfg, gemm_x3_kernel<true>, p) : cudaLaunchKernelEx(&cfg, gemm_x3_kernel<false>, p);
        if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaLaunchKernelEx(gemm_x3_kernel, cluster 2x1x1)");
    } else if (f16) gemm_x3_kernel<true><<<grid, kThreadsTc, kSmemBytes, (cudaStream_t)stream>>>(p);
    else gemm_x3_kernel<false><<<grid, kThreadsTc, kSmemBytes, (cudaStream_t)stream>>>(p);
    FUMI_CHECK_LAUNCH("gemm_x3_kernel");
    if (p.atomic == 2) {
        const int64_t total = M * N;
        const unsigned rgrid = unsigned((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        splitk_reduce_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(p.parts, splits, M, N, c, ldc, accumulate);
        FUMI_CHECK_LAUNCH("splitk_reduce_kernel");
    }
    return FUMI_OK;
}
}  // namespace

extern "C" int fumi_gemm_tf32x3(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo,
                                const float* bias, float* c, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                                int64_t ldc, int32_t act, int32_t accumulate, int32_t split_k, void* stream) {
    return launch_gemm_x3(false, a_hi, a_lo, b_hi, b_lo, nullptr, nullptr, bias, c, M, N, K, lda, ldb, ldc, act, accumulate,
                          split_k, stream);
}

extern "C" int fumi_gemm_f16x3(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                               const float* a_absmax, const float* b_absmax, const float* bias, float* c, int64_t M,
                               int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int32_t act,
                               int32_t accumulate, int32_t split_k, void* stream) {
    FUMI_CHECK_ARG(a_absmax && b_absmax, "null scale pointer");
    return launch_gemm_x3(true, a_hi, a_lo, b_hi, b_lo, a_absmax, b_absmax, bias, c, M, N, K, lda, ldb, ldc, act, accumulate,
                          split_k, stream);
}

extern "C" int fumi_absmax(const float* x, int64_t n, float* out, void* stream) {
    FUMI_CHECK_ARG(n >= 0 && out && (x || n == 0), "bad argument");
    cudaError_t e = cudaMemsetAsync(out, 0, 4, (cudaStream_t)stream);
    if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaMemsetAsync");
    if (n == 0) return FUMI_OK;
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    absmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, reinterpret_cast<unsigned int*>(out));
    FUMI_CHECK_LAUNCH("absmax_kernel");
    return FUMI_OK;
}

extern "C" int fumi_split_f16(const float* x, const float* absmax, void* hi, void* lo, int64_t n, void* stream) {
    FUMI_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return FUMI_OK;
    FUMI_CHECK_ARG(x && absmax && hi && lo, "null pointer");
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, static_cast<__half*>(hi), static_cast<__half*>(lo), n,
                                                                         absmax);
    FUMI_CHECK_LAUNCH("split_f16_kernel");
    return FUMI_OK;
}

extern "C" int fumi_transpose_split_f16(const float* x, const float* absmax, void* hiT, void* loT, int64_t R, int64_t C,
                                        int64_t ldt, void* stream) {
    FUMI_CHECK_ARG(R >= 1 && C >= 1 && ldt >= R && (ldt & 7) == 0, "need ldt >= R and ldt % 8 == 0");
    FUMI_CHECK_ARG(x && absmax && hiT && loT, "null pointer");
    dim3 grid((unsigned)((ldt + 31) / 32), (unsigned)((C + 31) / 32));
    FUMI_CHECK_ARG(grid.y <= 65535, "too many columns");
    transpose_split_f16_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, static_cast<__half*>(hiT),
                                                                               static_cast<__half*>(loT), R, C, ldt, absmax);
    FUMI_CHECK_LAUNCH("transpose_split_f16_kernel");
    return FUMI_OK;
}

// precision=1 entry points of fumi_linear_fwd / fumi_linear_wgrad take *unsplit* operands; the engine uses the
// prepared-plane API above instead, so these only exist to fail loudly if somebody asks for them.
int fumi_linear_fwd_tc(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*) {
    fumi_set_error("precision=1 runs on pre-split operand planes: use fumi_split_tf32 + fumi_gemm_tf32x3");
    return FUMI_ERR_UNSUPPORTED;
}
int fumi_linear_wgrad_tc(const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*) {
    fumi_set_error("precision=1 runs on pre-split operand planes: use fumi_transpose_split_tf32 + fumi_gemm_tf32x3");
    return FUMI_ERR_UNSUPPORTED;
}
