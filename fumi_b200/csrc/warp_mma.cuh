// Warp-level tensor-core building block for the per-task contractions of the episode kernels.
//
// The per-task matrices are small (NK = 25 rows at 5-way 5-shot): a tcgen05 tile (M >= 64 per CTA, operands
// staged as 4-byte TF32 planes) would idle most of the array and does not fit next to the task state in
// shared memory, so these contractions use warp-synchronous mma.sync.m16n8k8 TF32 on fragments read straight
// from the fp32 tiles in shared memory, with the same fp32-accurate 3-pass split as dense_tc.cu:
//      x = hi + lo,  hi = tf32(x);      a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
// Measured on B200 (tools/mma_bench.cu): mma.sync TF32 478 MAC/clk/SM vs FFMA 124, so 3 passes are ~1.3x the
// FFMA *peak* while needing ~10x fewer issue slots and half the shared-memory wavefronts of the FFMA loops
// (which ran LDS-bound at 16-19 % of the FMA pipe).
// The tensor core adds into its accumulator with truncation; to keep fp32-grade results every 32-wide slice
// of K starts from a zero accumulator (cross terms first, then hi*hi) and is added to the running sum with
// an ordinary round-to-nearest FADD.
#pragma once

#include <cstdint>

#ifdef FUMI_EMU
#include "cuda_emu.h"
static inline void fumi_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const fumi_half ah = fumi_f2h(a), bh = fumi_f2h(b);
    const fumi_half al = fumi_f2h(a - fumi_h2f(ah)), bl = fumi_f2h(b - fumi_h2f(bh));
    hi = uint32_t(ah.bits) | (uint32_t(bh.bits) << 16);
    lo = uint32_t(al.bits) | (uint32_t(bl.bits) << 16);
}
static inline void fumi_join2(uint32_t hi, uint32_t lo, float& a, float& b) {
    a = fumi_h2f(fumi_half{uint16_t(hi & 0xFFFFu)}) + fumi_h2f(fumi_half{uint16_t(lo & 0xFFFFu)});
    b = fumi_h2f(fumi_half{uint16_t(hi >> 16)}) + fumi_h2f(fumi_half{uint16_t(lo >> 16)});
}
static inline float fumi_fast_exp(float x) { return std::exp(x); }
static inline float fumi_fast_log(float x) { return std::log(x); }
static inline float fumi_fast_rcp(float x) { return 1.f / x; }
static inline void fumi_cp_async16(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 16); }
static inline void fumi_cp_async_wait() {}
static inline void fumi_cp_async_commit() {}
template <int N> static inline void fumi_cp_async_wait_n() {}
#else
#include <cuda_fp16.h>
// ---- fp16 plane primitives (see warp_gemm_f16x3 below) ------------------------------------------------------
typedef __half fumi_half;
__device__ __forceinline__ fumi_half fumi_f2h(float x) { return __float2half_rn(x); }
__device__ __forceinline__ float fumi_h2f(fumi_half h) { return __half2float(h); }
__device__ __forceinline__ void fumi_ldsm4(uint32_t (&r)[4], const fumi_half* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void fumi_ldsm4t(uint32_t (&r)[4], const fumi_half* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void fumi_ldsm2(uint32_t (&r)[2], const fumi_half* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void fumi_ldsm2t(uint32_t (&r)[2], const fumi_half* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void fumi_mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void fumi_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t fumi_tf32_hi(float x) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    return h;
}
// MUFU-based exp / log / reciprocal (softmax probabilities and the CE loss: ~1e-6 relative, far inside the 1e-4 bar;
// the argmax that decides the predictions is taken on the logits themselves)
__device__ __forceinline__ float fumi_fast_exp(float x) { return __expf(x); }
__device__ __forceinline__ float fumi_fast_log(float x) { return __logf(x); }
__device__ __forceinline__ float fumi_fast_rcp(float x) { return __frcp_rn(x); }
// 16-byte asynchronous global -> shared copy (LDGSTS) and the wait for all copies of this thread
__device__ __forceinline__ void fumi_cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void fumi_cp_async_wait() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fumi_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void fumi_cp_async_wait_n() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// (a, b) -> packed fp16 pairs hi = (fp16(a), fp16(b)) and lo = (fp16(a - hi.x), fp16(b - hi.y)): one F2FP per pair
__device__ __forceinline__ void fumi_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// packed (hi, lo) pairs -> (a, b) = hi + lo  (unscaled)
__device__ __forceinline__ void fumi_join2(uint32_t hi, uint32_t lo, float& a, float& b) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    a = hf.x + lf.x;
    b = hf.y + lf.y;
}
#endif

// hi = x with the 13 low mantissa bits cleared (one LOP3; cvt.rna.tf32 expands to ~8 instructions on sm_100),
// lo = x - hi exactly (<= 13 significant bits, of which the tensor core keeps 11: error 2^-21 relative to x).
__device__ __forceinline__ void fumi_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// acc[mt][nt][0..3] (+)= A[16*MT x K] . B[K x 8*NT] for one warp.
//   element A(m,k) = A[m*lda + k]  (ATRANS: A[k*lda + m]);   element B(k,n) = B[k*ldb + n]  (BTRANS: B[n*ldb + k])
//   bscale multiplies B on load (folds -alpha into an operand).  K is a multiple of 8; rows/cols beyond the
//   logical extent must hold zeros in shared memory.
// Fragment ownership (PTX m16n8k8.tf32): g = lane/4, t = lane%4
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);  b0 (k=t, n=g)  b1 (k=t+4, n=g)
//   c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
template <int MT, int NT, bool ATRANS, bool BTRANS, int ONEPASS = -1>
__device__ __forceinline__ void warp_gemm_3xtf32(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                 int K, float bscale, float (&acc)[MT][NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // Small tile counts: one pass over the fragments with two accumulators per tile (cross terms / hi*hi):
    // half the shared-memory reads and splits, and two independent MMA chains per tile.  Large tile counts
    // (register budget): two passes over the 32-wide slice into one accumulator.  ONEPASS = 0/1 forces a mode.
    constexpr bool kOnePass = ONEPASS < 0 ? (MT * NT <= 8) : (ONEPASS != 0);
    for (int k0 = 0; k0 < K; k0 += 32) {
        float part[MT][NT][4];
        float cross[kOnePass ? MT : 1][kOnePass ? NT : 1][4];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    part[i][j][q] = 0.f;
                    if (kOnePass) cross[i][j][q] = 0.f;
                }
        const int kend = (K - k0) < 32 ? (K - k0) : 32;
#pragma unroll
        for (int pass = 0; pass < (kOnePass ? 1 : 2); ++pass) {
#pragma unroll
            for (int kk = 0; kk < 32; kk += 8) {               // fixed trip count: the loads of the next k step overlap the MMAs
                if (kk >= kend) break;
                const int k = k0 + kk;
                uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    const int m = i * 16 + g;
                    float v0, v1, v2, v3;
                    if (ATRANS) {
                        v0 = A[(k + t) * lda + m]; v1 = A[(k + t) * lda + m + 8];
                        v2 = A[(k + t + 4) * lda + m]; v3 = A[(k + t + 4) * lda + m + 8];
                    } else {
                        v0 = A[m * lda + k + t]; v1 = A[(m + 8) * lda + k + t];
                        v2 = A[m * lda + k + t + 4]; v3 = A[(m + 8) * lda + k + t + 4];
                    }
                    fumi_split(v0, ahi[i][0], alo[i][0]);
                    fumi_split(v1, ahi[i][1], alo[i][1]);
                    fumi_split(v2, ahi[i][2], alo[i][2]);
                    fumi_split(v3, ahi[i][3], alo[i][3]);
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int n = j * 8 + g;
                    float w0, w1;
                    if (BTRANS) { w0 = B[n * ldb + k + t]; w1 = B[n * ldb + k + t + 4]; }
                    else        { w0 = B[(k + t) * ldb + n]; w1 = B[(k + t + 4) * ldb + n]; }
                    uint32_t bhi[2], blo[2];
                    fumi_split(w0 * bscale, bhi[0], blo[0]);
                    fumi_split(w1 * bscale, bhi[1], blo[1]);
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        if (kOnePass) {
                            fumi_mma_tf32(cross[i][j], alo[i], bhi);
                            fumi_mma_tf32(part[i][j], ahi[i], bhi);
                            fumi_mma_tf32(cross[i][j], ahi[i], blo);
                        } else if (pass == 0) {                    // cross terms first, then hi*hi
                            fumi_mma_tf32(part[i][j], alo[i], bhi);
                            fumi_mma_tf32(part[i][j], ahi[i], blo);
                        } else {
                            fumi_mma_tf32(part[i][j], ahi[i], bhi);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    acc[i][j][q] += kOnePass ? (part[i][j][q] + cross[i][j][q]) : part[i][j][q];
    }
}

// Visit the accumulator elements this lane owns: f(m, n, value&)
template <int MT, int NT, typename F>
__device__ __forceinline__ void warp_tile_foreach(float (&acc)[MT][NT][4], F f) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            f(i * 16 + g, j * 8 + 2 * t, acc[i][j][0]);
            f(i * 16 + g, j * 8 + 2 * t + 1, acc[i][j][1]);
            f(i * 16 + g + 8, j * 8 + 2 * t, acc[i][j][2]);
            f(i * 16 + g + 8, j * 8 + 2 * t + 1, acc[i][j][3]);
        }
}

// ================================================================================================================
// fp16 hi/lo PLANES: the operand matrices live in shared memory already split, x * 2^s = hi + lo with
// hi = fp16(x 2^s), lo = fp16(x 2^s - hi) (22 significant bits; same bytes as one fp32 tile), so the inner loop is
// ldmatrix + mma.m16n8k16 only: 3 MMAs per 16 k and no arithmetic, against 2 x (loads + LOP3/FADD splits + 3 MMAs)
// on fp32 tiles.  Measured (tools/warp_gemm_bench.cu, 16 warps, 32x64x256): 2,231 vs 5,782 cycles per GEMM per SM,
// 96 % of the fp16 tensor-pipe rate, relative error 2.3e-7 (tf32 split 3.9e-7, fp32 FMA chain 4.2e-7).
//   acc[mt][nt][0..3] += (A 2^sa) . (B 2^sb): the caller multiplies by 2^-(sa+sb).
//   A(m,k) = A[m*lda + k]  (ATRANS: A[k*lda + m]);   B(k,n) = B[k*ldb + n]  (BTRANS: B[n*ldb + k]);  lda / ldb in
//   halves, rows 16-byte aligned and 16 bytes apart mod 128 (ld = cols + 8 for cols % 64 == 0) -> conflict-free.
//   K % 16 == 0; the accumulator restarts every 64 k (the tensor core adds with truncation).
//   SINGLE: K <= 64 and acc holds zeros on entry (every warp-local product of the episode kernels: K = 16 / 32 / 64) --
//   the one 64-wide slice accumulates straight into acc, without the zeroed partial tile and the add that fold a
//   slice into a running sum (5-7 % of the episode kernels' instructions were those moves and adds).
template <int MT, int NT, bool ATRANS, bool BTRANS, bool SINGLE = false>
__device__ __forceinline__ void warp_gemm_f16x3(const fumi_half* Ahi, const fumi_half* Alo, int lda, const fumi_half* Bhi,
                                                const fumi_half* Blo, int ldb, int K, float (&acc)[MT][NT][4]) {
    static_assert(NT == 1 || NT % 2 == 0, "n tiles come in pairs (ldmatrix.x4) or alone");
    const int lane = threadIdx.x & 31;
    const int l7 = lane & 7, b3 = (lane >> 3) & 1, b4 = lane >> 4;
    // element offsets of this lane's ldmatrix row (relative to (m0 = 0 / n0 = 0, k = 0))
    const int aoff = ATRANS ? (l7 + 8 * b4) * lda + 8 * b3 : (l7 + 8 * b3) * lda + 8 * b4;
    const int boff = BTRANS ? (l7 + 8 * b4) * ldb + 8 * b3 : (l7 + 8 * b3) * ldb + 8 * b4;
    const int boff1 = BTRANS ? l7 * ldb + 8 * b3 : (l7 + 8 * b3) * ldb;           // single n tile (x2): lanes 0-15 count
    // few tiles per warp: the three products of a tile go to separate accumulators (three independent MMA chains
    // instead of one dependent chain of 3 K/16 MMAs -- the 64-wide layer has one tile per warp and was latency-bound)
    constexpr bool kSplitAcc = MT * NT <= 2;
#pragma unroll 1
    for (int k0 = 0; k0 < (SINGLE ? 1 : K); k0 += 64) {
        float part_[SINGLE ? 1 : MT][SINGLE ? 1 : NT][4];
        auto part = [&](int i, int j) -> float (&)[4] {
            if constexpr (SINGLE) return acc[i][j]; else return part_[i][j];
        };
        float c1[kSplitAcc ? MT : 1][kSplitAcc ? NT : 1][4], c2[kSplitAcc ? MT : 1][kSplitAcc ? NT : 1][4];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!SINGLE) part(i, j)[q] = 0.f;
                    if (kSplitAcc) { c1[i][j][q] = 0.f; c2[i][j][q] = 0.f; }
                }
#pragma unroll
        for (int kk = 0; kk < 64; kk += 16) {
            if (k0 + kk >= K) break;
            const int k = k0 + kk;
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int o = ATRANS ? k * lda + 16 * i + aoff : 16 * i * lda + k + aoff;
                if (ATRANS) { fumi_ldsm4t(ah[i], Ahi + o); fumi_ldsm4t(al[i], Alo + o); }
                else        { fumi_ldsm4(ah[i], Ahi + o);  fumi_ldsm4(al[i], Alo + o); }
            }
            if (NT == 1) {
                uint32_t bh[2], bl[2];
                const int o = BTRANS ? k + boff1 : k * ldb + boff1;
                if (BTRANS) { fumi_ldsm2(bh, Bhi + o);  fumi_ldsm2(bl, Blo + o); }
                else        { fumi_ldsm2t(bh, Bhi + o); fumi_ldsm2t(bl, Blo + o); }
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    fumi_mma_f16(kSplitAcc ? c1[i][0] : part(i, 0), al[i], bh[0], bh[1]);
                    fumi_mma_f16(kSplitAcc ? c2[i][0] : part(i, 0), ah[i], bl[0], bl[1]);
                    fumi_mma_f16(part(i, 0), ah[i], bh[0], bh[1]);
                }
            } else {
#pragma unroll
                for (int jp = 0; jp < NT / 2; ++jp) {
                    uint32_t bh[4], bl[4];           // {b0, b1} of n tile 2 jp, {b0, b1} of n tile 2 jp + 1
                    const int o = BTRANS ? 16 * jp * ldb + k + boff : k * ldb + 16 * jp + boff;
                    if (BTRANS) { fumi_ldsm4(bh, Bhi + o);  fumi_ldsm4(bl, Blo + o); }
                    else        { fumi_ldsm4t(bh, Bhi + o); fumi_ldsm4t(bl, Blo + o); }
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            fumi_mma_f16(kSplitAcc ? c1[i][2 * jp + h] : part(i, 2 * jp + h), al[i], bh[2 * h], bh[2 * h + 1]);
                            fumi_mma_f16(kSplitAcc ? c2[i][2 * jp + h] : part(i, 2 * jp + h), ah[i], bl[2 * h], bl[2 * h + 1]);
                            fumi_mma_f16(part(i, 2 * jp + h), ah[i], bh[2 * h], bh[2 * h + 1]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (SINGLE) { if (kSplitAcc) acc[i][j][q] += c1[i][j][q] + c2[i][j][q]; }
                    else acc[i][j][q] += kSplitAcc ? part(i, j)[q] + (c1[i][j][q] + c2[i][j][q]) : part(i, j)[q];
                }
    }
}

// ---- tile-outer building blocks: a warp holds the A fragments of its whole K range and walks the n tiles in pairs,
// so that the epilogue of one pair (ALU) is independent of -- and gets scheduled under -- the MMAs of the next pair.
// A fragments of a [16 x K] block, K = 16 KS.  ATRANS: A(m,k) = A[k*lda + m], else A[m*lda + k].
template <int KS, bool ATRANS>
__device__ __forceinline__ void warp_load_a(const fumi_half* Ahi, const fumi_half* Alo, int lda, uint32_t (&ah)[KS][4], uint32_t (&al)[KS][4]) {
    const int lane = threadIdx.x & 31;
    const int l7 = lane & 7, b3 = (lane >> 3) & 1, b4 = lane >> 4;
    const int aoff = ATRANS ? (l7 + 8 * b4) * lda + 8 * b3 : (l7 + 8 * b3) * lda + 8 * b4;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const int o = ATRANS ? 16 * k * lda + aoff : 16 * k + aoff;
        if (ATRANS) { fumi_ldsm4t(ah[k], Ahi + o); fumi_ldsm4t(al[k], Alo + o); }
        else        { fumi_ldsm4(ah[k], Ahi + o);  fumi_ldsm4(al[k], Alo + o); }
    }
}
// acc[h][0..3] = (A 2^sa) . (B 2^sb) for the n tile pair starting at column n0 (two 8-wide tiles), K = 16 KS <= 64 (one
// accumulator restart).  B(k,n) = B[k*ldb + n]  (BTRANS: B[n*ldb + k]).
template <int KS, bool BTRANS>
__device__ __forceinline__ void warp_mma_pair(const uint32_t (&ah)[KS][4], const uint32_t (&al)[KS][4], const fumi_half* Bhi,
                                              const fumi_half* Blo, int ldb, int n0, float (&acc)[2][4]) {
    const int lane = threadIdx.x & 31;
    const int l7 = lane & 7, b3 = (lane >> 3) & 1, b4 = lane >> 4;
    const int boff = BTRANS ? (l7 + 8 * b4) * ldb + 8 * b3 : (l7 + 8 * b3) * ldb + 8 * b4;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[h][q] = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        uint32_t bh[4], bl[4];
        const int o = BTRANS ? n0 * ldb + 16 * k + boff : 16 * k * ldb + n0 + boff;
        if (BTRANS) { fumi_ldsm4(bh, Bhi + o);  fumi_ldsm4(bl, Blo + o); }
        else        { fumi_ldsm4t(bh, Bhi + o); fumi_ldsm4t(bl, Blo + o); }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            fumi_mma_f16(acc[h], al[k], bh[2 * h], bh[2 * h + 1]);
            fumi_mma_f16(acc[h], ah[k], bl[2 * h], bl[2 * h + 1]);
            fumi_mma_f16(acc[h], ah[k], bh[2 * h], bh[2 * h + 1]);
        }
    }
}

// power-of-two plane scale from max |x| (held as the bit pattern of a non-negative float): the largest element lands
// in [2^13, 2^14).  Returns the exponent s (planes hold x 2^s).
__device__ __forceinline__ int fumi_plane_exp(uint32_t absmax_bits) {
    const int e = int(absmax_bits >> 23) - 126;              // |x| < 2^e  (0 and denormals: e <= -126)
    if (absmax_bits == 0u) return 0;
    const int s = 14 - e;
    return s > 100 ? 100 : (s < -100 ? -100 : s);
}
// same with the tracked max landing in [2^(target-1), 2^target): planes written with a LAGGED exponent (the one
// derived from the previous production of the same matrix) keep 16 - target binades of headroom before fp16 overflow
__device__ __forceinline__ int fumi_plane_exp_t(uint32_t absmax_bits, int target) {
    if (absmax_bits == 0u) return 0;
    const int s = target - (int(absmax_bits >> 23) - 126);
    return s > 100 ? 100 : (s < -100 ? -100 : s);
}
__device__ __forceinline__ float fumi_exp2i(int s) { return __uint_as_float(uint32_t(127 + s) << 23); }   // 2^s, |s| <= 126

// two adjacent plane elements with one 32-bit access each (off even, rows 16-byte aligned)
__device__ __forceinline__ void fumi_plane_load2(const fumi_half* hi, const fumi_half* lo, int off, float inv, float& a, float& b) {
    const uint32_t h = *reinterpret_cast<const uint32_t*>(hi + off), l = *reinterpret_cast<const uint32_t*>(lo + off);
    fumi_half h0, h1, l0, l1;
    *reinterpret_cast<uint16_t*>(&h0) = uint16_t(h & 0xFFFFu); *reinterpret_cast<uint16_t*>(&h1) = uint16_t(h >> 16);
    *reinterpret_cast<uint16_t*>(&l0) = uint16_t(l & 0xFFFFu); *reinterpret_cast<uint16_t*>(&l1) = uint16_t(l >> 16);
    a = (fumi_h2f(h0) + fumi_h2f(l0)) * inv;
    b = (fumi_h2f(h1) + fumi_h2f(l1)) * inv;
}
__device__ __forceinline__ void fumi_plane_store2(fumi_half* hi, fumi_half* lo, int off, float a, float b, float scale) {
    a *= scale; b *= scale;
    const fumi_half ah = fumi_f2h(a), bh = fumi_f2h(b);
    const fumi_half al = fumi_f2h(a - fumi_h2f(ah)), bl = fumi_f2h(b - fumi_h2f(bh));
    *reinterpret_cast<uint32_t*>(hi + off) = uint32_t(*reinterpret_cast<const uint16_t*>(&ah)) | (uint32_t(*reinterpret_cast<const uint16_t*>(&bh)) << 16);
    *reinterpret_cast<uint32_t*>(lo + off) = uint32_t(*reinterpret_cast<const uint16_t*>(&al)) | (uint32_t(*reinterpret_cast<const uint16_t*>(&bl)) << 16);
}
