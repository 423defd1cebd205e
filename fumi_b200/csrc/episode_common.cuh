// Definitions shared by the episode kernels' translation units (not part of the C ABI).
#pragma once
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"
#include "warp_mma.cuh"

namespace fumi_epi {

constexpr int kThreads = 256;
constexpr int kW1S = kH1 + 1;        // padded row stride of W1^T in shared memory (bank-conflict free)
constexpr int kGS = kMaxSupport + 4; // row stride of the Gram tile in shared memory
constexpr int kLS = kMaxWays;        // row stride of the logits tile

struct EpiParams {
    fumi_episode_cfg cfg;
    int64_t B;
    const float* proj;
    const int64_t* sup_rows;
    const int64_t* qry_rows;
    const int64_t* sup_y;
    const int64_t* qry_y;
    const float* gram;
    const float* b0;
    const float* w1;
    const float* b1;
    const float* head_table;
    const int64_t* head_rows;
    float* logits;
    int64_t* preds;
    float* task_loss;
    float* task_acc;
    float* stash;
    int save;              // 1: per-task stash with step records (backward / parity dumps); 0: per-CTA scratch
    int64_t slot_floats;
    // backward only
    float loss_scale;
    float* d_proj;
    float* d_head;
    float* d_b0_parts;
    float* d_w1_parts;
    float* d_b1_parts;
    unsigned long long* phase;   // 64 cycle counters of the phase profiler, or null
};

struct Layout {
    int64_t per_task, S0, S1, w1t, b0, b1, head, steps, per_step, oH0, oH1, oDZ1, oDL, oHP, qH0, qH1, qLG;
};

__host__ __device__ inline Layout make_layout(const fumi_episode_cfg& c) {
    Layout L;
    const int64_t n = c.num_support, N = c.num_ways;
    L.S0 = 0;
    L.S1 = n * kH0;
    L.w1t = 2 * n * kH0;
    L.b0 = L.w1t + int64_t(kH0) * kH1;
    L.b1 = L.b0 + kH0;
    L.head = L.b1 + kH1;
    L.steps = L.head + ((N * kHD + 3) / 4) * 4;
    L.oH0 = 0;
    L.oH1 = n * kH0;
    L.oDZ1 = L.oH1 + n * kH1;
    L.oDL = L.oDZ1 + n * kH1;
    L.oHP = L.oDL + ((n * N + 3) / 4) * 4;
    L.per_step = L.oHP + ((N * kHD + 3) / 4) * 4;
    // query activations of the final forward (tensor-core path: the backward does not recompute them)
    const int64_t mq = c.num_query;
    L.qH0 = L.steps + int64_t(c.steps) * L.per_step;
    L.qH1 = L.qH0 + mq * kH0;
    L.qLG = L.qH1 + mq * kH1;
    L.per_task = L.qLG + ((mq * N + 3) / 4) * 4;
    return L;
}

// Stash of the fp16-plane kernels (NK <= 32), in 4-byte words; every block starts on a 16-byte boundary.
//   adapted state (fp32, read by the parity tests and by the backward's prologue): S, W1^T, b0, b1, head
//   per inner step: H1, dL, head of the step (fp32), the plane exponents, and H0 / dZ1 AS THE FORWARD HELD THEM, i.e.
//                   fp16 hi/lo planes [n][256] / [n][64] (dense rows): the backward copies them straight into its
//                   operand tiles (no conversion, no max exchange)
//   query pass:     H1q, softmax - onehot (fp32), the H0q planes and one exponent per 32-row tile
struct LayoutF {
    int64_t per_task, S, w1t, b0, b1, head, steps, per_step;
    int64_t oH1, oDL, oHP, oEXP, oH0h, oH0l, oDZh, oDZl;      // within a step record
    int64_t qH1, qLG, qEXP, qH0h, qH0l;                        // query pass (task offsets)
};
__host__ __device__ inline int64_t pad4(int64_t x) { return (x + 3) & ~int64_t(3); }
__host__ __device__ inline LayoutF make_layout_f(const fumi_episode_cfg& c) {
    LayoutF L;
    const int64_t n = c.num_support, N = c.num_ways, m = c.num_query;
    L.S = 0;
    L.w1t = pad4(n * kH0);
    L.b0 = L.w1t + int64_t(kH0) * kH1;
    L.b1 = L.b0 + kH0;
    L.head = L.b1 + kH1;
    L.steps = L.head + pad4(N * kHD);
    L.oH1 = 0;
    L.oDL = pad4(n * kH1);
    L.oHP = L.oDL + pad4(n * N);
    L.oEXP = L.oHP + pad4(N * kHD);
    L.oH0h = L.oEXP + 4;
    L.oH0l = L.oH0h + pad4(n * kH0 / 2);
    L.oDZh = L.oH0l + pad4(n * kH0 / 2);
    L.oDZl = L.oDZh + pad4(n * kH1 / 2);
    L.per_step = L.oDZl + pad4(n * kH1 / 2);
    L.qH1 = L.steps + int64_t(c.steps) * L.per_step;
    L.qLG = L.qH1 + pad4(m * kH1);
    L.qEXP = L.qLG + pad4(m * N);
    L.qH0h = L.qEXP + pad4((m + 31) / 32);
    L.qH0l = L.qH0h + pad4(m * kH0 / 2);
    L.per_task = L.qH0l + pad4(m * kH0 / 2);
    return L;
}

__device__ inline float dropout_scale(const fumi_episode_cfg& c) {
    return c.dropout_p > 0.f ? 1.f / (1.f - c.dropout_p) : 1.f;
}
// one 32-bit hash per (row, column pair): even column -> low 16 bits, odd column -> high 16 bits
__device__ inline uint32_t dropout_base(const fumi_episode_cfg& c, int64_t task, int pass, int layer) {
    return fumi_mask_base(c.dropout_seed, uint64_t(task), uint32_t(pass), uint32_t(layer));
}
__device__ inline uint32_t dropout_bits(uint32_t base, int row, int col) {
    return fumi_mask_pair(base, uint32_t(row), uint32_t(col));
}
__device__ inline uint32_t dropout_thr(const fumi_episode_cfg& c) { return uint32_t(c.dropout_p * 65536.f); }
__device__ inline bool dropout_keep_bits(uint32_t bits, int col, uint32_t thr) {
    return ((col & 1) ? (bits >> 16) : (bits & 0xFFFFu)) >= thr;
}
constexpr int kS0 = kH0 + 4;     // row stride of [rows][H0] tiles (A operand: conflict-free fragment loads)
constexpr int kS1 = kH1 + 4;     // row stride of [rows][H1] tiles and of W1^T [H0][H1]
constexpr int kSS = kH0 + 8;     // row stride of S (B operand, k = row)
constexpr int kSG = 36;          // row stride of a Gram tile with up to 32 columns
constexpr int kMaxQueryRows = 640;

constexpr int kThreads16 = 512;

constexpr int kHW = kH1 + 8;      // half stride of W1^T [H0][H1] and dZ1 [rows][H1] planes
constexpr int kHS = kH0 + 8;      // half stride of S / H0 [rows][H0] planes
constexpr int kHG = 32 + 8;       // half stride of a Gram tile [rows][32] planes

// block-wide max of a non-negative value without atomics: every warp leaves its max in its own word of the slot
// (all 16 words are rewritten by each production, so a slot needs no reset), readers reduce the 16 words after
// the barrier that follows.
__device__ __forceinline__ void block_max_push(float* slot, float m) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = m;
}
__device__ __forceinline__ int block_max_exp(const float* slot) {          // plane exponent s for the slot's matrix
    const float4 a = *reinterpret_cast<const float4*>(slot), b = *reinterpret_cast<const float4*>(slot + 4);
    const float4 c = *reinterpret_cast<const float4*>(slot + 8), d = *reinterpret_cast<const float4*>(slot + 12);
    const float m = fmaxf(fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))),
                          fmaxf(fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w))));
    return fumi_plane_exp(__float_as_uint(m));
}
__device__ __forceinline__ void store_pair(fumi_half* hi, fumi_half* lo, int off, float a, float b, float scale) {
    fumi_plane_store2(hi, lo, off, a, b, scale);
}
__device__ __forceinline__ float plane_value(const fumi_half* hi, const fumi_half* lo, int off, float inv) {
    return (fumi_h2f(hi[off]) + fumi_h2f(lo[off])) * inv;
}

// ---- helpers of the fp16-plane kernels ------------------------------------------------------------------------------
constexpr int kTarget = 8;                 // lagged planes: tracked max in [2^7, 2^8), 8 binades of headroom
__device__ __forceinline__ float slot_max(const float* slot) {
    const float4 a = *reinterpret_cast<const float4*>(slot), b = *reinterpret_cast<const float4*>(slot + 4);
    const float4 c = *reinterpret_cast<const float4*>(slot + 8), d = *reinterpret_cast<const float4*>(slot + 12);
    return fmaxf(fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))),
                 fmaxf(fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w))));
}
__device__ __forceinline__ float warp_max(float m) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    return m;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ void st_planes2(fumi_half* hi, fumi_half* lo, int off, float a, float b, float scale) {
    uint32_t h, l;
    fumi_split2(a * scale, b * scale, h, l);
    *reinterpret_cast<uint32_t*>(hi + off) = h;
    *reinterpret_cast<uint32_t*>(lo + off) = l;
}
__device__ __forceinline__ void st_plane1(fumi_half* hi, fumi_half* lo, int off, float a, float scale) {
    const float v = a * scale;
    const fumi_half h = fumi_f2h(v);
    hi[off] = h;
    lo[off] = fumi_f2h(v - fumi_h2f(h));
}
__device__ __forceinline__ void ld_planes2(const fumi_half* hi, const fumi_half* lo, int off, float inv, float& a, float& b) {
    fumi_join2(*reinterpret_cast<const uint32_t*>(hi + off), *reinterpret_cast<const uint32_t*>(lo + off), a, b);
    a *= inv;
    b *= inv;
}
// does max * 2^e fit fp16 with margin?  (uniform across the block: every thread evaluates the same values)
__device__ __forceinline__ bool plane_overflow(float mx, int e) { return mx * fumi_exp2i(e) >= 60000.f; }

// predicated form: the row test lives inside the instruction (@p red...) instead of a divergent branch around an
// opaque asm statement (one BSSY / BRA / BSYNC region per call otherwise)
__device__ __forceinline__ void atomic_add2_if(bool pred, float* addr, float a, float b) {
#ifdef FUMI_EMU
    if (pred) { atomicAdd(addr, a); atomicAdd(addr + 1, b); }
#else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\n@p red.global.add.v2.f32 [%0], {%1, %2};\n}" ::"l"(addr), "f"(a), "f"(b),
                 "r"(int(pred)) : "memory");
#endif
}
__device__ __forceinline__ void atomic_add2(float* addr, float a, float b) {
#ifdef FUMI_EMU
    atomicAdd(addr, a);
    atomicAdd(addr + 1, b);
#else
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
#endif
}

// Phase profiler (diagnostics only; enabled by fumi_debug_phase_profile(1)): thread 0 of each CTA adds the SM cycles
// between consecutive marks to phase[id] (64 device counters handed in through EpiParams::phase; null = off).
struct PhaseClock {
    long long last;
    unsigned long long* ctr;
    __device__ __forceinline__ void start(unsigned long long* counters) {
#ifndef FUMI_EMU
        ctr = threadIdx.x == 0 ? counters : nullptr;
        if (ctr) last = clock64();
#else
        ctr = nullptr; last = 0;
#endif
    }
    __device__ __forceinline__ void mark(int id) {
#ifndef FUMI_EMU
        if (ctr) {
            const long long t = clock64();
            atomicAdd(&ctr[id], (unsigned long long)(t - last));
            last = t;
        }
#endif
    }
};

// launchers of the fp16-plane tensor-core kernels (episode_fwd_f16.cu / episode_bwd_f16.cu); NK <= 32
int launch_episode_fwd_f16(const EpiParams& P, int grid, void* stream);
int launch_episode_bwd_f16(const EpiParams& P, int grid, void* stream);
bool episode_f16_supported(const fumi_episode_cfg& c);
size_t episode_bwd_f16_smem_bytes(int class_bucket);
unsigned long long* episode_phase_counters();      // null unless fumi_debug_phase_profile(1)

}  // namespace fumi_epi
