// Device half of the episodic task sampler (fumi_sampler_expand).
//
// The reference draws, for every (task, class) pair, a fresh numpy RandomState seeded with
// (hash(class tuple) + class) % 2**32 and permutes ALL of the class's images with it, keeping the
// first K+Q (torchmeta 1.7.0 ClassSplitter_.__getitem__ as used at fumi/dataset/data.py:146-184;
// SURVEY.md Appendix B).  That is B*N independent jobs of ~3 k dependent integer operations each --
// 20 k jobs per 4096-task meta-batch, ~30 ms on 16 host cores, and the reason the end-to-end rate
// trailed the kernels.  Here one warp owns one job:
//   lane 0        init_genrand(seed): 624-step serial recurrence into shared memory
//   32 lanes      the MT19937 twist, 32 words at a time (reads complete before writes; word k needs
//                 the old k+1 and k+397, or the new k-227, so ascending 32-word groups are safe)
//   32 lanes      tempering of the 624 words into a second array (the serial loop below then costs a load,
//                 a mask and a compare per draw; tempering inside it was 40 % of its instructions)
//   lane 0        numpy's backward Fisher-Yates: j = masked-rejection draw <= i, swap perm[i], perm[j]
//   32 lanes      the K+Q picks -> image id, bank row and label, int64, written to HBM
// Serial chains of different warps overlap; a job costs its class size, so the host plan hands the jobs over
// longest class first (job_order) -- 0.36 ms for a 4096 x 5 batch.
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"

namespace {

constexpr int kWarps = 8;          // jobs per CTA
constexpr int kMT = 624;

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// In-place regeneration of the 624 state words by one warp, then the tempered outputs into draws[].
__device__ __forceinline__ void mt_twist_warp(uint32_t* key, uint32_t* draws, int lane) {
    for (int base = 0; base < kMT; base += 32) {
        const int k = base + lane;
        uint32_t v = 0;
        if (k < kMT) {
            const uint32_t y = (key[k] & 0x80000000u) | (key[k == kMT - 1 ? 0 : k + 1] & 0x7fffffffu);
            v = key[k < kMT - 397 ? k + 397 : k - (kMT - 397)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (k < kMT) { key[k] = v; draws[k] = mt_temper(v); }
        __syncwarp();
    }
}

template <typename PermT>
__global__ void __launch_bounds__(kWarps * 32)
sampler_expand_kernel(const int64_t* __restrict__ offsets, const int64_t* __restrict__ ids, int perm_stride,
                      const int64_t* __restrict__ classes, const int64_t* __restrict__ label_perm,
                      const uint32_t* __restrict__ perm_seed, const int32_t* __restrict__ picks,
                      const int32_t* __restrict__ job_order, int64_t jobs, int K, int Q,
                      int64_t* __restrict__ sup_ids, int64_t* __restrict__ qry_ids,
                      int64_t* __restrict__ sup_y, int64_t* __restrict__ qry_y,
                      int64_t* __restrict__ sup_rows, int64_t* __restrict__ qry_rows) {
    FUMI_DYN_SMEM(uint32_t, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* key = smem + warp * (2 * kMT);
    uint32_t* draws = key + kMT;
    PermT* perm = reinterpret_cast<PermT*>(smem + kWarps * 2 * kMT) + size_t(warp) * perm_stride;
    const int64_t slot = int64_t(blockIdx.x) * kWarps + warp;
    if (slot >= jobs) return;                      // whole warp leaves together; no block-wide barrier below
    const int64_t job = job_order ? job_order[slot] : slot;        // longest classes first when the plan says so

    const int64_t c = classes[job];
    const int64_t row0 = offsets[c];
    const int n = int(offsets[c + 1] - row0);

    if (lane == 0) {                               // init_genrand
        uint32_t s = perm_seed[job];
        for (int i = 0; i < kMT; ++i) {
            key[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + uint32_t(i) + 1u;
        }
    }
    for (int i = lane; i < n; i += 32) perm[i] = PermT(i);
    __syncwarp();
    mt_twist_warp(key, draws, lane);

    int pos = 0;
    int i = n - 1;
    while (i >= 1) {                               // RandomState.shuffle: for i in reversed(range(1, n))
        if (lane == 0) {
            // run until the shuffle ends or the 624 words are used up
            uint32_t mask = uint32_t(i);
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            uint32_t d = draws[pos];
            while (i >= 1 && pos < kMT) {
                const uint32_t v = d & mask;
                d = draws[++pos];                  // next draw in flight (one word past the array is readable shared memory)
                if (v <= uint32_t(i)) {            // accepted: j = v
                    const PermT t = perm[i];
                    perm[i] = perm[v];
                    perm[v] = t;
                    --i;
                    if (uint32_t(i) < ((mask >> 1) + 1u)) mask >>= 1;          // i dropped below the top bit
                }
            }
        }
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= 1) {                              // state exhausted mid-shuffle: next 624 words
            pos = 0;
            __syncwarp();
            mt_twist_warp(key, draws, lane);
        }
    }
    __syncwarp();

    const int KQ = K + Q;
    const int64_t label = label_perm[job];
    const int32_t* pk = picks + job * KQ;
    for (int e = lane; e < KQ; e += 32) {
        const int64_t row = row0 + int64_t(perm[pk[e]]);
        const int64_t id = ids[row];
        if (e < K) {
            const int64_t o = job * K + e;
            sup_rows[o] = row; sup_ids[o] = id; sup_y[o] = label;
        } else {
            const int64_t o = job * Q + (e - K);
            qry_rows[o] = row; qry_ids[o] = id; qry_y[o] = label;
        }
    }
}

}  // namespace

extern "C" int fumi_sampler_expand(const int64_t* class_offsets, const int64_t* class_image_ids,
                                   int64_t max_class_size, const int64_t* classes, const int64_t* label_perm,
                                   const uint32_t* perm_seed, const int32_t* picks, const int32_t* job_order,
                                   int64_t B, int32_t N, int32_t K, int32_t Q, int64_t* sup_ids, int64_t* qry_ids, int64_t* sup_y,
                                   int64_t* qry_y, int64_t* sup_rows, int64_t* qry_rows, void* stream) {
    FUMI_CHECK_ARG(B >= 0 && N > 0 && K > 0 && Q >= 0 && max_class_size >= K + Q, "bad sizes");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(class_offsets && class_image_ids && classes && label_perm && perm_seed && picks && sup_ids &&
                   qry_ids && sup_y && qry_y && sup_rows && qry_rows, "null pointer");
    const int64_t jobs = B * N;
    const bool wide = max_class_size > 65535;
    const int64_t stride = (max_class_size + 7) & ~int64_t(7);
    const size_t smem = size_t(kWarps) * 2 * kMT * 4 + size_t(kWarps) * stride * (wide ? 4 : 2);
    FUMI_CHECK_ARG(smem <= 227 * 1024, "class too large for the shared-memory permutation (max ~22k images per class)");
    const unsigned grid = unsigned((jobs + kWarps - 1) / kWarps);
    if (wide) {
#ifndef FUMI_EMU
        cudaFuncSetAttribute(sampler_expand_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
#endif
        FUMI_LAUNCH(sampler_expand_kernel<uint32_t>, grid, kWarps * 32, smem, stream, class_offsets, class_image_ids,
                    int(stride), classes, label_perm, perm_seed, picks, job_order, jobs, K, Q, sup_ids, qry_ids, sup_y, qry_y,
                    sup_rows, qry_rows);
    } else {
#ifndef FUMI_EMU
        cudaFuncSetAttribute(sampler_expand_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
#endif
        FUMI_LAUNCH(sampler_expand_kernel<uint16_t>, grid, kWarps * 32, smem, stream, class_offsets, class_image_ids,
                    int(stride), classes, label_perm, perm_seed, picks, job_order, jobs, K, Q, sup_ids, qry_ids, sup_y, qry_y,
                    sup_rows, qry_rows);
    }
    FUMI_CHECK_LAUNCH("fumi_sampler_expand");
    return FUMI_OK;
}
