// Episode Gram blocks: gram[b, i, j] = <x_i, x_j>, i over the task's NK support + NQ query rows,
// j over its NK support rows, with rows gathered straight from the HBM feature bank.
//
// This is the HBM-bound kernel of the path: it is the only place a sampled feature row
// (D fp32 = 8 KB at D = 2048) is read, once per task (algorithmic bytes 4 (NK+NQ) D per task,
// SURVEY.md section 8(d)).  It replaces the reference's per-step F.linear(x, W0_task) on
// materialised per-task weights (fumi.py:161,178) -- see episode.cu for how G is consumed --
// and the loader's per-sample feature copies (dataset/data.py:545,571-577).
//
// fp32 FMA version: one CTA = (task, 64-row i tile, 32-row j tile); both tiles are staged through
// shared memory in DC-wide slices of the feature dimension; 4x4 register tile per thread with
// float4 reads along d (row strides chosen so both reads are bank-conflict free).
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"

namespace {

constexpr int TI = 64, TJ = 32, DC = 64, ST = DC + 4, GT = 128;

__global__ void __launch_bounds__(GT) gram_kernel(const float* __restrict__ feats, int64_t D,
                                                  const int64_t* __restrict__ sup_rows,
                                                  const int64_t* __restrict__ qry_rows, int NK, int NQ,
                                                  float* __restrict__ gram) {
    __shared__ __align__(16) float Xs[TI * ST];
    __shared__ __align__(16) float Ss[TJ * ST];
    __shared__ long long xrow[TI], srow[TJ];
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.z;
    const int i0 = blockIdx.x * TI, j0 = blockIdx.y * TJ;
    const int ni = min(TI, NK + NQ - i0), nj = min(TJ, NK - j0);
    if (tid < TI) {
        const int i = i0 + tid;
        xrow[tid] = tid < ni ? (i < NK ? sup_rows[b * NK + i] : qry_rows[b * NQ + (i - NK)]) : -1;
    }
    if (tid < TJ) srow[tid] = tid < nj ? sup_rows[b * NK + j0 + tid] : -1;
    __syncthreads();
    // thread tile: rows i = ii*16 + ti (ti = tid/8), cols j = jj*8 + tj (tj = tid%8)
    const int ti = tid >> 3, tj = tid & 7;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    for (int64_t d0 = 0; d0 < D; d0 += DC) {
        // stage the slice: (TI + TJ) rows x DC floats, float4 per thread, coalesced along d
        for (int idx = tid; idx < (TI + TJ) * (DC / 4); idx += GT) {
            const int r = idx / (DC / 4), q = (idx - r * (DC / 4)) * 4;
            const long long row = r < TI ? xrow[r] : srow[r - TI];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0) {
                const int64_t d = d0 + q;
                const float* src = feats + row * D + d;
                if (d + 3 < D) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (d < D) v.x = src[0];
                    if (d + 1 < D) v.y = src[1];
                    if (d + 2 < D) v.z = src[2];
                }
            }
            float* dst = r < TI ? &Xs[r * ST + q] : &Ss[(r - TI) * ST + q];
            *reinterpret_cast<float4*>(dst) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int q = 0; q < DC; q += 4) {
            float4 x[4], s[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) x[a] = *reinterpret_cast<const float4*>(&Xs[(a * 16 + ti) * ST + q]);
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] = *reinterpret_cast<const float4*>(&Ss[(c * 8 + tj) * ST + q]);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v = acc[a][c];
                    v = fmaf(x[a].x, s[c].x, v);
                    v = fmaf(x[a].y, s[c].y, v);
                    v = fmaf(x[a].z, s[c].z, v);
                    v = fmaf(x[a].w, s[c].w, v);
                    acc[a][c] = v;
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = a * 16 + ti;
        if (i >= ni) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = c * 8 + tj;
            if (j < nj) gram[(b * int64_t(NK + NQ) + i0 + i) * NK + j0 + j] = acc[a][c];
        }
    }
}

}  // namespace

int fumi_gram_tc(const float*, int64_t, int64_t, const int64_t*, const int64_t*, int64_t, int32_t, int32_t, float*, void*);

extern "C" int fumi_gram(const float* feats, int64_t num_rows, int64_t D, const int64_t* sup_rows,
                         const int64_t* qry_rows, int64_t B, int32_t NK, int32_t NQ, float* gram, void* stream) {
    FUMI_CHECK_ARG(B >= 0 && NK >= 1 && NK <= kMaxSupport && NQ >= 0 && D >= 1 && num_rows >= 1, "bad shape");
    FUMI_CHECK_ARG((D & 3) == 0, "feature dim must be a multiple of 4 (float4 rows)");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(feats && sup_rows && (qry_rows || NQ == 0) && gram, "null pointer");
    FUMI_CHECK_ARG(B <= 65535, "at most 65535 tasks per call");
    dim3 grid((unsigned)((NK + NQ + TI - 1) / TI), (unsigned)((NK + TJ - 1) / TJ), (unsigned)B);
    FUMI_LAUNCH(gram_kernel, grid, GT, 0, stream, feats, D, sup_rows, qry_rows, int(NK), int(NQ), gram);
    FUMI_CHECK_LAUNCH("gram_kernel");
    return FUMI_OK;
}
