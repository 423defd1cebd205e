// Episode Gram blocks: gram[b, i, j] = <x_i, x_j>, i over the task's NK support + NQ query rows,
// j over its NK support rows, with rows gathered straight from the HBM feature bank.
//
// This is the HBM-bound kernel of the path: it is the only place a sampled feature row
// (D fp32 = 8 KB at D = 2048) is read, once per task (algorithmic bytes 4 (NK+NQ) D per task,
// SURVEY.md section 8(d)).  It replaces the reference's per-step F.linear(x, W0_task) on
// materialised per-task weights (fumi.py:161,178) -- see episode.cu for how G is consumed --
// and the loader's per-sample feature copies (dataset/data.py:545,571-577).
//
// One CTA = (task, 64-row i tile, 32-row j tile); both tiles are staged through shared memory in DC-wide
// slices of the feature dimension (coalesced float4 row reads); each of the 4 warps owns 16 rows x 32
// columns of the output and accumulates it with warp-level 3xTF32 tensor-core tiles (warp_mma.cuh:
// fp32-accurate split, accumulator restarted every 32 features).  Row stride 68 makes both fragment reads
// bank-conflict free.
#include <cstdint>
#include <cstdlib>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"
#include "warp_mma.cuh"

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
    // warp w: rows [16w, 16w+16) of the i tile, all 32 columns of the j tile
    const int w = tid >> 5;
    float acc[1][4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
    for (int64_t d0 = 0; d0 < D; d0 += DC) {
        // stage the slice: (TI + TJ) rows x DC floats, float4 per thread, coalesced along d
        {
            float4 v[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) {                     // (64 + 32) rows x 16 float4 = 1536 = 12 per thread x 128
                const int idx = tid + q * GT;
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < (TI + TJ) * (DC / 4)) {
                    const int r = idx / (DC / 4), c4 = (idx - r * (DC / 4)) * 4;
                    const long long row = r < TI ? xrow[r] : srow[r - TI];
                    const int64_t d = d0 + c4;
                    if (row >= 0 && d + 3 < D) v[q] = __ldg(reinterpret_cast<const float4*>(feats + row * D + d));
                }
            }
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const int idx = tid + q * GT;
                if (idx < (TI + TJ) * (DC / 4)) {
                    const int r = idx / (DC / 4), c4 = (idx - r * (DC / 4)) * 4;
                    float* dst = r < TI ? &Xs[r * ST + c4] : &Ss[(r - TI) * ST + c4];
                    *reinterpret_cast<float4*>(dst) = v[q];
                }
            }
        }
        __syncthreads();
        warp_gemm_3xtf32<1, 4, false, true>(Xs + 16 * w * ST, ST, Ss, ST, DC, 1.f, acc);
        __syncthreads();
    }
    warp_tile_foreach<1, 4>(acc, [&](int ii, int j, float& v) {
        const int i = 16 * w + ii;
        if (i < ni && j < nj) gram[(b * int64_t(NK + NQ) + i0 + i) * NK + j0 + j] = v;
    });
}

}  // namespace

int fumi_gram_tc(const float*, int64_t, int64_t, const int64_t*, const int64_t*, int64_t, int32_t, int32_t, float*, void*);

#ifndef FUMI_EMU
int fumi_gram_tc_launch(const float* feats, int64_t D, const int64_t* sup_rows, const int64_t* qry_rows, int64_t B,
                        int32_t NK, int32_t NQ, float* gram, void* stream);     // gram_tc.cu (tcgen05)
#endif

extern "C" int fumi_gram(const float* feats, int64_t num_rows, int64_t D, const int64_t* sup_rows,
                         const int64_t* qry_rows, int64_t B, int32_t NK, int32_t NQ, float* gram, void* stream) {
    FUMI_CHECK_ARG(B >= 0 && NK >= 1 && NK <= kMaxSupport && NQ >= 0 && D >= 1 && num_rows >= 1, "bad shape");
    FUMI_CHECK_ARG((D & 3) == 0, "feature dim must be a multiple of 4 (float4 rows)");
    FUMI_CHECK_ARG((uintptr_t(feats) & 15) == 0, "feature matrix must be 16-byte aligned");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(feats && sup_rows && (qry_rows || NQ == 0) && gram, "null pointer");
#ifndef FUMI_EMU
    {   // NK <= 32 and NK + NQ <= 192: tcgen05 kernel (FUMI_GRAM_TC=0 keeps the warp-level kernel, for A/B runs)
        static int use_tc = -1;
        if (use_tc < 0) { const char* e = getenv("FUMI_GRAM_TC"); use_tc = (e && atoi(e) == 0) ? 0 : 1; }
        if (use_tc) {
            const int rc = fumi_gram_tc_launch(feats, D, sup_rows, qry_rows, B, NK, NQ, gram, stream);
            if (rc <= 0) return rc;
        }
    }
#endif
    FUMI_CHECK_ARG(B <= 65535, "at most 65535 tasks per call");
    dim3 grid((unsigned)((NK + NQ + TI - 1) / TI), (unsigned)((NK + TJ - 1) / TJ), (unsigned)B);
    FUMI_LAUNCH(gram_kernel, grid, GT, 0, stream, feats, D, sup_rows, qry_rows, int(NK), int(NQ), gram);
    FUMI_CHECK_LAUNCH("gram_kernel");
    return FUMI_OK;
}
