// AM3 meta-test scoring (BASELINE config 4): prototype + text mixing, squared distances, argmin, CE.
// Reference: AM3.evaluate (fumi/models/am3.py:159-200) with utils.get_prototypes (utils.py:331-376:
// per-class scatter-mean of image embeddings, convex mix with the class text prototype by lamda),
// utils.prototypical_loss (utils.py:390-402: CE over -||p - q||^2) and utils.get_preds
// (utils.py:302-328: argmin, lowest index on ties).  Eval mode (dropout off).  One CTA per task.
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"

namespace {

constexpr int kMaxP = 128;

__global__ void __launch_bounds__(256) am3_score_kernel(const float* __restrict__ emb,
                                                        const float* __restrict__ text_proto,
                                                        const float* __restrict__ lamda,
                                                        const int64_t* __restrict__ sup_rows,
                                                        const int64_t* __restrict__ qry_rows,
                                                        const int64_t* __restrict__ sup_y,
                                                        const int64_t* __restrict__ qry_y,
                                                        const int64_t* __restrict__ class_rows, int N, int NK, int NQ,
                                                        int Pd, int lamda_fixed, float* __restrict__ protos,
                                                        float* __restrict__ dist, int64_t* __restrict__ preds,
                                                        float* __restrict__ task_loss) {
    __shared__ float pr[kMaxWays * kMaxP];
    __shared__ float part[256];
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    for (int idx = tid; idx < N * Pd; idx += 256) {
        const int c = idx / Pd, p = idx - c * Pd;
        float sum = 0.f;
        int cnt = 0;
        for (int i = 0; i < NK; ++i) {
            if (sup_y[b * NK + i] == c) { sum += emb[sup_rows[b * NK + i] * Pd + p]; ++cnt; }
        }
        const float ebar = sum / float(cnt > 0 ? cnt : 1);
        const int64_t cr = class_rows[b * N + c];
        float lam = lamda[cr];
        if (lamda_fixed == 0) lam = 0.f;
        else if (lamda_fixed == 1) lam = 1.f;
        const float v = cnt > 0 ? lam * ebar + (1.f - lam) * text_proto[cr * Pd + p] : 0.f;
        pr[idx] = v;
        protos[b * N * Pd + idx] = v;
    }
    __syncthreads();
    float loss = 0.f;
    for (int j = tid; j < NQ; j += 256) {
        const float* e = emb + qry_rows[b * NQ + j] * Pd;
        float d[kMaxWays];
        for (int c = 0; c < N; ++c) d[c] = 0.f;
        for (int p = 0; p < Pd; ++p) {
            const float ev = e[p];
            for (int c = 0; c < N; ++c) {
                const float t = pr[c * Pd + p] - ev;
                d[c] = fmaf(t, t, d[c]);
            }
        }
        int best = 0;
        float mn = d[0];
        for (int c = 1; c < N; ++c) if (d[c] < mn) { mn = d[c]; best = c; }       // first min (torch.min)
        float sum = 0.f;
        for (int c = 0; c < N; ++c) sum += expf(mn - d[c]);                         // logits = -d, max = -mn
        const int y = int(qry_y[b * NQ + j]);
        loss += (logf(sum) - mn) + d[y];
        preds[b * NQ + j] = best;
        for (int c = 0; c < N; ++c) dist[(b * NQ + j) * N + c] = d[c];
    }
    part[tid] = loss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < 256; ++i) t += part[i];
        task_loss[b] = t;
    }
}

}  // namespace

extern "C" int fumi_am3_score(const float* emb, const float* text_proto, const float* lamda, const int64_t* sup_rows,
                              const int64_t* qry_rows, const int64_t* sup_y, const int64_t* qry_y,
                              const int64_t* class_rows, int64_t B, int32_t N, int32_t NK, int32_t NQ, int32_t P,
                              int32_t lamda_fixed, float* protos, float* dist, int64_t* preds, float* task_loss,
                              void* stream) {
    FUMI_CHECK_ARG(B >= 0 && N >= 1 && N <= kMaxWays && NK >= 1 && NQ >= 1, "bad shape");
    FUMI_CHECK_ARG(P >= 1 && P <= kMaxP, "prototype_dim must be in [1,128]");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(emb && text_proto && lamda && sup_rows && qry_rows && sup_y && qry_y && class_rows && protos &&
                   dist && preds && task_loss, "null pointer");
    FUMI_LAUNCH(am3_score_kernel, (unsigned)B, 256, 0, stream, emb, text_proto, lamda, sup_rows, qry_rows, sup_y, qry_y,
                class_rows, int(N), int(NK), int(NQ), int(P), int(lamda_fixed), protos, dist, preds, task_loss);
    FUMI_CHECK_LAUNCH("am3_score_kernel");
    return FUMI_OK;
}
