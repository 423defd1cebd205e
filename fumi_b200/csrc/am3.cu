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

// Backward of the same scoring for AM3 meta-training (am3.py:154-196: loss.backward() through
// prototypical_loss / get_prototypes).  One CTA per task.  With G = (softmax(-d) - onehot) * loss_scale:
//   d proto_c = -2 sum_q G[q][c] (proto_c - e_q)          d e_q = 2 sum_c G[q][c] (proto_c - e_q)
//   d ebar_c = lam_c d proto_c  (each support row of class c gets d ebar_c / count_c)
//   d t_c = (1 - lam_c) d proto_c                          d lam_c = <d proto_c, ebar_c - t_c>
// d_emb rows are accumulated with atomics (rows repeat across tasks); d_tproto / d_lamda are per (task, class) and
// are scattered into the class tables by fumi_scatter_add_rows.
__global__ void __launch_bounds__(256) am3_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ text_proto,
                                                      const float* __restrict__ lamda, const int64_t* __restrict__ sup_rows,
                                                      const int64_t* __restrict__ qry_rows, const int64_t* __restrict__ sup_y,
                                                      const int64_t* __restrict__ qry_y, const int64_t* __restrict__ class_rows,
                                                      int N, int NK, int NQ, int Pd, int lamda_fixed,
                                                      const float* __restrict__ protos, const float* __restrict__ dist,
                                                      float loss_scale, float* __restrict__ d_emb,
                                                      float* __restrict__ d_tproto, float* __restrict__ d_lamda) {
    FUMI_DYN_SMEM(float, sm);
    float* pr = sm;                 // [N][Pd]
    float* G = sm + N * Pd;         // [NQ][N]
    float* dp = G + NQ * N;         // [N][Pd]
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    for (int idx = tid; idx < N * Pd; idx += 256) pr[idx] = protos[b * N * Pd + idx];
    for (int j = tid; j < NQ; j += 256) {
        const float* d = dist + (b * NQ + j) * N;
        float mn = d[0];
        for (int c = 1; c < N; ++c) mn = fminf(mn, d[c]);
        float sum = 0.f;
        for (int c = 0; c < N; ++c) sum += expf(mn - d[c]);
        const int y = int(qry_y[b * NQ + j]);
        for (int c = 0; c < N; ++c) G[j * N + c] = (expf(mn - d[c]) / sum - (c == y ? 1.f : 0.f)) * loss_scale;
    }
    __syncthreads();
    // d e_q (thread = (query, feature) pairs, features fastest: coalesced reads of the embedding row)
    for (int idx = tid; idx < NQ * Pd; idx += 256) {
        const int j = idx / Pd, p = idx - j * Pd;
        const int64_t row = qry_rows[b * NQ + j];
        const float ev = emb[row * Pd + p];
        float a = 0.f;
        for (int c = 0; c < N; ++c) a = fmaf(G[j * N + c], pr[c * Pd + p] - ev, a);
        atomicAdd(&d_emb[row * Pd + p], 2.f * a);
    }
    // d proto_c
    for (int idx = tid; idx < N * Pd; idx += 256) {
        const int c = idx / Pd, p = idx - c * Pd;
        const float pv = pr[idx];
        float a = 0.f;
        for (int j = 0; j < NQ; ++j) a = fmaf(G[j * N + c], pv - emb[qry_rows[b * NQ + j] * Pd + p], a);
        dp[idx] = -2.f * a;
    }
    __syncthreads();
    for (int idx = tid; idx < N * Pd; idx += 256) {
        const int c = idx / Pd, p = idx - c * Pd;
        float sum = 0.f;
        int cnt = 0;
        for (int i = 0; i < NK; ++i)
            if (sup_y[b * NK + i] == c) { sum += emb[sup_rows[b * NK + i] * Pd + p]; ++cnt; }
        const int64_t cr = class_rows[b * N + c];
        float lam = lamda[cr];
        if (lamda_fixed == 0) lam = 0.f;
        else if (lamda_fixed == 1) lam = 1.f;
        const float g = cnt > 0 ? dp[idx] : 0.f;
        d_tproto[(b * N + c) * Pd + p] = (1.f - lam) * g;
        if (cnt > 0) {
            const float de = lam * g / float(cnt);
            for (int i = 0; i < NK; ++i)
                if (sup_y[b * NK + i] == c) atomicAdd(&d_emb[sup_rows[b * NK + i] * Pd + p], de);
        }
        // <d proto_c, ebar_c - t_c>: per-feature terms parked in dp, summed below
        dp[idx] = (cnt > 0 && lamda_fixed < 0) ? g * (sum / float(cnt) - text_proto[cr * Pd + p]) : 0.f;
    }
    __syncthreads();
    for (int c = tid; c < N; c += 256) {
        float a = 0.f;
        for (int p = 0; p < Pd; ++p) a += dp[c * Pd + p];
        d_lamda[b * N + c] = a;
    }
}

// x[r][c] *= keep(r, c) / (1 - p) with the counter-based mask of common.cuh (stream id = `layer`); applying it to a
// gradient with the same (seed, layer) is the dropout backward.
__global__ void dropout_apply_kernel(float* __restrict__ x, int64_t rows, int64_t cols, uint64_t seed, uint32_t layer,
                                     uint32_t thr, float scale) {
    const uint32_t base = fumi_mask_base(seed, 0, 0, layer);
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < rows * cols; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const uint32_t bits = fumi_mask_pair(base, uint32_t(r), uint32_t(c));
        const uint32_t f = (c & 1) ? (bits >> 16) : (bits & 0xFFFFu);
        x[i] = f >= thr ? x[i] * scale : 0.f;
    }
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ y, float* __restrict__ dy, int64_t n) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        dy[i] *= y[i] * (1.f - y[i]);
}

}  // namespace

extern "C" int fumi_am3_bwd(const float* emb, const float* text_proto, const float* lamda, const int64_t* sup_rows,
                            const int64_t* qry_rows, const int64_t* sup_y, const int64_t* qry_y, const int64_t* class_rows,
                            int64_t B, int32_t N, int32_t NK, int32_t NQ, int32_t P, int32_t lamda_fixed,
                            const float* protos, const float* dist, float loss_scale, float* d_emb, float* d_tproto,
                            float* d_lamda, void* stream) {
    FUMI_CHECK_ARG(B >= 0 && N >= 1 && N <= kMaxWays && NK >= 1 && NQ >= 1, "bad shape");
    FUMI_CHECK_ARG(P >= 1 && P <= kMaxP, "prototype_dim must be in [1,128]");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(emb && text_proto && lamda && sup_rows && qry_rows && sup_y && qry_y && class_rows && protos && dist &&
                   d_emb && d_tproto && d_lamda, "null pointer");
    const size_t smem = (size_t(2) * N * P + size_t(NQ) * N) * sizeof(float);
    FUMI_CHECK_ARG(smem <= 48 * 1024, "num_query * num_ways too large for the AM3 backward");
    FUMI_LAUNCH(am3_bwd_kernel, (unsigned)B, 256, smem, stream, emb, text_proto, lamda, sup_rows, qry_rows, sup_y, qry_y,
                class_rows, int(N), int(NK), int(NQ), int(P), int(lamda_fixed), protos, dist, loss_scale, d_emb, d_tproto,
                d_lamda);
    FUMI_CHECK_LAUNCH("am3_bwd_kernel");
    return FUMI_OK;
}

extern "C" int fumi_dropout_apply(float* x, int64_t rows, int64_t cols, uint64_t seed, uint32_t layer, float p, void* stream) {
    FUMI_CHECK_ARG(x && rows >= 0 && cols >= 1 && p >= 0.f && p < 1.f, "bad argument");
    if (rows == 0 || p == 0.f) return FUMI_OK;
    const int64_t n = rows * cols;
    const unsigned grid = unsigned((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    FUMI_LAUNCH(dropout_apply_kernel, grid, 256, 0, stream, x, rows, cols, seed, layer, uint32_t(p * 65536.f), 1.f / (1.f - p));
    FUMI_CHECK_LAUNCH("dropout_apply_kernel");
    return FUMI_OK;
}

// counts[t * N + p] += 1 per (target t, prediction p): the sufficient statistics of sklearn's accuracy_score and
// precision_recall_fscore_support(average="macro") that AM3.evaluate / get_preds report (utils.py:302-328,
// am3.py:196-204) -- 8 N^2 bytes leave the device instead of a host pass over every prediction.
__global__ void __launch_bounds__(256) confusion_kernel(const int64_t* __restrict__ y, const int64_t* __restrict__ pred,
                                                        int64_t n, int N, unsigned long long* __restrict__ counts) {
    __shared__ int h[kMaxWays * kMaxWays];
    for (int i = threadIdx.x; i < N * N; i += 256) h[i] = 0;
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
        const int64_t t = y[i], p = pred[i];
        if (t >= 0 && t < N && p >= 0 && p < N) atomicAdd(&h[int(t) * N + int(p)], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * N; i += 256)
        if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}

extern "C" int fumi_confusion_counts(const int64_t* y, const int64_t* pred, int64_t n, int32_t N, int64_t* counts,
                                     void* stream) {
    FUMI_CHECK_ARG(n >= 0 && N >= 1 && N <= kMaxWays, "bad shape");
    if (n == 0) return FUMI_OK;
    FUMI_CHECK_ARG(y && pred && counts, "null pointer");
    const unsigned grid = unsigned((n + 256 * 16 - 1) / (256 * 16) < 592 ? (n + 256 * 16 - 1) / (256 * 16) : 592);
    FUMI_LAUNCH(confusion_kernel, grid, 256, 0, stream, y, pred, n, int(N), reinterpret_cast<unsigned long long*>(counts));
    FUMI_CHECK_LAUNCH("confusion_kernel");
    return FUMI_OK;
}

extern "C" int fumi_sigmoid_bwd(const float* y, float* dy, int64_t n, void* stream) {
    FUMI_CHECK_ARG(y && dy && n >= 0, "bad argument");
    if (n == 0) return FUMI_OK;
    const unsigned grid = unsigned((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    FUMI_LAUNCH(sigmoid_bwd_kernel, grid, 256, 0, stream, y, dy, n);
    FUMI_CHECK_LAUNCH("sigmoid_bwd_kernel");
    return FUMI_OK;
}

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
