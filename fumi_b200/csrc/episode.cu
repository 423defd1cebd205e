// Fused episodic inner loop, query scoring and their hand-written second-order backward.
//
// Reference semantics: fumi/models/fumi.py:148-193 and fumi/models/maml.py:158-191 (per-task Python
// loop: n_steps x [im_forward, cross_entropy, autograd.grad(create_graph=True), SGD update of the
// head AND of the whole image MLP], then query forward / argmax / CE, then outer backward).
//
// Formulation (DESIGN.md "Gram form"): support features X are fixed during adaptation, so the
// adapted first layer is W0_s = W0 - alpha * S_s^T X with S_s = sum_{t<s} dZ0_t  [NK x H0].  Hence
//      Z0_s = A - alpha * G S_s + b0_s,      A = X W0^T (rows of `proj`),  G = X X^T (`gram`)
//      Zq   = Aq - alpha * Gq S_S + b0_S     for the query rows.
// The 2 MB per-task W0 is never materialised; per-task state is S, W1 (64 KB), b0, b1 and the head.
//
// One CTA (256 threads == H0 columns) walks a task; rows are processed in tiles of TR so any
// NK <= 128 fits; W1^T lives in shared memory for the whole task, S in an L2-resident slot of the
// workspace (ping-pong across steps).  Thread mappings:
//   "column" ops (Z0, dZ0, S updates):  thread == hidden unit h, loop over tile rows (registers)
//   "A" ops   (Z1 = H0 W1^T):           thread == (o = tid%64, row group = tid/64)
//   "C" ops   (W1 -= a * dZ1^T H0):     thread == (o = tid%64, 64-wide k slab = tid/64)
// All reductions run in a fixed order (no atomics except the scatter into d_proj) so results are
// reproducible run to run.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"
#include "warp_mma.cuh"

#include "episode_common.cuh"

using namespace fumi_epi;

namespace {

// Shared-memory carve-up (floats).  Buffers hold TR = max(support tile, query tile) rows.
struct Smem {
    float* sS;     // [TRS][H0]   (forward, single support tile) first-layer coefficient S, updated in place
    float* sA;     // [TRS][H0]   (forward, single support tile) projected support rows A = X W0^T
    float* w1t;    // [H0][kW1S]  adapted W1^T
    float* aw1t;   // [H0][kW1S]  (backward) adjoint of W1^T
    float* h0t;    // [TR][H0]
    float* tt;     // [TR][H0]    (backward) r_dH0 / r_H0 / bar_Z0
    float* h1t;    // [TR][H1]
    float* dz1t;   // [TR][H1]
    float* rz1t;   // [TR][H1]    (backward) r_dZ1
    float* rh1t;   // [TR][H1]    (backward) r_H1 / r_Z1
    float* lt;     // [TR][kLS]   logits / dL
    float* rlt;    // [TR][kLS]   (backward) r_dL / r_L
    float* gt;     // [TR][kGS]   Gram tile
    float* hp;     // [N][HD]     head (current step)
    float* dhp;    // [N][HD]     forward: head gradient; backward: adjoint of head
    float* rhp;    // [N][HD]     (backward) this step's contribution to the head adjoint
    float* b1s;    // [H1]
    float* ab1;    // [H1]        (backward)
    float* rb1;    // [H1]        (backward)
    float* rowv;   // [TR]        per-row scalars (loss)
    float* rowc;   // [TR]        per-row scalars (correct)
    long long* rows;  // [TR]     gather rows of the tile
    int* ys;       // [TR]        labels of the tile
};

template <int TR, bool BWD, int TRC = 0>
__host__ __device__ inline size_t smem_floats() {
    size_t f = size_t(kH0) * kW1S + size_t(TR) * kH0 + 2 * size_t(TR) * kH1 + size_t(TR) * kLS + size_t(TR) * kGS +
               2 * size_t(kMaxWays) * kHD + kH1 + 2 * TR + 2 * TR /*rows (8B)*/ + TR;
    if (BWD) f += size_t(kH0) * kW1S + size_t(TR) * kH0 + 2 * size_t(TR) * kH1 + size_t(TR) * kLS +
                  size_t(kMaxWays) * kHD + 2 * kH1;
    f += 2 * size_t(TRC) * kH0;
    return f + 16;
}

template <int TR, bool BWD, int TRC = 0>
__device__ inline Smem carve(float* base) {
    Smem s;
    float* p = base;
    s.rows = reinterpret_cast<long long*>(p); p += 2 * TR;     // 8-byte aligned: first
    s.w1t = p; p += kH0 * kW1S;
    // keep float4-read buffers 16-byte aligned: kH0*kW1S = 16640 floats (multiple of 4)
    s.h0t = p; p += TR * kH0;
    s.h1t = p; p += TR * kH1;
    s.dz1t = p; p += TR * kH1;
    s.lt = p; p += TR * kLS;
    s.gt = p; p += TR * kGS;
    s.hp = p; p += kMaxWays * kHD;
    s.dhp = p; p += kMaxWays * kHD;
    s.b1s = p; p += kH1;
    s.rowv = p; p += TR;
    s.rowc = p; p += TR;
    s.ys = reinterpret_cast<int*>(p); p += TR;
    if (BWD) {
        s.tt = p; p += TR * kH0;
        s.rz1t = p; p += TR * kH1;
        s.rh1t = p; p += TR * kH1;
        s.rlt = p; p += TR * kLS;
        s.aw1t = p; p += kH0 * kW1S;
        s.rhp = p; p += kMaxWays * kHD;
        s.ab1 = p; p += kH1;
        s.rb1 = p; p += kH1;
    } else {
        s.tt = s.rz1t = s.rh1t = s.rlt = s.aw1t = s.rhp = s.ab1 = s.rb1 = nullptr;
    }
    s.sS = p; p += TRC * kH0;
    s.sA = p; p += TRC * kH0;
    return s;
}

// ---- tile loaders -----------------------------------------------------------------------------
// rows / labels / Gram rows of tile [r0, r0+tr) of the support (qry=false) or query set of task b.
template <int TR>
__device__ inline void load_tile_meta(const EpiParams& P, const Smem& s, int64_t b, bool qry, int r0, int tr) {
    const int n = P.cfg.num_support, m = P.cfg.num_query;
    const int tid = threadIdx.x;
    if (tid < TR) {
        long long row = 0;
        int y = 0;
        if (tid < tr) {
            if (qry) { row = P.qry_rows[b * m + r0 + tid]; y = int(P.qry_y[b * m + r0 + tid]); }
            else     { row = P.sup_rows[b * n + r0 + tid]; y = int(P.sup_y[b * n + r0 + tid]); }
        }
        s.rows[tid] = row;
        s.ys[tid] = y;
    }
    const int n4 = (n + 3) & ~3;
    const float* g = P.gram + (b * int64_t(n + m) + (qry ? n : 0) + r0) * n;
    for (int idx = tid; idx < TR * n4; idx += kThreads) {
        const int i = idx / n4, j = idx - i * n4;
        s.gt[i * kGS + j] = (i < tr && j < n) ? g[int64_t(i) * n + j] : 0.f;
    }
}

// (a) Z0 = A - alpha * G S + b0 -> H0 = relu(Z0) * dropout.   thread == column h.
// CACHED: S and the projected rows A come from shared memory (single support tile); else from global.
template <int TR, bool CACHED>
__device__ inline void tile_h0(const EpiParams& P, const Smem& s, int64_t task, const float* Scur, float b0h,
                               int r0, int tr, int pass) {
    const int h = threadIdx.x;
    const int n = P.cfg.num_support;
    float acc[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) acc[i] = 0.f;
    if (Scur != nullptr) {
        const int n4 = (n + 3) & ~3;
        for (int j = 0; j < n4; j += 4) {
            float sv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) sv[q] = (j + q < n) ? Scur[(j + q) * kH0 + h] : 0.f;
#pragma unroll
            for (int i = 0; i < TR; ++i) {
                const float4 g = *reinterpret_cast<const float4*>(&s.gt[i * kGS + j]);
                acc[i] = fmaf(g.x, sv[0], acc[i]);
                acc[i] = fmaf(g.y, sv[1], acc[i]);
                acc[i] = fmaf(g.z, sv[2], acc[i]);
                acc[i] = fmaf(g.w, sv[3], acc[i]);
            }
        }
    }
    const float alpha = P.cfg.step_size;
    const float sc = dropout_scale(P.cfg);
    const bool drop = P.cfg.dropout_p > 0.f;
    const uint32_t thr = dropout_thr(P.cfg);
    const uint32_t dbase = drop ? dropout_base(P.cfg, task, pass, 0) : 0u;
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        float v = 0.f;
        if (i < tr) {
            const float a = CACHED ? s.sA[i * kH0 + h] : P.proj[s.rows[i] * kH0 + h];
            const float z = a + b0h - alpha * acc[i];
            // (no short-circuit: a mask hashed under a data-dependent branch is a divergent branch per element)
            const bool keep = !drop | dropout_keep_bits(dropout_bits(dbase, r0 + i, h), h, thr);
            v = ((z > 0.f) & keep) ? z * sc : 0.f;
        }
        s.h0t[i * kH0 + h] = v;
    }
}

// (b) Z1 = H0 W1^T + b1 -> H1.   thread == (o, row group).
template <int TR>
__device__ inline void tile_h1(const EpiParams& P, const Smem& s, int64_t task, int r0, int tr, int pass) {
    constexpr int RPT = TR / 4;
    const int o = threadIdx.x & 63, ig = threadIdx.x >> 6;
    float acc[RPT];
#pragma unroll
    for (int ii = 0; ii < RPT; ++ii) acc[ii] = 0.f;
    for (int k = 0; k < kH0; k += 4) {
        const float w0 = s.w1t[(k + 0) * kW1S + o], w1 = s.w1t[(k + 1) * kW1S + o];
        const float w2 = s.w1t[(k + 2) * kW1S + o], w3 = s.w1t[(k + 3) * kW1S + o];
#pragma unroll
        for (int ii = 0; ii < RPT; ++ii) {
            const float4 hv = *reinterpret_cast<const float4*>(&s.h0t[(ig + 4 * ii) * kH0 + k]);
            acc[ii] = fmaf(hv.x, w0, acc[ii]);
            acc[ii] = fmaf(hv.y, w1, acc[ii]);
            acc[ii] = fmaf(hv.z, w2, acc[ii]);
            acc[ii] = fmaf(hv.w, w3, acc[ii]);
        }
    }
    const float sc = dropout_scale(P.cfg);
    const bool drop = P.cfg.dropout_p > 0.f;
    const uint32_t thr = dropout_thr(P.cfg);
    const uint32_t dbase = drop ? dropout_base(P.cfg, task, pass, 1) : 0u;
    const float b = s.b1s[o];
#pragma unroll
    for (int ii = 0; ii < RPT; ++ii) {
        const int i = ig + 4 * ii;
        float v = 0.f;
        if (i < tr) {
            const float z = acc[ii] + b;
            const bool keep = !drop | dropout_keep_bits(dropout_bits(dbase, r0 + i, o), o, thr);
            v = ((z > 0.f) & keep) ? z * sc : 0.f;
        }
        s.h1t[i * kH1 + o] = v;
    }
}

// (c) logits = H1 . head[:, :64]^T + head[:, 64]
template <int TR>
__device__ inline void tile_logits(const EpiParams& P, const Smem& s, int tr) {
    const int N = P.cfg.num_ways;
    for (int idx = threadIdx.x; idx < TR * N; idx += kThreads) {
        const int i = idx / N, c = idx - i * N;
        float l = 0.f;
        if (i < tr) {
            l = s.hp[c * kHD + kH1];
            for (int o = 0; o < kH1; ++o) l = fmaf(s.h1t[i * kH1 + o], s.hp[c * kHD + o], l);
        }
        s.lt[i * kLS + c] = l;
    }
}

// softmax statistics of row i (N <= 32): returns max, sum of exp, and fills e[] with exp(l - max)
__device__ inline void row_softmax(const float* l, int N, float& mx, float& sum) {
    mx = l[0];
#pragma unroll 1
    for (int c = 1; c < N; ++c) mx = fmaxf(mx, l[c]);
    sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < N; ++c) sum += expf(l[c] - mx);
}

// "C" op:  acc[k][o] = sum_i X[i][k] * Y[i][o]  for the thread's o and its 64-wide k slab, handed
// four k at a time to `sink(k, a0, a1, a2, a3)`.
template <int TR, typename Sink>
__device__ inline void outer_rows(const float* X /*[TR][H0]*/, const float* Y /*[TR][H1]*/, Sink sink) {
    const int o = threadIdx.x & 63, kg = threadIdx.x >> 6;
    float yr[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) yr[i] = Y[i * kH1 + o];
#pragma unroll
    for (int kk = 0; kk < 64; kk += 4) {   // fully unrolled: sinks index per-thread register arrays by kk
        const int k = kg * 64 + kk;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            const float4 xv = *reinterpret_cast<const float4*>(&X[i * kH0 + k]);
            a0 = fmaf(xv.x, yr[i], a0);
            a1 = fmaf(xv.y, yr[i], a1);
            a2 = fmaf(xv.z, yr[i], a2);
            a3 = fmaf(xv.w, yr[i], a3);
        }
        sink(kk, k, o, a0, a1, a2, a3);
    }
}

// "B" op: acc[i] (+)= sum_o Y[i][o] * Wt[h][o] * wscale   thread == h.
template <int TR>
__device__ inline void cols_from_h1(const float* Y /*[TR][H1]*/, const float* Wt /*[H0][kW1S]*/, float wscale,
                                    float (&acc)[TR]) {
    const int h = threadIdx.x;
    for (int o = 0; o < kH1; o += 4) {
        const float w0 = Wt[h * kW1S + o] * wscale, w1 = Wt[h * kW1S + o + 1] * wscale;
        const float w2 = Wt[h * kW1S + o + 2] * wscale, w3 = Wt[h * kW1S + o + 3] * wscale;
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            const float4 y = *reinterpret_cast<const float4*>(&Y[i * kH1 + o]);
            acc[i] = fmaf(y.x, w0, acc[i]);
            acc[i] = fmaf(y.y, w1, acc[i]);
            acc[i] = fmaf(y.z, w2, acc[i]);
            acc[i] = fmaf(y.w, w3, acc[i]);
        }
    }
}

// ------------------------------------------------------------------------------------ forward
// TRS / TRQ: rows per support / query tile.  MULTI: more than one support tile (NK > TRS): S ping-pongs in
// the L2-resident workspace and the W1 update is accumulated in registers until all tiles of the step are
// done.  Otherwise (the common case, NK <= 32) S and the projected support rows stay in shared memory for
// the whole task: the support set is staged once and reused by all inner steps.
template <int TRS, int TRQ, bool MULTI>
__global__ void __launch_bounds__(kThreads, 1) episode_fwd_kernel(EpiParams P) {
    constexpr int TRM = TRS > TRQ ? TRS : TRQ;
    FUMI_DYN_SMEM(float, smem_raw);
    const Smem s = carve<TRM, false, MULTI ? 0 : TRS>(smem_raw);
    const fumi_episode_cfg& c = P.cfg;
    const int tid = threadIdx.x;
    const int n = c.num_support, m = c.num_query, N = c.num_ways, steps = c.steps;
    const float alpha = c.step_size;
    const Layout L = make_layout(c);
    const int o_ = tid & 63, kg_ = tid >> 6;
    __shared__ float task_sum[2];

    for (int64_t b = blockIdx.x; b < P.B; b += gridDim.x) {
        const int64_t task = c.task_offset + b;
        float* slot = P.stash + (P.save ? b : int64_t(blockIdx.x)) * P.slot_floats;
        float* Sbuf[2] = {slot + L.S0, slot + L.S1};
        // ---- task prologue: W1^T, biases, head init (and the projected support rows)
        for (int idx = tid; idx < kH0 * kH1; idx += kThreads) {
            const int o = idx / kH0, k = idx - o * kH0;        // w1 is [H1][H0] row-major
            s.w1t[k * kW1S + o] = P.w1[idx];
        }
        float b0h = P.b0[tid];
        if (tid < kH1) s.b1s[tid] = P.b1[tid];
        for (int idx = tid; idx < N * kHD; idx += kThreads) {
            const int cc = idx / kHD, o = idx - cc * kHD;
            const int64_t r = P.head_rows ? P.head_rows[b * N + cc] : cc;
            s.hp[idx] = P.head_table[r * kHD + o];
        }
        if (!MULTI) {
            for (int i = 0; i < n; ++i) s.sA[i * kH0 + tid] = P.proj[P.sup_rows[b * n + i] * kH0 + tid];
        }
        if (tid == 0) { task_sum[0] = 0.f; task_sum[1] = 0.f; }
        __syncthreads();

        int cur = 0;
        for (int st = 0; st < steps; ++st) {
            const float* Scur = st > 0 ? (MULTI ? Sbuf[cur] : s.sS) : nullptr;       // S_0 == 0
            float* Snext = Sbuf[cur ^ 1];
            float* rec = P.save ? slot + L.steps + int64_t(st) * L.per_step : nullptr;
            float db0 = 0.f, db1 = 0.f;
            float dw1[MULTI ? 64 : 1];
            if (MULTI) {
#pragma unroll
                for (int q = 0; q < (MULTI ? 64 : 1); ++q) dw1[q] = 0.f;
            }
            for (int idx = tid; idx < N * kHD; idx += kThreads) s.dhp[idx] = 0.f;
            for (int r0 = 0; r0 < n; r0 += TRS) {
                const int tr = min(TRS, n - r0);
                load_tile_meta<TRS>(P, s, b, false, r0, tr);
                __syncthreads();
                tile_h0<TRS, !MULTI>(P, s, task, Scur, b0h, r0, tr, st);
                __syncthreads();
                tile_h1<TRS>(P, s, task, r0, tr, st);
                __syncthreads();
                tile_logits<TRS>(P, s, tr);
                __syncthreads();
                if (tid < TRS) {                                   // dL = (softmax - onehot) / n
                    float* l = &s.lt[tid * kLS];
                    if (tid < tr) {
                        float mx, sum;
                        row_softmax(l, N, mx, sum);
                        const float inv = 1.f / sum, invn = 1.f / float(n);
                        const int y = s.ys[tid];
                        for (int cc = 0; cc < N; ++cc) {
                            const float p = expf(l[cc] - mx) * inv;
                            l[cc] = (p - (cc == y ? 1.f : 0.f)) * invn;
                        }
                    } else {
                        for (int cc = 0; cc < N; ++cc) l[cc] = 0.f;
                    }
                }
                __syncthreads();
                // (d) head gradient, (e) dZ1 (uses the pre-update head)
                for (int idx = tid; idx < N * kHD; idx += kThreads) {
                    const int cc = idx / kHD, o = idx - cc * kHD;
                    float a = 0.f;
                    for (int i = 0; i < tr; ++i) a = fmaf(s.lt[i * kLS + cc], o < kH1 ? s.h1t[i * kH1 + o] : 1.f, a);
                    s.dhp[idx] += a;
                }
                {
                    const float sc = dropout_scale(c);
#pragma unroll
                    for (int ii = 0; ii < TRS / 4; ++ii) {
                        const int i = kg_ + 4 * ii;
                        float dz = 0.f;
                        if (i < tr && s.h1t[i * kH1 + o_] > 0.f) {
                            float dh = 0.f;
                            for (int cc = 0; cc < N; ++cc) dh = fmaf(s.lt[i * kLS + cc], s.hp[cc * kHD + o_], dh);
                            dz = dh * sc;
                        }
                        s.dz1t[i * kH1 + o_] = dz;
                    }
                }
                __syncthreads();
                if (tid < kH1) {
                    float a = 0.f;
                    for (int i = 0; i < tr; ++i) a += s.dz1t[i * kH1 + tid];
                    db1 += a;
                }
                // (f) dZ0 = (dZ1 W1) * mask ; S += dZ0
                {
                    float acc[TRS];
#pragma unroll
                    for (int i = 0; i < TRS; ++i) acc[i] = 0.f;
                    cols_from_h1<TRS>(s.dz1t, s.w1t, 1.f, acc);
                    const float sc = dropout_scale(c);
#pragma unroll
                    for (int i = 0; i < TRS; ++i) {
                        if (i < tr) {
                            const float dz0 = s.h0t[i * kH0 + tid] > 0.f ? acc[i] * sc : 0.f;
                            if (MULTI) {
                                const int64_t off = int64_t(r0 + i) * kH0 + tid;
                                Snext[off] = (Scur ? Scur[off] : 0.f) + dz0;
                            } else {            // single tile: every Z0 of this step is done, update in place
                                s.sS[i * kH0 + tid] = (st > 0 ? s.sS[i * kH0 + tid] : 0.f) + dz0;
                            }
                            db0 += dz0;
                        }
                    }
                }
                if (rec) {                                          // records for the backward
                    for (int i = 0; i < tr; ++i) rec[L.oH0 + int64_t(r0 + i) * kH0 + tid] = s.h0t[i * kH0 + tid];
                    for (int idx = tid; idx < tr * kH1; idx += kThreads) {
                        rec[L.oH1 + int64_t(r0) * kH1 + idx] = s.h1t[idx];
                        rec[L.oDZ1 + int64_t(r0) * kH1 + idx] = s.dz1t[idx];
                    }
                    for (int idx = tid; idx < tr * N; idx += kThreads) {
                        const int i = idx / N, cc = idx - i * N;
                        rec[L.oDL + int64_t(r0) * N + idx] = s.lt[i * kLS + cc];
                    }
                }
                __syncthreads();                                    // (f) done reading W1^T
                // (g) W1 -= alpha * dZ1^T H0
                if (MULTI) {
                    outer_rows<TRS>(s.h0t, s.dz1t, [&](int kk, int, int, float a0, float a1, float a2, float a3) {
                        dw1[kk] += a0; dw1[kk + 1] += a1; dw1[kk + 2] += a2; dw1[kk + 3] += a3;
                    });
                } else {
                    outer_rows<TRS>(s.h0t, s.dz1t, [&](int, int k, int o, float a0, float a1, float a2, float a3) {
                        s.w1t[(k + 0) * kW1S + o] -= alpha * a0;
                        s.w1t[(k + 1) * kW1S + o] -= alpha * a1;
                        s.w1t[(k + 2) * kW1S + o] -= alpha * a2;
                        s.w1t[(k + 3) * kW1S + o] -= alpha * a3;
                    });
                }
                __syncthreads();
            }
            // ---- end of step: apply the SGD update (all gradients were taken at the pre-update point)
            if (rec) for (int idx = tid; idx < N * kHD; idx += kThreads) rec[L.oHP + idx] = s.hp[idx];
            for (int idx = tid; idx < N * kHD; idx += kThreads) s.hp[idx] -= alpha * s.dhp[idx];
            if (tid < kH1) s.b1s[tid] -= alpha * db1;
            b0h -= alpha * db0;
            if (MULTI) {
#pragma unroll
                for (int kk = 0; kk < (MULTI ? 64 : 1); ++kk) s.w1t[(kg_ * 64 + kk) * kW1S + o_] -= alpha * dw1[kk];
            }
            cur ^= 1;
            __syncthreads();
        }

        // ---- query scoring
        const float* Sfin = steps > 0 ? (MULTI ? Sbuf[cur] : s.sS) : nullptr;
        for (int r0 = 0; r0 < m; r0 += TRQ) {
            const int tr = min(TRQ, m - r0);
            load_tile_meta<TRQ>(P, s, b, true, r0, tr);
            __syncthreads();
            tile_h0<TRQ, false>(P, s, task, Sfin, b0h, r0, tr, steps);
            __syncthreads();
            tile_h1<TRQ>(P, s, task, r0, tr, steps);
            __syncthreads();
            tile_logits<TRQ>(P, s, tr);
            __syncthreads();
            if (tid < tr) {
                const float* l = &s.lt[tid * kLS];
                float mx, sum;
                row_softmax(l, N, mx, sum);
                int best = 0;
                for (int cc = 1; cc < N; ++cc) if (l[cc] > l[best]) best = cc;   // first max (torch.max)
                const int y = s.ys[tid];
                s.rowv[tid] = (logf(sum) + mx) - l[y];
                s.rowc[tid] = best == y ? 1.f : 0.f;
                const int64_t q = b * m + r0 + tid;
                P.preds[q] = best;
                for (int cc = 0; cc < N; ++cc) P.logits[q * N + cc] = l[cc];
            }
            __syncthreads();
            if (tid == 0) {
                float a = task_sum[0], k = task_sum[1];
                for (int i = 0; i < tr; ++i) { a += s.rowv[i]; k += s.rowc[i]; }
                task_sum[0] = a; task_sum[1] = k;
            }
            __syncthreads();
        }
        if (tid == 0) {
            P.task_loss[b] = task_sum[0] / float(m);
            P.task_acc[b] = task_sum[1] / float(m);
        }
        if (P.save) {                                               // adapted state
            for (int idx = tid; idx < kH0 * kH1; idx += kThreads) {
                const int k = idx / kH1, o = idx - k * kH1;
                slot[L.w1t + idx] = s.w1t[k * kW1S + o];
            }
            slot[L.b0 + tid] = b0h;
            if (tid < kH1) slot[L.b1 + tid] = s.b1s[tid];
            for (int idx = tid; idx < N * kHD; idx += kThreads) slot[L.head + idx] = s.hp[idx];
            if (!MULTI) {                                           // final S where the backward expects it
                float* Sout = Sbuf[steps & 1];
                for (int i = 0; i < n; ++i) Sout[int64_t(i) * kH0 + tid] = steps > 0 ? s.sS[i * kH0 + tid] : 0.f;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ forward (tensor core)
// Single-support-tile forward (NK <= 32, the common case): all GEMM-shaped ops run as warp-level 3xTF32
// mma.sync tiles (warp_mma.cuh) on the fp32 tiles in shared memory.  MT = 1 for NK <= 16, else 2
// (rows padded with zeros to 16*MT).  Warp w of 8 owns hidden units [32w, 32w+32) of the 256-wide ops and
// output units [8w, 8w+8) of the 64-wide op.  Query rows go through in tiles of 32.


// 512 threads: the phases are latency-bound with 2 warps per scheduler (an 8-warp version measured issue
// active 22-28 %, tensor pipe 20 %), so the work of every phase is split over 16 warps:
// warp w owns hidden units [16w, 16w+16) of the 256-wide ops and the (m tile w/8, n tile w%8) block of the
// 64-wide op.  Launch bound 512 threads -> 128 registers per thread.

template <typename SM, typename Sink>
__device__ __forceinline__ void m16_tile_logits_softmax(const EpiParams& P, const SM& s, int rows, int tr, Sink sink) {
    const int N = P.cfg.num_ways;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rpw = 32 / N, li = lane / N, c = lane - li * N;
    for (int rbase = w * rpw; rbase < rows; rbase += (kThreads16 / 32) * rpw) {      // uniform within a warp
        const int i = rbase + li;
        const bool act = li < rpw && i < rows;
        float l = 0.f;
        if (act && i < tr) {
            float l0 = s.hp[c * kHD + kH1], l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll 4
            for (int o = 0; o < kH1; o += 4) {
                l0 = fmaf(s.h1t[i * kS1 + o], s.hp[c * kHD + o], l0);
                l1 = fmaf(s.h1t[i * kS1 + o + 1], s.hp[c * kHD + o + 1], l1);
                l2 = fmaf(s.h1t[i * kS1 + o + 2], s.hp[c * kHD + o + 2], l2);
                l3 = fmaf(s.h1t[i * kS1 + o + 3], s.hp[c * kHD + o + 3], l3);
            }
            l = (l0 + l1) + (l2 + l3);
        }
        if (act) s.lt[i * kLS + c] = l;
        __syncwarp();
        float mx = 0.f, sum = 1.f;
        if (act && i < tr) row_softmax(&s.lt[i * kLS], N, mx, sum);
        __syncwarp();
        if (act) sink(i, c, l, mx, sum);
    }
}

// ------------------------------------------------------------------------------------ forward, fp16 planes
// Same algorithm and launch shape as episode_fwd_mma16_kernel, but every GEMM operand lives in shared memory as
// PRE-SPLIT fp16 hi/lo planes (x 2^s = hi + lo, 22 significant bits, the bytes of one fp32 tile): the inner loops
// are ldmatrix + mma.m16n8k16 only (warp_gemm_f16x3; 2.6x the throughput of splitting fp32 fragments per use,
// tools/warp_gemm_bench.cu).  A matrix is produced in registers, its max |x| goes through a shared-memory
// atomicMax (one extra barrier), the power-of-two scale 2^s puts that max in [2^13, 2^14), and the planes are
// written once; consumers undo the two scales exactly in their epilogues.  W1^T and S are updated in place by
// reconstructing (hi + lo) 2^-s at the thread's own accumulator positions.  The projected rows A are not staged in
// shared memory any more: each thread loads the 8-byte pieces at its own accumulator positions from `proj`.

struct SmemF {
    fumi_half *w1h, *w1l, *sh, *sl, *h0h, *h0l, *dzh, *dzl, *gsh, *gsl, *gqh, *gql;
    float *h1t, *dz1t, *lt, *hp, *dhp, *b0s, *db0s, *b1s, *rowv, *rowc;
    float* mx;                     // max |x| of each matrix: 16 per-warp partials per slot, two slots per matrix
    long long *rowsQ, *rowsS;
    int *ysQ, *ysS;
};
enum { MX_W1 = 0, MX_S = 2, MX_H0 = 4, MX_DZ = 6, MX_GS = 8, MX_GQ = 10, MX_COUNT = 12 };   // slot indices (x 16 floats)
// lays the buffers out from `base` (16-byte aligned pieces) and returns the total size; base = nullptr sizes it
__host__ __device__ inline size_t carve_f(char* base, SmemF& s) {
    char* p = base;
    auto take = [&](size_t bytes) { char* r = p; p += (bytes + 15) & ~size_t(15); return r; };
    s.rowsQ = reinterpret_cast<long long*>(take(kMaxQueryRows * 8));
    s.rowsS = reinterpret_cast<long long*>(take(32 * 8));
    s.ysQ = reinterpret_cast<int*>(take(kMaxQueryRows * 4));
    s.ysS = reinterpret_cast<int*>(take(32 * 4));
    s.mx = reinterpret_cast<float*>(take(MX_COUNT * 16 * 4));
    // everything from here on is zero-filled once per kernel (plane pads must be finite)
    s.w1h = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));  s.w1l = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));
    s.sh = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));    s.sl = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));
    s.h0h = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));   s.h0l = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));
    s.dzh = reinterpret_cast<fumi_half*>(take(32 * kHW * 2));   s.dzl = reinterpret_cast<fumi_half*>(take(32 * kHW * 2));
    s.gsh = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));   s.gsl = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));
    s.gqh = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));   s.gql = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));
    s.h1t = reinterpret_cast<float*>(take(32 * kS1 * 4));
    s.dz1t = reinterpret_cast<float*>(take(32 * kS1 * 4));
    s.lt = reinterpret_cast<float*>(take(32 * kLS * 4));
    s.hp = reinterpret_cast<float*>(take(kMaxWays * kHD * 4));
    s.dhp = reinterpret_cast<float*>(take(kMaxWays * kHD * 4));
    s.b0s = reinterpret_cast<float*>(take(kH0 * 4));
    s.db0s = reinterpret_cast<float*>(take(kH0 * 4));
    s.b1s = reinterpret_cast<float*>(take(kH1 * 4));
    s.rowv = reinterpret_cast<float*>(take(32 * 4));
    s.rowc = reinterpret_cast<float*>(take(32 * 4));
    return size_t(p - base);
}
inline size_t smem_f_bytes() { SmemF t; return carve_f(nullptr, t); }


template <int MT>
__global__ void __launch_bounds__(kThreads16, 1) episode_fwd_f16_kernel(EpiParams P) {
    constexpr int RS = 16 * MT;
    constexpr int NT_ = kThreads16;
    FUMI_DYN_SMEM(float, smem_raw);
    SmemF s;
    const size_t smem_total = carve_f(reinterpret_cast<char*>(smem_raw), s);
    const fumi_episode_cfg& c = P.cfg;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int col = tid & 255, half = tid >> 8;
    const int n = c.num_support, m = c.num_query, N = c.num_ways, steps = c.steps;
    const float alpha = c.step_size;
    const Layout L = make_layout(c);
    const int o_ = tid & 63, kg_ = tid >> 6;
    const float dsc = dropout_scale(c);
    const bool drop = c.dropout_p > 0.f;
    const uint32_t thr = dropout_thr(c);
    __shared__ float task_sum[2];
    PhaseClock pc;
    pc.start(P.phase);
    int e_w1 = 0, e_s = 0, e_h0 = 0, e_dz = 0, e_gs = 0;              // plane exponents of the current W1^T, S, H0, dZ1, G_support

    // planes start as zeros: pad rows / columns that no phase writes must be finite
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(s.w1h);
        const int nz = int((reinterpret_cast<char*>(smem_raw) + smem_total - reinterpret_cast<char*>(s.w1h)) / 4);
        for (int idx = tid; idx < nz; idx += NT_) z[idx] = 0u;
    }
    __syncthreads();

    // the H0 epilogue shared by support steps and query tiles: acc (raw G.S product) -> activations in acc, planes
    // written after a block-wide max.  rows: global bank rows of the tile's 32 rows (smem), gscale: 2^-(sG + sS)
    // projected rows at this thread's accumulator positions, straight from `proj` (issued a phase ahead of their use)
    float2 ap[2][2][2];
    auto h0_load = [&](const long long* rows, int tr) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i == 0 || tr > 16)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                    const int r = 16 * i + g + 8 * hq, h = 16 * w + 8 * j + 2 * t;
                    ap[i][j][hq] = r < tr ? __ldg(reinterpret_cast<const float2*>(&P.proj[rows[r] * kH0 + h])) : make_float2(0.f, 0.f);
                }
    };
    auto h0_tile = [&](const fumi_half* gh, const fumi_half* gl, int r0, int tr, bool use_s, float ginv, int pass, int64_t task,
                       int prod) {
        float acc[2][2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
        const bool two = tr > 16;                           // rows 16..31 in use (support: NK > 16; last query tile may be short)
        if (use_s) {
            if (two) warp_gemm_f16x3<2, 2, false, false>(gh, gl, kHG, s.sh + 16 * w, s.sl + 16 * w, kHS, 32, acc);
            else warp_gemm_f16x3<1, 2, false, false>(gh, gl, kHG, s.sh + 16 * w, s.sl + 16 * w, kHS, MT == 1 ? 16 : 32,
                                                     reinterpret_cast<float(&)[1][2][4]>(acc));
        }
        const uint32_t dbase = drop ? dropout_base(c, task, pass, 0) : 0u;
        const float gs = use_s ? alpha * ginv : 0.f;
        float b0v[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            b0v[j][0] = s.b0s[16 * w + 8 * j + 2 * t];
            b0v[j][1] = s.b0s[16 * w + 8 * j + 2 * t + 1];
        }
        float mxv = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i == 0 || two)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {                 // one column pair (h, h + 1) of row r
                    const int r = 16 * i + g + 8 * hq, h = 16 * w + 8 * j + 2 * t;
                    const float z0 = ap[i][j][hq].x + b0v[j][0] - gs * acc[i][j][2 * hq];
                    const float z1 = ap[i][j][hq].y + b0v[j][1] - gs * acc[i][j][2 * hq + 1];
                    bool k0 = r < tr && z0 > 0.f, k1 = r < tr && z1 > 0.f;
                    if (drop) {
                        const uint32_t bits = dropout_bits(dbase, r0 + r, h);
                        k0 = k0 & ((bits & 0xFFFFu) >= thr);
                        k1 = k1 & ((bits >> 16) >= thr);
                    }
                    const float v0 = k0 ? z0 * dsc : 0.f, v1 = k1 ? z1 * dsc : 0.f;
                    acc[i][j][2 * hq] = v0;
                    acc[i][j][2 * hq + 1] = v1;
                    mxv = fmaxf(mxv, fmaxf(v0, v1));
                }
        block_max_push(s.mx + 16 * (MX_H0 + (prod & 1)), mxv);
        __syncthreads();
        e_h0 = block_max_exp(s.mx + 16 * (MX_H0 + (prod & 1)));
        const float sc = fumi_exp2i(e_h0);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (i == 0 || two)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq)
                    store_pair(s.h0h, s.h0l, (16 * i + g + 8 * hq) * kHS + 16 * w + 8 * j + 2 * t, acc[i][j][2 * hq],
                               acc[i][j][2 * hq + 1], sc);
        __syncthreads();
    };
    // H1 = relu(H0 W1^T + b1) (dropout) -> fp32 tile; warp (m tile w / 8, n tile w % 8)
    auto h1_tile = [&](int r0, int tr, int mtiles, int pass, int64_t task, int prod_h0, int prod_w1) {
        const int nt = w & 7, mt = w >> 3;
        if (mt >= mtiles) return;
        float acc[1][1][4] = {{{0.f, 0.f, 0.f, 0.f}}};
        warp_gemm_f16x3<1, 1, false, false>(s.h0h + 16 * mt * kHS, s.h0l + 16 * mt * kHS, kHS, s.w1h + 8 * nt, s.w1l + 8 * nt, kHW,
                                            kH0, acc);
        const float inv = fumi_exp2i(-e_h0) * fumi_exp2i(-e_w1);
        const uint32_t dbase = drop ? dropout_base(c, task, pass, 1) : 0u;
        uint32_t bits = 0;
        warp_tile_foreach<1, 1>(acc, [&](int ii, int oo, float& cv) {
            const int i = 16 * mt + ii, o = 8 * nt + oo;
            float v = 0.f;
            if (i < tr) {
                const float z = cv * inv + s.b1s[o];
                if (drop && (oo & 1) == 0) bits = dropout_bits(dbase, r0 + i, o);
                v = ((z > 0.f) & (!drop | dropout_keep_bits(bits, o, thr))) ? z * dsc : 0.f;
            }
            s.h1t[i * kS1 + o] = v;
        });
    };

    for (int64_t b = blockIdx.x; b < P.B; b += gridDim.x) {
        const int64_t task = c.task_offset + b;
        float* slot = P.stash + (P.save ? b : int64_t(blockIdx.x)) * P.slot_floats;
        int prod_w1 = 0, prod_s = 0, prod_h0 = 0, prod_dz = 0;            // productions so far (slot parity)
        // ---- task prologue
        float w1v[32];                                                    // W1[o][h] for h = col, o in [32 half, 32 half + 32)
        {
            float mxv = 0.f;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                w1v[q] = __ldg(&P.w1[(half * 32 + q) * kH0 + col]);
                mxv = fmaxf(mxv, fabsf(w1v[q]));
            }
            block_max_push(s.mx + 16 * (MX_W1), mxv);
        }
        float gv[2];                                                      // support Gram block, <= 2 entries per thread
        {
            float mxv = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                gv[q] = (i < n && j < n) ? __ldg(&P.gram[(b * int64_t(n + m) + i) * n + j]) : 0.f;
                mxv = fmaxf(mxv, fabsf(gv[q]));
            }
            block_max_push(s.mx + 16 * (MX_GS), mxv);
        }
        if (tid < kH0) s.b0s[tid] = __ldg(&P.b0[tid]);
        if (tid < kH1) s.b1s[tid] = __ldg(&P.b1[tid]);
        for (int idx = tid; idx < N * kHD; idx += NT_) {
            const int cc = idx / kHD, o = idx - cc * kHD;
            const int64_t r = P.head_rows ? __ldg(&P.head_rows[b * N + cc]) : cc;
            s.hp[idx] = __ldg(&P.head_table[r * kHD + o]);
        }
        if (tid < 32) {
            s.ysS[tid] = tid < n ? int(P.sup_y[b * n + tid]) : 0;
            s.rowsS[tid] = tid < n ? P.sup_rows[b * n + tid] : 0;
        }
        for (int idx = tid; idx < m; idx += NT_) {
            s.rowsQ[idx] = P.qry_rows[b * m + idx];
            s.ysQ[idx] = int(P.qry_y[b * m + idx]);
        }
        if (tid == 0) { task_sum[0] = 0.f; task_sum[1] = 0.f; }
        __syncthreads();
        {
            e_w1 = block_max_exp(s.mx + 16 * MX_W1);
            const float sc = fumi_exp2i(e_w1);
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {                    // 8 halves = 16 bytes per store, rows 144 bytes apart
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float a = w1v[8 * q8 + 2 * q] * sc, bb = w1v[8 * q8 + 2 * q + 1] * sc;
                    const fumi_half ah = fumi_f2h(a), bh = fumi_f2h(bb);
                    const fumi_half al = fumi_f2h(a - fumi_h2f(ah)), bl = fumi_f2h(bb - fumi_h2f(bh));
                    ph[q] = uint32_t(*reinterpret_cast<const uint16_t*>(&ah)) | (uint32_t(*reinterpret_cast<const uint16_t*>(&bh)) << 16);
                    pl[q] = uint32_t(*reinterpret_cast<const uint16_t*>(&al)) | (uint32_t(*reinterpret_cast<const uint16_t*>(&bl)) << 16);
                }
                *reinterpret_cast<uint4*>(&s.w1h[col * kHW + half * 32 + 8 * q8]) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(&s.w1l[col * kHW + half * 32 + 8 * q8]) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
            e_gs = block_max_exp(s.mx + 16 * MX_GS);
            const float sg = fumi_exp2i(e_gs);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                const float v = gv[q] * sg;
                const fumi_half hh = fumi_f2h(v);
                s.gsh[i * kHG + j] = hh;
                s.gsl[i * kHG + j] = fumi_f2h(v - fumi_h2f(hh));
            }
        }
        if (steps > 0) h0_load(s.rowsS, n);
        __syncthreads();
        pc.mark(20);    // prologue

        for (int st = 0; st < steps; ++st) {
            float* rec = P.save ? slot + L.steps + int64_t(st) * L.per_step : nullptr;
            // ---- H0 = relu(A + b0 - alpha G S) (dropout): planes
            {
                const float ginv = st > 0 ? fumi_exp2i(-e_gs) * fumi_exp2i(-e_s) : 0.f;
                h0_tile(s.gsh, s.gsl, 0, n, st > 0, ginv, st, task, prod_h0);
                ++prod_h0;
            }
            pc.mark(21);    // s: H0
            h1_tile(0, n, MT, st, task, prod_h0 - 1, prod_w1);
            __syncthreads();
            pc.mark(22);    // s: H1 (K=256)
            {
                const float invn = 1.f / float(n);
                m16_tile_logits_softmax(P, s, RS, n, [&](int i, int cc, float l, float mx, float sum) {
                    float dl = 0.f;
                    if (i < n) dl = (expf(l - mx) * (1.f / sum) - (cc == s.ysS[i] ? 1.f : 0.f)) * invn;
                    s.lt[i * kLS + cc] = dl;
                });
            }
            __syncthreads();
            pc.mark(23);    // s: logits + softmax
            // ---- head gradient; dZ1 (fp32 copy for the stash / column sums, planes for the two GEMMs)
            for (int idx = tid; idx < N * kHD; idx += NT_) {
                const int cc = idx / kHD, o = idx - cc * kHD;
                float a = 0.f;
#pragma unroll 2
                for (int i = 0; i < n; ++i) a = fmaf(s.lt[i * kLS + cc], o < kH1 ? s.h1t[i * kS1 + o] : 1.f, a);
                s.dhp[idx] = a;
            }
            float dzv[4];
            {
                float mxv = 0.f;
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int i = kg_ + 8 * ii;
                    float dz = 0.f;
                    if (i < n && s.h1t[i * kS1 + o_] > 0.f) {
                        float dh = 0.f;
#pragma unroll 1
                        for (int cc = 0; cc < N; ++cc) dh = fmaf(s.lt[i * kLS + cc], s.hp[cc * kHD + o_], dh);
                        dz = dh * dsc;
                    }
                    dzv[ii] = dz;
                    s.dz1t[i * kS1 + o_] = dz;
                    mxv = fmaxf(mxv, fabsf(dz));
                }
                block_max_push(s.mx + 16 * (MX_DZ + (prod_dz & 1)), mxv);
            }
            __syncthreads();
            {
                e_dz = block_max_exp(s.mx + 16 * (MX_DZ + (prod_dz & 1)));
                const float sc = fumi_exp2i(e_dz);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const float v = dzv[ii] * sc;
                    const fumi_half hh = fumi_f2h(v);
                    s.dzh[(kg_ + 8 * ii) * kHW + o_] = hh;
                    s.dzl[(kg_ + 8 * ii) * kHW + o_] = fumi_f2h(v - fumi_h2f(hh));
                }
                ++prod_dz;
            }
            __syncthreads();
            pc.mark(25);    // s: dhp, dZ1
            float db1 = 0.f;
            if (tid < kH1) for (int i = 0; i < n; ++i) db1 += s.dz1t[i * kS1 + tid];
            // ---- dZ0 = (dZ1 W1) * gate ;  S += dZ0 ;  db0 = column sums of dZ0
            {
                float acc[2][2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
                warp_gemm_f16x3<MT, 2, false, true>(s.dzh, s.dzl, kHW, s.w1h + 16 * w * kHW, s.w1l + 16 * w * kHW, kHW, kH1,
                                                    reinterpret_cast<float(&)[MT][2][4]>(acc));
                const float inv = fumi_exp2i(-e_dz) * fumi_exp2i(-e_w1);
                const float sinv = st > 0 ? fumi_exp2i(-e_s) : 0.f;
                float colsum[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
                float mxv = 0.f;
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            const int r = 16 * i + g + 8 * hq, off = r * kHS + 16 * w + 8 * j + 2 * t;
                            float h0a, h0b, s0 = 0.f, s1 = 0.f;
                            fumi_plane_load2(s.h0h, s.h0l, off, 1.f, h0a, h0b);                 // only the sign is used
                            if (st > 0 && r < n) fumi_plane_load2(s.sh, s.sl, off, sinv, s0, s1);
                            const float d0 = (r < n && h0a > 0.f) ? acc[i][j][2 * hq] * inv * dsc : 0.f;
                            const float d1 = (r < n && h0b > 0.f) ? acc[i][j][2 * hq + 1] * inv * dsc : 0.f;
                            colsum[j][0] += d0;
                            colsum[j][1] += d1;
                            acc[i][j][2 * hq] = s0 + d0;
                            acc[i][j][2 * hq + 1] = s1 + d1;
                            mxv = fmaxf(mxv, fmaxf(fabsf(s0 + d0), fabsf(s1 + d1)));
                        }
                block_max_push(s.mx + 16 * (MX_S + (prod_s & 1)), mxv);
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float v = colsum[j][q];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if ((lane >> 2) == 0) s.db0s[16 * w + 8 * j + 2 * (lane & 3) + q] = v;
                    }
                if (rec) {                                          // records for the backward (H0 from its planes)
                    const float hinv = fumi_exp2i(-e_h0);
                    for (int i = tid >> 7; i < n; i += 4) {           // two columns per thread
                        float a0, a1;
                        fumi_plane_load2(s.h0h, s.h0l, i * kHS + 2 * (tid & 127), hinv, a0, a1);
                        *reinterpret_cast<float2*>(&rec[L.oH0 + int64_t(i) * kH0 + 2 * (tid & 127)]) = make_float2(a0, a1);
                    }
                    for (int idx = tid; idx < n * kH1; idx += NT_) {
                        const int i = idx / kH1, o = idx - i * kH1;
                        rec[L.oH1 + idx] = s.h1t[i * kS1 + o];
                        rec[L.oDZ1 + idx] = s.dz1t[i * kS1 + o];
                    }
                    for (int idx = tid; idx < n * N; idx += NT_) {
                        const int i = idx / N, cc = idx - i * N;
                        rec[L.oDL + idx] = s.lt[i * kLS + cc];
                    }
                    for (int idx = tid; idx < N * kHD; idx += NT_) rec[L.oHP + idx] = s.hp[idx];
                }
                __syncthreads();                                    // all reads of the old S planes are done
                e_s = block_max_exp(s.mx + 16 * (MX_S + (prod_s & 1)));
                const float sc = fumi_exp2i(e_s);
#pragma unroll
                for (int i = 0; i < MT; ++i)                  // NK <= 16: rows 16..31 of the S planes stay zero
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq)
                            store_pair(s.sh, s.sl, (16 * i + g + 8 * hq) * kHS + 16 * w + 8 * j + 2 * t, acc[i][j][2 * hq],
                                       acc[i][j][2 * hq + 1], sc);
                ++prod_s;
            }
            pc.mark(26);    // s: dZ0 gemm, S update, stash
            // ---- W1 -= alpha dZ1^T H0 (rows h of W1^T owned by this warp), in place on the planes
            {
                float acc[1][8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
                warp_gemm_f16x3<1, 8, true, false>(s.h0h + 16 * w, s.h0l + 16 * w, kHS, s.dzh, s.dzl, kHW, RS, acc);
                const float inv = alpha * fumi_exp2i(-e_h0) * fumi_exp2i(-e_dz);
                const float winv = fumi_exp2i(-e_w1);
                float mxv = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        float o0, o1;
                        fumi_plane_load2(s.w1h, s.w1l, (16 * w + g + 8 * hq) * kHW + 8 * j + 2 * t, winv, o0, o1);
                        acc[0][j][2 * hq] = o0 - acc[0][j][2 * hq] * inv;
                        acc[0][j][2 * hq + 1] = o1 - acc[0][j][2 * hq + 1] * inv;
                        mxv = fmaxf(mxv, fmaxf(fabsf(acc[0][j][2 * hq]), fabsf(acc[0][j][2 * hq + 1])));
                    }
                block_max_push(s.mx + 16 * (MX_W1 + ((prod_w1 + 1) & 1)), mxv);
                __syncthreads();                                    // all reads of the old W1 planes are done
                e_w1 = block_max_exp(s.mx + 16 * (MX_W1 + ((prod_w1 + 1) & 1)));
                const float sc = fumi_exp2i(e_w1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    store_pair(s.w1h, s.w1l, (16 * w + g) * kHW + 8 * j + 2 * t, acc[0][j][0], acc[0][j][1], sc);
                    store_pair(s.w1h, s.w1l, (16 * w + g + 8) * kHW + 8 * j + 2 * t, acc[0][j][2], acc[0][j][3], sc);
                }
                ++prod_w1;
            }
            for (int idx = tid; idx < N * kHD; idx += NT_) s.hp[idx] -= alpha * s.dhp[idx];
            if (tid < kH1) s.b1s[tid] -= alpha * db1;
            if (tid < kH0) s.b0s[tid] -= alpha * s.db0s[tid];
            if (st + 1 < steps) h0_load(s.rowsS, n);            // the same rows every step; latency hides behind the barrier
            __syncthreads();
            pc.mark(27);    // s: W1 update gemm
        }

        // ---- query scoring, 32 rows per tile; the Gram tile of the next tile travels in registers, and its max is
        // pushed one tile ahead so that its plane scale needs no barrier of its own
        float qg[2];
        int qtile = 0;
        auto q_load = [&](int r0) {                              // loads only: the values are first touched a tile later
            const int tr = min(32, m - r0);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                qg[q] = (i < tr && j < n) ? __ldg(&P.gram[(b * int64_t(n + m) + n + r0 + i) * n + j]) : 0.f;
            }
        };
        q_load(0);
        block_max_push(s.mx + 16 * MX_GQ, fmaxf(fabsf(qg[0]), fabsf(qg[1])));
        h0_load(s.rowsQ, min(32, m));
        __syncthreads();
        const float sinv_fin = steps > 0 ? fumi_exp2i(-e_s) : 0.f;
        for (int r0 = 0; r0 < m; r0 += 32, ++qtile) {
            const int tr = min(32, m - r0);
            const int gexp = block_max_exp(s.mx + 16 * (MX_GQ + (qtile & 1)));
            {
                const float sg = fumi_exp2i(gexp);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                    const float v = qg[q] * sg;
                    const fumi_half hh = fumi_f2h(v);
                    s.gqh[i * kHG + j] = hh;
                    s.gql[i * kHG + j] = fumi_f2h(v - fumi_h2f(hh));
                }
            }
            __syncthreads();
            if (r0 > 0 && w == 0) {                             // loss / accuracy of the previous tile (rowv, rowc)
                float rv = s.rowv[lane], rc = s.rowc[lane];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    rv += __shfl_xor_sync(0xffffffffu, rv, off);
                    rc += __shfl_xor_sync(0xffffffffu, rc, off);
                }
                if (lane == 0) { task_sum[0] += rv; task_sum[1] += rc; }
            }
            pc.mark(28);    // q: loads
            h0_tile(s.gqh, s.gql, r0, tr, steps > 0, fumi_exp2i(-gexp) * sinv_fin, steps, task, prod_h0);
            ++prod_h0;
            if (r0 + 32 < m) h0_load(s.rowsQ + r0 + 32, min(32, m - r0 - 32));     // next tile's rows, a tile ahead
            if (r0 + 32 < m) q_load(r0 + 32);                   // issued here, consumed at the end of the tile
            pc.mark(29);    // q: H0
            h1_tile(r0, tr, 2, steps, task, prod_h0 - 1, prod_w1);
            __syncthreads();
            pc.mark(30);    // q: H1
            m16_tile_logits_softmax(P, s, 32, tr, [&](int i, int cc, float l, float mx, float sum) {
                if (i >= tr) {
                    if (cc == 0) { s.rowv[i] = 0.f; s.rowc[i] = 0.f; }
                    return;
                }
                const int y = s.ysQ[r0 + i];
                const int64_t q = b * m + r0 + i;
                P.logits[q * N + cc] = l;
                if (P.save)
                    slot[L.qLG + int64_t(r0 + i) * N + cc] = expf(l - mx) * (1.f / sum) - (cc == y ? 1.f : 0.f);
                if (cc == 0) {
                    const float* lr = &s.lt[i * kLS];
                    int best = 0;
#pragma unroll 1
                    for (int k = 1; k < N; ++k) if (lr[k] > lr[best]) best = k;      // first max (torch.max)
                    s.rowv[i] = (logf(sum) + mx) - lr[y];
                    s.rowc[i] = best == y ? 1.f : 0.f;
                    P.preds[q] = best;
                }
            });
            pc.mark(31);    // q: logits + softmax
            if (P.save) {
                const float hinv = fumi_exp2i(-e_h0);
                for (int i = tid >> 7; i < tr; i += 4) {
                    float a0, a1;
                    fumi_plane_load2(s.h0h, s.h0l, i * kHS + 2 * (tid & 127), hinv, a0, a1);
                    *reinterpret_cast<float2*>(&slot[L.qH0 + int64_t(r0 + i) * kH0 + 2 * (tid & 127)]) = make_float2(a0, a1);
                }
                for (int idx = tid; idx < tr * kH1; idx += NT_) {
                    const int i = idx / kH1, o = idx - i * kH1;
                    slot[L.qH1 + int64_t(r0) * kH1 + idx] = s.h1t[i * kS1 + o];
                }
            }
            if (r0 + 32 < m) block_max_push(s.mx + 16 * (MX_GQ + ((qtile + 1) & 1)), fmaxf(fabsf(qg[0]), fabsf(qg[1])));
            __syncthreads();
            pc.mark(32);    // q: stash, loss
        }
        if (w == 0) {
            float rv = s.rowv[lane], rc = s.rowc[lane];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                rv += __shfl_xor_sync(0xffffffffu, rv, off);
                rc += __shfl_xor_sync(0xffffffffu, rc, off);
            }
            if (lane == 0) {
                P.task_loss[b] = (task_sum[0] + rv) / float(m);
                P.task_acc[b] = (task_sum[1] + rc) / float(m);
            }
        }
        if (P.save) {                                               // adapted state (fp32, as the backward expects it)
            const float winv = fumi_exp2i(-e_w1);
            for (int idx = tid; idx < kH0 * kH1; idx += NT_) {
                const int k = idx / kH1, o = idx - k * kH1;
                slot[L.w1t + idx] = plane_value(s.w1h, s.w1l, k * kHW + o, winv);
            }
            if (tid < kH0) slot[L.b0 + tid] = s.b0s[tid];
            if (tid < kH1) slot[L.b1 + tid] = s.b1s[tid];
            for (int idx = tid; idx < N * kHD; idx += NT_) slot[L.head + idx] = s.hp[idx];
            float* Sout = slot + ((steps & 1) ? L.S1 : L.S0);
            for (int i = half; i < n; i += 2)
                Sout[int64_t(i) * kH0 + col] = steps > 0 ? plane_value(s.sh, s.sl, i * kHS + col, sinv_fin) : 0.f;
        }
        __syncthreads();
        pc.mark(33);    // epilogue
    }
}

// ------------------------------------------------------------------------------------ backward (tensor core)
// Reverse sweep for NK <= 32 on warp-level 3xTF32 tiles.  Same recursion as episode_bwd_kernel /
// oracle/episode_np.py; differences: rows go through in tiles of 16 (MT = 1) because W1^T and its adjoint
// (2 x 68 KB) share the SM's shared memory with the tiles; the query activations come from the forward's
// stash instead of being recomputed; the adjoint of S lives in the task's (L2-resident) workspace slot.
struct SmemB {
    float *w1t, *aw1t, *h0t, *tt, *h1t, *dz1t, *rzh, *lt, *rlt, *gS, *gQ, *hp, *ahp, *rhp, *b1s, *ab1, *rb1, *ab0s, *rb0s,
        *gb0s;
    long long* rows;
    int* ys;
};
__host__ __device__ inline size_t smem_b_floats() {
    return 2 * size_t(kH0) * kS1 + 2 * 16 * kS0 + 3 * 16 * kS1 + 2 * 16 * kLS + 32 * kSG + 16 * kSG + 3 * kMaxWays * kHD +
           3 * kH1 + 3 * kH0 + 32 + 16 + 16;
}
__device__ inline SmemB carve_b(float* p) {
    SmemB s;
    s.rows = reinterpret_cast<long long*>(p); p += 32;
    s.w1t = p; p += kH0 * kS1;
    s.aw1t = p; p += kH0 * kS1;
    s.h0t = p; p += 16 * kS0;          // h0t and tt are adjacent: together a [32][kS0] buffer
    s.tt = p; p += 16 * kS0;
    s.dz1t = p; p += 16 * kS1;         // dz1t and rzh are adjacent: together a [32][kS1] buffer
    s.rzh = p; p += 16 * kS1;
    s.h1t = p; p += 16 * kS1;
    s.lt = p; p += 16 * kLS;
    s.rlt = p; p += 16 * kLS;
    s.gS = p; p += 32 * kSG;
    s.gQ = p; p += 16 * kSG;
    s.hp = p; p += kMaxWays * kHD;
    s.ahp = p; p += kMaxWays * kHD;
    s.rhp = p; p += kMaxWays * kHD;
    s.b1s = p; p += kH1;
    s.ab1 = p; p += kH1;
    s.rb1 = p; p += kH1;
    s.ab0s = p; p += kH0;
    s.rb0s = p; p += kH0;
    s.gb0s = p; p += kH0;
    s.ys = reinterpret_cast<int*>(p); p += 16;
    return s;
}


// ------------------------------------------------------------------------------------ backward, 16 warps
// 512 threads.  Warp w owns rows [16w, 16w+16) of the
// W1^T-shaped adjoints, hidden units [16w, 16w+16) of the 256-wide ops, and in the 64-wide reductions over
// the 256 hidden units the (n tile w%8, K half w/8) block; the two K halves meet in shared memory.
template <int MT>
__device__ __forceinline__ void slab16_colsum(float (&acc)[MT][2][4], float* dst, bool accumulate) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < MT; ++i) v += acc[i][j][q] + acc[i][j][q + 2];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if ((lane >> 2) == 0) {
                float* d = dst + 16 * w + 8 * j + 2 * (lane & 3) + q;
                *d = accumulate ? *d + v : v;
            }
        }
}

__global__ void __launch_bounds__(kThreads16, 1) episode_bwd_mma16_kernel(EpiParams P) {
    constexpr int NT_ = kThreads16;
    FUMI_DYN_SMEM(float, smem_raw);
    const SmemB s = carve_b(smem_raw);
    const fumi_episode_cfg& c = P.cfg;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int col = tid & 255, half = tid >> 8;
    const int n = c.num_support, m = c.num_query, N = c.num_ways, steps = c.steps;
    const float alpha = c.step_size;
    const Layout L = make_layout(c);
    const int o_ = tid & 63, kg_ = tid >> 6;          // 8 row groups x 64 outputs
    const float sc = dropout_scale(c);
    PhaseClock pc;
    pc.start(P.phase);

    for (int idx = tid; idx < 32 * kSG; idx += NT_) s.gS[idx] = 0.f;
    for (int idx = tid; idx < 16 * kSG; idx += NT_) s.gQ[idx] = 0.f;
    __syncthreads();

    for (int64_t b = blockIdx.x; b < P.B; b += gridDim.x) {
        float* slot = P.stash + b * P.slot_floats;
        float* aS = slot + L.S0;           // adjoint of S  [n][H0]
        float* bZ = slot + L.S1;           // bar_Z0 rows of the current step [n][H0]
        // ---- adapted state, zeroed adjoints
        {   // W1_S^T [H0][H1] from the stash: coalesced 16-byte loads, 8 per thread in flight
            float4 v[8];
            const float4* src = reinterpret_cast<const float4*>(slot + L.w1t);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __ldg(&src[tid + q * NT_]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int idx = tid + q * NT_, k = idx >> 4, o = (idx & 15) * 4;
                *reinterpret_cast<float4*>(&s.w1t[k * kS1 + o]) = v[q];
                *reinterpret_cast<float4*>(&s.aw1t[k * kS1 + o]) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (tid < kH0) s.ab0s[tid] = 0.f;
        if (tid < kH1) { s.b1s[tid] = slot[L.b1 + tid]; s.ab1[tid] = 0.f; }
        for (int idx = tid; idx < N * kHD; idx += NT_) { s.hp[idx] = slot[L.head + idx]; s.ahp[idx] = 0.f; }
        for (int idx = tid; idx < n * n; idx += NT_) {
            const int i = idx / n, j = idx - i * n;
            s.gS[i * kSG + j] = __ldg(&P.gram[(b * int64_t(n + m) + i) * n + j]);
        }
        __syncthreads();
        pc.mark(0);     // prologue

        // ---- query pass (activations from the forward's stash)
        const float qscale = P.loss_scale / float(m);
        float aSq[2][2][4];                  // Gq^T dZ0q summed over the query tiles (this warp's 16 columns)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) aSq[i][j][q] = 0.f;
        float aWq[1][8][4];                  // dZ1q^T H0q summed over the query tiles (this warp's 16 rows of W1^T)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) aWq[0][j][q] = 0.f;
        // software prefetch: the next tile's stash rows travel in registers while the current tile is processed
        float pv[8], pu[2], pg = 0.f, pl = 0.f;
        long long prow = 0;
        int py = 0;
        auto q_load = [&](int r0) {
            const int tr = min(16, m - r0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = half * 8 + q;
                pv[q] = i < tr ? __ldg(&slot[L.qH0 + int64_t(r0 + i) * kH0 + col]) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_;
                pu[q] = idx < tr * kH1 ? __ldg(&slot[L.qH1 + int64_t(r0) * kH1 + idx]) : 0.f;
            }
            {
                const int i = tid / n, j = tid - i * n;
                pg = (tid < 16 * n && i < tr) ? __ldg(&P.gram[(b * int64_t(n + m) + n + r0 + i) * n + j]) : 0.f;
            }
            pl = tid < tr * N ? __ldg(&slot[L.qLG + int64_t(r0) * N + tid]) : 0.f;
            if (tid < 16) {
                prow = tid < tr ? P.qry_rows[b * m + r0 + tid] : 0;
                py = tid < tr ? int(P.qry_y[b * m + r0 + tid]) : 0;
            }
        };
        q_load(0);
        for (int r0 = 0; r0 < m; r0 += 16) {
            const int tr = min(16, m - r0);
#pragma unroll
            for (int q = 0; q < 8; ++q) s.h0t[(half * 8 + q) * kS0 + col] = pv[q];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_;
                s.h1t[(idx / kH1) * kS1 + (idx % kH1)] = pu[q];
            }
            if (tid < 16 * n) s.gQ[(tid / n) * kSG + (tid % n)] = pg;
            if (tid < 16 * N) s.lt[(tid / N) * kLS + (tid % N)] = pl * qscale;     // dLq (softmax - onehot from the forward)
            if (tid < 16) { s.rows[tid] = prow; s.ys[tid] = py; }
            __syncthreads();
            if (r0 + 16 < m) q_load(r0 + 16);
            pc.mark(1);     // q: tile loads
            for (int idx = tid; idx < N * kHD; idx += NT_) {              // a_head += dLq^T [H1q | 1]
                const int cc = idx / kHD, o = idx - cc * kHD;
                float a = 0.f;
                for (int i = 0; i < tr; ++i) a = fmaf(s.lt[i * kLS + cc], o < kH1 ? s.h1t[i * kS1 + o] : 1.f, a);
                s.ahp[idx] += a;
            }
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {                                   // dZ1q
                const int i = kg_ + 8 * ii;
                float dz = 0.f;
                if (i < tr && s.h1t[i * kS1 + o_] > 0.f) {
                    float dh = 0.f;
                    for (int cc = 0; cc < N; ++cc) dh = fmaf(s.lt[i * kLS + cc], s.hp[cc * kHD + o_], dh);
                    dz = dh * sc;
                }
                s.dz1t[i * kS1 + o_] = dz;
            }
            __syncthreads();
            pc.mark(2);     // q: a_head, dZ1q
            if (tid < kH1) {
                float a = 0.f;
                for (int i = 0; i < tr; ++i) a += s.dz1t[i * kS1 + tid];
                s.ab1[tid] += a;
            }
            // a_W1 += dZ1q^T H0q   (rows h of W1^T owned by this warp): summed in registers over the query tiles
            warp_gemm_3xtf32<1, 8, true, false, 0>(s.h0t + 16 * w, kS0, s.dz1t, kS1, 16, 1.f, aWq);
            {   // dZ0q = (dZ1q W1) * gate  -> tt, d_proj, a_b0
                float acc[1][2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
                warp_gemm_3xtf32<1, 2, false, true>(s.dz1t, kS1, s.w1t + 16 * w * kS1, kS1, kH1, 1.f, acc);
                warp_tile_foreach<1, 2>(acc, [&](int i, int hh, float& cv) {
                    const int h = 16 * w + hh;
                    cv = (i < tr && s.h0t[i * kS0 + h] > 0.f) ? cv * sc : 0.f;
                    s.tt[i * kS0 + h] = cv;
                });
                const int g = lane >> 2, t = lane & 3;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int h = 16 * w + 8 * j + 2 * t;
                    if (g < tr) atomic_add2(&P.d_proj[s.rows[g] * kH0 + h], acc[0][j][0], acc[0][j][1]);
                    if (g + 8 < tr) atomic_add2(&P.d_proj[s.rows[g + 8] * kH0 + h], acc[0][j][2], acc[0][j][3]);
                }
                slab16_colsum<1>(acc, s.ab0s, true);
            }
            __syncthreads();
            pc.mark(3);     // q: a_W1 gemm, dZ0q gemm, atomics
            if (steps > 0)     // a_S -= alpha * Gq^T dZ0q  (kept in registers until the last query tile)
                warp_gemm_3xtf32<2, 2, true, false>(s.gQ, kSG, s.tt + 16 * w, kS0, 16, 1.f, aSq);
            __syncthreads();
            pc.mark(4);     // q: a_S gemm
        }
        if (steps > 0) {
            warp_tile_foreach<2, 2>(aSq, [&](int j, int hh, float& cv) {
                if (j < n) aS[int64_t(j) * kH0 + 16 * w + hh] = -alpha * cv;          // a_S starts from zero
            });
        }
        warp_tile_foreach<1, 8>(aWq, [&](int hh, int o, float& cv) { s.aw1t[(16 * w + hh) * kS1 + o] = cv; });   // was zero
        __syncthreads();

        // ---- inner steps in reverse
        if (!c.first_order) {
            for (int st = steps - 1; st >= 0; --st) {
                const float* rec = slot + L.steps + int64_t(st) * L.per_step;
                for (int idx = tid; idx < N * kHD; idx += NT_) { s.hp[idx] = rec[L.oHP + idx]; s.rhp[idx] = 0.f; }
                if (tid < kH1) s.rb1[tid] = 0.f;
                if (tid < kH0) { s.rb0s[tid] = 0.f; s.gb0s[tid] = -alpha * s.ab0s[tid]; }
                {   // all rows at once: (h0t|tt) is a [32][kS0] buffer and (dz1t|rzh) a [32][kS1] buffer
                    float v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int i = half * 16 + q;
                        v[q] = i < n ? __ldg(&rec[L.oH0 + int64_t(i) * kH0 + col]) : 0.f;
                    }
                    float u[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int idx = tid + q * NT_;
                        u[q] = idx < n * kH1 ? __ldg(&rec[L.oDZ1 + idx]) : 0.f;
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) s.h0t[(half * 16 + q) * kS0 + col] = v[q];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int idx = tid + q * NT_;
                        s.dz1t[(idx / kH1) * kS1 + (idx % kH1)] = u[q];
                    }
                }
                // software prefetch of the row tiles of this step (registers), first tile issued before the undo GEMM
                float tv[8], ta[8], tu[2], td[2], tl = 0.f;
                long long trow = 0;
                int ty = 0;
                auto t_load = [&](int r0) {
                    const int tr = min(16, n - r0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int i = half * 8 + q;
                        tv[q] = i < tr ? __ldg(&rec[L.oH0 + int64_t(r0 + i) * kH0 + col]) : 0.f;
                        ta[q] = i < tr ? aS[int64_t(r0 + i) * kH0 + col] : 0.f;
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int idx = tid + q * NT_;
                        tu[q] = idx < tr * kH1 ? __ldg(&rec[L.oH1 + int64_t(r0) * kH1 + idx]) : 0.f;
                        td[q] = idx < tr * kH1 ? __ldg(&rec[L.oDZ1 + int64_t(r0) * kH1 + idx]) : 0.f;
                    }
                    {
                        const int i = tid / kLS, cc = tid - i * kLS;
                        tl = (i < tr && cc < N) ? __ldg(&rec[L.oDL + int64_t(r0 + i) * N + cc]) : 0.f;
                    }
                    if (tid < 16) {
                        trow = tid < tr ? P.sup_rows[b * n + r0 + tid] : 0;
                        ty = tid < tr ? int(P.sup_y[b * n + r0 + tid]) : 0;
                    }
                };
                __syncthreads();
                t_load(0);
                {   // undo W1_{s+1} = W1_s - alpha dZ1^T H0
                    float acc[1][8][4];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
                    warp_gemm_3xtf32<1, 8, true, false, 0>(s.h0t + 16 * w, kS0, s.dz1t, kS1, 32, 1.f, acc);
                    warp_tile_foreach<1, 8>(acc, [&](int hh, int o, float& cv) { s.w1t[(16 * w + hh) * kS1 + o] += alpha * cv; });
                }
                float rw[1][8][4];                       // this step's contribution to a_W1 (own rows), over both tiles
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) rw[0][j][q] = 0.f;
                __syncthreads();
                pc.mark(5);     // s: loads + undo W1

                for (int r0 = 0; r0 < n; r0 += 16) {
                    const int tr = min(16, n - r0);
                    {
                        const float gb0 = s.gb0s[col];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int i = half * 8 + q;
                            s.h0t[i * kS0 + col] = tv[q];
                            s.tt[i * kS0 + col] = (i < tr && tv[q] > 0.f) ? (ta[q] + gb0) * sc : 0.f;      // (12r) r_dH0
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int idx = tid + q * NT_;
                            s.h1t[(idx / kH1) * kS1 + (idx % kH1)] = tu[q];
                            s.dz1t[(idx / kH1) * kS1 + (idx % kH1)] = td[q];
                        }
                        s.lt[tid] = tl;                                  // 16 x kLS == 512 == one element per thread
                        if (tid < 16) { s.rows[tid] = trow; s.ys[tid] = ty; }
                    }
                    __syncthreads();
                    if (r0 + 16 < n) t_load(r0 + 16);
                    pc.mark(6);     // s: tile loads
                    // (11r)+(10r): r_dZ1 = r_dH0 W1^T + H0 (g_W1)^T over this warp's half of the 256 hidden units
                    float rz[1][1][4] = {{{0.f, 0.f, 0.f, 0.f}}};
                    {
                        const int nt = w & 7, kh = w >> 3;
                        warp_gemm_3xtf32<1, 1, false, false>(s.tt + 128 * kh, kS0, s.w1t + 128 * kh * kS1 + 8 * nt, kS1, 128,
                                                             1.f, rz);
                        warp_gemm_3xtf32<1, 1, false, false>(s.h0t + 128 * kh, kS0, s.aw1t + 128 * kh * kS1 + 8 * nt, kS1, 128,
                                                             -alpha, rz);
                        if (kh == 1)
                            warp_tile_foreach<1, 1>(rz, [&](int i, int oo, float& cv) { s.rzh[i * kS1 + 8 * nt + oo] = cv; });
                    }
                    // r_W1 += dZ1^T r_dH0 ;  r_H0 = dZ1 g_W1 (kept in registers)
                    warp_gemm_3xtf32<1, 8, true, false, 0>(s.tt + 16 * w, kS0, s.dz1t, kS1, 16, 1.f, rw);
                    float rh0[1][2][4];
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) rh0[0][j][q] = 0.f;
                    warp_gemm_3xtf32<1, 2, false, true>(s.dz1t, kS1, s.aw1t + 16 * w * kS1, kS1, kH1, -alpha, rh0);
                    __syncthreads();
                    if ((w >> 3) == 0) {                   // (9r) r_dH1 = gate1 * (both K halves + g_b1)
                        const int nt = w & 7;
                        warp_tile_foreach<1, 1>(rz, [&](int i, int oo, float& cv) {
                            const int o = 8 * nt + oo;
                            s.rzh[i * kS1 + o] = (i < tr && s.h1t[i * kS1 + o] > 0.f)
                                                     ? (cv + s.rzh[i * kS1 + o] - alpha * s.ab1[o]) * sc : 0.f;
                        });
                    }
                    __syncthreads();
                    pc.mark(7);     // s: r_dH1 gemms (K=256 x2), r_W1 gemm, r_H0 gemm
                    // (8r)-(4r) for one row per warp, no block barrier in between: the chain
                    //   r_dL = r_dH1 Wh^T + H1 (g_Wh)^T + g_bh -> r_L = P (r_dL - <P, r_dL>) / n ->
                    //   r_H1 = dL g_Wh + r_L Wh -> r_Z1 = r_H1 * gate1
                    // only couples the 64 hidden units and N classes of the same row.  Lane l owns hidden units l, l + 32
                    // and (l < N) class l.  r_Z1 goes to dz1t (free since the GEMMs above); rzh keeps r_dH1 and rlt gets
                    // r_L for the cross-row head sums of the next phase.
                    {
                        const int i = w;
                        const bool arow = i < tr;
                        const float rd0 = s.rzh[i * kS1 + lane], rd1 = s.rzh[i * kS1 + lane + 32];
                        const float h10 = s.h1t[i * kS1 + lane], h11 = s.h1t[i * kS1 + lane + 32];
                        float rdl = 0.f;                                    // r_dL of class `lane`
                        float ra0 = 0.f, ra1 = 0.f;                         // r_H1 of the two hidden units
#pragma unroll 1
                        for (int cc = 0; cc < N; ++cc) {
                            const float* wh = &s.hp[cc * kHD];
                            const float* ah = &s.ahp[cc * kHD];
                            float a = fmaf(rd0, wh[lane], rd1 * wh[lane + 32]) - alpha * fmaf(h10, ah[lane], h11 * ah[lane + 32]);
#pragma unroll
                            for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
                            if (lane == cc) rdl = a - alpha * ah[kH1];
                            const float dl = s.lt[i * kLS + cc];
                            ra0 = fmaf(dl, -alpha * ah[lane], ra0);
                            ra1 = fmaf(dl, -alpha * ah[lane + 32], ra1);
                        }
                        const float pl = (arow && lane < N) ? s.lt[i * kLS + lane] * float(n) + (lane == s.ys[i] ? 1.f : 0.f) : 0.f;
                        float dot = pl * rdl;                               // <P, r_dL>
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
                        const float rl = pl * (rdl - dot) / float(n);       // r_L of class `lane`
#pragma unroll 1
                        for (int cc = 0; cc < N; ++cc) {
                            const float rlc = __shfl_sync(0xffffffffu, rl, cc);
                            ra0 = fmaf(rlc, s.hp[cc * kHD + lane], ra0);
                            ra1 = fmaf(rlc, s.hp[cc * kHD + lane + 32], ra1);
                        }
                        if (lane < N) s.rlt[i * kLS + lane] = arow ? rl : 0.f;
                        s.dz1t[i * kS1 + lane] = (arow && h10 > 0.f) ? ra0 * sc : 0.f;
                        s.dz1t[i * kS1 + lane + 32] = (arow && h11 > 0.f) ? ra1 * sc : 0.f;
                    }
                    __syncthreads();
                    pc.mark(8);     // s: r_dL .. r_Z1 (one row per warp)
                    // cross-row sums: r_head += dL^T r_dH1 + r_L^T [H1 | 1]
                    for (int idx = tid; idx < N * kHD; idx += NT_) {
                        const int cc = idx / kHD, o = idx - cc * kHD;
                        float a = 0.f;
                        if (o < kH1) {
                            for (int i = 0; i < tr; ++i)
                                a = fmaf(s.lt[i * kLS + cc], s.rzh[i * kS1 + o], fmaf(s.rlt[i * kLS + cc], s.h1t[i * kS1 + o], a));
                        } else {
                            for (int i = 0; i < tr; ++i) a += s.rlt[i * kLS + cc];
                        }
                        s.rhp[idx] += a;
                    }
                    // (3r) r_H0 += r_Z1 W1 ; r_W1 += r_Z1^T H0 ; r_b1 += sum r_Z1
                    if (tid < kH1) {
                        float a = 0.f;
                        for (int i = 0; i < tr; ++i) a += s.dz1t[i * kS1 + tid];
                        s.rb1[tid] += a;
                    }
                    warp_gemm_3xtf32<1, 2, false, true>(s.dz1t, kS1, s.w1t + 16 * w * kS1, kS1, kH1, 1.f, rh0);
                    warp_gemm_3xtf32<1, 8, true, false, 0>(s.h0t + 16 * w, kS0, s.dz1t, kS1, 16, 1.f, rw);
                    // (2r) bar_Z0 = r_H0 * gate0 ; (1r) a_A (d_proj), a_b0 ; bar_Z0 rows parked in the workspace
                    {
                        const int g = lane >> 2, t = lane & 3;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int h = 16 * w + 8 * j + 2 * t;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int i = g + (q >> 1) * 8;
                                float& cv = rh0[0][j][q];
                                cv = (i < tr && s.h0t[i * kS0 + h + (q & 1)] > 0.f) ? cv * sc : 0.f;
                            }
                            if (g < tr) {
                                atomic_add2(&P.d_proj[s.rows[g] * kH0 + h], rh0[0][j][0], rh0[0][j][1]);
                                *reinterpret_cast<float2*>(&bZ[int64_t(r0 + g) * kH0 + h]) = make_float2(rh0[0][j][0], rh0[0][j][1]);
                            }
                            if (g + 8 < tr) {
                                atomic_add2(&P.d_proj[s.rows[g + 8] * kH0 + h], rh0[0][j][2], rh0[0][j][3]);
                                *reinterpret_cast<float2*>(&bZ[int64_t(r0 + g + 8) * kH0 + h]) =
                                    make_float2(rh0[0][j][2], rh0[0][j][3]);
                            }
                        }
                        slab16_colsum<1>(rh0, s.rb0s, true);
                    }
                    __syncthreads();
                    pc.mark(11);    // s: r_H0 gemm, r_W1 gemm, bar_Z0, atomics
                }
                // ---- end of reversed step: fold the contributions into the adjoints; a_S -= alpha G bar_Z0
                warp_tile_foreach<1, 8>(rw, [&](int hh, int o, float& cv) { s.aw1t[(16 * w + hh) * kS1 + o] += cv; });
                for (int idx = tid; idx < N * kHD; idx += NT_) s.ahp[idx] += s.rhp[idx];
                if (tid < kH1) s.ab1[tid] += s.rb1[tid];
                if (tid < kH0) s.ab0s[tid] += s.rb0s[tid];
                {
                    float v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int i = half * 16 + q;
                        v[q] = i < n ? bZ[int64_t(i) * kH0 + col] : 0.f;
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) s.h0t[(half * 16 + q) * kS0 + col] = v[q];
                }
                __syncthreads();
                pc.mark(12);    // s: fold adjoints, reload bar_Z0
                {
                    float acc[2][2][4];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
                    warp_gemm_3xtf32<2, 2, false, false>(s.gS, kSG, s.h0t + 16 * w, kS0, 32, 1.f, acc);
                    const int g = lane >> 2, t = lane & 3;
                    float2 old[2][2][2];                             // batched read-modify-write of a_S
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const int row = i * 16 + g + q * 8;
                                old[i][j][q] = row < n ? *reinterpret_cast<const float2*>(
                                                             &aS[int64_t(row) * kH0 + 16 * w + j * 8 + 2 * t])
                                                       : make_float2(0.f, 0.f);
                            }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const int row = i * 16 + g + q * 8;
                                if (row < n)
                                    *reinterpret_cast<float2*>(&aS[int64_t(row) * kH0 + 16 * w + j * 8 + 2 * t]) =
                                        make_float2(old[i][j][q].x - alpha * acc[i][j][2 * q],
                                                    old[i][j][q].y - alpha * acc[i][j][2 * q + 1]);
                            }
                }
                __syncthreads();
                pc.mark(13);    // s: a_S gemm
            }
        }

        // ---- task epilogue: head gradient per task, shared-parameter gradients into this CTA's partials
        for (int idx = tid; idx < N * kHD; idx += NT_) P.d_head[b * N * kHD + idx] = s.ahp[idx];
        float* pw = P.d_w1_parts + int64_t(blockIdx.x) * kH0 * kH1;
#pragma unroll 1
        for (int base = tid; base < kH0 * kH1; base += 8 * NT_) {               // 8 partial-sum loads in flight per thread
            float gsum[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) gsum[q] = pw[base + q * NT_];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int idx = base + q * NT_, o = idx / kH0, k = idx - o * kH0;   // [H1][H0] like linear1.weight
                pw[idx] = gsum[q] + s.aw1t[k * kS1 + o];
            }
        }
        if (tid < kH0) P.d_b0_parts[int64_t(blockIdx.x) * kH0 + tid] += s.ab0s[tid];
        if (tid < kH1) P.d_b1_parts[int64_t(blockIdx.x) * kH1 + tid] += s.ab1[tid];
        __syncthreads();
        pc.mark(14);    // epilogue
    }
}

// ------------------------------------------------------------------------------------ backward
// Reverse sweep through the unrolled inner loop with hand-derived Hessian-vector products
// (oracle/episode_np.py is the line-by-line CPU statement of the same recursion).
template <int TR>
__global__ void __launch_bounds__(kThreads, 1) episode_bwd_kernel(EpiParams P) {
    FUMI_DYN_SMEM(float, smem_raw);
    const Smem s = carve<TR, true, 0>(smem_raw);
    const fumi_episode_cfg& c = P.cfg;
    const int tid = threadIdx.x;
    const int n = c.num_support, m = c.num_query, N = c.num_ways, steps = c.steps;
    const float alpha = c.step_size;
    const Layout L = make_layout(c);
    const int o_ = tid & 63, kg_ = tid >> 6;
    const float sc = dropout_scale(c);
    constexpr int RPT = TR / 4;

    for (int64_t b = blockIdx.x; b < P.B; b += gridDim.x) {
        const int64_t task = c.task_offset + b;
        float* slot = P.stash + b * P.slot_floats;
        const int fin = steps & 1;
        const float* Sfin = steps > 0 ? slot + (fin ? L.S1 : L.S0) : nullptr;
        float* aS[2] = {slot + (fin ? L.S0 : L.S1), slot + (fin ? L.S1 : L.S0)};   // aS[1] aliases S_final
        // ---- load the adapted state, zero the adjoints
        for (int idx = tid; idx < kH0 * kH1; idx += kThreads) {
            const int k = idx / kH1, o = idx - k * kH1;
            s.w1t[k * kW1S + o] = slot[L.w1t + idx];
            s.aw1t[k * kW1S + o] = 0.f;
        }
        const float b0h = slot[L.b0 + tid];
        float ab0 = 0.f;
        if (tid < kH1) { s.b1s[tid] = slot[L.b1 + tid]; s.ab1[tid] = 0.f; s.rb1[tid] = 0.f; }
        for (int idx = tid; idx < N * kHD; idx += kThreads) { s.hp[idx] = slot[L.head + idx]; s.dhp[idx] = 0.f; s.rhp[idx] = 0.f; }
        for (int j = 0; j < n; ++j) aS[0][int64_t(j) * kH0 + tid] = 0.f;
        __syncthreads();

        // ---- query pass: recompute the forward tile, then its backward
        const float qscale = P.loss_scale / float(m);
        for (int r0 = 0; r0 < m; r0 += TR) {
            const int tr = min(TR, m - r0);
            load_tile_meta<TR>(P, s, b, true, r0, tr);
            __syncthreads();
            tile_h0<TR, false>(P, s, task, Sfin, b0h, r0, tr, steps);
            __syncthreads();
            tile_h1<TR>(P, s, task, r0, tr, steps);
            __syncthreads();
            tile_logits<TR>(P, s, tr);
            __syncthreads();
            if (tid < TR) {
                float* l = &s.lt[tid * kLS];
                if (tid < tr) {
                    float mx, sum;
                    row_softmax(l, N, mx, sum);
                    const float inv = 1.f / sum;
                    const int y = s.ys[tid];
                    for (int cc = 0; cc < N; ++cc) l[cc] = (expf(l[cc] - mx) * inv - (cc == y ? 1.f : 0.f)) * qscale;
                } else {
                    for (int cc = 0; cc < N; ++cc) l[cc] = 0.f;
                }
            }
            __syncthreads();
            for (int idx = tid; idx < N * kHD; idx += kThreads) {              // a_head += dLq^T [H1q | 1]
                const int cc = idx / kHD, o = idx - cc * kHD;
                float a = 0.f;
                for (int i = 0; i < tr; ++i) a = fmaf(s.lt[i * kLS + cc], o < kH1 ? s.h1t[i * kH1 + o] : 1.f, a);
                s.dhp[idx] += a;
            }
#pragma unroll
            for (int ii = 0; ii < RPT; ++ii) {                                  // dZ1q
                const int i = kg_ + 4 * ii;
                float dz = 0.f;
                if (i < tr && s.h1t[i * kH1 + o_] > 0.f) {
                    float dh = 0.f;
                    for (int cc = 0; cc < N; ++cc) dh = fmaf(s.lt[i * kLS + cc], s.hp[cc * kHD + o_], dh);
                    dz = dh * sc;
                }
                s.dz1t[i * kH1 + o_] = dz;
            }
            __syncthreads();
            if (tid < kH1) {
                float a = 0.f;
                for (int i = 0; i < tr; ++i) a += s.dz1t[i * kH1 + tid];
                s.ab1[tid] += a;
            }
            outer_rows<TR>(s.h0t, s.dz1t, [&](int, int k, int o, float a0, float a1, float a2, float a3) {
                s.aw1t[(k + 0) * kW1S + o] += a0;
                s.aw1t[(k + 1) * kW1S + o] += a1;
                s.aw1t[(k + 2) * kW1S + o] += a2;
                s.aw1t[(k + 3) * kW1S + o] += a3;
            });
            {
                float acc[TR];
#pragma unroll
                for (int i = 0; i < TR; ++i) acc[i] = 0.f;
                cols_from_h1<TR>(s.dz1t, s.w1t, 1.f, acc);
                float z[TR];
#pragma unroll
                for (int i = 0; i < TR; ++i) {
                    z[i] = (i < tr && s.h0t[i * kH0 + tid] > 0.f) ? acc[i] * sc : 0.f;      // dZ0q
                    if (i < tr) {
                        ab0 += z[i];
                        atomicAdd(&P.d_proj[s.rows[i] * kH0 + tid], z[i]);
                    }
                }
                if (steps > 0) {                                                // a_S -= alpha * Gq^T dZ0q
                    for (int j = 0; j < n; ++j) {
                        float a = 0.f;
#pragma unroll
                        for (int i = 0; i < TR; ++i) a = fmaf(s.gt[i * kGS + j], z[i], a);
                        aS[0][int64_t(j) * kH0 + tid] -= alpha * a;
                    }
                }
            }
            __syncthreads();
        }

        // ---- inner steps in reverse
        int cur = 0;
        if (!c.first_order) {
            for (int st = steps - 1; st >= 0; --st) {
                const float* rec = slot + L.steps + int64_t(st) * L.per_step;
                // head of this step (pre-update) and W1 of this step: undo W1_{s+1} = W1_s - alpha dZ1^T H0
                for (int idx = tid; idx < N * kHD; idx += kThreads) s.hp[idx] = rec[L.oHP + idx];
                for (int r0 = 0; r0 < n; r0 += TR) {
                    const int tr = min(TR, n - r0);
                    for (int i = 0; i < TR; ++i)
                        s.h0t[i * kH0 + tid] = i < tr ? rec[L.oH0 + int64_t(r0 + i) * kH0 + tid] : 0.f;
                    for (int idx = tid; idx < TR * kH1; idx += kThreads)
                        s.dz1t[idx] = idx < tr * kH1 ? rec[L.oDZ1 + int64_t(r0) * kH1 + idx] : 0.f;
                    __syncthreads();
                    outer_rows<TR>(s.h0t, s.dz1t, [&](int, int k, int o, float a0, float a1, float a2, float a3) {
                        s.w1t[(k + 0) * kW1S + o] += alpha * a0;
                        s.w1t[(k + 1) * kW1S + o] += alpha * a1;
                        s.w1t[(k + 2) * kW1S + o] += alpha * a2;
                        s.w1t[(k + 3) * kW1S + o] += alpha * a3;
                    });
                    __syncthreads();
                }
                const float* aScur = aS[cur];
                float* aSnew = aS[cur ^ 1];
                for (int j = 0; j < n; ++j) aSnew[int64_t(j) * kH0 + tid] = aScur[int64_t(j) * kH0 + tid];
                float rb0 = 0.f;
                float rw1[64];
#pragma unroll
                for (int q = 0; q < 64; ++q) rw1[q] = 0.f;
                const float gb0 = -alpha * ab0;

                for (int r0 = 0; r0 < n; r0 += TR) {
                    const int tr = min(TR, n - r0);
                    load_tile_meta<TR>(P, s, b, false, r0, tr);
                    for (int i = 0; i < TR; ++i) {
                        const float h0 = i < tr ? rec[L.oH0 + int64_t(r0 + i) * kH0 + tid] : 0.f;
                        s.h0t[i * kH0 + tid] = h0;
                        // (12r) r_dH0 = (a_S' + g_b0) * M0
                        s.tt[i * kH0 + tid] = (i < tr && h0 > 0.f) ? (aScur[int64_t(r0 + i) * kH0 + tid] + gb0) * sc : 0.f;
                    }
                    for (int idx = tid; idx < TR * kH1; idx += kThreads) {
                        const bool in = idx < tr * kH1;
                        s.h1t[idx] = in ? rec[L.oH1 + int64_t(r0) * kH1 + idx] : 0.f;
                        s.dz1t[idx] = in ? rec[L.oDZ1 + int64_t(r0) * kH1 + idx] : 0.f;
                    }
                    for (int idx = tid; idx < TR * kLS; idx += kThreads) {
                        const int i = idx / kLS, cc = idx - i * kLS;
                        s.lt[idx] = (i < tr && cc < N) ? rec[L.oDL + int64_t(r0 + i) * N + cc] : 0.f;
                    }
                    __syncthreads();
                    // (11r)+(10r): r_dZ1 = r_dH0 W1^T + H0 (g_W1)^T + g_b1 ,  g_W1 = -alpha a_W1
                    {
                        float acc[RPT];
#pragma unroll
                        for (int ii = 0; ii < RPT; ++ii) acc[ii] = 0.f;
                        for (int k = 0; k < kH0; k += 4) {
                            float w[4], g[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                w[q] = s.w1t[(k + q) * kW1S + o_];
                                g[q] = -alpha * s.aw1t[(k + q) * kW1S + o_];
                            }
#pragma unroll
                            for (int ii = 0; ii < RPT; ++ii) {
                                const int i = kg_ + 4 * ii;
                                const float4 t4 = *reinterpret_cast<const float4*>(&s.tt[i * kH0 + k]);
                                const float4 h4 = *reinterpret_cast<const float4*>(&s.h0t[i * kH0 + k]);
                                float a = acc[ii];
                                a = fmaf(t4.x, w[0], a); a = fmaf(t4.y, w[1], a); a = fmaf(t4.z, w[2], a); a = fmaf(t4.w, w[3], a);
                                a = fmaf(h4.x, g[0], a); a = fmaf(h4.y, g[1], a); a = fmaf(h4.z, g[2], a); a = fmaf(h4.w, g[3], a);
                                acc[ii] = a;
                            }
                        }
                        const float gb1 = -alpha * s.ab1[o_];
#pragma unroll
                        for (int ii = 0; ii < RPT; ++ii) {
                            const int i = kg_ + 4 * ii;
                            // (9r) r_dH1 = r_dZ1 * M1
                            s.rz1t[i * kH1 + o_] = (i < tr && s.h1t[i * kH1 + o_] > 0.f) ? (acc[ii] + gb1) * sc : 0.f;
                        }
                    }
                    // r_W1 += dZ1^T r_dH0   (registers)
                    outer_rows<TR>(s.tt, s.dz1t, [&](int kk, int, int, float a0, float a1, float a2, float a3) {
                        rw1[kk] += a0; rw1[kk + 1] += a1; rw1[kk + 2] += a2; rw1[kk + 3] += a3;
                    });
                    __syncthreads();
                    // (10r) r_H0 = dZ1 g_W1  -> tt      thread == k
                    {
                        float acc[TR];
#pragma unroll
                        for (int i = 0; i < TR; ++i) acc[i] = 0.f;
                        cols_from_h1<TR>(s.dz1t, s.aw1t, -alpha, acc);
#pragma unroll
                        for (int i = 0; i < TR; ++i) s.tt[i * kH0 + tid] = acc[i];
                    }
                    // (8r)+(7r): r_dL = r_dH1 Wh^T + H1 (g_Wh)^T + g_bh ,  g_head = -alpha a_head
                    for (int idx = tid; idx < TR * N; idx += kThreads) {
                        const int i = idx / N, cc = idx - i * N;
                        float a = 0.f;
                        if (i < tr) {
                            a = -alpha * s.dhp[cc * kHD + kH1];
                            for (int o = 0; o < kH1; ++o) {
                                a = fmaf(s.rz1t[i * kH1 + o], s.hp[cc * kHD + o], a);
                                a = fmaf(s.h1t[i * kH1 + o], -alpha * s.dhp[cc * kHD + o], a);
                            }
                        }
                        s.rlt[i * kLS + cc] = a;
                    }
                    // r_head += dL^T r_dH1
                    for (int idx = tid; idx < N * kH1; idx += kThreads) {
                        const int cc = idx / kH1, o = idx - cc * kH1;
                        float a = 0.f;
                        for (int i = 0; i < tr; ++i) a = fmaf(s.lt[i * kLS + cc], s.rz1t[i * kH1 + o], a);
                        s.rhp[cc * kHD + o] += a;
                    }
                    // (7r) r_H1 = dL g_Wh
#pragma unroll
                    for (int ii = 0; ii < RPT; ++ii) {
                        const int i = kg_ + 4 * ii;
                        float a = 0.f;
                        if (i < tr)
                            for (int cc = 0; cc < N; ++cc) a = fmaf(s.lt[i * kLS + cc], -alpha * s.dhp[cc * kHD + o_], a);
                        s.rh1t[i * kH1 + o_] = a;
                    }
                    __syncthreads();
                    // (6r) r_L = P * (r_dL - <P, r_dL>) / n ,  P = n dL + Y
                    if (tid < tr) {
                        const int y = s.ys[tid];
                        float dot = 0.f;
                        for (int cc = 0; cc < N; ++cc) {
                            const float p = s.lt[tid * kLS + cc] * float(n) + (cc == y ? 1.f : 0.f);
                            dot = fmaf(p, s.rlt[tid * kLS + cc], dot);
                        }
                        for (int cc = 0; cc < N; ++cc) {
                            const float p = s.lt[tid * kLS + cc] * float(n) + (cc == y ? 1.f : 0.f);
                            s.rlt[tid * kLS + cc] = p * (s.rlt[tid * kLS + cc] - dot) / float(n);
                        }
                    }
                    __syncthreads();
                    // (5r) r_H1 += r_L Wh ; r_head += r_L^T [H1 | 1] ; (4r) r_Z1 = r_H1 * M1
#pragma unroll
                    for (int ii = 0; ii < RPT; ++ii) {
                        const int i = kg_ + 4 * ii;
                        float a = s.rh1t[i * kH1 + o_];
                        if (i < tr) {
                            for (int cc = 0; cc < N; ++cc) a = fmaf(s.rlt[i * kLS + cc], s.hp[cc * kHD + o_], a);
                        }
                        s.rh1t[i * kH1 + o_] = (i < tr && s.h1t[i * kH1 + o_] > 0.f) ? a * sc : 0.f;
                    }
                    for (int idx = tid; idx < N * kHD; idx += kThreads) {
                        const int cc = idx / kHD, o = idx - cc * kHD;
                        float a = 0.f;
                        for (int i = 0; i < tr; ++i) a = fmaf(s.rlt[i * kLS + cc], o < kH1 ? s.h1t[i * kH1 + o] : 1.f, a);
                        s.rhp[idx] += a;
                    }
                    __syncthreads();
                    // (3r) r_H0 += r_Z1 W1 ; r_W1 += r_Z1^T H0 ; r_b1 += sum r_Z1
                    if (tid < kH1) {
                        float a = 0.f;
                        for (int i = 0; i < tr; ++i) a += s.rh1t[i * kH1 + tid];
                        s.rb1[tid] += a;
                    }
                    outer_rows<TR>(s.h0t, s.rh1t, [&](int kk, int, int, float a0, float a1, float a2, float a3) {
                        rw1[kk] += a0; rw1[kk + 1] += a1; rw1[kk + 2] += a2; rw1[kk + 3] += a3;
                    });
                    {
                        float acc[TR];
#pragma unroll
                        for (int i = 0; i < TR; ++i) acc[i] = s.tt[i * kH0 + tid];
                        cols_from_h1<TR>(s.rh1t, s.w1t, 1.f, acc);
                        // (2r) bar_Z0 = r_H0 * M0 ; (1r) a_A, a_b0, a_S
                        float z[TR];
#pragma unroll
                        for (int i = 0; i < TR; ++i) {
                            z[i] = (i < tr && s.h0t[i * kH0 + tid] > 0.f) ? acc[i] * sc : 0.f;
                            if (i < tr) {
                                rb0 += z[i];
                                atomicAdd(&P.d_proj[s.rows[i] * kH0 + tid], z[i]);
                            }
                        }
                        for (int j = 0; j < n; ++j) {
                            float a = 0.f;
#pragma unroll
                            for (int i = 0; i < TR; ++i) a = fmaf(s.gt[i * kGS + j], z[i], a);
                            aSnew[int64_t(j) * kH0 + tid] -= alpha * a;
                        }
                    }
                    __syncthreads();
                }
                // ---- end of reversed step: fold this step's contributions into the adjoints
#pragma unroll
                for (int kk = 0; kk < 64; ++kk) s.aw1t[(kg_ * 64 + kk) * kW1S + o_] += rw1[kk];
                for (int idx = tid; idx < N * kHD; idx += kThreads) { s.dhp[idx] += s.rhp[idx]; s.rhp[idx] = 0.f; }
                if (tid < kH1) { s.ab1[tid] += s.rb1[tid]; s.rb1[tid] = 0.f; }
                ab0 += rb0;
                cur ^= 1;
                __syncthreads();
            }
        }

        // ---- task epilogue: head gradient per task, shared-parameter gradients into this CTA's partials
        for (int idx = tid; idx < N * kHD; idx += kThreads) P.d_head[b * N * kHD + idx] = s.dhp[idx];
        float* pw = P.d_w1_parts + int64_t(blockIdx.x) * kH0 * kH1;
        for (int idx = tid; idx < kH0 * kH1; idx += kThreads) {
            const int o = idx / kH0, k = idx - o * kH0;                         // [H1][H0] like linear1.weight
            pw[idx] += s.aw1t[k * kW1S + o];
        }
        P.d_b0_parts[int64_t(blockIdx.x) * kH0 + tid] += ab0;
        if (tid < kH1) P.d_b1_parts[int64_t(blockIdx.x) * kH1 + tid] += s.ab1[tid];
        __syncthreads();
    }
}

int check_cfg(const fumi_episode_cfg* c) {
    FUMI_CHECK_ARG(c != nullptr, "cfg is null");
    if (c->hid0 != kH0 || c->hid1 != kH1) {
        fumi_set_error("only --im_hid_dim 256 64 (the reference default) is compiled into this build");
        return FUMI_ERR_UNSUPPORTED;
    }
    FUMI_CHECK_ARG(c->num_ways >= 1 && c->num_ways <= kMaxWays, "num_ways must be in [1,32]");
    FUMI_CHECK_ARG(c->num_support >= 1 && c->num_support <= kMaxSupport, "num_support must be in [1,128]");
    FUMI_CHECK_ARG(c->num_query >= 1, "num_query must be >= 1");
    FUMI_CHECK_ARG(c->steps >= 0, "steps must be >= 0");
    FUMI_CHECK_ARG(c->dropout_p >= 0.f && c->dropout_p < 1.f, "dropout_p must be in [0,1)");
    return FUMI_OK;
}

int grid_for(int64_t B) {
    int sms = fumi_device_sm_count();
    if (sms <= 0) return sms;
    return int(B < sms ? B : sms);
}

#ifdef FUMI_EMU
#define FUMI_SET_SMEM_ATTR(kern, bytes) ((void)0)
#else
#define FUMI_SET_SMEM_ATTR(kern, bytes)                                                                        \
    do {                                                                                                       \
        cudaError_t e__ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)); \
        if (e__ != cudaSuccess) return fumi_cuda_fail(e__, "cudaFuncSetAttribute(episode kernel)");            \
    } while (0)
#endif

constexpr int kTRF = 32;   // forward query rows per tile
constexpr int kTRB = 16;   // backward rows per tile (two W1-sized buffers live in shared memory)

}  // namespace

// v1 (the round-1 kernels: fp16-plane forward, 3xTF32 backward) stays selectable for A/B runs while FUMI_EPISODE_V1=1
static bool use_v1() {
    static int v1 = -1;
    if (v1 < 0) { const char* e = getenv("FUMI_EPISODE_V1"); v1 = (e && atoi(e) != 0) ? 1 : 0; }
    return v1 != 0;
}
// 2: fp16-plane kernels with plane-format step records; 1: v1 tensor-core kernels; 0: fp32 FMA kernels (NK > 32)
static int episode_path(const fumi_episode_cfg& c) {
    const bool small = c.num_support <= 32;
    if (small && !use_v1() && episode_f16_supported(c)) return 2;
    if (small && c.num_query <= kMaxQueryRows) return 1;
    return 0;
}

extern "C" int64_t fumi_episode_stash_floats(const fumi_episode_cfg* cfg) {
    if (check_cfg(cfg) != FUMI_OK) return FUMI_ERR_ARG;
    return episode_path(*cfg) == 2 ? make_layout_f(*cfg).per_task : make_layout(*cfg).per_task;
}

extern "C" int fumi_stash_layout(const fumi_episode_cfg* cfg, fumi_stash_layout_t* out) {
    int rc = check_cfg(cfg);
    if (rc != FUMI_OK) return rc;
    FUMI_CHECK_ARG(out != nullptr, "out is null");
    if (episode_path(*cfg) == 2) {
        const LayoutF F = make_layout_f(*cfg);
        out->per_task = F.per_task; out->S = F.S; out->w1t = F.w1t; out->b0 = F.b0; out->b1 = F.b1; out->head = F.head;
        out->steps = F.steps; out->per_step = F.per_step;
        out->format = 1;
        out->rec_h1 = F.oH1; out->rec_exp = F.oEXP; out->rec_h0_hi = F.oH0h; out->rec_h0_lo = F.oH0l;
        return FUMI_OK;
    }
    const Layout L = make_layout(*cfg);
    out->per_task = L.per_task;
    out->S = (cfg->steps & 1) ? L.S1 : L.S0;
    out->w1t = L.w1t;
    out->b0 = L.b0;
    out->b1 = L.b1;
    out->head = L.head;
    out->steps = L.steps;
    out->per_step = L.per_step;
    out->format = 0;
    out->rec_h1 = L.oH1; out->rec_exp = -1; out->rec_h0_hi = L.oH0; out->rec_h0_lo = -1;
    return FUMI_OK;
}

extern "C" int fumi_episode_bwd_parts(void) { return fumi_device_sm_count(); }

extern "C" int fumi_episode_fwd(const fumi_episode_cfg* cfg, int64_t B, const float* proj, const int64_t* sup_rows,
                                const int64_t* qry_rows, const int64_t* sup_y, const int64_t* qry_y,
                                const float* gram, const float* b0, const float* w1, const float* b1,
                                const float* head_table, const int64_t* head_rows, float* logits, int64_t* preds,
                                float* task_loss, float* task_acc, float* stash, void* stream) {
    int rc = check_cfg(cfg);
    if (rc != FUMI_OK) return rc;
    FUMI_CHECK_ARG(B >= 0, "B < 0");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(proj && sup_rows && qry_rows && sup_y && qry_y && gram && b0 && w1 && b1 && head_table &&
                   logits && preds && task_loss && task_acc && stash, "null pointer");
    EpiParams P;
    std::memset(&P, 0, sizeof(P));
    P.cfg = *cfg; P.B = B; P.proj = proj; P.sup_rows = sup_rows; P.qry_rows = qry_rows; P.sup_y = sup_y;
    P.qry_y = qry_y; P.gram = gram; P.b0 = b0; P.w1 = w1; P.b1 = b1; P.head_table = head_table;
    P.head_rows = head_rows; P.logits = logits; P.preds = preds; P.task_loss = task_loss; P.task_acc = task_acc;
    P.stash = stash;
    P.save = cfg->reserved != 0;
    P.slot_floats = make_layout(*cfg).per_task;
    const int grid = grid_for(B);
    if (grid <= 0) return grid;
    const int nk = cfg->num_support;
#define FUMI_FWD_CASE(TRS, MULTI)                                                                              \
    do {                                                                                                       \
        constexpr int TRM_ = (TRS) > kTRF ? (TRS) : kTRF;                                                      \
        const size_t smem = smem_floats<TRM_, false, (MULTI) ? 0 : (TRS)>() * sizeof(float);                   \
        auto kern = episode_fwd_kernel<(TRS), kTRF, (MULTI)>;                                                   \
        FUMI_SET_SMEM_ATTR(kern, smem);                                                                        \
        FUMI_LAUNCH(kern, grid, kThreads, smem, stream, P);                                                    \
    } while (0)
    const int path = episode_path(*cfg);
    const bool use_mma = path == 1;
    P.phase = episode_phase_counters();
    if (path == 2) {
        P.slot_floats = make_layout_f(*cfg).per_task;
        return launch_episode_fwd_f16(P, grid, stream);
    } else if (use_mma) {
        const size_t smem = smem_f_bytes();
        if (nk <= 16) {
            FUMI_SET_SMEM_ATTR(episode_fwd_f16_kernel<1>, smem);
            FUMI_LAUNCH(episode_fwd_f16_kernel<1>, grid, kThreads16, smem, stream, P);
        } else {
            FUMI_SET_SMEM_ATTR(episode_fwd_f16_kernel<2>, smem);
            FUMI_LAUNCH(episode_fwd_f16_kernel<2>, grid, kThreads16, smem, stream, P);
        }
    } else if (nk > 32) FUMI_FWD_CASE(32, true);
    else if (nk > 28) FUMI_FWD_CASE(32, false);
    else if (nk > 16) FUMI_FWD_CASE(28, false);
    else if (nk > 8) FUMI_FWD_CASE(16, false);
    else FUMI_FWD_CASE(8, false);
#undef FUMI_FWD_CASE
    FUMI_CHECK_LAUNCH("episode_fwd_kernel");
    return FUMI_OK;
}

extern "C" int fumi_episode_bwd(const fumi_episode_cfg* cfg, int64_t B, const float* proj, const int64_t* sup_rows,
                                const int64_t* qry_rows, const int64_t* sup_y, const int64_t* qry_y,
                                const float* gram, const float* stash, float loss_scale, float* d_proj,
                                float* d_head, float* d_b0_parts, float* d_w1_parts, float* d_b1_parts,
                                void* stream) {
    int rc = check_cfg(cfg);
    if (rc != FUMI_OK) return rc;
    FUMI_CHECK_ARG(B >= 0, "B < 0");
    if (B == 0) return FUMI_OK;
    FUMI_CHECK_ARG(cfg->reserved != 0, "the forward pass must have run with save-for-backward (cfg.reserved = 1)");
    FUMI_CHECK_ARG(proj && sup_rows && qry_rows && sup_y && qry_y && gram && stash && d_proj && d_head &&
                   d_b0_parts && d_w1_parts && d_b1_parts, "null pointer");
    EpiParams P;
    std::memset(&P, 0, sizeof(P));
    P.cfg = *cfg; P.B = B; P.proj = proj; P.sup_rows = sup_rows; P.qry_rows = qry_rows; P.sup_y = sup_y;
    P.qry_y = qry_y; P.gram = gram;
    P.stash = const_cast<float*>(stash);     // the S ping-pong slots are reused for the adjoint of S
    P.save = 1;
    P.slot_floats = make_layout(*cfg).per_task;
    P.loss_scale = loss_scale; P.d_proj = d_proj; P.d_head = d_head;
    P.d_b0_parts = d_b0_parts; P.d_w1_parts = d_w1_parts; P.d_b1_parts = d_b1_parts;
    const int grid = grid_for(B);
    if (grid <= 0) return grid;
    P.phase = episode_phase_counters();
    if (episode_path(*cfg) == 2) {
        P.slot_floats = make_layout_f(*cfg).per_task;
        return launch_episode_bwd_f16(P, grid, stream);
    }
    if (episode_path(*cfg) == 1) {      // v1 tensor-core path (pairs with episode_fwd_f16_kernel)
        const size_t smem_b = smem_b_floats() * sizeof(float);
        FUMI_SET_SMEM_ATTR(episode_bwd_mma16_kernel, smem_b);
        FUMI_LAUNCH(episode_bwd_mma16_kernel, grid, kThreads16, smem_b, stream, P);
        FUMI_CHECK_LAUNCH("episode_bwd_mma16_kernel");
        return FUMI_OK;
    }
    const size_t smem = smem_floats<kTRB, true>() * sizeof(float);
#ifndef FUMI_EMU
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(episode_bwd_kernel<kTRB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaFuncSetAttribute(episode_bwd)");
        attr_done = true;
    }
#endif
    FUMI_LAUNCH((episode_bwd_kernel<kTRB>), grid, kThreads, smem, stream, P);
    FUMI_CHECK_LAUNCH("episode_bwd_kernel");
    return FUMI_OK;
}

// ---- diagnostics (not part of the reference-facing surface): per-phase SM-cycle counters of the episode kernels
static unsigned long long* g_phase_dev = nullptr;
static bool g_phase_on = false;
namespace fumi_epi {
unsigned long long* episode_phase_counters() { return g_phase_on ? g_phase_dev : nullptr; }
}
extern "C" int fumi_debug_phase_profile(int enable) {
#ifndef FUMI_EMU
    if (!g_phase_dev) {
        cudaError_t e = cudaMalloc(&g_phase_dev, sizeof(unsigned long long) * 64);
        if (e != cudaSuccess) return fumi_cuda_fail(e, "fumi_debug_phase_profile");
    }
    cudaError_t e = cudaMemset(g_phase_dev, 0, sizeof(unsigned long long) * 64);
    if (e != cudaSuccess) return fumi_cuda_fail(e, "fumi_debug_phase_profile");
#endif
    g_phase_on = enable != 0;
    return FUMI_OK;
}
extern "C" int fumi_debug_read_phases(unsigned long long* out64 /* HOST, 64 entries */) {
#ifndef FUMI_EMU
    if (!g_phase_dev) { for (int i = 0; i < 64; ++i) out64[i] = 0; return FUMI_OK; }
    cudaError_t e = cudaMemcpy(out64, g_phase_dev, sizeof(unsigned long long) * 64, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fumi_cuda_fail(e, "fumi_debug_read_phases");
#else
    for (int i = 0; i < 64; ++i) out64[i] = 0;
#endif
    return FUMI_OK;
}
