// Fused inner loop + query scoring on fp16 hi/lo operand planes, NK <= 32 (fumi/models/fumi.py:148-185,
// fumi/models/maml.py:158-183).  One persistent CTA of 16 warps walks a task; warp w owns hidden units
// [16w, 16w+16) of the 256-wide layer.
//
// What shapes the kernel is the dependency structure of one SGD step in the Gram form:
//     (a) H0 = act(A + b0 - alpha G S)          columns h of H0 need columns h of S            -> warp-local
//     (b) Z1 = H0 W1^T                          contracts over all 256 hidden units              -> block-wide
//     (c) logits, softmax, dL, dZ1              couples the 64 units / N classes of ONE row      -> one row per warp
//     (d) dZ0 = (dZ1 W1) * gate ; S += dZ0      columns h need rows h of W1^T                    -> warp-local
//     (e) W1 -= alpha dZ1^T H0                  rows h of W1^T need columns h of H0              -> warp-local
// so a step costs THREE block barriers (before b, before c, before d); (d) -> (e) -> (a of the next step) is one
// uninterrupted stream of 120 MMAs per warp with no block synchronisation.  Every GEMM operand lives in shared
// memory as pre-split fp16 planes x 2^e = hi + lo (warp_mma.cuh: ldmatrix + 3 x mma.m16n8k16 per 16 k, fp32-grade
// accuracy).  Plane exponents of matrices that all warps consume (H0, dZ1, W1^T, the query Gram tiles) are LAGGED:
// a production is written with the exponent derived from the previous production's max (target 2^8, i.e. 8
// binades of headroom; the new max travels through 16 per-warp words and is read after the barrier that follows
// anyway), which removes the "exchange the max, then write" barrier per matrix.  A matrix growing more than 250x
// between two consecutive inner steps would overflow fp16: the task's loss is then reported as NaN (the inner
// loop has diverged); shrinking only costs precision gracefully.  S and b0 are warp-private (exact scales, registers).
//
// In train mode the step records are stashed as the planes themselves (LayoutF): the backward copies them into its
// operand tiles with no conversion.
//
// Query scoring (fumi.py:178-185) has no such coupling -- S, W1, b0, b1 and the head are final -- so it runs with NO
// block barrier: a warp owns 16 query rows, walks the 256 hidden units in 16-column slabs, and the activations of a slab
// (accumulator layout of the Z0 MMAs) are used directly as the A fragments of one k step of Z1q = H0q W1^T.  The
// projected rows arrive through a per-warp cp.async ring three slabs ahead.
#include "episode_common.cuh"

namespace fumi_epi {
namespace {

enum { FX_W1 = 0, FX_H0 = 2, FX_DZ = 4, FX_GS = 6, FX_HP = 7, FX_COUNT = 8 };   // x 16 floats (H0 / dZ1: two parities)

constexpr int kRingRow = 24;                 // words per row of a ring slab (64 data bytes + pad: conflict-free LDS.64)
constexpr int kRingSlab = 16 * kRingRow;     // one slab: 16 rows

struct SmemV {
    fumi_half *w1h, *w1l, *sh, *sl, *h0h, *h0l, *dzh, *dzl, *gsh, *gsl;
    float *z1p, *h1t, *dz1t, *lt, *hp, *dhp, *b1s, *b0s, *mx, *red;
    int *ysS, *es;
};
__host__ __device__ inline size_t carve_v(char* base, SmemV& s, int N) {
    char* p = base;
    auto take = [&](size_t bytes) { char* r = p; p += (bytes + 15) & ~size_t(15); return r; };
    s.ysS = reinterpret_cast<int*>(take(32 * 4));
    s.mx = reinterpret_cast<float*>(take(FX_COUNT * 16 * 4));
    s.red = reinterpret_cast<float*>(take(4 * 16 * 4));
    s.es = reinterpret_cast<int*>(take(16 * 4));
    s.b0s = reinterpret_cast<float*>(take(kH0 * 4));
    // zero-filled once per kernel from here (plane pads must be finite)
    s.w1h = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));  s.w1l = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));
    s.sh = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));    s.sl = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));
    s.h0h = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));   s.h0l = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));
    s.dzh = reinterpret_cast<fumi_half*>(take(32 * kHW * 2));   s.dzl = reinterpret_cast<fumi_half*>(take(32 * kHW * 2));
    s.gsh = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));   s.gsl = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));
    s.z1p = reinterpret_cast<float*>(take(32 * kS1 * 4));
    s.h1t = reinterpret_cast<float*>(take(32 * kS1 * 4));
    s.dz1t = reinterpret_cast<float*>(take(32 * kS1 * 4));
    s.lt = reinterpret_cast<float*>(take(32 * kLS * 4));
    s.hp = reinterpret_cast<float*>(take(size_t(N) * kHD * 4));
    s.dhp = reinterpret_cast<float*>(take(size_t(N) * kHD * 4));
    s.b1s = reinterpret_cast<float*>(take(kH1 * 4));
    return size_t(p - base);
}

// MT: 16-row tiles of the support set (NK <= 16 MT); kNC: compile-time class count (>= N; the head buffer is zero-padded
// to kNC rows so that the per-class loops carry no guards)
// SAVE = false: meta-test (no stash): the record code is compiled out; DROP = false (meta-test, MAML): the mask code too
template <int MT, int kNC, bool SAVE, bool DROP>
__global__ void __launch_bounds__(kThreads16, 1) episode_fwd_v2_kernel(EpiParams P) {
    constexpr int RS = 16 * MT;
    constexpr int NT_ = kThreads16;
    FUMI_DYN_SMEM(float, smem_raw);
    const fumi_episode_cfg& c = P.cfg;
    SmemV s;
    const size_t smem_total = carve_v(reinterpret_cast<char*>(smem_raw), s, kNC);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n = c.num_support, m = c.num_query, N = c.num_ways, steps = c.steps;
    const float alpha = c.step_size;
    const LayoutF L = make_layout_f(c);
    const float dsc = dropout_scale(c);
    constexpr bool drop = DROP;
    const uint32_t thr = dropout_thr(c);
    const int hc = 16 * w + 2 * t;                       // this thread's column pairs: hc + 8 j + {0, 1}
    PhaseClock pc;
    pc.start(P.phase);

    {   // planes start as zeros: pad rows / columns that no phase writes must be finite
        uint32_t* z = reinterpret_cast<uint32_t*>(s.w1h);
        const int nz = int((reinterpret_cast<char*>(s.z1p) - reinterpret_cast<char*>(s.w1h)) / 4);
        for (int idx = tid; idx < nz; idx += NT_) z[idx] = 0u;
        (void)smem_total;
    }
    __syncthreads();

    for (int64_t b = blockIdx.x; b < P.B; b += gridDim.x) {
        const int64_t task = c.task_offset + b;
        float* slot = (SAVE && P.save) ? P.stash + b * P.slot_floats : nullptr;
        bool bad = false;                                 // a lagged plane exponent overflowed fp16
        int e_w1, e_h0 = 0, e_dz, e_gs, e_s = 0;
        int e_h0n = 0, e_dzn = 0;                         // exponents for the NEXT production (from the last observed max)
        int par_h0 = 0, par_dz = 0;

        // ------------------------------------------------------------------ prologue
        // fp32 master of this warp's 16 rows of W1^T at the thread's accumulator positions (the layout of the W1 update
        // GEMM), kept pre-multiplied by the plane scale 2^e_w1: the update is an FFMA and the planes are re-split from it
        float w1m[8][4];
        {
            float mxv = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    w1m[j][q] = __ldg(&P.w1[(8 * j + 2 * t + (q & 1)) * kH0 + 16 * w + g + 8 * (q >> 1)]);
                    mxv = fmaxf(mxv, fabsf(w1m[j][q]));
                }
            block_max_push(s.mx + 16 * FX_W1, mxv);
        }
        float gv[2];                                      // support Gram block
        {
            float mxv = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                gv[q] = (i < n && j < n) ? __ldg(&P.gram[(b * int64_t(n + m) + i) * n + j]) : 0.f;
                mxv = fmaxf(mxv, fabsf(gv[q]));
            }
            block_max_push(s.mx + 16 * FX_GS, mxv);
        }
        {
            float mxv = 0.f;
            for (int idx = tid; idx < kNC * kHD; idx += NT_) {
                const int cc = idx / kHD, o = idx - cc * kHD;
                float v = 0.f;                            // rows N .. kNC-1 stay zero (inert classes)
                if (cc < N) {
                    const int64_t r = P.head_rows ? __ldg(&P.head_rows[b * N + cc]) : cc;
                    v = __ldg(&P.head_table[r * kHD + o]);
                }
                s.hp[idx] = v;
                if (o < kH1) mxv = fmaxf(mxv, fabsf(v));
            }
            block_max_push(s.mx + 16 * FX_HP, mxv);
        }
        if (tid < kH1) s.b1s[tid] = __ldg(&P.b1[tid]);
        if (tid < 32) s.ysS[tid] = tid < n ? int(P.sup_y[b * n + tid]) : 0;
        // the previous task's query pass used [h0 planes, hp) as its ring: planes whose pad rows no phase writes start as zeros
        for (int idx = tid; idx < 2 * 32 * kHW * 2 / 16; idx += NT_) reinterpret_cast<uint4*>(s.dzh)[idx] = make_uint4(0u, 0u, 0u, 0u);
        float b0r[2][2];                                  // adapted linear0.bias at this thread's columns (warp-private)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            b0r[j][0] = __ldg(&P.b0[hc + 8 * j]);
            b0r[j][1] = __ldg(&P.b0[hc + 8 * j + 1]);
        }
        float Sr[MT][2][4];                               // S = sum of dZ0 so far at this thread's positions (fp32 master)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) Sr[i][j][q] = 0.f;
        uint32_t gate0 = 0u;                              // ReLU/dropout gates of H0 at this thread's 16 positions

        // projected rows at this thread's accumulator positions, straight from `proj`
        float2 ap[2][2][2];
        // (row ids travel in registers: the dependent index load is off the critical path)
        auto h0_rows = [&](const int64_t* rows, int tr, int (&rid)[2][2]) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                    const int r = 16 * i + g + 8 * hq;
                    rid[i][hq] = r < tr ? int(__ldg(&rows[r])) : -1;
                }
        };
        auto h0_load = [&](const int (&rid)[2][2], int mt) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < mt)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        ap[i][j][hq] = rid[i][hq] >= 0
                                           ? __ldg(reinterpret_cast<const float2*>(&P.proj[int64_t(rid[i][hq]) * kH0 + hc + 8 * j]))
                                           : make_float2(0.f, 0.f);
        };
        int srow[2][2];
        h0_rows(P.sup_rows + b * n, n, srow);
        // H0 epilogue shared by support steps and query tiles: acc = raw G.S product (scaled) -> activations written as
        // planes with the lagged exponent e_h0n; returns the gates.  mt: m tiles in use (support: MT, query: 2)
        auto h0_finish = [&](float (&acc)[2][2][4], int mt, int r0, int tr, float gs, int pass) -> uint32_t {
            const uint32_t dbase = drop ? dropout_base(c, task, pass, 0) : 0u;
            const float sc = fumi_exp2i(e_h0n);
            uint32_t gates = 0u;
            float mxv = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (i < mt)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const int r = 16 * i + g + 8 * hq, h = hc + 8 * j;
                        const float z0 = ap[i][j][hq].x + b0r[j][0] - gs * acc[i][j][2 * hq];
                        const float z1 = ap[i][j][hq].y + b0r[j][1] - gs * acc[i][j][2 * hq + 1];
                        bool k0 = (r < tr) & (z0 > 0.f), k1 = (r < tr) & (z1 > 0.f);      // (&, not &&: no divergent branches)
                        if (drop) {
                            const uint32_t bits = dropout_bits(dbase, r0 + r, h);
                            k0 = k0 & ((bits & 0xFFFFu) >= thr);
                            k1 = k1 & ((bits >> 16) >= thr);
                        }
                        const float v0 = k0 ? z0 * dsc : 0.f, v1 = k1 ? z1 * dsc : 0.f;
                        gates |= (uint32_t(k0) | (uint32_t(k1) << 1)) << (2 * (4 * i + 2 * j + hq));
                        mxv = fmaxf(mxv, fmaxf(v0, v1));
                        st_planes2(s.h0h, s.h0l, r * kHS + h, v0, v1, sc);
                    }
            block_max_push(s.mx + 16 * (FX_H0 + par_h0), mxv);
            return gates;
        };
        // bookkeeping after the barrier that follows a production: the exponent just used becomes current, the observed
        // max gives the next one
#define FUMI_ADOPT(cur, next, par, slot_id)                                     \
        do {                                                                    \
            const float mx__ = slot_max(s.mx + 16 * ((slot_id) + (par)));       \
            cur = next;                                                         \
            bad = bad || plane_overflow(mx__, cur);                             \
            if (mx__ > 0.f) next = fumi_plane_exp_t(__float_as_uint(mx__), kTarget); \
            par ^= 1;                                                           \
        } while (0)

        // step 0's H0 has no G.S term: its values (and max) are known before the first barrier
        float hv[2][2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) hv[i][j][q] = 0.f;
        if (steps > 0) {
            h0_load(srow, MT);
            const uint32_t dbase = drop ? dropout_base(c, task, 0, 0) : 0u;
            float mxv = 0.f;
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const int r = 16 * i + g + 8 * hq, h = hc + 8 * j;
                        const float z0 = ap[i][j][hq].x + b0r[j][0], z1 = ap[i][j][hq].y + b0r[j][1];
                        bool k0 = (r < n) & (z0 > 0.f), k1 = (r < n) & (z1 > 0.f);
                        if (drop) {
                            const uint32_t bits = dropout_bits(dbase, r, h);
                            k0 = k0 & ((bits & 0xFFFFu) >= thr);
                            k1 = k1 & ((bits >> 16) >= thr);
                        }
                        hv[i][j][2 * hq] = k0 ? z0 * dsc : 0.f;
                        hv[i][j][2 * hq + 1] = k1 ? z1 * dsc : 0.f;
                        gate0 |= (uint32_t(k0) | (uint32_t(k1) << 1)) << (2 * (4 * i + 2 * j + hq));
                        mxv = fmaxf(mxv, fmaxf(hv[i][j][2 * hq], hv[i][j][2 * hq + 1]));
                    }
            block_max_push(s.mx + 16 * FX_H0, mxv);
        }
        __syncthreads();                                  // P1: maxes of W1, G, hp, H0(step 0)
        {
            const float mw = slot_max(s.mx + 16 * FX_W1), mg = slot_max(s.mx + 16 * FX_GS), mh = slot_max(s.mx + 16 * FX_HP);
            e_w1 = fumi_plane_exp_t(__float_as_uint(mw), kTarget);    // fixed for the task: 8 binades of headroom
            e_gs = fumi_plane_exp_t(__float_as_uint(mg), kTarget);
            // |dZ1| <= dsc * sum_c |dL_c| |hp_c| <= dsc * (2 / n) * max |hp|: bound for the first production
            e_dz = e_dzn = fumi_plane_exp_t(__float_as_uint(fmaxf(mh * dsc * 2.f / float(n), 1e-30f)), kTarget);
            if (steps > 0) {
                const float m0 = slot_max(s.mx + 16 * FX_H0);
                e_h0 = e_h0n = fumi_plane_exp_t(__float_as_uint(m0), kTarget);
                par_h0 = 1;
            }
            const float sc = fumi_exp2i(e_w1);
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                    w1m[j][2 * hq] *= sc;
                    w1m[j][2 * hq + 1] *= sc;
                    st_planes2(s.w1h, s.w1l, (16 * w + g + 8 * hq) * kHW + 8 * j + 2 * t, w1m[j][2 * hq], w1m[j][2 * hq + 1], 1.f);
                }
            const float sg = fumi_exp2i(e_gs);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                st_plane1(s.gsh, s.gsl, i * kHG + j, gv[q], sg);
            }
            if (steps > 0) {
                const float sh0 = fumi_exp2i(e_h0);
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq)
                            st_planes2(s.h0h, s.h0l, (16 * i + g + 8 * hq) * kHS + hc + 8 * j, hv[i][j][2 * hq],
                                       hv[i][j][2 * hq + 1], sh0);
            }
        }
        pc.mark(20);
        float loss_sum = 0.f, corr_sum = 0.f;             // query loss / correct predictions of this warp's rows (lane 0)

        // ------------------------------------------------------------------ row-per-warp pieces
        // H1 of row i (lane owns units lane, lane + 32) from the Z1 partial sums, bias, ReLU, dropout
        auto h1_row = [&](int i, int grow, int parts, int pass, float& h1a, float& h1b) {
            float za = s.z1p[i * kS1 + lane] + s.b1s[lane], zb = s.z1p[i * kS1 + lane + 32] + s.b1s[lane + 32];
            if (parts == 2) { za += s.z1p[(16 + i) * kS1 + lane]; zb += s.z1p[(16 + i) * kS1 + lane + 32]; }
            bool ka = za > 0.f, kb = zb > 0.f;
            if (drop) {
                const uint32_t dbase = dropout_base(c, task, pass, 1);
                ka = ka & dropout_keep_bits(dropout_bits(dbase, grow, lane), lane, thr);
                kb = kb & dropout_keep_bits(dropout_bits(dbase, grow, lane + 32), lane + 32, thr);
            }
            h1a = ka ? za * dsc : 0.f;
            h1b = kb ? zb * dsc : 0.f;
        };
        // all N logits of the row in every lane's registers: N partial dot products, then ONE interleaved butterfly
        // (independent shuffle chains; softmax / argmax / dZ1 then need no further exchange)
        auto logits_row = [&](float h1a, float h1b, float (&lg)[kNC]) {
#pragma unroll
            for (int cc = 0; cc < kNC; ++cc) lg[cc] = fmaf(h1a, s.hp[cc * kHD + lane], h1b * s.hp[cc * kHD + lane + 32]);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
                for (int cc = 0; cc < kNC; ++cc) lg[cc] += __shfl_xor_sync(0xffffffffu, lg[cc], off);
#pragma unroll
            for (int cc = 0; cc < kNC; ++cc) lg[cc] = cc < N ? lg[cc] + s.hp[cc * kHD + kH1] : -3.0e38f;   // inert classes
        };
        // Z1 partial sums of the block-wide layer: warp (w >> 3, w & 7).  parts == 1: m tile w >> 3, all K;
        // parts == 2 (16 rows): m tile 0, K half w >> 3.  Stored unscaled in z1p[(16 (w >> 3) + row)][o].
        auto z1_gemm = [&](int parts, int mtiles) {
            const int nt = w & 7, hi8 = w >> 3;
            if (parts == 1 && hi8 >= mtiles) return;
            float acc[1][1][4] = {{{0.f, 0.f, 0.f, 0.f}}};
            const int mrow = parts == 1 ? 16 * hi8 : 0, k0 = parts == 1 ? 0 : 128 * hi8;
            warp_gemm_f16x3<1, 1, false, false>(s.h0h + mrow * kHS + k0, s.h0l + mrow * kHS + k0, kHS, s.w1h + k0 * kHW + 8 * nt,
                                                s.w1l + k0 * kHW + 8 * nt, kHW, parts == 1 ? kH0 : 128, acc);
            const float inv = fumi_exp2i(-e_h0) * fumi_exp2i(-e_w1);
            float* dst = s.z1p + (16 * hi8 + g) * kS1 + 8 * nt + 2 * t;
            *reinterpret_cast<float2*>(dst) = make_float2(acc[0][0][0] * inv, acc[0][0][1] * inv);
            *reinterpret_cast<float2*>(dst + 8 * kS1) = make_float2(acc[0][0][2] * inv, acc[0][0][3] * inv);
        };

        // ------------------------------------------------------------------ inner steps
        for (int st = 0; st < steps; ++st) {
            float* rec = slot ? slot + L.steps + int64_t(st) * L.per_step : nullptr;
            pc.mark(45);
            __syncthreads();                              // B1: H0 planes, W1 planes, b1 / head of this step
            if (st > 0) FUMI_ADOPT(e_h0, e_h0n, par_h0, FX_H0);
            pc.mark(21);
            z1_gemm(MT == 1 ? 2 : 1, MT);
            pc.mark(22);
            __syncthreads();                              // B2: Z1 partial sums
            pc.mark(40);
            // ---- (c) one row per warp: H1, logits, softmax, dL, dZ1 (planes with the lagged exponent)
            {
                const float invn = 1.f / float(n), sc = fumi_exp2i(e_dzn);
                float mxv = 0.f;
                // the warp's rows (w, w + 16) are processed unconditionally and fully unrolled -- two independent
                // instruction streams for the scheduler to interleave; pad rows (i >= n) only skip their stores
#pragma unroll
                for (int rr = 0; rr < MT; ++rr) {
                    const int i = w + 16 * rr;
                    const bool live = i < n;
                    float h1a, h1b;
                    h1_row(i, i, MT == 1 ? 2 : 1, st, h1a, h1b);
                    if (live) {
                        s.h1t[i * kS1 + lane] = h1a;
                        s.h1t[i * kS1 + lane + 32] = h1b;
                    }
                    float lg[kNC];
                    logits_row(h1a, h1b, lg);
                    float mxl = lg[0];
#pragma unroll
                    for (int cc = 1; cc < kNC; ++cc) mxl = fmaxf(mxl, lg[cc]);
                    float sum = 0.f;
#pragma unroll
                    for (int cc = 0; cc < kNC; ++cc) { lg[cc] = fumi_fast_exp(lg[cc] - mxl); sum += lg[cc]; }   // inert: exp(-3e38) = 0
                    const float rs = fumi_fast_rcp(sum);
                    const int y = s.ysS[i];               // (zero for pad rows)
                    float da = 0.f, db = 0.f, dlw = 0.f;
#pragma unroll
                    for (int cc = 0; cc < kNC; ++cc) {
                        const float dl = (lg[cc] * rs - (cc == y ? 1.f : 0.f)) * invn;
                        if (lane == cc) dlw = dl;
                        da = fmaf(dl, s.hp[cc * kHD + lane], da);
                        db = fmaf(dl, s.hp[cc * kHD + lane + 32], db);
                    }
                    da = (live & (h1a > 0.f)) ? da * dsc : 0.f;
                    db = (live & (h1b > 0.f)) ? db * dsc : 0.f;
                    if (live) {
                        if (lane < N) s.lt[i * kLS + lane] = dlw;
                        s.dz1t[i * kS1 + lane] = da;
                        s.dz1t[i * kS1 + lane + 32] = db;
                        st_plane1(s.dzh, s.dzl, i * kHW + lane, da, sc);
                        st_plane1(s.dzh, s.dzl, i * kHW + lane + 32, db, sc);
                    }
                    mxv = fmaxf(mxv, fmaxf(fabsf(da), fabsf(db)));
                }
                mxv = warp_max(mxv);
                if (lane == 0) s.mx[16 * (FX_DZ + par_dz) + w] = mxv;
            }
            pc.mark(41);
            if (rec) {                                    // H0 planes of this step: stable between B1 and B3
                for (int idx = tid; idx < n * 32; idx += NT_) {         // 16-byte pieces: 32 per row and plane
                    const int i = idx >> 5, q = idx & 31;
                    reinterpret_cast<uint4*>(rec + L.oH0h)[idx] = *reinterpret_cast<const uint4*>(&s.h0h[i * kHS + 8 * q]);
                    reinterpret_cast<uint4*>(rec + L.oH0l)[idx] = *reinterpret_cast<const uint4*>(&s.h0l[i * kHS + 8 * q]);
                }
            }
            pc.mark(42);
            __syncthreads();                              // B3: dZ1 planes, dL, H1
            FUMI_ADOPT(e_dz, e_dzn, par_dz, FX_DZ);
            pc.mark(23);
            // ---- (d) dZ0 = (dZ1 W1) * gate ; S += dZ0 ; b0 -= alpha colsum(dZ0): warp-local
            {
                float acc[MT][2][4];
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
                warp_gemm_f16x3<MT, 2, false, true, true>(s.dzh, s.dzl, kHW, s.w1h + 16 * w * kHW, s.w1l + 16 * w * kHW, kHW, kH1, acc);
                // (the MMAs above are in flight: the head-gradient sums of the first threads run under them)
                // ---- small updates by the first threads: head, b1 (their results are first read after the next B1)
                for (int idx = tid; idx < N * kHD; idx += NT_) {
                    const int cc = idx / kHD, o = idx - cc * kHD;
                    const float* hcol = o < kH1 ? s.h1t + o : nullptr;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;       // four independent chains over the rows
                    int i = 0;
                    for (; i + 4 <= n; i += 4) {
                        a0 = fmaf(s.lt[i * kLS + cc], hcol ? hcol[i * kS1] : 1.f, a0);
                        a1 = fmaf(s.lt[(i + 1) * kLS + cc], hcol ? hcol[(i + 1) * kS1] : 1.f, a1);
                        a2 = fmaf(s.lt[(i + 2) * kLS + cc], hcol ? hcol[(i + 2) * kS1] : 1.f, a2);
                        a3 = fmaf(s.lt[(i + 3) * kLS + cc], hcol ? hcol[(i + 3) * kS1] : 1.f, a3);
                    }
                    for (; i < n; ++i) a0 = fmaf(s.lt[i * kLS + cc], hcol ? hcol[i * kS1] : 1.f, a0);
                    s.dhp[idx] = (a0 + a1) + (a2 + a3);
                }
                const float inv = fumi_exp2i(-e_dz) * fumi_exp2i(-e_w1) * dsc;
                float colsum[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
                float mxv = 0.f;
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            const uint32_t gt = gate0 >> (2 * (4 * i + 2 * j + hq));
                            const float d0 = (gt & 1u) ? acc[i][j][2 * hq] * inv : 0.f;
                            const float d1 = (gt & 2u) ? acc[i][j][2 * hq + 1] * inv : 0.f;
                            colsum[j][0] += d0;
                            colsum[j][1] += d1;
                            Sr[i][j][2 * hq] += d0;
                            Sr[i][j][2 * hq + 1] += d1;
                            mxv = fmaxf(mxv, fmaxf(fabsf(Sr[i][j][2 * hq]), fabsf(Sr[i][j][2 * hq + 1])));
                        }
                mxv = warp_max(mxv);
                e_s = fumi_plane_exp(__float_as_uint(mxv));                   // exact: S is this warp's own operand
                const float ssc = fumi_exp2i(e_s);
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq)
                            st_planes2(s.sh, s.sl, (16 * i + g + 8 * hq) * kHS + hc + 8 * j, Sr[i][j][2 * hq], Sr[i][j][2 * hq + 1], ssc);
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float v = colsum[j][q];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        b0r[j][q] -= alpha * v;
                    }
            }
            pc.mark(24);
            if (st + 1 < steps) h0_load(srow, MT);       // next step's projected rows (L2): the loads fly under (e)
            // ---- (e) W1 -= alpha dZ1^T H0 on this warp's rows of W1^T: fp32 master in registers, planes re-split from it
            {
                // tile-outer: the A fragments (this warp's columns of H0, all rows) stay in registers; each pair of n tiles is
                // 6 MMAs per k step followed by its own update / re-split, which overlaps the next pair's MMAs
                uint32_t ah[MT][4], al[MT][4];
                warp_load_a<MT, true>(s.h0h + 16 * w, s.h0l + 16 * w, kHS, ah, al);
                const float inv = alpha * fumi_exp2i(-e_h0) * fumi_exp2i(-e_dz) * fumi_exp2i(e_w1);
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    float acc[2][4];
                    warp_mma_pair<MT, false>(ah, al, s.dzh, s.dzl, kHW, 16 * jp, acc);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            const int j = 2 * jp + h;
                            w1m[j][2 * hq] = fmaf(-inv, acc[h][2 * hq], w1m[j][2 * hq]);
                            w1m[j][2 * hq + 1] = fmaf(-inv, acc[h][2 * hq + 1], w1m[j][2 * hq + 1]);
                            st_planes2(s.w1h, s.w1l, (16 * w + g + 8 * hq) * kHW + 8 * j + 2 * t, w1m[j][2 * hq], w1m[j][2 * hq + 1], 1.f);
                        }
                }
            }
            pc.mark(25);
            // ---- step records (dZ1 planes, H1, dL, head before its update), then the head / b1 updates
            if (rec) {
                for (int idx = tid; idx < n * 8; idx += NT_) {          // 16-byte pieces: 8 per row and plane
                    const int i = idx >> 3, q = idx & 7;
                    reinterpret_cast<uint4*>(rec + L.oDZh)[idx] = *reinterpret_cast<const uint4*>(&s.dzh[i * kHW + 8 * q]);
                    reinterpret_cast<uint4*>(rec + L.oDZl)[idx] = *reinterpret_cast<const uint4*>(&s.dzl[i * kHW + 8 * q]);
                }
                for (int idx = tid; idx < n * kH1; idx += NT_) rec[L.oH1 + idx] = s.h1t[(idx >> 6) * kS1 + (idx & 63)];
                for (int idx = tid; idx < n * N; idx += NT_) {
                    const int i = idx / N, cc = idx - i * N;
                    rec[L.oDL + idx] = s.lt[i * kLS + cc];
                }
                if (tid == 0) {
                    reinterpret_cast<int*>(rec + L.oEXP)[0] = e_h0;
                    reinterpret_cast<int*>(rec + L.oEXP)[1] = e_dz;
                }
            }
            for (int idx = tid; idx < N * kHD; idx += NT_) {            // the thread that computed dhp[idx] above
                if (rec) rec[L.oHP + idx] = s.hp[idx];
                s.hp[idx] -= alpha * s.dhp[idx];
            }
            if (tid >= NT_ - kH1) {                       // the last two warps: the first eleven carry the head update
                const int o = tid - (NT_ - kH1);
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;           // four independent chains over the rows
                int i = 0;
                for (; i + 4 <= n; i += 4) {
                    d0 += s.dz1t[i * kS1 + o];
                    d1 += s.dz1t[(i + 1) * kS1 + o];
                    d2 += s.dz1t[(i + 2) * kS1 + o];
                    d3 += s.dz1t[(i + 3) * kS1 + o];
                }
                for (; i < n; ++i) d0 += s.dz1t[i * kS1 + o];
                s.b1s[o] -= alpha * ((d0 + d1) + (d2 + d3));
            }
            // ---- (a) of the next step: H0 = act(A + b0 - alpha G S), warp-local (S, b0 are this warp's own)
            if (st + 1 < steps) {
                float acc[2][2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
                __syncwarp();                             // this warp's S planes are complete
                warp_gemm_f16x3<MT, 2, false, false, true>(s.gsh, s.gsl, kHG, s.sh + 16 * w, s.sl + 16 * w, kHS, RS,
                                                     reinterpret_cast<float(&)[MT][2][4]>(acc));
                gate0 = h0_finish(acc, MT, 0, n, alpha * fumi_exp2i(-e_gs) * fumi_exp2i(-e_s), st + 1);
            }
            pc.mark(26);
        }

        // ------------------------------------------------------------------ adapted state is final
        // The query pass below is barrier-free and owns ROWS, not hidden units, so what was warp-private goes public
        // here: b0 and the exponents of the S planes of each 16-column slab; the fp32 masters leave the registers.
        {
            float mxv = 0.f;        // W1 planes keep the exponent of the initial weights: they overflow if the weights grew 250x
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) mxv = fmaxf(mxv, fabsf(w1m[j][q]));
            mxv = warp_max(mxv);
            if (lane == 0) { s.red[32 + w] = mxv; s.es[w] = e_s; }
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    *reinterpret_cast<float2*>(&s.b0s[hc + 8 * j]) = make_float2(b0r[j][0], b0r[j][1]);
            }
        }
        if (slot) {                                       // adapted state (fp32): parity dumps, backward prologue
            float* sl = const_cast<float*>(slot);
            const float winv = fumi_exp2i(-e_w1);
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq)
                    *reinterpret_cast<float2*>(&sl[L.w1t + (16 * w + g + 8 * hq) * kH1 + 8 * j + 2 * t]) =
                        make_float2(w1m[j][2 * hq] * winv, w1m[j][2 * hq + 1] * winv);
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    *reinterpret_cast<float2*>(&sl[L.b0 + hc + 8 * j]) = make_float2(b0r[j][0], b0r[j][1]);
            }
            if (tid >= NT_ - kH1) sl[L.b1 + tid - (NT_ - kH1)] = s.b1s[tid - (NT_ - kH1)];       // (the thread that updated it)
            for (int idx = tid; idx < N * kHD; idx += NT_) sl[L.head + idx] = s.hp[idx];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const int r = 16 * i + g + 8 * hq;
                        if (r < n)
                            *reinterpret_cast<float2*>(&sl[L.S + int64_t(r) * kH0 + hc + 8 * j]) =
                                make_float2(Sr[i][j][2 * hq], Sr[i][j][2 * hq + 1]);
                    }
        }
        // plane exponent of the query activations: the one derived from the last support H0 (8 binades of headroom);
        // without inner steps, from a bound on the first rows' pre-activations
        if (steps == 0) {
            int rid0[2][2];
            h0_rows(P.qry_rows + b * m, min(32, m), rid0);
            h0_load(rid0, 2);
            float mxv = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq)
                        mxv = fmaxf(mxv, fmaxf(fabsf(ap[i][j][hq].x + b0r[j][0]), fabsf(ap[i][j][hq].y + b0r[j][1])) * dsc);
            block_max_push(s.mx + 16 * FX_H0, mxv);
        }
        int rid_first[2] = {-1, -1};                      // row ids of this warp's first query tile: loaded ahead of the barrier
#pragma unroll
        for (int hq = 0; hq < 2; ++hq)
            if (16 * w + g + 8 * hq < m) rid_first[hq] = int(__ldg(&P.qry_rows[b * m + 16 * w + g + 8 * hq]));
        pc.mark(43);
        __syncthreads();                                  // QB: S / W1 planes, head, b1, b0, exponents of the adapted model
        if (steps == 0) e_h0n = fumi_plane_exp_t(__float_as_uint(fmaxf(slot_max(s.mx + 16 * FX_H0), 1e-30f)), kTarget);
        pc.mark(28);

        // ------------------------------------------------------------------ query scoring: one warp per 16 query rows
        // Z0q = Aq + b0 - alpha Gq S goes through in 16-column slabs; the activations of a slab ARE the A fragments of one
        // k step of Z1q = H0q W1^T (accumulator layout == A layout of m16n8k16), so H0q never touches shared memory and the
        // whole pass needs no block barrier: per slab 12 MMAs (Gram fragments held in registers x S planes) + 24 MMAs
        // (H0q fragments x W1^T planes).  The Z0 MMAs of slab c+1 are issued ahead of the Z1 MMAs of slab c.
        {
            const int64_t* qrows = P.qry_rows + b * m;
            const int e_hq = e_h0n;
            const float hsc = fumi_exp2i(e_hq);
            const uint32_t db0 = drop ? dropout_base(c, task, steps, 0) : 0u;
            const uint32_t db1 = drop ? dropout_base(c, task, steps, 1) : 0u;
            float* sl = const_cast<float*>(slot);
            if (slot && tid < (m + 31) / 32) reinterpret_cast<int*>(sl + L.qEXP)[tid] = e_hq;
            const int l7 = lane & 7, b3 = (lane >> 3) & 1, b4 = lane >> 4;
            const int boffw = (l7 + 8 * b3) * kHW + 8 * b4;          // ldmatrix row of this lane within a W1^T slab
            float hmax = 0.f;
            for (int mtile = w; 16 * mtile < m; mtile += 16) {
                const int r0 = 16 * mtile, tr = min(16, m - r0);
                int rid[2];
                bool live[2];
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                    live[hq] = g + 8 * hq < tr;
                    rid[hq] = mtile == w ? rid_first[hq] : (live[hq] ? int(__ldg(&qrows[r0 + g + 8 * hq])) : -1);
                }
                // Gram rows of the tile as A fragments (exact per-warp exponent)
                uint32_t gah[MT][4], gal[MT][4];
                int e_g = 0;
                if (steps > 0) {
                    float gvv[MT][4][2];
                    float mxv = 0.f;
#pragma unroll
                    for (int ks = 0; ks < MT; ++ks)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int k = 16 * ks + 8 * (q >> 1) + 2 * t + e;
                                const bool ok = live[q & 1] && k < n;
                                gvv[ks][q][e] = ok ? __ldg(&P.gram[(b * int64_t(n + m) + n + r0 + g + 8 * (q & 1)) * n + k]) : 0.f;
                                mxv = fmaxf(mxv, fabsf(gvv[ks][q][e]));
                            }
                    mxv = warp_max(mxv);
                    e_g = fumi_plane_exp(__float_as_uint(mxv));
                    const float gsc = fumi_exp2i(e_g);
#pragma unroll
                    for (int ks = 0; ks < MT; ++ks)
#pragma unroll
                        for (int q = 0; q < 4; ++q) fumi_split2(gvv[ks][q][0] * gsc, gvv[ks][q][1] * gsc, gah[ks][q], gal[ks][q]);
                } else {
#pragma unroll
                    for (int ks = 0; ks < MT; ++ks)
#pragma unroll
                        for (int q = 0; q < 4; ++q) gah[ks][q] = gal[ks][q] = 0u;
                }
                const float ginv = steps > 0 ? alpha * fumi_exp2i(-e_g) : 0.f;
                // projected rows of the tile: 64-byte pieces (one slab of one row) travel through a per-warp ring of three
                // slabs in the shared memory the support steps no longer need (cp.async, three slabs ahead: the rows are
                // random 1 KB lines of a 119 MB matrix, i.e. DRAM latency); lane (g, t) copies piece t of rows g and g + 8
                float* ring = reinterpret_cast<float*>(s.h0h) + w * (3 * kRingSlab);
                auto a_issue = [&](int slab) {
                    if (slab < 16) {
                        float* dst = ring + (slab % 3) * kRingSlab;
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq)
                            fumi_cp_async16(dst + (g + 8 * hq) * kRingRow + 4 * t,
                                            &P.proj[int64_t(rid[hq] >= 0 ? rid[hq] : 0) * kH0 + 16 * slab + 4 * t]);
                    }
                    fumi_cp_async_commit();
                };
                float2 apq[2][2];                         // ... at this thread's positions [n tile][row half]
                auto a_load = [&](int slab, float2 (&dst)[2][2]) {
                    const float* src = ring + (slab % 3) * kRingSlab;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) dst[h][hq] = *reinterpret_cast<const float2*>(src + (g + 8 * hq) * kRingRow + 8 * h + 2 * t);
                };
                uint32_t ah[4], al[4];                    // H0q fragments of the current slab
                uint32_t rb0[2];                          // row part of the dropout counter
                uint32_t *sth[2], *stl[2];                // stash rows of the H0q planes
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                    rb0[hq] = db0 + uint32_t(r0 + g + 8 * hq) * 0xC2B2AE35u;
                    sth[hq] = slot ? reinterpret_cast<uint32_t*>(sl + L.qH0h) + int64_t(r0 + g + 8 * hq) * (kH0 / 2) + t : nullptr;
                    stl[hq] = slot ? reinterpret_cast<uint32_t*>(sl + L.qH0l) + int64_t(r0 + g + 8 * hq) * (kH0 / 2) + t : nullptr;
                }
                auto h0_slab = [&](int slab, const float (&acc)[2][4], const float2 (&a)[2][2]) {
                    const float gs = steps > 0 ? ginv * fumi_exp2i(-s.es[slab]) : 0.f;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int col = 16 * slab + 8 * h + 2 * t;
                        const float2 bb = *reinterpret_cast<const float2*>(&s.b0s[col]);
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            const float z0 = a[h][hq].x + bb.x - gs * acc[h][2 * hq];
                            const float z1 = a[h][hq].y + bb.y - gs * acc[h][2 * hq + 1];
                            bool k0 = live[hq] & (z0 > 0.f), k1 = live[hq] & (z1 > 0.f);
                            if (drop) {
                                const uint32_t bits = fumi_lowbias32(rb0[hq] + uint32_t(col >> 1) * 0x27D4EB2Fu);   // == dropout_bits
                                k0 = k0 & ((bits & 0xFFFFu) >= thr);
                                k1 = k1 & ((bits >> 16) >= thr);
                            }
                            const float v0 = k0 ? z0 * dsc : 0.f, v1 = k1 ? z1 * dsc : 0.f;
                            hmax = fmaxf(hmax, fmaxf(v0, v1));
                            fumi_split2(v0 * hsc, v1 * hsc, ah[2 * h + hq], al[2 * h + hq]);
                            if (slot && live[hq]) {
                                sth[hq][8 * slab + 4 * h] = ah[2 * h + hq];
                                stl[hq][8 * slab + 4 * h] = al[2 * h + hq];
                            }
                        }
                    }
                };
                auto z0_slab = [&](int slab, float (&acc)[2][4]) {
                    if (steps > 0) {
                        warp_mma_pair<MT, false>(gah, gal, s.sh, s.sl, kHS, 16 * slab, acc);
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[h][q] = 0.f;
                    }
                };
                float tot[8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) tot[j][q] = 0.f;
                float acc[2][4];
                __syncwarp();                             // every lane is done with the ring (previous tile)
                a_issue(0);
                a_issue(1);
                a_issue(2);
                z0_slab(0, acc);
                fumi_cp_async_wait_n<2>();
                __syncwarp();
                a_load(0, apq);
                __syncwarp();
                a_issue(3);
                h0_slab(0, acc, apq);
                // one accumulator over all 256 hidden units (48 accumulations: ~3e-6 relative from the tensor core's
                // truncating adds, and nothing compounds on the query side)
#pragma unroll 1
                for (int slab = 0; slab < 16; ++slab) {
                    if (slab + 1 < 16) z0_slab(slab + 1, acc);                     // next slab's Z0 MMAs fly under this slab's Z1 MMAs
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        uint32_t bh[4], bl[4];
                        const int o = 16 * slab * kHW + 16 * jp + boffw;
                        fumi_ldsm4t(bh, s.w1h + o);
                        fumi_ldsm4t(bl, s.w1l + o);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            fumi_mma_f16(tot[2 * jp + h], al, bh[2 * h], bh[2 * h + 1]);
                            fumi_mma_f16(tot[2 * jp + h], ah, bl[2 * h], bl[2 * h + 1]);
                            fumi_mma_f16(tot[2 * jp + h], ah, bh[2 * h], bh[2 * h + 1]);
                        }
                    }
                    if (slab + 1 < 16) {
                        fumi_cp_async_wait_n<2>();        // the copies of slab + 1 have landed (two younger groups may be in flight)
                        __syncwarp();
                        a_load(slab + 1, apq);
                        __syncwarp();
                        a_issue(slab + 4);                // into the ring slot just read
                        h0_slab(slab + 1, acc, apq);
                    }
                }
                fumi_cp_async_wait_n<0>();
                // ---- H1q at the accumulator positions (rows g / g + 8, units 8 j + 2 t + {0, 1}), then the N logits of both
                // rows: per-lane partial dot products over its 16 units, summed over the 4 lanes of the quad
                const float zinv = fumi_exp2i(-e_hq) * fumi_exp2i(-e_w1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 bb = *reinterpret_cast<const float2*>(&s.b1s[8 * j + 2 * t]);
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const float z0 = fmaf(tot[j][2 * hq], zinv, bb.x), z1 = fmaf(tot[j][2 * hq + 1], zinv, bb.y);
                        bool k0 = z0 > 0.f, k1 = z1 > 0.f;
                        if (drop) {
                            const uint32_t bits = dropout_bits(db1, r0 + g + 8 * hq, 8 * j + 2 * t);
                            k0 = k0 & ((bits & 0xFFFFu) >= thr);
                            k1 = k1 & ((bits >> 16) >= thr);
                        }
                        tot[j][2 * hq] = k0 ? z0 * dsc : 0.f;
                        tot[j][2 * hq + 1] = k1 ? z1 * dsc : 0.f;
                        if (slot && live[hq])
                            *reinterpret_cast<float2*>(&sl[L.qH1 + int64_t(r0 + g + 8 * hq) * kH1 + 8 * j + 2 * t]) =
                                make_float2(tot[j][2 * hq], tot[j][2 * hq + 1]);
                    }
                }
                float lg[2][kNC];
#pragma unroll
                for (int cc = 0; cc < kNC; ++cc) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float h0 = s.hp[cc * kHD + 8 * j + 2 * t], h1 = s.hp[cc * kHD + 8 * j + 2 * t + 1];
                        s0 = fmaf(tot[j][0], h0, fmaf(tot[j][1], h1, s0));
                        s1 = fmaf(tot[j][2], h0, fmaf(tot[j][3], h1, s1));
                    }
                    lg[0][cc] = s0;
                    lg[1][cc] = s1;
                }
#pragma unroll
                for (int off = 1; off <= 2; off <<= 1)
#pragma unroll
                    for (int cc = 0; cc < kNC; ++cc) {
                        lg[0][cc] += __shfl_xor_sync(0xffffffffu, lg[0][cc], off);
                        lg[1][cc] += __shfl_xor_sync(0xffffffffu, lg[1][cc], off);
                    }
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                    const int64_t q = b * m + r0 + g + 8 * hq;
                    const int y = live[hq] ? int(__ldg(&P.qry_y[q])) : 0;
#pragma unroll
                    for (int cc = 0; cc < kNC; ++cc) lg[hq][cc] = cc < N ? lg[hq][cc] + s.hp[cc * kHD + kH1] : -3.0e38f;   // inert classes
                    float mxl = lg[hq][0], ly = 0.f;
                    int bi = 0;                           // argmax with ties to the lowest index (torch.max, fumi.py:180)
#pragma unroll
                    for (int cc = 0; cc < kNC; ++cc) {
                        if (lg[hq][cc] > mxl) { mxl = lg[hq][cc]; bi = cc; }
                        if (cc == y) ly = lg[hq][cc];
                        if (live[hq] && cc < N && (cc & 3) == t) P.logits[q * N + cc] = lg[hq][cc];
                    }
                    float sum = 0.f;
#pragma unroll
                    for (int cc = 0; cc < kNC; ++cc) { lg[hq][cc] = fumi_fast_exp(lg[hq][cc] - mxl); sum += lg[hq][cc]; }
                    if (slot && live[hq]) {
                        const float rs = fumi_fast_rcp(sum);
#pragma unroll
                        for (int cc = 0; cc < kNC; ++cc)
                            if (cc < N && (cc & 3) == t) sl[L.qLG + int64_t(r0 + g + 8 * hq) * N + cc] = lg[hq][cc] * rs - (cc == y ? 1.f : 0.f);
                    }
                    if (live[hq] && t == 0) {
                        P.preds[q] = bi;
                        loss_sum += (fumi_fast_log(sum) + mxl) - ly;
                        corr_sum += bi == y ? 1.f : 0.f;
                    }
                }
            }
            loss_sum = warp_sum(loss_sum);
            corr_sum = warp_sum(corr_sum);
            hmax = warp_max(hmax);
            if (lane == 0) {
                s.red[w] = loss_sum;
                s.red[16 + w] = corr_sum;
                s.red[48 + w] = plane_overflow(hmax, e_hq) ? 1.f : 0.f;
            }
        }
        pc.mark(30);
        __syncthreads();
        if (tid == 0) {
            float ls = 0.f, cs = 0.f, wm = 0.f, ov = 0.f;
            for (int q = 0; q < 16; ++q) { ls += s.red[q]; cs += s.red[16 + q]; wm = fmaxf(wm, s.red[32 + q]); ov += s.red[48 + q]; }
            bad = bad || !(wm < 60000.f) || ov > 0.f;
            P.task_loss[b] = bad ? __uint_as_float(0x7FC00000u) : ls / float(m);
            P.task_acc[b] = cs / float(m);
        }
        pc.mark(33);
    }
#undef FUMI_ADOPT
}

size_t smem_v_bytes(int N) { SmemV t; return carve_v(nullptr, t, N); }
int class_bucket(int N) { return N <= 5 ? 5 : (N <= 8 ? 8 : 11); }

}  // namespace

bool episode_f16_supported(const fumi_episode_cfg& c) {
    // both kernels size their head buffers by the class bucket (5 / 8 / 11); next to the backward's two W1-shaped plane
    // pairs 11 classes is what 227 KB of shared memory holds
    return c.num_support <= 32 && c.num_ways <= 11 && smem_v_bytes(class_bucket(c.num_ways)) <= 227 * 1024 &&
           episode_bwd_f16_smem_bytes(class_bucket(c.num_ways)) <= 227 * 1024;
}

int launch_episode_fwd_f16(const EpiParams& P, int grid, void* stream) {
    const size_t smem = smem_v_bytes(class_bucket(P.cfg.num_ways));
#ifndef FUMI_EMU
#define FUMI_SMEM_ATTR(kern)                                                                                       \
    do {                                                                                                            \
        cudaError_t e__ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));       \
        if (e__ != cudaSuccess) return fumi_cuda_fail(e__, "cudaFuncSetAttribute(episode_fwd_v2_kernel)");          \
    } while (0)
#else
#define FUMI_SMEM_ATTR(kern) ((void)0)
#endif
#define FUMI_FWD_LAUNCH(MT_, NC_)                                                   \
    do {                                                                            \
        if (P.cfg.dropout_p > 0.f) {                                                \
            FUMI_SMEM_ATTR((episode_fwd_v2_kernel<MT_, NC_, true, true>));          \
            FUMI_LAUNCH((episode_fwd_v2_kernel<MT_, NC_, true, true>), grid, kThreads16, smem, stream, P); \
        } else if (P.save) {                                                        \
            FUMI_SMEM_ATTR((episode_fwd_v2_kernel<MT_, NC_, true, false>));         \
            FUMI_LAUNCH((episode_fwd_v2_kernel<MT_, NC_, true, false>), grid, kThreads16, smem, stream, P); \
        } else {                                                                    \
            FUMI_SMEM_ATTR((episode_fwd_v2_kernel<MT_, NC_, false, false>));        \
            FUMI_LAUNCH((episode_fwd_v2_kernel<MT_, NC_, false, false>), grid, kThreads16, smem, stream, P); \
        }                                                                           \
    } while (0)
    const int nc = class_bucket(P.cfg.num_ways);
    if (P.cfg.num_support <= 16) {
        if (nc == 5) FUMI_FWD_LAUNCH(1, 5); else if (nc == 8) FUMI_FWD_LAUNCH(1, 8); else FUMI_FWD_LAUNCH(1, 11);
    } else {
        if (nc == 5) FUMI_FWD_LAUNCH(2, 5); else if (nc == 8) FUMI_FWD_LAUNCH(2, 8); else FUMI_FWD_LAUNCH(2, 11);
    }
#undef FUMI_FWD_LAUNCH
#undef FUMI_SMEM_ATTR
    FUMI_CHECK_LAUNCH("episode_fwd_v2_kernel");
    return FUMI_OK;
}

}  // namespace fumi_epi
