// Hand-written second-order backward of the fused episode (outer_loss.backward() through create_graph=True inner
// steps: fumi/models/fumi.py:165-176,190-192; fumi/models/maml.py:173-177,188-190) on fp16 hi/lo operand planes,
// NK <= 32.  Same recursion as oracle/episode_np.py's reverse sweep, in the Gram form of DESIGN.md section 2.
//
// One persistent CTA of 16 warps walks a task; warp w owns hidden units [16w, 16w+16): rows 16w.. of W1^T and of its
// adjoint, columns 16w.. of every [rows][256] quantity, and -- in registers -- its slab of the adjoint of S, of the
// adjoint of b0 and of this step's contribution to the adjoint of W1^T.  Products that contract over the hidden units
// (r_dZ1) go through shared memory and block barriers; everything whose output is a slab of hidden units is a
// warp-local chain of mma.sync tiles between them.  All GEMM operands are fp16 planes (warp_mma.cuh); the step records
// arrive from the forward already as planes (LayoutF) and are copied with cp.async.  Exponents of the planes produced
// here and consumed block-wide (r_dH0, r_Z1, W1_s, a_W1, dZ1q, Gram tiles) are lagged as in episode_fwd_f16.cu.
//
// Query pass: 32-row tiles (the region of the a_W1 planes, unused until the sweep starts, holds a second H0q buffer,
// the dZ0q planes and the Gram tile).  Reverse sweep: 16-row tiles, because W1_s and a_W1 (2 x 72 KB of planes) share
// the SM's shared memory with the tiles.
#include "episode_common.cuh"

namespace fumi_epi {
namespace {

constexpr int kLB = 16;                     // row stride of the dL / r_L tiles (N <= 11 on this path)
enum { BX_W1 = 0, BX_AW = 2, BX_TT = 4, BX_RZ = 6, BX_DZ = 8, BX_GQ = 10, BX_GS = 12, BX_HP = 13, BX_COUNT = 14 };

struct SmemW {
    fumi_half *w1h, *w1l, *awh, *awl, *h0h, *h0l, *dzh, *dzl, *rzh, *rzl, *gsh, *gsl;
    fumi_half *h0bh, *h0bl, *bzh, *bzl, *gqh, *gql;       // query pass only: aliases of the a_W1 plane region
    float *h1t, *rzp, *lt, *rlt, *hp, *ahp, *rhp, *b1s, *ab1, *rb1, *mx;
    long long *rows, *srows;
    int *ys, *sys;
};
__host__ __device__ inline size_t carve_w(char* base, SmemW& s, int N) {
    char* p = base;
    auto take = [&](size_t bytes) { char* r = p; p += (bytes + 15) & ~size_t(15); return r; };
    s.rows = reinterpret_cast<long long*>(take(32 * 8));
    s.ys = reinterpret_cast<int*>(take(32 * 4));
    s.srows = reinterpret_cast<long long*>(take(32 * 8));
    s.sys = reinterpret_cast<int*>(take(32 * 4));
    s.mx = reinterpret_cast<float*>(take(BX_COUNT * 16 * 4));
    // zero-filled once per kernel from here
    s.w1h = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));  s.w1l = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));
    s.awh = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));  s.awl = reinterpret_cast<fumi_half*>(take(kH0 * kHW * 2));
    s.h0h = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));   s.h0l = reinterpret_cast<fumi_half*>(take(32 * kHS * 2));
    s.dzh = reinterpret_cast<fumi_half*>(take(32 * kHW * 2));   s.dzl = reinterpret_cast<fumi_half*>(take(32 * kHW * 2));
    s.rzh = reinterpret_cast<fumi_half*>(take(16 * kHW * 2));   s.rzl = reinterpret_cast<fumi_half*>(take(16 * kHW * 2));
    s.gsh = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));   s.gsl = reinterpret_cast<fumi_half*>(take(32 * kHG * 2));
    s.h1t = reinterpret_cast<float*>(take(32 * kS1 * 4));
    s.rzp = reinterpret_cast<float*>(take(2 * 16 * kS1 * 4));
    s.lt = reinterpret_cast<float*>(take(32 * kLB * 4));
    s.rlt = reinterpret_cast<float*>(take(16 * kLB * 4));
    s.hp = reinterpret_cast<float*>(take(size_t(N) * kHD * 4));
    s.ahp = reinterpret_cast<float*>(take(size_t(N) * kHD * 4));
    s.rhp = reinterpret_cast<float*>(take(size_t(N) * kHD * 4));
    s.b1s = reinterpret_cast<float*>(take(kH1 * 4));
    s.ab1 = reinterpret_cast<float*>(take(kH1 * 4));
    s.rb1 = reinterpret_cast<float*>(take(kH1 * 4));
    // query-pass aliases inside [awh, awl + ...): 2 x 16,896 + 2 x 16,896 + 2 x 2,560 = 72,704 <= 73,728 bytes
    char* q = reinterpret_cast<char*>(s.awh);
    s.h0bh = reinterpret_cast<fumi_half*>(q); q += 32 * kHS * 2;
    s.h0bl = reinterpret_cast<fumi_half*>(q); q += 32 * kHS * 2;
    s.bzh = reinterpret_cast<fumi_half*>(q);  q += 32 * kHS * 2;
    s.bzl = reinterpret_cast<fumi_half*>(q);  q += 32 * kHS * 2;
    s.gqh = reinterpret_cast<fumi_half*>(q);  q += 32 * kHG * 2;
    s.gql = reinterpret_cast<fumi_half*>(q);
    return size_t(p - base);
}

// gates of a pair of plane elements: the stored activation is positive iff its hi or lo half is non-zero
__device__ __forceinline__ uint32_t plane_gates2(const fumi_half* hi, const fumi_half* lo, int off) {
    const uint32_t h = *reinterpret_cast<const uint32_t*>(hi + off), l = *reinterpret_cast<const uint32_t*>(lo + off);
    const uint32_t v = h | l;
    return uint32_t((v & 0x7FFFu) != 0u) | (uint32_t((v & 0x7FFF0000u) != 0u) << 1);
}

// MT: 16-row tiles of the support set; kNC: compile-time class count (>= N; head buffers zero-padded to kNC rows)
template <int MT, int kNC>
__global__ void __launch_bounds__(kThreads16, 1) episode_bwd_v2_kernel(EpiParams P) {
    constexpr int RS = 16 * MT;
    constexpr int NT_ = kThreads16;
    FUMI_DYN_SMEM(float, smem_raw);
    const fumi_episode_cfg& c = P.cfg;
    SmemW s;
    carve_w(reinterpret_cast<char*>(smem_raw), s, kNC);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n = c.num_support, m = c.num_query, N = c.num_ways, steps = c.steps;
    const float alpha = c.step_size;
    const LayoutF L = make_layout_f(c);
    const float sc = dropout_scale(c);
    const int hc = 16 * w + 2 * t;                       // this thread's column pairs: hc + 8 j + {0, 1}
    PhaseClock pc;
    pc.start(P.phase);

    {   // planes (pads must be finite) and the dL tiles (columns N .. of inert classes must read as zero)
        uint32_t* z = reinterpret_cast<uint32_t*>(s.w1h);
        const int nz = int((reinterpret_cast<char*>(s.h1t) - reinterpret_cast<char*>(s.w1h)) / 4);
        for (int idx = tid; idx < nz; idx += NT_) z[idx] = 0u;
        for (int idx = tid; idx < 32 * kLB; idx += NT_) s.lt[idx] = 0.f;
    }
    __syncthreads();

#define FUMI_ADOPT(cur, next, par, slot_id)                                     \
    do {                                                                        \
        const float mx__ = slot_max(s.mx + 16 * ((slot_id) + (par)));           \
        cur = next;                                                             \
        bad = bad || plane_overflow(mx__, cur);                                 \
        if (mx__ > 0.f) next = fumi_plane_exp_t(__float_as_uint(mx__), kTarget); \
        par ^= 1;                                                               \
    } while (0)

    for (int64_t b = blockIdx.x; b < P.B; b += gridDim.x) {
        const float* slot = P.stash + b * P.slot_floats;
        bool bad = false;
        int e_w1, e_w1n, e_aw = 0, e_awn = 0, e_tt = 0, e_ttn = 0, e_rz = 0, e_rzn = 0, e_dz, e_dzn, e_gs, e_gq, e_gqn, e_h0 = 0;
        int par_w1 = 0, par_aw = 0, par_tt = 0, par_rz = 0, par_dz = 0, par_gq = 0;

        // ------------------------------------------------------------------ prologue: adapted state, zero adjoints
        float w1v[32];                                    // W1_S^T row k = tid >> 1, columns 32 (tid & 1) ..
        {
            const float4* src = reinterpret_cast<const float4*>(slot + L.w1t) + tid * 8;
            float mxv = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = __ldg(&src[q]);
                w1v[4 * q] = v.x; w1v[4 * q + 1] = v.y; w1v[4 * q + 2] = v.z; w1v[4 * q + 3] = v.w;
                mxv = fmaxf(mxv, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
            block_max_push(s.mx + 16 * BX_W1, mxv);
        }
        float gv[2];
        {
            float mxv = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                gv[q] = (i < n && j < n) ? __ldg(&P.gram[(b * int64_t(n + m) + i) * n + j]) : 0.f;
                mxv = fmaxf(mxv, fabsf(gv[q]));
            }
            block_max_push(s.mx + 16 * BX_GS, mxv);
        }
        {
            float mxv = 0.f;
            for (int idx = tid; idx < kNC * kHD; idx += NT_) {
                const float v = idx < N * kHD ? slot[L.head + idx] : 0.f;      // rows N .. kNC-1: inert classes
                s.hp[idx] = v;
                s.ahp[idx] = 0.f;
                if (idx % kHD < kH1) mxv = fmaxf(mxv, fabsf(v));
            }
            block_max_push(s.mx + 16 * BX_HP, mxv);
        }
        if (tid < kH1) { s.b1s[tid] = slot[L.b1 + tid]; s.ab1[tid] = 0.f; }
        if (tid < 32) {
            s.srows[tid] = tid < n ? P.sup_rows[b * n + tid] : 0;
            s.sys[tid] = tid < n ? int(P.sup_y[b * n + tid]) : 0;
        }
        float aS[2][2][4];                                // adjoint of S at this thread's positions (rows x own columns)
        float wacc[8][4];                                 // adjoint of W1^T rows [16w, 16w+16): fp32 master (planes are re-split from it)
        float ab0r[2][2] = {{0.f, 0.f}, {0.f, 0.f}};      // adjoint of b0 at this thread's columns
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) aS[i][j][q] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) wacc[j][q] = 0.f;

        const float qscale = P.loss_scale / float(m);
        const int ntq = (m + 31) / 32;
        // query tile loads.  H0q planes: cp.async into buffer `buf` (0: h0h/h0l, 1: the alias in the a_W1 region);
        // H1q, dLq, Gram rows, rows / labels: registers now, shared memory at the start of the tile.
        float pu[4], pl = 0.f, pg[2];
        long long prow = 0;
        int py = 0;
        auto q_issue = [&](int tile) {
            const int r0 = 32 * tile, tr = min(32, m - r0);
            fumi_half* dh = (tile & 1) ? s.h0bh : s.h0h;
            fumi_half* dl = (tile & 1) ? s.h0bl : s.h0l;
            const uint4* sh = reinterpret_cast<const uint4*>(slot + L.qH0h) + int64_t(r0) * 32;
            const uint4* sl = reinterpret_cast<const uint4*>(slot + L.qH0l) + int64_t(r0) * 32;
            for (int idx = tid; idx < 32 * 32; idx += NT_) {
                const int i = idx >> 5, q = idx & 31;
                if (i < tr) {
                    fumi_cp_async16(dh + i * kHS + 8 * q, sh + idx);
                    fumi_cp_async16(dl + i * kHS + 8 * q, sl + idx);
                } else {
                    *reinterpret_cast<uint4*>(dh + i * kHS + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(dl + i * kHS + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = tid + q * NT_;
                pu[q] = idx < tr * kH1 ? __ldg(&slot[L.qH1 + int64_t(r0) * kH1 + idx]) : 0.f;
            }
            pl = tid < tr * N ? __ldg(&slot[L.qLG + int64_t(r0) * N + tid]) : 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                pg[q] = (i < tr && j < n) ? __ldg(&P.gram[(b * int64_t(n + m) + n + r0 + i) * n + j]) : 0.f;
            }
            if (tid < 32) {
                prow = tid < tr ? P.qry_rows[b * m + r0 + tid] : 0;
                py = tid < tr ? int(P.qry_y[b * m + r0 + tid]) : 0;
            }
        };
        q_issue(0);
        __syncthreads();                                  // maxes of W1, G, head
        {
            const float mw = slot_max(s.mx + 16 * BX_W1), mg = slot_max(s.mx + 16 * BX_GS), mh = slot_max(s.mx + 16 * BX_HP);
            e_w1 = e_w1n = fumi_plane_exp_t(__float_as_uint(mw), kTarget);
            e_gs = fumi_plane_exp_t(__float_as_uint(mg), kTarget);
            e_gq = e_gqn = e_gs;
            // |dZ1q| <= sc * sum_c |dLq_c| * max |head| <= sc * 2 qscale * max |head|
            e_dz = e_dzn = fumi_plane_exp_t(__float_as_uint(fmaxf(mh * sc * 2.f * qscale, 1e-30f)), kTarget);
            par_w1 = 1;
            const float scw = fumi_exp2i(e_w1);
            const int k = tid >> 1, o0 = 32 * (tid & 1);
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {
                uint32_t ph[4], plo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) fumi_split2(w1v[8 * q8 + 2 * q] * scw, w1v[8 * q8 + 2 * q + 1] * scw, ph[q], plo[q]);
                *reinterpret_cast<uint4*>(&s.w1h[k * kHW + o0 + 8 * q8]) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(&s.w1l[k * kHW + o0 + 8 * q8]) = make_uint4(plo[0], plo[1], plo[2], plo[3]);
            }
            const float sg = fumi_exp2i(e_gs);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                st_plane1(s.gsh, s.gsl, i * kHG + j, gv[q], sg);
            }
        }
        pc.mark(0);

        // ------------------------------------------------------------------ query pass (32-row tiles)
        for (int tile = 0; tile < ntq; ++tile) {
            const int r0 = 32 * tile, tr = min(32, m - r0);
            const fumi_half* qh = (tile & 1) ? s.h0bh : s.h0h;
            const fumi_half* ql = (tile & 1) ? s.h0bl : s.h0l;
            // registers -> shared memory (H1q, dLq scaled, Gram planes with the lagged exponent, rows / labels)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = tid + q * NT_;
                s.h1t[(idx >> 6) * kS1 + (idx & 63)] = pu[q];
            }
            if (tid < 32 * N) s.lt[(tid / N) * kLB + (tid % N)] = pl * qscale;
            {
                const float sg = fumi_exp2i(e_gqn);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int idx = tid + q * NT_, i = idx >> 5, j = idx & 31;
                    st_plane1(s.gqh, s.gql, i * kHG + j, pg[q], sg);
                }
                block_max_push(s.mx + 16 * (BX_GQ + par_gq), fmaxf(fabsf(pg[0]), fabsf(pg[1])));
            }
            if (tid < 32) { s.rows[tid] = prow; s.ys[tid] = py; }
            e_h0 = reinterpret_cast<const int*>(slot + L.qEXP)[tile];
            fumi_cp_async_wait();
            __syncthreads();                              // Q0: tile data
            FUMI_ADOPT(e_gq, e_gqn, par_gq, BX_GQ);
            if (tile + 1 < ntq) q_issue(tile + 1);        // next tile: planes into the other buffer, the rest in registers
            pc.mark(1);
            // one row per warp: dZ1q = gate1 * sc * (dLq . head)  -> planes (lagged exponent)
            {
                const float scd = fumi_exp2i(e_dzn);
                float mxv = 0.f;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int i = w + 16 * rr;
                    float da = 0.f, db = 0.f;
                    if (i < tr) {
#pragma unroll
                        for (int cc = 0; cc < kNC; ++cc) {            // inert classes: zero head rows
                            const float dlc = s.lt[i * kLB + cc];
                            da = fmaf(dlc, s.hp[cc * kHD + lane], da);
                            db = fmaf(dlc, s.hp[cc * kHD + lane + 32], db);
                        }
                        da = s.h1t[i * kS1 + lane] > 0.f ? da * sc : 0.f;
                        db = s.h1t[i * kS1 + lane + 32] > 0.f ? db * sc : 0.f;
                    }
                    st_plane1(s.dzh, s.dzl, i * kHW + lane, da, scd);
                    st_plane1(s.dzh, s.dzl, i * kHW + lane + 32, db, scd);
                    mxv = fmaxf(mxv, fmaxf(fabsf(da), fabsf(db)));
                }
                mxv = warp_max(mxv);
                if (lane == 0) s.mx[16 * (BX_DZ + par_dz) + w] = mxv;
            }
            __syncthreads();                              // Q1: dZ1q planes
            FUMI_ADOPT(e_dz, e_dzn, par_dz, BX_DZ);
            pc.mark(2);
            // a_head += dLq^T [H1q | 1] ; a_b1 += column sums of dZ1q   (first threads)
            for (int idx = tid; idx < N * kHD; idx += NT_) {
                const int cc = idx / kHD, o = idx - cc * kHD;
                const float* hcol = o < kH1 ? s.h1t + o : nullptr;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int i = 0;
                for (; i + 4 <= tr; i += 4) {
                    a0 = fmaf(s.lt[i * kLB + cc], hcol ? hcol[i * kS1] : 1.f, a0);
                    a1 = fmaf(s.lt[(i + 1) * kLB + cc], hcol ? hcol[(i + 1) * kS1] : 1.f, a1);
                    a2 = fmaf(s.lt[(i + 2) * kLB + cc], hcol ? hcol[(i + 2) * kS1] : 1.f, a2);
                    a3 = fmaf(s.lt[(i + 3) * kLB + cc], hcol ? hcol[(i + 3) * kS1] : 1.f, a3);
                }
                for (; i < tr; ++i) a0 = fmaf(s.lt[i * kLB + cc], hcol ? hcol[i * kS1] : 1.f, a0);
                s.ahp[idx] += (a0 + a1) + (a2 + a3);
            }
            if (tid >= NT_ - kH1) {                       // the last two warps: the first ones carry the head sums
                const int o = tid - (NT_ - kH1);
                const float inv = fumi_exp2i(-e_dz);
                float a = 0.f;
                for (int i = 0; i < tr; ++i) a += plane_value(s.dzh, s.dzl, i * kHW + o, inv);
                s.ab1[o] += a;
            }
            // a_W1^T[own rows] += H0q^T dZ1q
            {
                float acc[1][8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
                warp_gemm_f16x3<1, 8, true, false, true>(qh + 16 * w, ql + 16 * w, kHS, s.dzh, s.dzl, kHW, 32, acc);
                const float inv = fumi_exp2i(-e_h0) * fumi_exp2i(-e_dz);
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) wacc[j][q] = fmaf(acc[0][j][q], inv, wacc[j][q]);
            }
            // dZ0q = (dZ1q W1) * gate0 -> d_proj, a_b0, planes of this warp's columns; a_S -= alpha Gq^T dZ0q
            {
                float acc[2][2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
                warp_gemm_f16x3<2, 2, false, true, true>(s.dzh, s.dzl, kHW, s.w1h + 16 * w * kHW, s.w1l + 16 * w * kHW, kHW, kH1, acc);
                const float inv = fumi_exp2i(-e_dz) * fumi_exp2i(-e_w1) * sc;
                float colsum[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
                float mxv = 0.f;
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            const int r = 16 * i + g + 8 * hq, h = hc + 8 * j;
                            const uint32_t gt = plane_gates2(qh, ql, r * kHS + h);          // (pad rows hold zeros: gates 0)
                            const float d0 = (gt & 1u) ? acc[i][j][2 * hq] * inv : 0.f;
                            const float d1 = (gt & 2u) ? acc[i][j][2 * hq + 1] * inv : 0.f;
                            acc[i][j][2 * hq] = d0;
                            acc[i][j][2 * hq + 1] = d1;
                            colsum[j][0] += d0;
                            colsum[j][1] += d1;
                            mxv = fmaxf(mxv, fmaxf(fabsf(d0), fabsf(d1)));
                            atomic_add2_if(r < tr, &P.d_proj[s.rows[r] * kH0 + h], d0, d1);
                        }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float v = colsum[j][q];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        ab0r[j][q] += v;
                    }
                if (steps > 0 && !c.first_order) {
                    mxv = warp_max(mxv);
                    const int e_bz = fumi_plane_exp(__float_as_uint(mxv));
                    const float bsc = fumi_exp2i(e_bz);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int hq = 0; hq < 2; ++hq)
                                st_planes2(s.bzh, s.bzl, (16 * i + g + 8 * hq) * kHS + hc + 8 * j, acc[i][j][2 * hq],
                                           acc[i][j][2 * hq + 1], bsc);
                    __syncwarp();
                    float as[MT][2][4];
#pragma unroll
                    for (int i = 0; i < MT; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) as[i][j][q] = 0.f;
                    warp_gemm_f16x3<MT, 2, true, false, true>(s.gqh, s.gql, kHG, s.bzh + 16 * w, s.bzl + 16 * w, kHS, 32, as);
                    const float ainv = -alpha * fumi_exp2i(-e_gq) * fumi_exp2i(-e_bz);
#pragma unroll
                    for (int i = 0; i < MT; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) aS[i][j][q] = fmaf(as[i][j][q], ainv, aS[i][j][q]);
                }
            }
            __syncthreads();                              // Q2: tile buffers free
            pc.mark(3);
        }

        // a_W1 planes from the query-pass sum (exact exponent; the region stops being the query pass's scratch here)
        {
            float mxv = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) mxv = fmaxf(mxv, fabsf(wacc[j][q]));
            block_max_push(s.mx + 16 * BX_AW, mxv);
            __syncthreads();
            e_aw = e_awn = fumi_plane_exp_t(__float_as_uint(slot_max(s.mx + 16 * BX_AW)), kTarget);
            par_aw = 1;
            const float asc = fumi_exp2i(e_aw);
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int hq = 0; hq < 2; ++hq)
                    st_planes2(s.awh, s.awl, (16 * w + g + 8 * hq) * kHW + 8 * j + 2 * t, wacc[j][2 * hq], wacc[j][2 * hq + 1], asc);
        }
        pc.mark(4);

        // ------------------------------------------------------------------ inner steps in reverse (16-row tiles)
        if (!c.first_order && steps > 0) {
            bool tt_first = true, rz_first = true;
            // step records -> shared memory: H0 planes of all rows into (h0 | tt) = h0h rows 0..31, dZ1 planes, head
            // One step's records -> shared memory, issued a whole step ahead (cp.async; the small dL / head blocks travel in
            // registers and are stored at the start of the step): H0 planes of all rows into (h0 | tt) = h0h rows 0..31,
            // dZ1 planes, H1.
            float dl_pref = 0.f, hp_pref[2] = {0.f, 0.f};
            int ex_pref[2] = {0, 0};
            auto s_issue = [&](int st) {
                const float* rec = slot + L.steps + int64_t(st) * L.per_step;
                const uint4* sh = reinterpret_cast<const uint4*>(rec + L.oH0h);
                const uint4* sl = reinterpret_cast<const uint4*>(rec + L.oH0l);
                for (int idx = tid; idx < RS * 32; idx += NT_) {
                    const int i = idx >> 5, q = idx & 31;
                    if (i < n) {
                        fumi_cp_async16(s.h0h + i * kHS + 8 * q, sh + idx);
                        fumi_cp_async16(s.h0l + i * kHS + 8 * q, sl + idx);
                    } else {
                        *reinterpret_cast<uint4*>(s.h0h + i * kHS + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(s.h0l + i * kHS + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                const uint4* zh = reinterpret_cast<const uint4*>(rec + L.oDZh);
                const uint4* zl = reinterpret_cast<const uint4*>(rec + L.oDZl);
                for (int idx = tid; idx < RS * 8; idx += NT_) {
                    const int i = idx >> 3, q = idx & 7;
                    if (i < n) {
                        fumi_cp_async16(s.dzh + i * kHW + 8 * q, zh + idx);
                        fumi_cp_async16(s.dzl + i * kHW + 8 * q, zl + idx);
                    } else {
                        *reinterpret_cast<uint4*>(s.dzh + i * kHW + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(s.dzl + i * kHW + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                const uint4* h1 = reinterpret_cast<const uint4*>(rec + L.oH1);
                for (int idx = tid; idx < RS * 16; idx += NT_) {          // H1 [n][64] fp32: 16 pieces per row
                    const int i = idx >> 4, q = idx & 15;
                    if (i < n) fumi_cp_async16(s.h1t + i * kS1 + 4 * q, h1 + idx);
                    else *reinterpret_cast<uint4*>(s.h1t + i * kS1 + 4 * q) = make_uint4(0u, 0u, 0u, 0u);
                }
                ex_pref[0] = __ldg(reinterpret_cast<const int*>(rec + L.oEXP));
                ex_pref[1] = __ldg(reinterpret_cast<const int*>(rec + L.oEXP) + 1);
                dl_pref = tid < n * N ? __ldg(&rec[L.oDL + tid]) : 0.f;
#pragma unroll
                for (int q = 0; q < 2; ++q) hp_pref[q] = tid + q * NT_ < N * kHD ? __ldg(&rec[L.oHP + tid + q * NT_]) : 0.f;
            };
            // second row tile of a step: its H0 planes (rows 16..31 of the records, L2-resident since s_issue) replace the
            // first tile's in rows 0..15
            auto t_reload = [&](int st) {
                const float* rec = slot + L.steps + int64_t(st) * L.per_step;
                const uint4* sh = reinterpret_cast<const uint4*>(rec + L.oH0h) + 16 * 32;
                const uint4* sl = reinterpret_cast<const uint4*>(rec + L.oH0l) + 16 * 32;
                for (int idx = tid; idx < 16 * 32; idx += NT_) {
                    const int i = idx >> 5, q = idx & 31;
                    if (i < n - 16) {
                        fumi_cp_async16(s.h0h + i * kHS + 8 * q, sh + idx);
                        fumi_cp_async16(s.h0l + i * kHS + 8 * q, sl + idx);
                    } else {
                        *reinterpret_cast<uint4*>(s.h0h + i * kHS + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(s.h0l + i * kHS + 8 * q) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            };
            s_issue(steps - 1);
            for (int st = steps - 1; st >= 0; --st) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (tid + q * NT_ < N * kHD) { s.hp[tid + q * NT_] = hp_pref[q]; s.rhp[tid + q * NT_] = 0.f; }   // rows >= N stay 0
                if (tid < n * N) s.lt[(tid / N) * kLB + (tid % N)] = dl_pref;
                if (tid < kH1) s.rb1[tid] = 0.f;
                e_h0 = ex_pref[0];
                const int e_dzs = ex_pref[1];
                // both row tiles of a step see the adjoint of S as it was BEFORE the step: the second tile's rows are
                // snapshotted here, the a_S updates of the first tile then go straight into aS
                float aS1[2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) aS1[j][q] = aS[1][j][q];
                float gb0[2][2], rb0r[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
                for (int j = 0; j < 2; ++j) { gb0[j][0] = -alpha * ab0r[j][0]; gb0[j][1] = -alpha * ab0r[j][1]; }
                // (wacc is the fp32 master of this warp's rows of a_W1^T and runs on across the steps: the planes keep the
                // value from before the step until they are re-split from it at the end of the step)
                fumi_cp_async_wait();
                __syncthreads();                          // U0: records of the step in shared memory; a_W1 planes folded
                if (st < steps - 1) FUMI_ADOPT(e_aw, e_awn, par_aw, BX_AW);
                pc.mark(5);
                // ---- undo W1_{s+1} = W1_s - alpha dZ1^T H0 on this warp's rows, in place on the planes
                {
                    float acc[1][8][4];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
                    warp_gemm_f16x3<1, 8, true, false, true>(s.h0h + 16 * w, s.h0l + 16 * w, kHS, s.dzh, s.dzl, kHW, RS, acc);
                    const float inv = alpha * fumi_exp2i(-e_h0) * fumi_exp2i(-e_dzs);
                    const float winv = fumi_exp2i(-e_w1), wsc = fumi_exp2i(e_w1n);
                    float mxv = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            const int off = (16 * w + g + 8 * hq) * kHW + 8 * j + 2 * t;
                            float o0, o1;
                            ld_planes2(s.w1h, s.w1l, off, winv, o0, o1);
                            o0 = fmaf(acc[0][j][2 * hq], inv, o0);
                            o1 = fmaf(acc[0][j][2 * hq + 1], inv, o1);
                            mxv = fmaxf(mxv, fmaxf(fabsf(o0), fabsf(o1)));
                            st_planes2(s.w1h, s.w1l, off, o0, o1, wsc);
                        }
                    block_max_push(s.mx + 16 * (BX_W1 + par_w1), mxv);
                }
                pc.mark(6);
                for (int r0 = 0; r0 < n; r0 += 16) {
                    const int tr = min(16, n - r0), ti = r0 >> 4;
                    if (r0 > 0) {                          // second tile: its H0 planes replace the first tile's
                        t_reload(st);
                        fumi_cp_async_wait();
                        __syncthreads();
                    }
                    // ---- r_dH0 = gate0 * sc * (a_S + g_b0) on this warp's columns -> planes in rows 16..31 of (h0 | tt)
                    uint32_t gate0 = 0u;
                    {
                        float tv[2][2][2];
                        float mxv = 0.f;
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int hq = 0; hq < 2; ++hq) {
                                const int r = g + 8 * hq, h = hc + 8 * j;
                                const uint32_t gt = plane_gates2(s.h0h, s.h0l, r * kHS + h);    // (pad rows hold zeros: gates 0)
                                gate0 |= gt << (2 * (2 * j + hq));
                                const float a0 = ti == 0 ? aS[0][j][2 * hq] : aS1[j][2 * hq];
                                const float a1 = ti == 0 ? aS[0][j][2 * hq + 1] : aS1[j][2 * hq + 1];
                                tv[j][hq][0] = (gt & 1u) ? (a0 + gb0[j][0]) * sc : 0.f;
                                tv[j][hq][1] = (gt & 2u) ? (a1 + gb0[j][1]) * sc : 0.f;
                                mxv = fmaxf(mxv, fmaxf(fabsf(tv[j][hq][0]), fabsf(tv[j][hq][1])));
                            }
                        block_max_push(s.mx + 16 * (BX_TT + par_tt), mxv);
                        if (tt_first) {                   // no history yet: exchange the max first
                            __syncthreads();
                            const float mm = slot_max(s.mx + 16 * (BX_TT + par_tt));
                            if (mm > 0.f) e_ttn = fumi_plane_exp_t(__float_as_uint(mm), kTarget);
                            tt_first = false;
                        }
                        const float tsc = fumi_exp2i(e_ttn);
                        __syncwarp();                     // this warp's reads of rows 16..31 (undo GEMM) are done
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int hq = 0; hq < 2; ++hq)
                                st_planes2(s.h0h, s.h0l, (16 + g + 8 * hq) * kHS + hc + 8 * j, tv[j][hq][0], tv[j][hq][1], tsc);
                    }
                    __syncthreads();                      // T1: r_dH0 planes, W1_s planes (undo), a_W1 planes
                    FUMI_ADOPT(e_tt, e_ttn, par_tt, BX_TT);
                    if (r0 == 0) FUMI_ADOPT(e_w1, e_w1n, par_w1, BX_W1);
                    pc.mark(7);
                    // ---- r_dZ1 partial sums over this warp's half of the 256 hidden units: r_dH0 W1_s^T - alpha H0 a_W1^T
                    {
                        const int nt = w & 7, kh = w >> 3;
                        float a1[1][1][4] = {{{0.f, 0.f, 0.f, 0.f}}}, a2[1][1][4] = {{{0.f, 0.f, 0.f, 0.f}}};
                        warp_gemm_f16x3<1, 1, false, false>(s.h0h + 16 * kHS + 128 * kh, s.h0l + 16 * kHS + 128 * kh, kHS,
                                                            s.w1h + 128 * kh * kHW + 8 * nt, s.w1l + 128 * kh * kHW + 8 * nt, kHW, 128, a1);
                        warp_gemm_f16x3<1, 1, false, false>(s.h0h + 128 * kh, s.h0l + 128 * kh, kHS, s.awh + 128 * kh * kHW + 8 * nt,
                                                            s.awl + 128 * kh * kHW + 8 * nt, kHW, 128, a2);
                        const float i1 = fumi_exp2i(-e_tt) * fumi_exp2i(-e_w1), i2 = -alpha * fumi_exp2i(-e_h0) * fumi_exp2i(-e_aw);
                        float* dst = s.rzp + (16 * kh + g) * kS1 + 8 * nt + 2 * t;
                        *reinterpret_cast<float2*>(dst) = make_float2(fmaf(a2[0][0][0], i2, a1[0][0][0] * i1), fmaf(a2[0][0][1], i2, a1[0][0][1] * i1));
                        *reinterpret_cast<float2*>(dst + 8 * kS1) = make_float2(fmaf(a2[0][0][2], i2, a1[0][0][2] * i1), fmaf(a2[0][0][3], i2, a1[0][0][3] * i1));
                    }
                    // ---- warp-local: r_W1[own rows] += r_dH0^T dZ1 ;  r_H0 = -alpha dZ1 a_W1^T[own rows]
                    float rh0[2][4];
                    {
                        float acc[1][8][4];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
                        warp_gemm_f16x3<1, 8, true, false, true>(s.h0h + 16 * kHS + 16 * w, s.h0l + 16 * kHS + 16 * w, kHS, s.dzh + r0 * kHW,
                                                            s.dzl + r0 * kHW, kHW, 16, acc);
                        const float inv = fumi_exp2i(-e_tt) * fumi_exp2i(-e_dzs);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) wacc[j][q] = fmaf(acc[0][j][q], inv, wacc[j][q]);
                        float a2[1][2][4];
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) a2[0][j][q] = 0.f;
                        warp_gemm_f16x3<1, 2, false, true, true>(s.dzh + r0 * kHW, s.dzl + r0 * kHW, kHW, s.awh + 16 * w * kHW, s.awl + 16 * w * kHW, kHW, kH1, a2);
                        const float i2 = -alpha * fumi_exp2i(-e_dzs) * fumi_exp2i(-e_aw);
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) rh0[j][q] = a2[0][j][q] * i2;
                    }
                    __syncthreads();                      // T2: r_dZ1 partial sums
                    pc.mark(8);
                    // ---- one row per warp: r_dH1 -> r_dL -> r_L -> r_H1 -> r_Z1 (couples only the units / classes of a row)
                    {
                        const int i = w;
                        const bool arow = i < tr;
                        const float h10 = s.h1t[(r0 + i) * kS1 + lane], h11 = s.h1t[(r0 + i) * kS1 + lane + 32];
                        float rd0 = s.rzp[i * kS1 + lane] + s.rzp[(16 + i) * kS1 + lane] - alpha * s.ab1[lane];
                        float rd1 = s.rzp[i * kS1 + lane + 32] + s.rzp[(16 + i) * kS1 + lane + 32] - alpha * s.ab1[lane + 32];
                        rd0 = (arow & (h10 > 0.f)) ? rd0 * sc : 0.f;           // r_dH1  (&, not &&: selects, no divergent branches)
                        rd1 = (arow & (h11 > 0.f)) ? rd1 * sc : 0.f;
                        // r_dL of every class in every lane: N partial sums, one interleaved butterfly
                        float rdl[kNC];
#pragma unroll
                        for (int cc = 0; cc < kNC; ++cc) {
                            const float* wh = &s.hp[cc * kHD];
                            const float* ah = &s.ahp[cc * kHD];
                            rdl[cc] = fmaf(rd0, wh[lane], rd1 * wh[lane + 32]) - alpha * fmaf(h10, ah[lane], h11 * ah[lane + 32]);
                        }
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
                            for (int cc = 0; cc < kNC; ++cc) rdl[cc] += __shfl_xor_sync(0xffffffffu, rdl[cc], off);
                        const int y = s.sys[r0 + i];
                        const float fn = float(n), invn = 1.f / fn;
                        float dlr[kNC];                    // dL of the row (zero for inert classes and pad rows)
                        float dot = 0.f;                   // <P, r_dL> with P = softmax = n dL + onehot
#pragma unroll
                        for (int cc = 0; cc < kNC; ++cc) {
                            dlr[cc] = (arow && cc < N) ? s.lt[(r0 + i) * kLB + cc] : 0.f;
                            rdl[cc] -= alpha * s.ahp[cc * kHD + kH1];
                            const float pc_ = arow ? fmaf(dlr[cc], fn, cc == y ? 1.f : 0.f) : 0.f;
                            dot = fmaf(pc_, rdl[cc], dot);
                        }
                        float ra0 = 0.f, ra1 = 0.f, rl = 0.f;
#pragma unroll
                        for (int cc = 0; cc < kNC; ++cc) {
                            const float pc_ = arow ? fmaf(dlr[cc], fn, cc == y ? 1.f : 0.f) : 0.f;
                            const float rlc = pc_ * (rdl[cc] - dot) * invn;                 // r_L of class cc
                            if (lane == cc) rl = rlc;
                            ra0 = fmaf(dlr[cc], -alpha * s.ahp[cc * kHD + lane], fmaf(rlc, s.hp[cc * kHD + lane], ra0));
                            ra1 = fmaf(dlr[cc], -alpha * s.ahp[cc * kHD + lane + 32], fmaf(rlc, s.hp[cc * kHD + lane + 32], ra1));
                        }
                        if (lane < N) s.rlt[i * kLB + lane] = arow ? rl : 0.f;
                        s.rzp[i * kS1 + lane] = rd0;                           // r_dH1 for the head sums
                        s.rzp[i * kS1 + lane + 32] = rd1;
                        const float z0 = (arow & (h10 > 0.f)) ? ra0 * sc : 0.f, z1 = (arow & (h11 > 0.f)) ? ra1 * sc : 0.f;
                        const float mxv = warp_max(fmaxf(fabsf(z0), fabsf(z1)));
                        if (lane == 0) s.mx[16 * (BX_RZ + par_rz) + w] = mxv;
                        if (rz_first) {
                            __syncthreads();
                            const float mm = slot_max(s.mx + 16 * (BX_RZ + par_rz));
                            if (mm > 0.f) e_rzn = fumi_plane_exp_t(__float_as_uint(mm), kTarget);
                            rz_first = false;
                        }
                        const float rsc = fumi_exp2i(e_rzn);
                        st_plane1(s.rzh, s.rzl, i * kHW + lane, z0, rsc);
                        st_plane1(s.rzh, s.rzl, i * kHW + lane + 32, z1, rsc);
                    }
                    __syncthreads();                      // T3: r_Z1 planes, r_L, r_dH1
                    FUMI_ADOPT(e_rz, e_rzn, par_rz, BX_RZ);
                    pc.mark(9);
                    // ---- cross-row sums (first threads): r_head += dL^T r_dH1 + r_L^T [H1 | 1] ;  r_b1 += sum r_Z1
                    for (int idx = tid; idx < N * kHD; idx += NT_) {
                        const int cc = idx / kHD, o = idx - cc * kHD;
                        float a0 = 0.f, a1 = 0.f;          // two independent chains over the rows
                        // one loop for weight and bias columns (the bias column reads r_dH1 = 0, H1 = 1): a warp that holds
                        // both kinds would otherwise run two loops one after the other
                        const bool wcol = o < kH1;
                        const int oc = wcol ? o : 0;
                        int i = 0;
                        for (; i + 2 <= tr; i += 2) {
                            const float z0 = wcol ? s.rzp[i * kS1 + oc] : 0.f, z1 = wcol ? s.rzp[(i + 1) * kS1 + oc] : 0.f;
                            const float u0 = wcol ? s.h1t[(r0 + i) * kS1 + oc] : 1.f, u1 = wcol ? s.h1t[(r0 + i + 1) * kS1 + oc] : 1.f;
                            a0 = fmaf(s.lt[(r0 + i) * kLB + cc], z0, fmaf(s.rlt[i * kLB + cc], u0, a0));
                            a1 = fmaf(s.lt[(r0 + i + 1) * kLB + cc], z1, fmaf(s.rlt[(i + 1) * kLB + cc], u1, a1));
                        }
                        if (i < tr) {
                            const float z0 = wcol ? s.rzp[i * kS1 + oc] : 0.f, u0 = wcol ? s.h1t[(r0 + i) * kS1 + oc] : 1.f;
                            a0 = fmaf(s.lt[(r0 + i) * kLB + cc], z0, fmaf(s.rlt[i * kLB + cc], u0, a0));
                        }
                        s.rhp[idx] += a0 + a1;
                    }
                    if (tid >= NT_ - kH1) {
                        const int o = tid - (NT_ - kH1);
                        const float inv = fumi_exp2i(-e_rz);
                        float a = 0.f;
                        for (int i = 0; i < tr; ++i) a += plane_value(s.rzh, s.rzl, i * kHW + o, inv);
                        s.rb1[o] += a;
                    }
                    // ---- warp-local: r_H0 += r_Z1 W1_s ; r_W1 += H0^T r_Z1 ; bar_Z0 ; d_proj ; a_S -= alpha G bar_Z0
                    {
                        float a2[1][2][4];
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) a2[0][j][q] = 0.f;
                        warp_gemm_f16x3<1, 2, false, true, true>(s.rzh, s.rzl, kHW, s.w1h + 16 * w * kHW, s.w1l + 16 * w * kHW, kHW, kH1, a2);
                        const float i2 = fumi_exp2i(-e_rz) * fumi_exp2i(-e_w1);
                        float acc[1][8][4];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[0][j][q] = 0.f;
                        warp_gemm_f16x3<1, 8, true, false, true>(s.h0h + 16 * w, s.h0l + 16 * w, kHS, s.rzh, s.rzl, kHW, 16, acc);
                        const float inv = fumi_exp2i(-e_h0) * fumi_exp2i(-e_rz);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int q = 0; q < 4; ++q) wacc[j][q] = fmaf(acc[0][j][q], inv, wacc[j][q]);
                        float colsum[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
                        float mxv = 0.f;
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int hq = 0; hq < 2; ++hq) {
                                const int r = g + 8 * hq, h = hc + 8 * j;
                                const uint32_t gt = gate0 >> (2 * (2 * j + hq));
                                const float d0 = (gt & 1u) ? fmaf(a2[0][j][2 * hq], i2, rh0[j][2 * hq]) * sc : 0.f;
                                const float d1 = (gt & 2u) ? fmaf(a2[0][j][2 * hq + 1], i2, rh0[j][2 * hq + 1]) * sc : 0.f;
                                rh0[j][2 * hq] = d0;
                                rh0[j][2 * hq + 1] = d1;
                                colsum[j][0] += d0;
                                colsum[j][1] += d1;
                                mxv = fmaxf(mxv, fmaxf(fabsf(d0), fabsf(d1)));
                                atomic_add2_if(r < tr, &P.d_proj[s.srows[min(r0 + r, 31)] * kH0 + h], d0, d1);
                            }
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                float v = colsum[j][q];
                                v += __shfl_xor_sync(0xffffffffu, v, 4);
                                v += __shfl_xor_sync(0xffffffffu, v, 8);
                                v += __shfl_xor_sync(0xffffffffu, v, 16);
                                rb0r[j][q] += v;
                            }
                        mxv = warp_max(mxv);
                        const int e_bz = fumi_plane_exp(__float_as_uint(mxv));
                        const float bsc = fumi_exp2i(e_bz);
                        __syncwarp();                     // this warp's reads of its r_dH0 columns are done
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int hq = 0; hq < 2; ++hq)
                                st_planes2(s.h0h, s.h0l, (16 + g + 8 * hq) * kHS + hc + 8 * j, rh0[j][2 * hq], rh0[j][2 * hq + 1], bsc);
                        __syncwarp();
                        float as[MT][2][4];
#pragma unroll
                        for (int i = 0; i < MT; ++i)
#pragma unroll
                            for (int j = 0; j < 2; ++j)
#pragma unroll
                                for (int q = 0; q < 4; ++q) as[i][j][q] = 0.f;
                        warp_gemm_f16x3<MT, 2, false, false, true>(s.gsh + r0, s.gsl + r0, kHG, s.h0h + 16 * kHS + 16 * w, s.h0l + 16 * kHS + 16 * w,
                                                             kHS, 16, as);
                        const float ainv = -alpha * fumi_exp2i(-e_gs) * fumi_exp2i(-e_bz);
#pragma unroll
                        for (int i = 0; i < MT; ++i)
#pragma unroll
                            for (int j = 0; j < 2; ++j)
#pragma unroll
                                for (int q = 0; q < 4; ++q) aS[i][j][q] = fmaf(as[i][j][q], ainv, aS[i][j][q]);
                    }
                    __syncthreads();                      // T4: tile buffers free
                    pc.mark(11);
                }
                // ---- end of the reversed step: next step's records start travelling, then the adjoints are folded
                if (st > 0) s_issue(st - 1);
                {
                    const float asc = fumi_exp2i(e_awn);
                    float mxv = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            const int off = (16 * w + g + 8 * hq) * kHW + 8 * j + 2 * t;
                            mxv = fmaxf(mxv, fmaxf(fabsf(wacc[j][2 * hq]), fabsf(wacc[j][2 * hq + 1])));
                            st_planes2(s.awh, s.awl, off, wacc[j][2 * hq], wacc[j][2 * hq + 1], asc);
                        }
                    block_max_push(s.mx + 16 * (BX_AW + par_aw), mxv);
                }
                for (int idx = tid; idx < N * kHD; idx += NT_) s.ahp[idx] += s.rhp[idx];
                if (tid < kH1) s.ab1[tid] += s.rb1[tid];
#pragma unroll
                for (int j = 0; j < 2; ++j) { ab0r[j][0] += rb0r[j][0]; ab0r[j][1] += rb0r[j][1]; }
                pc.mark(12);
            }
            __syncthreads();
            FUMI_ADOPT(e_aw, e_awn, par_aw, BX_AW);
        } else {
            __syncthreads();
        }

        // ------------------------------------------------------------------ task epilogue
        for (int idx = tid; idx < N * kHD; idx += NT_)
            P.d_head[b * N * kHD + idx] = bad ? __uint_as_float(0x7FC00000u) : s.ahp[idx];
        {
            float* pw = P.d_w1_parts + int64_t(blockIdx.x) * kH0 * kH1;
            const float ainv = fumi_exp2i(-e_aw);
#pragma unroll 1
            for (int base = tid; base < kH0 * kH1; base += 8 * NT_) {           // 8 partial-sum loads in flight per thread
                float gsum[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) gsum[q] = pw[base + q * NT_];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int idx = base + q * NT_, o = idx / kH0, k = idx - o * kH0;   // [H1][H0] like linear1.weight
                    pw[idx] = gsum[q] + plane_value(s.awh, s.awl, k * kHW + o, ainv);
                }
            }
        }
        if (g == 0) {
            float* pb = P.d_b0_parts + int64_t(blockIdx.x) * kH0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                pb[hc + 8 * j] += ab0r[j][0];
                pb[hc + 8 * j + 1] += ab0r[j][1];
            }
        }
        if (tid < kH1) P.d_b1_parts[int64_t(blockIdx.x) * kH1 + tid] += s.ab1[tid];
        __syncthreads();
        pc.mark(14);
    }
#undef FUMI_ADOPT
}

size_t smem_w_bytes(int N) { SmemW t; return carve_w(nullptr, t, N); }

}  // namespace

int launch_episode_bwd_f16(const EpiParams& P, int grid, void* stream) {
    const int nc = P.cfg.num_ways <= 5 ? 5 : (P.cfg.num_ways <= 8 ? 8 : 11);
    const size_t smem = smem_w_bytes(nc);
    if (smem > 227 * 1024) {
        fumi_set_error("episode backward: shared-memory budget exceeded for this num_ways");
        return FUMI_ERR_UNSUPPORTED;
    }
#ifndef FUMI_EMU
#define FUMI_SMEM_ATTR(kern)                                                                                       \
    do {                                                                                                            \
        cudaError_t e__ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));       \
        if (e__ != cudaSuccess) return fumi_cuda_fail(e__, "cudaFuncSetAttribute(episode_bwd_v2_kernel)");          \
    } while (0)
#else
#define FUMI_SMEM_ATTR(kern) ((void)0)
#endif
#define FUMI_BWD_LAUNCH(MT_, NC_)                                                   \
    do {                                                                            \
        FUMI_SMEM_ATTR((episode_bwd_v2_kernel<MT_, NC_>));                          \
        FUMI_LAUNCH((episode_bwd_v2_kernel<MT_, NC_>), grid, kThreads16, smem, stream, P); \
    } while (0)
    if (P.cfg.num_support <= 16) {
        if (nc == 5) FUMI_BWD_LAUNCH(1, 5); else if (nc == 8) FUMI_BWD_LAUNCH(1, 8); else FUMI_BWD_LAUNCH(1, 11);
    } else {
        if (nc == 5) FUMI_BWD_LAUNCH(2, 5); else if (nc == 8) FUMI_BWD_LAUNCH(2, 8); else FUMI_BWD_LAUNCH(2, 11);
    }
#undef FUMI_BWD_LAUNCH
#undef FUMI_SMEM_ATTR
    FUMI_CHECK_LAUNCH("episode_bwd_v2_kernel");
    return FUMI_OK;
}

size_t episode_bwd_f16_smem_bytes(int N) { return smem_w_bytes(N); }

}  // namespace fumi_epi
