// Dense layers of the path, fp32 FMA version (precision 0).
//   fumi_linear_fwd   y = act(x w^T + b)      first image layer over feature rows, hypernetwork, AM3 encoders
//   fumi_linear_wgrad dw (+)= dy^T x          their weight gradients (dW0 = d_proj^T X is the large one)
//   fumi_linear_dgrad dx = (dy w) * relu-gate hypernetwork hidden layer
// Reference: F.linear / nn.Linear forward+backward in fumi/models/fumi.py:70-107,109-113,214-218 and
// am3.py:105-126 executed by torch (MKL / cuBLAS).  Classic shared-memory tiled SGEMM (128x128x16,
// 8x8 register tile); the tcgen05 3xTF32 path (precision 1, dense_tc.cu) replaces it for the large
// contractions.
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, PAD = 4;

__device__ inline float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return tanhf(v);
    if (act == 3) return 1.f / (1.f + expf(-v));
    return v;
}

// C[M,N] = act(A[M,K] B[N,K]^T + bias): both operands K-contiguous.
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                         const float* __restrict__ bias, float* __restrict__ C,
                                                         int64_t M, int64_t N, int64_t K, int act) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = int64_t(blockIdx.y) * BM, n0 = int64_t(blockIdx.x) * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const bool vec = (K & 3) == 0;
    for (int64_t k0 = 0; k0 < K; k0 += BK) {
        // 128 rows x 16 k = 512 float4 per operand, 2 per thread
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int idx = tid + r * 256;
            const int row = idx >> 2, kq = (idx & 3) * 4;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            const int64_t gm = m0 + row, gn = n0 + row, gk = k0 + kq;
            if (gm < M) {
                if (vec && gk + 3 < K) a = *reinterpret_cast<const float4*>(&A[gm * K + gk]);
                else {
                    if (gk < K) a.x = A[gm * K + gk];
                    if (gk + 1 < K) a.y = A[gm * K + gk + 1];
                    if (gk + 2 < K) a.z = A[gm * K + gk + 2];
                    if (gk + 3 < K) a.w = A[gm * K + gk + 3];
                }
            }
            if (gn < N) {
                if (vec && gk + 3 < K) b = *reinterpret_cast<const float4*>(&Bm[gn * K + gk]);
                else {
                    if (gk < K) b.x = Bm[gn * K + gk];
                    if (gk + 1 < K) b.y = Bm[gn * K + gk + 1];
                    if (gk + 2 < K) b.z = Bm[gn * K + gk + 2];
                    if (gk + 3 < K) b.w = Bm[gn * K + gk + 3];
                }
            }
            As[kq][row] = a.x; As[kq + 1][row] = a.y; As[kq + 2][row] = a.z; As[kq + 3][row] = a.w;
            Bs[kq][row] = b.x; Bs[kq + 1][row] = b.y; Bs[kq + 2][row] = b.z; Bs[kq + 3][row] = b.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t gm = m0 + ty * TM + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t gn = n0 + tx * TN + j;
            if (gn < N) C[gm * N + gn] = apply_act(acc[i][j] + (bias ? bias[gn] : 0.f), act);
        }
    }
}

// C[N,K] += A[M,N]^T B[M,K] over the M-slab of blockIdx.z (atomic accumulate; C zeroed by the host).
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                           float* __restrict__ C, int64_t M, int64_t N, int64_t K,
                                                           int64_t slab) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t n0 = int64_t(blockIdx.y) * BM, k0 = int64_t(blockIdx.x) * BN;
    const int64_t mb = int64_t(blockIdx.z) * slab, me = min(M, mb + slab);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int64_t m0 = mb; m0 < me; m0 += BK) {
        // 16 rows x 128 cols per operand = 2048 floats, 8 per thread, coalesced along the row
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int idx = tid + r * 256;
            const int row = idx >> 7, col = idx & 127;
            const int64_t gm = m0 + row;
            As[row][col] = (gm < me && n0 + col < N) ? A[gm * N + n0 + col] : 0.f;
            Bs[row][col] = (gm < me && k0 + col < K) ? Bm[gm * K + k0 + col] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t gn = n0 + ty * TM + i;
        if (gn >= N) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t gk = k0 + tx * TN + j;
            if (gk < K) atomicAdd(&C[gn * K + gk], acc[i][j]);
        }
    }
}

// out[n] (+)= sum_m dy[m, n]   one block per 32 columns, fixed-order tree over the rows
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, float* __restrict__ out,
                                                     int64_t M, int64_t N, int accumulate) {
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t col = int64_t(blockIdx.x) * 32 + lane;
    float a = 0.f;
    if (col < N)
        for (int64_t mrow = w; mrow < M; mrow += 8) a += dy[mrow * N + col];
    part[w][lane] = a;
    __syncthreads();
    if (w == 0 && col < N) {
        float t = 0.f;
        for (int q = 0; q < 8; ++q) t += part[q][lane];
        out[col] = accumulate ? out[col] + t : t;
    }
}

__global__ void zero_kernel(float* p, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) p[i] = 0.f;
}

// dx[M,K] = (dy[M,N] w[N,K]) * (gate > 0)
__global__ void __launch_bounds__(256) linear_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                           const float* __restrict__ gate, float* __restrict__ dx,
                                                           int64_t M, int64_t N, int64_t K) {
    const int64_t mrow = blockIdx.y;
    const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float a = 0.f;
    for (int64_t nn = 0; nn < N; ++nn) a = fmaf(dy[mrow * N + nn], w[nn * K + k], a);
    if (gate && !(gate[mrow * K + k] > 0.f)) a = 0.f;
    dx[mrow * K + k] = a;
}

__global__ void tanh_bwd_kernel(const float* __restrict__ y, float* __restrict__ dy, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        dy[i] *= 1.f - y[i] * y[i];
}

}  // namespace

int fumi_linear_fwd_tc(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*);
int fumi_linear_wgrad_tc(const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*);

extern "C" int fumi_linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t M, int64_t N,
                               int64_t K, int32_t act, int32_t precision, void* stream) {
    FUMI_CHECK_ARG(M >= 0 && N > 0 && K > 0, "bad shape");
    FUMI_CHECK_ARG(act >= 0 && act <= 3, "act must be 0..3");
    if (M == 0) return FUMI_OK;
    FUMI_CHECK_ARG(x && w && y, "null pointer");
    if (precision == 1) return fumi_linear_fwd_tc(x, w, bias, y, M, N, K, act, stream);
    FUMI_CHECK_ARG(precision == 0, "precision must be 0 (fp32) or 1 (tcgen05 3xTF32)");
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
    FUMI_LAUNCH(linear_fwd_kernel, grid, 256, 0, stream, x, w, bias, y, M, N, K, act);
    FUMI_CHECK_LAUNCH("linear_fwd_kernel");
    return FUMI_OK;
}

extern "C" int fumi_linear_wgrad(const float* dy, const float* x, float* dw, float* db, int64_t M, int64_t N,
                                 int64_t K, int32_t accumulate, int32_t precision, void* stream) {
    FUMI_CHECK_ARG(M >= 0 && N > 0 && K > 0, "bad shape");
    FUMI_CHECK_ARG(dy && x && dw, "null pointer");
    if (db) {
        FUMI_LAUNCH(colsum_kernel, (unsigned)((N + 31) / 32), 256, 0, stream, dy, db, M, N, accumulate);
        FUMI_CHECK_LAUNCH("colsum_kernel");
    }
    if (precision == 1) return fumi_linear_wgrad_tc(dy, x, dw, M, N, K, accumulate, stream);
    FUMI_CHECK_ARG(precision == 0, "precision must be 0 (fp32) or 1 (tcgen05 3xTF32)");
    if (!accumulate) {
        FUMI_LAUNCH(zero_kernel, 256, 256, 0, stream, dw, N * K);
        FUMI_CHECK_LAUNCH("zero_kernel");
    }
    if (M == 0) return FUMI_OK;
    const int64_t tiles = ((N + BM - 1) / BM) * ((K + BN - 1) / BN);
    int sms = fumi_device_sm_count();
    if (sms <= 0) return sms;
    int64_t split = (4 * int64_t(sms) + tiles - 1) / tiles;
    const int64_t max_split = (M + 4 * BK - 1) / (4 * BK);
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    int64_t slab = (M + split - 1) / split;
    slab = ((slab + BK - 1) / BK) * BK;
    split = (M + slab - 1) / slab;
    dim3 grid((unsigned)((K + BN - 1) / BN), (unsigned)((N + BM - 1) / BM), (unsigned)split);
    FUMI_LAUNCH(linear_wgrad_kernel, grid, 256, 0, stream, dy, x, dw, M, N, K, slab);
    FUMI_CHECK_LAUNCH("linear_wgrad_kernel");
    return FUMI_OK;
}

extern "C" int fumi_linear_dgrad(const float* dy, const float* w, const float* gate, float* dx, int64_t M,
                                 int64_t N, int64_t K, void* stream) {
    FUMI_CHECK_ARG(M >= 0 && N > 0 && K > 0, "bad shape");
    if (M == 0) return FUMI_OK;
    FUMI_CHECK_ARG(dy && w && dx, "null pointer");
    FUMI_CHECK_ARG(M < 65536, "M too large for this kernel (used for the hypernetwork hidden layer only)");
    dim3 grid((unsigned)((K + 255) / 256), (unsigned)M);
    FUMI_LAUNCH(linear_dgrad_kernel, grid, 256, 0, stream, dy, w, gate, dx, M, N, K);
    FUMI_CHECK_LAUNCH("linear_dgrad_kernel");
    return FUMI_OK;
}

extern "C" int fumi_tanh_bwd(const float* y, float* dy, int64_t n, void* stream) {
    FUMI_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return FUMI_OK;
    FUMI_CHECK_ARG(y && dy, "null pointer");
    FUMI_LAUNCH(tanh_bwd_kernel, 256, 256, 0, stream, y, dy, n);
    FUMI_CHECK_LAUNCH("tanh_bwd_kernel");
    return FUMI_OK;
}
