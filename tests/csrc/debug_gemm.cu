// TEST INFRASTRUCTURE (not part of libfumi_b200.so / the public header): one warp runs warp_gemm_f16x3 (csrc/warp_mma.cuh) on a small problem in
// every operand-layout variant, so the fp16-plane primitive (ldmatrix addressing, fragment order, plane scaling)
// is tested on its own -- on the GPU and under the host emulation -- before the episode kernels depend on it.
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "../../fumi_b200/csrc/common.cuh"
#include "../../fumi_b200/csrc/launch.cuh"
#include "../../fumi_b200/csrc/warp_mma.cuh"

extern "C" int fumi_debug_gemm_f16(const float* A, const float* B, int32_t variant, int32_t M, int32_t N, int32_t K,
                                   float* out, void* stream);

namespace {

constexpr int kDM = 32, kDN = 64, kDK = 64;       // largest problem; strides cols + 8 halves

// A logical [M][K], B logical [K][N] (row-major fp32 in global memory); ATRANS stores A as [k][m], BTRANS stores B as [n][k]
template <int MT, int NT, bool ATRANS, bool BTRANS>
__global__ void __launch_bounds__(32) debug_gemm_f16_kernel(const float* __restrict__ A, const float* __restrict__ B, int K,
                                                            float* __restrict__ out) {
    constexpr int M = 16 * MT, N = 8 * NT;
    __shared__ __align__(16) fumi_half ah[kDM * (kDK + 8) > kDK * (kDM + 8) ? kDM * (kDK + 8) : kDK * (kDM + 8)];
    __shared__ __align__(16) fumi_half al[kDM * (kDK + 8) > kDK * (kDM + 8) ? kDM * (kDK + 8) : kDK * (kDM + 8)];
    __shared__ __align__(16) fumi_half bh[kDK * (kDN + 8) > kDN * (kDK + 8) ? kDK * (kDN + 8) : kDN * (kDK + 8)];
    __shared__ __align__(16) fumi_half bl[kDK * (kDN + 8) > kDN * (kDK + 8) ? kDK * (kDN + 8) : kDN * (kDK + 8)];
    __shared__ unsigned mx[2];
    const int lane = threadIdx.x;
    if (lane < 2) mx[lane] = 0u;
    __syncwarp();
    float ma = 0.f, mb = 0.f;
    for (int i = lane; i < M * K; i += 32) ma = fmaxf(ma, fabsf(A[i]));
    for (int i = lane; i < K * N; i += 32) mb = fmaxf(mb, fabsf(B[i]));
    atomicMax(&mx[0], __float_as_uint(ma));
    atomicMax(&mx[1], __float_as_uint(mb));
    __syncwarp();
    const int sa = fumi_plane_exp(mx[0]), sb = fumi_plane_exp(mx[1]);
    const int lda = ATRANS ? M + 8 : K + 8, ldb = BTRANS ? K + 8 : N + 8;
    for (int i = lane; i < M * K; i += 32) {
        const int m = i / K, k = i - m * K;
        const float v = A[i] * fumi_exp2i(sa);
        const fumi_half h = fumi_f2h(v);
        const int o = ATRANS ? k * lda + m : m * lda + k;
        ah[o] = h;
        al[o] = fumi_f2h(v - fumi_h2f(h));
    }
    for (int i = lane; i < K * N; i += 32) {
        const int k = i / N, n = i - k * N;
        const float v = B[i] * fumi_exp2i(sb);
        const fumi_half h = fumi_f2h(v);
        const int o = BTRANS ? n * ldb + k : k * ldb + n;
        bh[o] = h;
        bl[o] = fumi_f2h(v - fumi_h2f(h));
    }
    __syncwarp();
    float acc[MT][NT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    warp_gemm_f16x3<MT, NT, ATRANS, BTRANS>(ah, al, lda, bh, bl, ldb, K, acc);
    const float inv = fumi_exp2i(-sa) * fumi_exp2i(-sb);
    warp_tile_foreach<MT, NT>(acc, [&](int m, int n, float& v) { out[m * N + n] = v * inv; });
}

template <int MT, int NT>
int launch_variant(int variant, const float* A, const float* B, int K, float* out, void* stream) {
    switch (variant) {
        case 0: FUMI_LAUNCH((debug_gemm_f16_kernel<MT, NT, false, false>), 1, 32, 0, stream, A, B, K, out); break;
        case 1: FUMI_LAUNCH((debug_gemm_f16_kernel<MT, NT, false, true>), 1, 32, 0, stream, A, B, K, out); break;
        case 2: FUMI_LAUNCH((debug_gemm_f16_kernel<MT, NT, true, false>), 1, 32, 0, stream, A, B, K, out); break;
        default: FUMI_LAUNCH((debug_gemm_f16_kernel<MT, NT, true, true>), 1, 32, 0, stream, A, B, K, out); break;
    }
    FUMI_CHECK_LAUNCH("debug_gemm_f16_kernel");
    return FUMI_OK;
}

}  // namespace

// out[M][N] = A[M][K] . B[K][N] through fp16 hi/lo planes; variant bit 0: B stored [n][k], bit 1: A stored [k][m].
// Supported (M, N): (16, 8), (32, 16), (16, 64), (32, 8); K in {16, 32, 48, 64}.
extern "C" int fumi_debug_gemm_f16(const float* A, const float* B, int32_t variant, int32_t M, int32_t N, int32_t K,
                                   float* out, void* stream) {
    FUMI_CHECK_ARG(A && B && out && variant >= 0 && variant < 4, "bad argument");
    FUMI_CHECK_ARG(K >= 16 && K <= kDK && K % 16 == 0, "K must be 16, 32, 48 or 64");
    if (M == 16 && N == 8) return launch_variant<1, 1>(variant, A, B, K, out, stream);
    if (M == 32 && N == 16) return launch_variant<2, 2>(variant, A, B, K, out, stream);
    if (M == 16 && N == 64) return launch_variant<1, 8>(variant, A, B, K, out, stream);
    if (M == 32 && N == 8) return launch_variant<2, 1>(variant, A, B, K, out, stream);
    fumi_set_error("fumi_debug_gemm_f16: unsupported (M, N)");
    return FUMI_ERR_ARG;
}
