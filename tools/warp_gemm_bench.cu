// Microbenchmark + accuracy check (B200) of two ways to run the per-task small GEMMs of the episode kernels on
// the warp-level tensor path, 16 warps per CTA, one CTA per SM, operands resident in shared memory:
//   (a) today:   fp32 tiles, every consumer warp splits its fragments into tf32 hi/lo (LOP3 + FADD per element
//                per use) and issues 3 x mma.m16n8k8.tf32 per k8 step                    (csrc/warp_mma.cuh)
//   (b) planned: operands kept as PRE-SPLIT fp16 hi/lo planes (same bytes as one fp32 tile), fragments fetched
//                with ldmatrix(.trans), 3 x mma.m16n8k16.f16 per k16 step, no arithmetic in the loop.
// Shape of the test: D[32 x 64] = A[32 x 256] . Wt[256 x 64]  (H1 = H0 . W1^T of one task), warp w owns the
// (m tile w / 8, n tile w % 8) block like episode_fwd_mma16_kernel, or larger register tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -o /tmp/wgb tools/warp_gemm_bench.cu && /tmp/wgb
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../fumi_b200/csrc/warp_mma.cuh"

constexpr int M = 32, N = 64, K = 256;
constexpr int kSA = K + 4, kSB = N + 4;            // fp32 strides of (a)
constexpr int kHA = K + 8, kHB = N + 8;            // half strides of (b): rows 16 bytes apart mod 128 -> ldmatrix conflict-free

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const __half* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const __half* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// D[16 MT x 16 NP] += A[rows m0.., k] . B[k][n0..]: A planes row-major [m][k] (k contiguous), B planes [k][n]
// (n contiguous, fetched with ldmatrix.trans); NP = pairs of n tiles.  Accumulator restarted every 64 k.
template <int MT, int NP>
__device__ __forceinline__ void warp_gemm_f16x3(const __half* Ahi, const __half* Alo, int lda, const __half* Bhi,
                                                const __half* Blo, int ldb, int Kdim, float (&acc)[MT][2 * NP][4]) {
    const int lane = threadIdx.x & 31;
    const int arow = (lane & 7) + 8 * ((lane >> 3) & 1), acol = 8 * (lane >> 4);        // A: (M0 M1 | M2 M3) = (rows, rows+8 | k+8)
    const int brow = (lane & 7) + 8 * ((lane >> 3) & 1), bcol = 8 * (lane >> 4);        // B^T: (k, k+8 | n+8)
    for (int k0 = 0; k0 < Kdim; k0 += 64) {
        float part[MT][2 * NP][4];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < 2 * NP; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) part[i][j][q] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 64; kk += 16) {
            const int k = k0 + kk;
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                ldsm4(ah[i], Ahi + (16 * i + arow) * lda + k + acol);
                ldsm4(al[i], Alo + (16 * i + arow) * lda + k + acol);
            }
#pragma unroll
            for (int jp = 0; jp < NP; ++jp) {
                uint32_t bh[4], bl[4];                             // {b0, b1} of n tile 2 jp, {b0, b1} of n tile 2 jp + 1
                ldsm4t(bh, Bhi + (k + brow) * ldb + 16 * jp + bcol);
                ldsm4t(bl, Blo + (k + brow) * ldb + 16 * jp + bcol);
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    mma_f16(part[i][2 * jp], al[i], bh[0], bh[1]);
                    mma_f16(part[i][2 * jp], ah[i], bl[0], bl[1]);
                    mma_f16(part[i][2 * jp], ah[i], bh[0], bh[1]);
                    mma_f16(part[i][2 * jp + 1], al[i], bh[2], bh[3]);
                    mma_f16(part[i][2 * jp + 1], ah[i], bl[2], bl[3]);
                    mma_f16(part[i][2 * jp + 1], ah[i], bh[2], bh[3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < 2 * NP; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][j][q] += part[i][j][q];
    }
}

// ---- (a) fp32 tiles, split per use --------------------------------------------------------------------------
template <int MT, int NT>
__global__ void __launch_bounds__(512, 1) bench_tf32(const float* gA, const float* gB, float* out, int iters) {
    extern __shared__ float sm[];
    float* A = sm;                    // [M][kSA]
    float* B = sm + M * kSA;          // [K][kSB]
    for (int i = threadIdx.x; i < M * K; i += 512) A[(i / K) * kSA + i % K] = gA[i];
    for (int i = threadIdx.x; i < K * N; i += 512) B[(i / N) * kSB + i % N] = gB[i];
    __syncthreads();
    const int w = threadIdx.x >> 5;
    constexpr int WM = M / (16 * MT), WN = N / (8 * NT);          // warps that tile the output once
    const int tile = w % (WM * WN), mt = tile / WN, nt = tile % WN;
    float acc[MT][NT][4];
    for (int i = 0; i < MT; ++i) for (int j = 0; j < NT; ++j) for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    for (int it = 0; it < iters; ++it)
        warp_gemm_3xtf32<MT, NT, false, false>(A + 16 * MT * mt * kSA, kSA, B + 8 * NT * nt, kSB, K, 1.f, acc);
    if (blockIdx.x == 0 && w < WM * WN)
        warp_tile_foreach<MT, NT>(acc, [&](int m, int n, float& v) { out[(16 * MT * mt + m) * N + 8 * NT * nt + n] = v / float(iters); });
}

// ---- (b) pre-split fp16 planes ------------------------------------------------------------------------------
template <int MT, int NP>
__global__ void __launch_bounds__(512, 1) bench_f16(const float* gA, const float* gB, float sa, float sb, float* out, int iters) {
    extern __shared__ __half smh[];
    __half* Ahi = smh;                       // [M][kHA]
    __half* Alo = Ahi + M * kHA;
    __half* Bhi = Alo + M * kHA;             // [K][kHB]
    __half* Blo = Bhi + K * kHB;
    for (int i = threadIdx.x; i < M * K; i += 512) {
        const float v = gA[i] * sa;
        const __half h = __float2half_rn(v);
        Ahi[(i / K) * kHA + i % K] = h;
        Alo[(i / K) * kHA + i % K] = __float2half_rn(v - __half2float(h));
    }
    for (int i = threadIdx.x; i < K * N; i += 512) {
        const float v = gB[i] * sb;
        const __half h = __float2half_rn(v);
        Bhi[(i / N) * kHB + i % N] = h;
        Blo[(i / N) * kHB + i % N] = __float2half_rn(v - __half2float(h));
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    constexpr int WM = M / (16 * MT), WN = N / (16 * NP);
    const int tile = w % (WM * WN), mt = tile / WN, nt = tile % WN;
    float acc[MT][2 * NP][4];
    for (int i = 0; i < MT; ++i) for (int j = 0; j < 2 * NP; ++j) for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    for (int it = 0; it < iters; ++it)
        warp_gemm_f16x3<MT, NP>(Ahi + 16 * MT * mt * kHA, Alo + 16 * MT * mt * kHA, kHA, Bhi + 16 * NP * nt, Blo + 16 * NP * nt, kHB,
                                K, acc);
    const float inv = 1.f / (sa * sb * float(iters));
    if (blockIdx.x == 0 && w < WM * WN)
        warp_tile_foreach<MT, 2 * NP>(acc, [&](int m, int n, float& v) { out[(16 * MT * mt + m) * N + 16 * NP * nt + n] = v * inv; });
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    std::vector<float> hA(M * K), hB(K * N);
    srand(3);
    auto rnd = [] { return float(rand()) / RAND_MAX * 2.f - 1.f; };
    for (auto& v : hA) v = fmaxf(rnd() * 3.f, 0.f);                 // post-ReLU activations
    for (auto& v : hB) v = rnd() * 0.06f;                           // weights ~ U(-1/16, 1/16)
    std::vector<double> ref(M * N, 0.0);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += double(hA[m * K + k]) * hB[k * N + n]; ref[m * N + n] = s; }
    {   // what plain fp32 arithmetic gives on the same data
        double num = 0, den = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            float s32 = 0.f;
            for (int k = 0; k < K; ++k) s32 = fmaf(hA[m * K + k], hB[k * N + n], s32);
            num = fmax(num, fabs(s32 - ref[m * N + n])); den = fmax(den, fabs(ref[m * N + n]));
        }
        printf("fp32 FMA chain on the host: relerr %.2e\n", num / den);
    }
    float *dA, *dB, *dO;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dO, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 2000;
    auto relerr = [&]() {                                  // of the last launch (made with iters = 1)
        std::vector<float> o(M * N);
        cudaDeviceSynchronize();
        cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0;
        for (int i = 0; i < M * N; ++i) { num = fmax(num, fabs(o[i] - ref[i])); den = fmax(den, fabs(ref[i])); }
        return num / den;
    };
    double err = 0;
    auto check = [&](const char* name, float ms, int warps_per_gemm) {
        const double num = err, den = 1;
        const double gemms = double(iters) * (16.0 / warps_per_gemm);        // GEMMs per CTA (all 16 warps busy)
        const double cyc = ms * 1e-3 * clk * 1e3;
        printf("%-34s %8.3f ms  %7.0f cycles per 32x64x256 GEMM per SM  %6.0f fp32-equivalent MAC/clk/SM  relerr %.2e\n", name, ms,
               cyc / gemms, double(M) * N * K * gemms / cyc, num / den);
    };
    const size_t smA = (M * kSA + K * kSB) * 4, smB = (2 * M * kHA + 2 * K * kHB) * 2;
#define RUN_TF32(MT, NT)                                                                                           \
    {                                                                                                              \
        cudaFuncSetAttribute(bench_tf32<MT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smA));            \
        bench_tf32<MT, NT><<<sms, 512, smA>>>(dA, dB, dO, 1);                                                       \
        err = relerr();                                                                                            \
        float ms = time_ms([&] { bench_tf32<MT, NT><<<sms, 512, smA>>>(dA, dB, dO, iters); });                      \
        check("tf32 split-per-use MT=" #MT " NT=" #NT, ms, (M / (16 * MT)) * (N / (8 * NT)));                       \
    }
#define RUN_F16(MT, NP)                                                                                            \
    {                                                                                                              \
        cudaFuncSetAttribute(bench_f16<MT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smB));             \
        bench_f16<MT, NP><<<sms, 512, smB>>>(dA, dB, 2048.f, 65536.f, dO, 1);                                       \
        err = relerr();                                                                                            \
        float ms = time_ms([&] { bench_f16<MT, NP><<<sms, 512, smB>>>(dA, dB, 2048.f, 65536.f, dO, iters); });     \
        check("fp16 planes + ldmatrix MT=" #MT " NP=" #NP, ms, (M / (16 * MT)) * (N / (16 * NP)));                  \
    }
    RUN_TF32(1, 1) RUN_TF32(1, 2) RUN_TF32(2, 2) RUN_TF32(2, 4) RUN_TF32(1, 8)
    RUN_F16(1, 1) RUN_F16(2, 1) RUN_F16(2, 2) RUN_F16(1, 4)
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
