// Microbenchmark (B200): issue rate of the warp-level mma.sync path for TF32 / BF16 and of plain FFMA,
// to decide how the per-task small GEMMs of the episode kernels should be computed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void tf32_kernel(float* out, int iters) {
    float c[NACC][4];
    for (int a = 0; a < NACC; ++a) for (int j = 0; j < 4; ++j) c[a][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x + 1, a2 = threadIdx.x + 2, a3 = threadIdx.x + 3, b0 = 5, b1 = 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int a = 0; a < NACC; ++a) for (int j = 0; j < 4; ++j) s += c[a][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void bf16_kernel(float* out, int iters) {
    float c[NACC][4];
    for (int a = 0; a < NACC; ++a) for (int j = 0; j < 4; ++j) c[a][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x + 1, a2 = threadIdx.x + 2, a3 = threadIdx.x + 3, b0 = 5, b1 = 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int a = 0; a < NACC; ++a) for (int j = 0; j < 4; ++j) s += c[a][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float* out, int iters) {
    float c[16];
    for (int a = 0; a < 16; ++a) c[a] = a;
    float x = threadIdx.x * 1e-3f, y = 1.0001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int a = 0; a < 16; ++a) c[a] = fmaf(c[a], y, x);
    }
    float s = 0.f;
    for (int a = 0; a < 16; ++a) s += c[a];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32;
        float ms = time_ms([&] { tf32_kernel<8><<<sms, threads>>>(out, iters); });
        double mac = double(iters) * 8 * 16 * 8 * 8 * warps * sms;
        printf("tf32 m16n8k8  warps/SM=%2d: %.3f ms  %.1f TFLOP/s  %.0f MAC/clk/SM (at %d MHz)\n", warps, ms, 2 * mac / ms / 1e9,
               mac / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
        ms = time_ms([&] { bf16_kernel<8><<<sms, threads>>>(out, iters); });
        mac = double(iters) * 8 * 16 * 8 * 16 * warps * sms;
        printf("bf16 m16n8k16 warps/SM=%2d: %.3f ms  %.1f TFLOP/s  %.0f MAC/clk/SM\n", warps, ms, 2 * mac / ms / 1e9,
               mac / (ms * 1e-3) / sms / (khz * 1e3));
        ms = time_ms([&] { ffma_kernel<<<sms, threads>>>(out, iters); });
        mac = double(iters) * 16 * 32 * warps * sms;
        printf("ffma          warps/SM=%2d: %.3f ms  %.1f TFLOP/s  %.0f MAC/clk/SM\n", warps, ms, 2 * mac / ms / 1e9,
               mac / (ms * 1e-3) / sms / (khz * 1e3));
    }
    return 0;
}
