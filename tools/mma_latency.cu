// Dependent-issue latency and throughput of mma.sync.m16n8k16 (f16 -> f32) on this GPU: one warp, chains of dependent
// MMAs on 1 / 2 / 4 / 8 independent accumulators.   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_build/mma_latency tools/mma_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int NACC>
__global__ void k(long long* out, float* sink, int iters) {
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
    uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
    float c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma(c[i], a, b0, b1);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < NACC; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int NACC> void run(int warps) {
    long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4 * 32 * warps);
    const int iters = 2000;
    k<NACC><<<1, 32 * warps>>>(d, s, iters); k<NACC><<<1, 32 * warps>>>(d, s, iters);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d  independent accumulators %d: %.1f cycles per MMA per warp  (%.1f cycles per dependent step)\n", warps, NACC,
           double(h) / (double(iters) * NACC), double(h) / iters);
    cudaFree(d); cudaFree(s);
}
int main() {
    for (int w : {1, 4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
