// TEST INFRASTRUCTURE ONLY: a minimal host emulation of the CUDA execution model, used to run the
// product's kernel *bodies* (fumi_b200/csrc/*.cu compiled with g++ -DFUMI_EMU) on CPU threads in the
// GPU-less development container.  One OS thread per CUDA thread of a block, blocks run one after
// another, __syncthreads() is a pthread barrier.  It exists to exercise indexing / synchronisation /
// math of the kernels before spending GPU time; it is never loaded by the fumi_b200 package and
// proves nothing about performance.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__
#define FUMI_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(fumi_emu::dyn_smem)

struct uint3_emu { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorName(cudaError_t) { return "emu"; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

namespace fumi_emu {
extern thread_local uint3_emu t_threadIdx;
extern uint3_emu g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern char* dyn_smem;
extern pthread_barrier_t g_barrier;

void launch_impl(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem);
template <typename F>
inline void launch(F&& f, dim3 grid, dim3 block, size_t smem) { launch_impl(std::function<void()>(f), grid, block, smem); }
}  // namespace fumi_emu

#define threadIdx (fumi_emu::t_threadIdx)
#define blockIdx (fumi_emu::g_blockIdx)
#define blockDim (fumi_emu::g_blockDim)
#define gridDim (fumi_emu::g_gridDim)

static inline void __syncthreads() { pthread_barrier_wait(&fumi_emu::g_barrier); }

static inline float atomicAdd(float* addr, float v) {
    uint32_t* p = reinterpret_cast<uint32_t*>(addr);
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED), nw;
    float f;
    do {
        std::memcpy(&f, &old, 4);
        f += v;
        std::memcpy(&nw, &f, 4);
    } while (!__atomic_compare_exchange_n(p, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    std::memcpy(&f, &old, 4);
    return f;
}

using std::max;
using std::min;
static inline long min(long a, long long b) { return a < b ? a : long(b); }
