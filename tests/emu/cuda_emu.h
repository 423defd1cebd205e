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
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
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

static inline int atomicAdd(int* addr, int v) { return __atomic_fetch_add(addr, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* addr, unsigned long long v) { return __atomic_fetch_add(addr, v, __ATOMIC_RELAXED); }

static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
#define __forceinline__ inline
template <typename T> static inline T __ldg(const T* p) { return *p; }

// ---- warp-collective emulation: tf32 rounding and mma.sync.m16n8k8 (TF32 inputs, fp32 accumulate)
namespace fumi_emu {
extern pthread_barrier_t g_warp_barrier[64];
struct WarpXchg { uint32_t a[32][4]; uint32_t b[32][2]; };
extern WarpXchg g_xchg[64];
}
static inline void __syncwarp() { pthread_barrier_wait(&fumi_emu::g_warp_barrier[fumi_emu::t_threadIdx.x >> 5]); }
static inline uint32_t fumi_tf32_hi(float x) {               // cvt.rna.tf32.f32: round to nearest, ties away
    uint32_t u = __float_as_uint(x);
    u += 0x1000u;
    return u & 0xFFFFE000u;
}
static inline float fumi_emu_tf32_trunc(uint32_t u) { return __uint_as_float(u & 0xFFFFE000u); }
static inline void fumi_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    const int w = fumi_emu::t_threadIdx.x >> 5, lane = fumi_emu::t_threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    fumi_emu::WarpXchg& x = fumi_emu::g_xchg[w];
    for (int i = 0; i < 4; ++i) x.a[lane][i] = a[i];
    for (int i = 0; i < 2; ++i) x.b[lane][i] = b[i];
    __syncwarp();
    // A(m,k): lane (m%8)*4 + k%4, reg (m>=8) + 2*(k>=4);  B(k,n): lane n*4 + k%4, reg (k>=4)
    auto A = [&](int m, int k) { return fumi_emu_tf32_trunc(x.a[(m & 7) * 4 + (k & 3)][(m >> 3) + 2 * (k >> 2)]); };
    auto B = [&](int k, int n) { return fumi_emu_tf32_trunc(x.b[n * 4 + (k & 3)][k >> 2]); };
    const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
    float out[4];
    for (int q = 0; q < 4; ++q) {
        float s = c[q];
        for (int k = 0; k < 8; ++k) s += A(rows[q], k) * B(k, cols[q]);
        out[q] = s;
    }
    __syncwarp();
    for (int q = 0; q < 4; ++q) c[q] = out[q];
}

static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    const int w = fumi_emu::t_threadIdx.x >> 5, lane = fumi_emu::t_threadIdx.x & 31;
    fumi_emu::WarpXchg& x = fumi_emu::g_xchg[w];
    x.a[lane][0] = __float_as_uint(v);
    __syncwarp();
    const float r = __uint_as_float(x.a[lane ^ lane_mask][0]);
    __syncwarp();
    return r;
}

static inline int __shfl_xor_sync(unsigned, int v, int lane_mask) {
    const int w = fumi_emu::t_threadIdx.x >> 5, lane = fumi_emu::t_threadIdx.x & 31;
    fumi_emu::WarpXchg& x = fumi_emu::g_xchg[w];
    x.a[lane][0] = uint32_t(v);
    __syncwarp();
    const int r = int(x.a[lane ^ lane_mask][0]);
    __syncwarp();
    return r;
}

static inline float __shfl_sync(unsigned, float v, int src_lane) {
    const int w = fumi_emu::t_threadIdx.x >> 5, lane = fumi_emu::t_threadIdx.x & 31;
    fumi_emu::WarpXchg& x = fumi_emu::g_xchg[w];
    x.a[lane][0] = __float_as_uint(v);
    __syncwarp();
    const float r = __uint_as_float(x.a[src_lane][0]);
    __syncwarp();
    return r;
}
static inline int __shfl_sync(unsigned, int v, int src_lane) {
    const int w = fumi_emu::t_threadIdx.x >> 5, lane = fumi_emu::t_threadIdx.x & 31;
    fumi_emu::WarpXchg& x = fumi_emu::g_xchg[w];
    x.a[lane][0] = uint32_t(v);
    __syncwarp();
    const int r = int(x.a[src_lane][0]);
    __syncwarp();
    return r;
}

// ---- fp16 planes (warp_mma.cuh): a 16-bit storage type, ldmatrix and mma.m16n8k16.f16 --------------------------
struct fumi_half { uint16_t bits; };
static inline fumi_half fumi_f2h(float f) {                 // round to nearest even, like __float2half_rn
    uint32_t x = __float_as_uint(f);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    uint16_t h;
    if (x >= 0x7F800000u) h = uint16_t(x > 0x7F800000u ? 0x7E00u : 0x7C00u);
    else if (x >= 0x477FF000u) h = 0x7C00u;                  // rounds to >= 65520: inf
    else if (x >= 0x38800000u) {                             // normal
        const uint32_t mant = x & 0x7FFFFFu, e = (x >> 23) - 112u;
        uint32_t v = (e << 10) | (mant >> 13);
        const uint32_t rem = mant & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (v & 1u))) ++v;
        h = uint16_t(v);
    } else if (x >= 0x33000000u) {                           // subnormal
        const uint32_t e = x >> 23, mant = (x & 0x7FFFFFu) | 0x800000u;
        const uint32_t shift = 126u - e;                     // 14 .. 24
        uint32_t v = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (v & 1u))) ++v;
        h = uint16_t(v);
    } else h = 0;
    return fumi_half{uint16_t(h | sign)};
}
static inline float fumi_h2f(fumi_half hh) {
    const uint32_t h = hh.bits, sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    if (e == 0) return __uint_as_float(sign) + (sign ? -1.f : 1.f) * float(m) * 5.9604644775390625e-8f;   // m * 2^-24
    if (e == 31) return __uint_as_float(sign | 0x7F800000u | (m << 13));
    return __uint_as_float(sign | ((e + 112u) << 23) | (m << 13));
}
static inline unsigned atomicMax(unsigned* addr, unsigned v) {
    unsigned old = __atomic_load_n(addr, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(addr, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
namespace fumi_emu { extern const void* g_ptr_xchg[64][32]; }
// ldmatrix.x4: lane 8 i + r supplies the address of row r of 8x8 matrix i; lane (g, t) receives, per matrix, the pair
// (row g, cols 2t, 2t+1) -- or with .trans the pair (rows 2t, 2t+1; col g).
template <bool TRANS, int NMAT>
static inline void fumi_emu_ldsm(uint32_t* r, const fumi_half* p) {
    const int w = fumi_emu::t_threadIdx.x >> 5, lane = fumi_emu::t_threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    fumi_emu::g_ptr_xchg[w][lane] = p;
    __syncwarp();
    for (int i = 0; i < NMAT; ++i) {
        uint16_t a, b;
        if (TRANS) {
            a = static_cast<const fumi_half*>(fumi_emu::g_ptr_xchg[w][8 * i + 2 * t])[g].bits;
            b = static_cast<const fumi_half*>(fumi_emu::g_ptr_xchg[w][8 * i + 2 * t + 1])[g].bits;
        } else {
            const fumi_half* row = static_cast<const fumi_half*>(fumi_emu::g_ptr_xchg[w][8 * i + g]);
            a = row[2 * t].bits;
            b = row[2 * t + 1].bits;
        }
        r[i] = uint32_t(a) | (uint32_t(b) << 16);
    }
    __syncwarp();
}
static inline void fumi_ldsm4(uint32_t (&r)[4], const fumi_half* p) { fumi_emu_ldsm<false, 4>(r, p); }
static inline void fumi_ldsm4t(uint32_t (&r)[4], const fumi_half* p) { fumi_emu_ldsm<true, 4>(r, p); }
static inline void fumi_ldsm2(uint32_t (&r)[2], const fumi_half* p) { fumi_emu_ldsm<false, 2>(r, p); }
static inline void fumi_ldsm2t(uint32_t (&r)[2], const fumi_half* p) { fumi_emu_ldsm<true, 2>(r, p); }
static inline void fumi_mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    const int w = fumi_emu::t_threadIdx.x >> 5, lane = fumi_emu::t_threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    fumi_emu::WarpXchg& x = fumi_emu::g_xchg[w];
    for (int i = 0; i < 4; ++i) x.a[lane][i] = a[i];
    x.b[lane][0] = b0;
    x.b[lane][1] = b1;
    __syncwarp();
    auto half_of = [](uint32_t reg, int k) { return fumi_h2f(fumi_half{uint16_t(k & 1 ? reg >> 16 : reg & 0xFFFFu)}); };
    // A(m,k): lane (m%8)*4 + (k%8)/2, reg (m>=8) + 2*(k>=8);  B(k,n): lane n*4 + (k%8)/2, reg (k>=8)
    auto A = [&](int m, int k) { return half_of(x.a[(m & 7) * 4 + ((k & 7) >> 1)][(m >> 3) + 2 * (k >> 3)], k); };
    auto B = [&](int k, int n) { return half_of(x.b[n * 4 + ((k & 7) >> 1)][k >> 3], k); };
    const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
    float out[4];
    for (int q = 0; q < 4; ++q) {
        float s = c[q];
        for (int k = 0; k < 16; ++k) s += A(rows[q], k) * B(k, cols[q]);
        out[q] = s;
    }
    __syncwarp();
    for (int q = 0; q < 4; ++q) c[q] = out[q];
}

using std::max;
using std::min;
static inline long min(long a, long long b) { return a < b ? a : long(b); }
