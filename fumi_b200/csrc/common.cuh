// Internal helpers shared by the kernels of libfumi_b200.so (not part of the C ABI).
#pragma once
#ifdef FUMI_EMU
#include "cuda_emu.h"      // tests/emu: host emulation of the kernels (test infrastructure only)
#else
#include <cuda_runtime.h>
#endif

#include <cstdint>
#include <string>

void fumi_set_error(const std::string& msg);
int fumi_cuda_fail(cudaError_t e, const char* what);

#define FUMI_CHECK_ARG(cond, msg)                                  \
    do {                                                           \
        if (!(cond)) {                                             \
            fumi_set_error(std::string(__func__) + ": " + (msg));  \
            return FUMI_ERR_ARG;                                   \
        }                                                          \
    } while (0)

#define FUMI_CHECK_LAUNCH(what)                                    \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return fumi_cuda_fail(e__, what);  \
    } while (0)

// Compiled hidden sizes of the adapted image MLP (reference default --im_hid_dim 256 64).
constexpr int kH0 = 256;
constexpr int kH1 = 64;
constexpr int kHD = kH1 + 1;      // head row: 64 weights + 1 bias (fumi.py:76-79)
constexpr int kMaxWays = 32;
constexpr int kMaxSupport = 128;  // NK rows per task

// Counter-based dropout mask shared by forward and backward (and mirrored in fumi_b200/dropout.py for parity
// tests).  32-bit integer hashing only (the 64-bit multiplies of a splitmix were ~25 instructions per mask):
//   base  = mix(seed, task, pass, layer)                       once per tile
//   h32   = lowbias32(base + row * 0xC2B2AE35 + (col >> 1) * 0x27D4EB2F)
//   field = (col & 1) ? h32 >> 16 : h32 & 0xFFFF ;  keep iff field >= floor(p * 65536)
// One hash serves the two adjacent columns a lane owns in the MMA accumulator layout.
__host__ __device__ inline uint32_t fumi_lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
__host__ __device__ inline uint32_t fumi_mask_base(uint64_t seed, uint64_t task, uint32_t pass, uint32_t layer) {
    uint32_t x = fumi_lowbias32(uint32_t(seed) ^ 0x9E3779B9u);
    x = fumi_lowbias32(x ^ uint32_t(seed >> 32));
    x = fumi_lowbias32(x + uint32_t(task) * 0x85EBCA6Bu);
    x = fumi_lowbias32(x ^ uint32_t(task >> 32));
    x = fumi_lowbias32(x + pass * 0x9E3779B1u + layer * 0x61C88647u);
    return x;
}
__host__ __device__ inline uint32_t fumi_mask_pair(uint32_t base, uint32_t row, uint32_t col) {
    return fumi_lowbias32(base + row * 0xC2B2AE35u + (col >> 1) * 0x27D4EB2Fu);
}
