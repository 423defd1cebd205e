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

// Counter-based dropout mask shared by forward and backward (and mirrored in fumi_b200/dropout.py for
// parity tests).  One 64-bit hash covers the four columns 4g..4g+3 of a row: column c uses the 16-bit field
// (c & 3) and is kept iff field >= floor(p * 65536).
__host__ __device__ inline uint64_t fumi_mask_hash64(uint64_t seed, uint64_t task, uint32_t pass, uint32_t layer,
                                                     uint32_t row, uint32_t col_group) {
    uint64_t x = seed ^ (task * 0x9E3779B97F4A7C15ULL);
    x += (uint64_t(pass) << 40) ^ (uint64_t(layer) << 32) ^ (uint64_t(row) << 12) ^ uint64_t(col_group);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;       // splitmix64 finaliser
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
