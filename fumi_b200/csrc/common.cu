// Error plumbing and device queries shared by every translation unit of libfumi_b200.so.
#include <cuda_runtime.h>

#include <string>

#include "../../include/fumi_b200.h"
#include "common.cuh"

static thread_local std::string g_last_error;

void fumi_set_error(const std::string& msg) { g_last_error = msg; }

int fumi_cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return FUMI_ERR_CUDA;
}

extern "C" int fumi_abi_version(void) { return FUMI_B200_ABI_VERSION; }

extern "C" const char* fumi_last_error(void) { return g_last_error.c_str(); }

extern "C" int fumi_device_sm_count(void) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaGetDevice");
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return fumi_cuda_fail(e, "cudaDeviceGetAttribute");
    return sms;
}
