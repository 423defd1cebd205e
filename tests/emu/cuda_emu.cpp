// TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

#include <string>

namespace fumi_emu {
thread_local uint3_emu t_threadIdx;
uint3_emu g_blockIdx;
dim3 g_blockDim, g_gridDim;
char* dyn_smem = nullptr;
pthread_barrier_t g_barrier;
pthread_barrier_t g_warp_barrier[64];
WarpXchg g_xchg[64];
const void* g_ptr_xchg[64][32];

void launch_impl(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem) {
    const unsigned nthreads = block.x * block.y * block.z;
    g_blockDim = block;
    g_gridDim = grid;
    void* mem = nullptr;
    if (posix_memalign(&mem, 128, smem + 128) != 0) abort();
    // poison the dynamic shared memory so reads of unwritten locations show up as NaN
    std::memset(mem, 0xFF, smem + 128);
    dyn_smem = static_cast<char*>(mem);
    pthread_barrier_init(&g_barrier, nullptr, nthreads);
    const unsigned nwarps = (nthreads + 31) / 32;
    for (unsigned w = 0; w < nwarps; ++w) {
        const unsigned cnt = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
        pthread_barrier_init(&g_warp_barrier[w], nullptr, cnt);
    }
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t]() {
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        if (t == 0) { g_blockIdx.x = bx; g_blockIdx.y = by; g_blockIdx.z = bz; }
                        pthread_barrier_wait(&g_barrier);
                        body();
                        pthread_barrier_wait(&g_barrier);
                    }
        });
    }
    for (auto& th : pool) th.join();
    pthread_barrier_destroy(&g_barrier);
    for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
    free(mem);
    dyn_smem = nullptr;
}
}  // namespace fumi_emu

// ---- stand-ins for fumi_b200/csrc/common.cu (which needs the CUDA runtime)
#include "../../include/fumi_b200.h"
static thread_local std::string g_last_error;
void fumi_set_error(const std::string& msg) { g_last_error = msg; }
int fumi_cuda_fail(cudaError_t, const char* what) { g_last_error = what; return FUMI_ERR_CUDA; }
extern "C" int fumi_abi_version(void) { return FUMI_B200_ABI_VERSION; }
extern "C" const char* fumi_last_error(void) { return g_last_error.c_str(); }
extern "C" int fumi_device_sm_count(void) {
    const char* e = getenv("FUMI_EMU_SMS");       // fewer "SMs" than tasks exercises the per-CTA task loop
    return e ? atoi(e) : 2;
}

// tcgen05 entry points cannot be emulated: they fail loudly here (GPU tests cover them).
int fumi_linear_fwd_tc(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
int fumi_linear_wgrad_tc(const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_split_tf32(const float*, float*, float*, int64_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_transpose_split_tf32(const float*, float*, float*, int64_t, int64_t, int64_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_gemm_tf32x3(const float*, const float*, const float*, const float*, const float*, float*, int64_t,
                                int64_t, int64_t, int64_t, int64_t, int64_t, int32_t, int32_t, int32_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_absmax(const float*, int64_t, float*, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_split_f16(const float*, const float*, void*, void*, int64_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_transpose_split_f16(const float*, const float*, void*, void*, int64_t, int64_t, int64_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_gemm_f16x3(const void*, const void*, const void*, const void*, const float*, const float*, const float*,
                               float*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int32_t, int32_t, int32_t, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
extern "C" int fumi_gram_f16(const void*, const void*, const float*, int64_t, int64_t, const int64_t*, const int64_t*, int64_t,
                             int32_t, int32_t, float*, void*) {
    g_last_error = "tcgen05 path is not emulated"; return FUMI_ERR_UNSUPPORTED; }
