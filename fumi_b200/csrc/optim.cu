// Outer-loop step and the small deterministic reductions around the episode kernels.
//   fumi_adam_step        torch.optim.Adam / AdamW single-tensor update, one fused launch over the
//                         flat parameter buffer (reference: utils/utils.py:280-290; fumi.py:191-193)
//   fumi_reduce_parts     sum of the per-CTA partial meta-gradients, fixed order
//   fumi_scatter_add_rows per-task head gradients -> per-class hypernetwork-output gradient
//   fumi_reduce_loss_acc  outer_loss / B and accuracy / B  (fumi.py:187-188)
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "common.cuh"
#include "launch.cuh"

namespace {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   float lr, float b1, float b2, float eps, float wd,
                                                   float bc1, float bc2_sqrt, int decoupled) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        float pi = p[i], gi = g[i];
        if (decoupled) pi *= 1.f - lr * wd;          // AdamW: param.mul_(1 - lr * weight_decay)
        else gi = fmaf(wd, pi, gi);                  // Adam: grad = grad.add(param, alpha=weight_decay)
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(1.f - b2, gi * gi, v[i] * b2);       // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
        m[i] = mi;
        v[i] = vi;
    }
}

__global__ void __launch_bounds__(256) reduce_parts_kernel(const float* __restrict__ parts, int64_t P, int64_t n,
                                                           float* __restrict__ out, int accumulate) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int64_t p = 0; p < P; ++p) a += parts[p * n + i];
    out[i] = accumulate ? out[i] + a : a;
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ src,
                                                               const int64_t* __restrict__ rows, int64_t n,
                                                               int64_t width, float* __restrict__ table) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n * width) return;
    const int64_t i = idx / width, c = idx - i * width;
    atomicAdd(&table[rows[i] * width + c], src[idx]);
}

// one block; fixed-order tree: 256 strided partials, then a sequential sum by thread 0
__global__ void __launch_bounds__(256) reduce_loss_acc_kernel(const float* __restrict__ loss,
                                                              const float* __restrict__ acc, int64_t B,
                                                              float* __restrict__ out) {
    __shared__ float pl[256], pa[256];
    float l = 0.f, a = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += 256) { l += loss[i]; a += acc[i]; }
    pl[threadIdx.x] = l;
    pa[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tl = 0.f, ta = 0.f;
        for (int i = 0; i < 256; ++i) { tl += pl[i]; ta += pa[i]; }
        out[0] = tl / float(B);
        out[1] = ta / float(B);
    }
}

}  // namespace

extern "C" int fumi_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                              int32_t decoupled, void* stream) {
    FUMI_CHECK_ARG(n >= 0 && step >= 1, "n < 0 or step < 1 (step is 1-based)");
    if (n == 0) return FUMI_OK;
    FUMI_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "null pointer");
    const double bc1 = 1.0 - pow(double(beta1), double(step));
    const double bc2 = 1.0 - pow(double(beta2), double(step));
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    FUMI_LAUNCH(adam_kernel, (unsigned)blocks, 256, 0, stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                eps, weight_decay, float(bc1), float(sqrt(bc2)), int(decoupled));
    FUMI_CHECK_LAUNCH("adam_kernel");
    return FUMI_OK;
}

extern "C" int fumi_reduce_parts(const float* parts, int64_t P, int64_t n, float* out, int32_t accumulate,
                                 void* stream) {
    FUMI_CHECK_ARG(P >= 0 && n >= 0, "bad shape");
    if (n == 0) return FUMI_OK;
    FUMI_CHECK_ARG(parts && out, "null pointer");
    FUMI_LAUNCH(reduce_parts_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, parts, P, n, out, int(accumulate));
    FUMI_CHECK_LAUNCH("reduce_parts_kernel");
    return FUMI_OK;
}

extern "C" int fumi_scatter_add_rows(const float* src, const int64_t* rows, int64_t n, int64_t width, float* table,
                                     void* stream) {
    FUMI_CHECK_ARG(n >= 0 && width >= 1, "bad shape");
    if (n == 0) return FUMI_OK;
    FUMI_CHECK_ARG(src && rows && table, "null pointer");
    FUMI_LAUNCH(scatter_add_rows_kernel, (unsigned)((n * width + 255) / 256), 256, 0, stream, src, rows, n, width, table);
    FUMI_CHECK_LAUNCH("scatter_add_rows_kernel");
    return FUMI_OK;
}

extern "C" int fumi_reduce_loss_acc(const float* task_loss, const float* task_acc, int64_t B, float* out,
                                    void* stream) {
    FUMI_CHECK_ARG(B >= 1, "B < 1");
    FUMI_CHECK_ARG(task_loss && task_acc && out, "null pointer");
    FUMI_LAUNCH(reduce_loss_acc_kernel, 1, 256, 0, stream, task_loss, task_acc, B, out);
    FUMI_CHECK_LAUNCH("reduce_loss_acc_kernel");
    return FUMI_OK;
}
