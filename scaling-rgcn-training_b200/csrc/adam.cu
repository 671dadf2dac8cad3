// adam.cu — one-pass Adam with L2 weight decay: the update the
// reference builds at model/modelTrainer.py:44 (torch.optim.Adam(lr, weight_decay), defaults
// betas = (0.9, 0.999), eps = 1e-8, no amsgrad) applied to one parameter tensor in a single
// kernel — 4 streams in, 3 out (28 B/element) instead of torch's multi-kernel foreach chain.
// At AM-shape the [N,63] embedding alone is 105 M elements.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace rgcn {
namespace {

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, int64_t n, float lr_over_bc1, float beta1,
                                              float beta2, float eps, float wd, float inv_sqrt_bc2,
                                              const int64_t* __restrict__ step_dev, float lr) {
    if (step_dev) {   // graph-replayable form: the step count lives on the device, same double-precision formulas
        const double t = (double)*step_dev;
        lr_over_bc1 = (float)((double)lr / (1.0 - pow((double)beta1, t)));
        inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, t)));
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float pi = p[i];
        const float gi = fmaf(wd, pi, g[i]);                     // grad + weight_decay * param
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);  // exp_avg
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        const float denom = fmaf(sqrtf(vi), inv_sqrt_bc2, eps);  // sqrt(v)/sqrt(1-b2^t) + eps
        p[i] = pi - lr_over_bc1 * (mi / denom);
    }
}

}  // namespace
}  // namespace rgcn

extern "C" int rgcn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream) {
    using namespace rgcn;
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_adam_step: bad argument");
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    note_launch(1);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), beta1, beta2, eps,
                                                   weight_decay, (float)(1.0 / sqrt(bc2)), nullptr, lr);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int rgcn_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, const int64_t* step_dev,
                                  void* stream) {
    using namespace rgcn;
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || !step_dev)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_adam_step_dev: bad argument");
    if (n == 0) return 0;
    note_launch(1);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, 0.f, beta1, beta2, eps,
                                                   weight_decay, 0.f, step_dev, lr);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
