// eval_counts.cu — device side of the reference's `evaluate` (model/evaluation.py:14-31): turn the
// model output into predictions (argmax -> one-hot for the CE path, round() for the sigmoid/BCE
// path) for the evaluated rows and count, per class, true positives / false positives / false
// negatives, plus the number of rows predicted exactly — everything sklearn's accuracy_score and
// f1_score(average='weighted' | 'macro') need on multilabel-indicator input.  Only 3C+1 integers
// cross to the host (the reference hands whole device tensors to sklearn, which fails on CUDA).
#include "common.cuh"

namespace rgcn {
namespace {

__global__ void __launch_bounds__(256) k_eval_counts(const float* __restrict__ pred, int64_t ldp, int C,
                                                     const int64_t* __restrict__ idx, int64_t n,
                                                     const int64_t* __restrict__ y, int mode,
                                                     unsigned long long* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const float* p = pred + idx[r] * ldp;
    int arg = 0;
    if (mode == 0) {   // first index of the maximum, like torch.argmax
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = p[c];
            if (v > best || (v == best && c < bi)) {
                best = v;
                bi = c;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) {
                best = ob;
                bi = oi;
            }
        }
        arg = bi;
    }
    bool all_equal = true;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        bool pc = false, yc = false;
        if (c < C) {
            pc = mode == 0 ? (c == arg) : (rintf(p[c]) != 0.f);   // torch.round: half to even
            yc = y[r * C + c] != 0;
            if (pc && yc) atomicAdd(counts + c, 1ull);
            if (pc && !yc) atomicAdd(counts + C + c, 1ull);
            if (!pc && yc) atomicAdd(counts + 2 * C + c, 1ull);
        }
        all_equal = all_equal && __all_sync(0xffffffffu, c >= C || pc == yc);
    }
    if (lane == 0 && all_equal) atomicAdd(counts + 3 * C, 1ull);
}

}  // namespace
}  // namespace rgcn

extern "C" int rgcn_eval_counts(const float* pred, int64_t ldp, int32_t num_classes, const int64_t* idx, int64_t n,
                                const int64_t* y, int32_t mode, int64_t* counts, void* stream) {
    using namespace rgcn;
    if (!pred || !idx || !y || !counts || num_classes <= 0 || ldp < num_classes || n < 0 || mode < 0 || mode > 1)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_eval_counts: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    RGCN_CUDA(cudaMemsetAsync(counts, 0, (size_t)(3 * num_classes + 1) * sizeof(int64_t), st));
    if (n == 0) return 0;
    note_launch(1);
    k_eval_counts<<<(int)((n + 7) / 8), 256, 0, st>>>(pred, ldp, num_classes, idx, n, y, mode,
                                                     reinterpret_cast<unsigned long long*>(counts));
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
