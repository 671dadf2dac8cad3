// map_gather.cu — K5: summary -> original embedding map-gather with the three combiners.
// Replaces the O(N*S) Python dict walk of reference model/embeddingTricks.py:8-25 and
// sum/concat/stack (:27-49).  The string->index resolution stays on the host (exact integer
// work, done once); the device does the row traffic: one warp per original node.
// Sum order is s = 0..S-1 with the same fp32 adds as python's sum() -> bit-exact.
#include "common.cuh"

namespace rgcn {
namespace {

constexpr int MAX_SUMS = 16;
struct MapArgs {
    const float* emb[MAX_SUMS];
    const int32_t* idx[MAX_SUMS];
    const float* fb[MAX_SUMS];
    int num_sums;
    int64_t n;
    int feat;
    int mode;
    float* out;
};

__global__ void __launch_bounds__(256) k_map_gather(const MapArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= a.n) return;
    if (a.mode == 0) {
        for (int c = lane; c < a.feat; c += 32) {
            float acc = 0.f;
            for (int s = 0; s < a.num_sums; ++s) {
                const int32_t j = a.idx[s][i];
                const float v = j >= 0 ? a.emb[s][(int64_t)j * a.feat + c] : a.fb[s][i * a.feat + c];
                acc = s == 0 ? v : acc + v;
            }
            a.out[i * a.feat + c] = acc;
        }
    } else {
        for (int s = 0; s < a.num_sums; ++s) {
            const int32_t j = a.idx[s][i];
            const float* rp = j >= 0 ? a.emb[s] + (int64_t)j * a.feat : a.fb[s] + i * a.feat;
            float* op = a.mode == 1 ? a.out + (i * a.num_sums + s) * a.feat : a.out + ((int64_t)s * a.n + i) * a.feat;
            for (int c = lane; c < a.feat; c += 32) op[c] = rp[c];
        }
    }
}

}  // namespace
}  // namespace rgcn

extern "C" int rgcn_map_gather(const float* const* host_emb, const int32_t* const* host_idx,
                               const float* const* host_fallback, int32_t num_sums, int64_t num_nodes, int32_t feat,
                               int32_t mode, float* out, void* stream) {
    using namespace rgcn;
    if (!host_emb || !host_idx || !host_fallback || !out || num_sums <= 0 || num_sums > MAX_SUMS || num_nodes < 0 ||
        feat <= 0 || mode < 0 || mode > 2)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_map_gather: bad argument (1 <= num_sums <= 16, mode in 0..2)");
    if (num_nodes == 0) return 0;
    MapArgs a{};
    for (int s = 0; s < num_sums; ++s) {
        if (!host_emb[s] || !host_idx[s] || !host_fallback[s])
            return fail(RGCN_ERR_INVALID_ARG, "rgcn_map_gather: null table pointer");
        a.emb[s] = host_emb[s];
        a.idx[s] = host_idx[s];
        a.fb[s] = host_fallback[s];
    }
    a.num_sums = num_sums;
    a.n = num_nodes;
    a.feat = feat;
    a.mode = mode;
    a.out = out;
    const int wpb = 8;
    ProfScope prof(TAG_MAP, feat, num_sums, (cudaStream_t)stream);
    note_launch(1);
    k_map_gather<<<(int)((num_nodes + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
