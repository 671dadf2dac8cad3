// map_gather.cu — K5: summary -> original embedding map-gather with the three combiners.
// Replaces the O(N*S) Python dict walk of reference model/embeddingTricks.py:8-25 and
// sum/concat/stack (:27-49).  The string->index resolution stays on the host (exact integer
// work, done once); the device does the row traffic: one warp per 32 original nodes, flat 128-byte walks.
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

// One warp per block of 32 consecutive original nodes.  The block's output is ONE contiguous span of
// 32 * feat floats (sum / stack; per-summary spans of feat for concat), so the warp walks it 32 floats at a
// time: every store instruction writes a full, aligned 128-byte line whatever the row width (63 floats = 252 B
// rows are not even 16-byte aligned, a warp-per-row walk leaves every second store half empty), and the loads
// of one instruction fall into at most two table rows.  The block's map entries sit in shared memory; the row
// of an element is (j + 0.5) / feat in float, exact for j < 32 * feat <= 2^16.
__global__ void __launch_bounds__(256) k_map_gather(const MapArgs a) {
    __shared__ int32_t s_idx[8][MAX_SUMS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t row0 = blk * 32;
    if (row0 >= a.n) return;
    const int rows = (int)min((int64_t)32, a.n - row0);
    const int feat = a.feat, span = rows * feat, ns = a.num_sums;
    const float inv_feat = 1.0f / (float)feat;
    for (int s = 0; s < ns; ++s) s_idx[warp][s][lane] = lane < rows ? a.idx[s][row0 + lane] : -1;
    __syncwarp();
    auto fetch = [&](int s, int r, int c) -> float {
        const int32_t j = s_idx[warp][s][r];
        const float* p = j >= 0 ? a.emb[s] + (int64_t)j * feat + c : a.fb[s] + (row0 + r) * feat + c;
        return __ldg(p);
    };
    // two 32-float steps per trip: the loads of both are in flight together
    for (int j0 = 0; j0 < span; j0 += 64) {
        int r[2], c[2];
        bool ok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = j0 + 32 * h + lane;
            ok[h] = j < span;
            const int jj = ok[h] ? j : 0;
            r[h] = min((int)(((float)jj + 0.5f) * inv_feat), rows - 1);
            c[h] = jj - r[h] * feat;
        }
        if (a.mode == 0) {
            float acc[2] = {0.f, 0.f};
            for (int s = 0; s < ns; ++s) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float v = ok[h] ? fetch(s, r[h], c[h]) : 0.f;
                    acc[h] = s == 0 ? v : acc[h] + v;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (ok[h]) a.out[row0 * feat + j0 + 32 * h + lane] = acc[h];
        } else {
            for (int s = 0; s < ns; ++s) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!ok[h]) continue;
                    const float v = fetch(s, r[h], c[h]);
                    // stack: out[s][node][c] ; concat: out[node][s * feat + c]
                    const int64_t o = a.mode == 2 ? ((int64_t)s * a.n + row0 + r[h]) * feat + c[h]
                                                  : ((row0 + r[h]) * ns + s) * (int64_t)feat + c[h];
                    a.out[o] = v;
                }
            }
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
    if (feat > 2048) return fail(RGCN_ERR_UNSUPPORTED, "rgcn_map_gather: feat > 2048");
    const int wpb = 8;
    const int64_t blocks32 = (num_nodes + 31) / 32;
    ProfScope prof(TAG_MAP, feat, num_sums, (cudaStream_t)stream);
    note_launch(1);
    k_map_gather<<<(int)((blocks32 + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
