// attn_head.cu — the per-node core of the attention transfer head (reference model/layers.py:59-61:
// nn.MultiheadAttention(embed_dim = emb, num_heads = S, dropout = 0.2) over the S stacked summary embeddings,
// q = k = v, sequence length L = S, batch = all N nodes, and only attn_output[0] is used).
//
// The two projections on either side are dense contractions and run on the tcgen05 kernel (gemm_tc.cu).  What
// is left per (node, head) is an S-way softmax over dot products of head_dim (= 21) floats — for query position
// 0 only, because the reference keeps attn_output[0] alone:
//     score[s] = q[b, h] . k[s, b, h] / sqrt(head_dim)      p = softmax_s(score)      (dropout on p when training)
//     o[b, h]  = sum_s p[s] v[s, b, h]
// One thread per (node, head): HBM-bound streaming of q, k, v rows (adjacent threads read adjacent 84-byte head
// slices, so a warp touches ~11 contiguous rows).  Backward is the same walk with the softmax Jacobian.
#include <algorithm>

#include "common.cuh"

namespace rgcn {
namespace {

constexpr int MAX_S = 8, MAX_D = 64;

struct AttnArgs {
    const float* q; int64_t ldq;        // [N, heads * d]      (query rows of summary 0, bias added, unscaled)
    const float* kv; int64_t ldkv;      // [S * N, 2 * heads * d]: k columns first, then v
    int S, heads, d;
    int64_t N;
    const float* keep;                  // nullable [N, heads, S]: dropout keep mask already scaled by 1 / (1 - p)
    float* probs;                       // [N, heads, S] softmax BEFORE dropout (saved for backward)
    float* o; int64_t ldo;              // [N, heads * d]
    // backward
    const float* go; int64_t ldgo;      // dL/do
    float* gq; int64_t ldgq;
    float* gkv; int64_t ldgkv;
};

// DT > 0: head_dim known at compile time (the loops unroll and the row slices live in registers); 0 = generic
template <int DT>
__global__ void __launch_bounds__(256) k_attn_fwd(const AttnArgs a) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= a.N * a.heads) return;
    const int64_t b = i / a.heads;
    const int h = (int)(i - b * a.heads), d = DT > 0 ? DT : a.d, hd = a.heads * d;
    const float scale = rsqrtf((float)d);
    float qv[DT > 0 ? DT : MAX_D];
    const float* qp = a.q + b * a.ldq + h * d;
    _Pragma("unroll") for (int j = 0; j < d; ++j) qv[j] = qp[j] * scale;
    float sc[MAX_S], mx = -INFINITY;
    for (int s = 0; s < a.S; ++s) {
        const float* kp = a.kv + ((int64_t)s * a.N + b) * a.ldkv + h * d;
        float acc = 0.f;
        _Pragma("unroll") for (int j = 0; j < d; ++j) acc = fmaf(qv[j], kp[j], acc);
        sc[s] = acc;
        mx = fmaxf(mx, acc);
    }
    float sum = 0.f;
    for (int s = 0; s < a.S; ++s) {
        sc[s] = expf(sc[s] - mx);
        sum += sc[s];
    }
    const float inv = 1.f / sum;
    float ov[DT > 0 ? DT : MAX_D];
    _Pragma("unroll") for (int j = 0; j < d; ++j) ov[j] = 0.f;
    for (int s = 0; s < a.S; ++s) {
        const float p = sc[s] * inv;
        a.probs[i * a.S + s] = p;
        const float pd = a.keep ? p * a.keep[i * a.S + s] : p;
        const float* vp = a.kv + ((int64_t)s * a.N + b) * a.ldkv + hd + h * d;
        _Pragma("unroll") for (int j = 0; j < d; ++j) ov[j] = fmaf(pd, vp[j], ov[j]);
    }
    float* op = a.o + b * a.ldo + h * d;
    _Pragma("unroll") for (int j = 0; j < d; ++j) op[j] = ov[j];
}

template <int DT>
__global__ void __launch_bounds__(256) k_attn_bwd(const AttnArgs a) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= a.N * a.heads) return;
    const int64_t b = i / a.heads;
    const int h = (int)(i - b * a.heads), d = DT > 0 ? DT : a.d, hd = a.heads * d;
    const float scale = rsqrtf((float)d);
    float g[DT > 0 ? DT : MAX_D], qv[DT > 0 ? DT : MAX_D];
    const float* gp = a.go + b * a.ldgo + h * d;
    const float* qp = a.q + b * a.ldq + h * d;
    _Pragma("unroll") for (int j = 0; j < d; ++j) {
        g[j] = gp[j];
        qv[j] = qp[j];
    }
    // dL/dp (through the dropout scaling), then the softmax Jacobian on the pre-dropout probabilities
    float p[MAX_S], dp[MAX_S], dot = 0.f;
    for (int s = 0; s < a.S; ++s) {
        p[s] = a.probs[i * a.S + s];
        const float keep = a.keep ? a.keep[i * a.S + s] : 1.f;
        const float* vp = a.kv + ((int64_t)s * a.N + b) * a.ldkv + hd + h * d;
        float acc = 0.f;
        _Pragma("unroll") for (int j = 0; j < d; ++j) acc = fmaf(g[j], vp[j], acc);
        dp[s] = acc * keep;
        dot = fmaf(p[s], dp[s], dot);
        // dL/dv[s] = (p dropped) * g
        float* gv = a.gkv + ((int64_t)s * a.N + b) * a.ldgkv + hd + h * d;
        const float pd = p[s] * keep;
        _Pragma("unroll") for (int j = 0; j < d; ++j) gv[j] = pd * g[j];
    }
    float gq[DT > 0 ? DT : MAX_D];
    _Pragma("unroll") for (int j = 0; j < d; ++j) gq[j] = 0.f;
    for (int s = 0; s < a.S; ++s) {
        const float ds = p[s] * (dp[s] - dot) * scale;       // dL/d(q.k)
        const float* kp = a.kv + ((int64_t)s * a.N + b) * a.ldkv + h * d;
        float* gk = a.gkv + ((int64_t)s * a.N + b) * a.ldgkv + h * d;
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            gq[j] = fmaf(ds, kp[j], gq[j]);
            gk[j] = ds * qv[j];
        }
    }
    float* gqp = a.gq + b * a.ldgq + h * d;
    _Pragma("unroll") for (int j = 0; j < d; ++j) gqp[j] = gq[j];
}

int check(const AttnArgs& a, const char* who) {
    if (!a.q || !a.kv || !a.probs || a.S <= 0 || a.S > MAX_S || a.heads <= 0 || a.d <= 0 || a.d > MAX_D || a.N < 0 ||
        a.ldq < a.heads * a.d || a.ldkv < 2 * a.heads * a.d)
        return fail(RGCN_ERR_INVALID_ARG, std::string(who) + ": bad argument (1 <= S <= 8, head_dim <= 64)");
    return 0;
}

}  // namespace
}  // namespace rgcn

using namespace rgcn;

extern "C" int rgcn_attn_head_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, int32_t num_sums,
                                  int64_t num_nodes, int32_t heads, int32_t head_dim, const float* keep_scaled, float* probs,
                                  float* o, int64_t ldo, void* stream) {
    AttnArgs a{};
    a.q = q; a.ldq = ldq; a.kv = kv; a.ldkv = ldkv; a.S = num_sums; a.heads = heads; a.d = head_dim; a.N = num_nodes;
    a.keep = keep_scaled; a.probs = probs; a.o = o; a.ldo = ldo;
    int rc = check(a, "rgcn_attn_head_fwd");
    if (rc) return rc;
    if (!o || ldo < heads * head_dim) return fail(RGCN_ERR_INVALID_ARG, "rgcn_attn_head_fwd: bad output");
    if (num_nodes == 0) return 0;
    const int64_t total = num_nodes * heads;
    note_launch(1);
    const int grid = (int)((total + 255) / 256);
    if (head_dim == 21) k_attn_fwd<21><<<grid, 256, 0, (cudaStream_t)stream>>>(a);        // emb 63, 3 summaries
    else if (head_dim == 16) k_attn_fwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else k_attn_fwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int rgcn_attn_head_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, int32_t num_sums,
                                  int64_t num_nodes, int32_t heads, int32_t head_dim, const float* keep_scaled,
                                  const float* probs, const float* go, int64_t ldgo, float* gq, int64_t ldgq, float* gkv,
                                  int64_t ldgkv, void* stream) {
    AttnArgs a{};
    a.q = q; a.ldq = ldq; a.kv = kv; a.ldkv = ldkv; a.S = num_sums; a.heads = heads; a.d = head_dim; a.N = num_nodes;
    a.keep = keep_scaled; a.probs = const_cast<float*>(probs);
    a.go = go; a.ldgo = ldgo; a.gq = gq; a.ldgq = ldgq; a.gkv = gkv; a.ldgkv = ldgkv;
    int rc = check(a, "rgcn_attn_head_bwd");
    if (rc) return rc;
    if (!go || !gq || !gkv || ldgo < heads * head_dim || ldgq < heads * head_dim || ldgkv < 2 * heads * head_dim)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_attn_head_bwd: bad gradient buffers");
    if (num_nodes == 0) return 0;
    const int64_t total = num_nodes * heads;
    note_launch(1);
    const int grid = (int)((total + 255) / 256);
    if (head_dim == 21) k_attn_bwd<21><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else if (head_dim == 16) k_attn_bwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else k_attn_bwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
