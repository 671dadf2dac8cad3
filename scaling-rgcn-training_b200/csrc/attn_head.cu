// attn_head.cu — the per-node core of the attention transfer head (reference model/layers.py:59-61:
// nn.MultiheadAttention(embed_dim = emb, num_heads = S, dropout = 0.2) over the S stacked summary embeddings,
// q = k = v, sequence length L = S, batch = all N nodes, and only attn_output[0] is used).
//
// The two projections on either side are dense contractions and run on the tcgen05 kernel (gemm_tc.cu).  What
// is left per (node, head) is an S-way softmax over dot products of head_dim (= 21) floats — for query position
// 0 only, because the reference keeps attn_output[0] alone:
//     score[s] = q[b, h] . k[s, b, h] / sqrt(head_dim)      p = softmax_s(score)      (dropout on p when training)
//     o[b, h]  = sum_s p[s] v[s, b, h]
// One thread per (node, head).  A head slice is 84 bytes: read or written straight from the rows it would be 4-byte
// accesses scattered over ~11 rows per warp (the first version did that: forward 1.38 ms, backward 4.1 ms at
// AM shape, the dK / dV stores being the cost).  The staged kernels (k_attn_fwd_st / k_attn_bwd_st) give every warp
// floor(32 / heads) whole nodes, bring their q / K_s / V_s rows into shared memory with coalesced 4-byte cp.async
// (rows packed tightly, so lane l's slice starts at float l * head_dim: conflict-free for odd head_dim), compute
// from there, write results over the staged inputs and store the tiles back row by row.  The per-thread kernels
// remain as the fallback for shapes whose tiles do not fit shared memory.
#include <algorithm>

#include "common.cuh"

namespace rgcn {
namespace {

constexpr int MAX_S = 8, MAX_D = 64;

struct AttnArgs {
    const float* q; int64_t ldq;        // [N, heads * d]      (query rows of summary 0, bias added, unscaled)
    const float* kv; int64_t ldkv;      // [S * N, >= voff + heads * d]: k columns at 0, v columns at voff
    int64_t voff;
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
    const int h = (int)(i - b * a.heads), d = DT > 0 ? DT : a.d;
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
        const float* vp = a.kv + ((int64_t)s * a.N + b) * a.ldkv + a.voff + h * d;
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
    const int h = (int)(i - b * a.heads), d = DT > 0 ? DT : a.d;
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
        const float* vp = a.kv + ((int64_t)s * a.N + b) * a.ldkv + a.voff + h * d;
        float acc = 0.f;
        _Pragma("unroll") for (int j = 0; j < d; ++j) acc = fmaf(g[j], vp[j], acc);
        dp[s] = acc * keep;
        dot = fmaf(p[s], dp[s], dot);
        // dL/dv[s] = (p dropped) * g
        float* gv = a.gkv + ((int64_t)s * a.N + b) * a.ldgkv + a.voff + h * d;
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

// ---------------------------------------------------------------------------------------------
// staged kernels
// ---------------------------------------------------------------------------------------------
constexpr int AW = 4;   // warps per block

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct WarpTile {
    int64_t b0;        // first node of the warp
    int rows;          // nodes of the warp that exist
    int qp, rs, tile;  // quads per row (ceil4(heads * d) / 4), floats per staged row (4 qp + 4), floats per tile
    float* base;       // (1 + S) tiles
    bool active;       // this lane owns a (node, head)
    int64_t item;      // its global (node, head) index
    int slice;         // float offset of the lane's head slice inside a tile
};
// tile[r * rs + 4 c4 ..] <- g[r * ld + 4 c4 ..]   rows x qp quads, 16-byte cp.async (two rows per warp instruction)
__device__ __forceinline__ void stage_rows(const WarpTile& w, float* tile, const float* g, int64_t ld, int lane) {
    for (int i = lane; i < w.rows * w.qp; i += 32) {
        const int r = i / w.qp, c = (i - r * w.qp) * 4;
        cp_async16(tile + r * w.rs + c, g + r * ld + c);
    }
}
__device__ __forceinline__ void store_rows(const WarpTile& w, float* g, int64_t ld, const float* tile, int lane) {
    for (int i = lane; i < w.rows * w.qp; i += 32) {
        const int r = i / w.qp, c = (i - r * w.qp) * 4;
        *reinterpret_cast<float4*>(g + r * ld + c) = *reinterpret_cast<const float4*>(tile + r * w.rs + c);
    }
}
template <int DT>
__device__ __forceinline__ bool warp_tile(const AttnArgs& a, float* smem, WarpTile& w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = DT > 0 ? DT : a.d;
    const int npw = 32 / a.heads;
    w.qp = (a.heads * d + 3) >> 2;
    w.rs = 4 * w.qp + 4;     // + 4 floats: rows stay 16-byte aligned and the head slices of different nodes spread over the banks
    w.tile = npw * w.rs;
    w.base = smem + (size_t)warp * (1 + a.S) * w.tile;
    w.b0 = ((int64_t)blockIdx.x * AW + warp) * npw;
    if (w.b0 >= a.N) return false;
    w.rows = (int)min((int64_t)npw, a.N - w.b0);
    w.active = lane < w.rows * a.heads;
    w.item = w.b0 * a.heads + lane;
    w.slice = (lane / a.heads) * w.rs + (lane % a.heads) * d;
    return true;
}

template <int DT>
__global__ void __launch_bounds__(AW * 32) k_attn_fwd_st(const AttnArgs a) {
    extern __shared__ __align__(16) float attn_smem[];
    const int lane = threadIdx.x & 31;
    const int d = DT > 0 ? DT : a.d;
    constexpr int DM = DT > 0 ? DT : MAX_D;
    WarpTile w;
    if (!warp_tile<DT>(a, attn_smem, w)) return;
    // q and the K part of every summary
    stage_rows(w, w.base, a.q + w.b0 * a.ldq, a.ldq, lane);
    for (int s = 0; s < a.S; ++s)
        stage_rows(w, w.base + (1 + s) * w.tile, a.kv + ((int64_t)s * a.N + w.b0) * a.ldkv, a.ldkv, lane);
    cp_async_wait_all();
    __syncwarp();
    const float scale = rsqrtf((float)d);
    float sc[MAX_S], mx = -INFINITY;
    if (w.active) {
        float qv[DM];
        const float* qs = w.base + w.slice;
        _Pragma("unroll") for (int j = 0; j < d; ++j) qv[j] = qs[j] * scale;
        for (int s = 0; s < a.S; ++s) {
            const float* ks = w.base + (1 + s) * w.tile + w.slice;
            float acc = 0.f;
            _Pragma("unroll") for (int j = 0; j < d; ++j) acc = fmaf(qv[j], ks[j], acc);
            sc[s] = acc;
            mx = fmaxf(mx, acc);
        }
    }
    __syncwarp();   // every lane is done with the K tiles: the V parts take their place
    for (int s = 0; s < a.S; ++s)
        stage_rows(w, w.base + (1 + s) * w.tile, a.kv + ((int64_t)s * a.N + w.b0) * a.ldkv + a.voff, a.ldkv, lane);
    float pd[MAX_S];
    if (w.active) {
        float sum = 0.f;
        for (int s = 0; s < a.S; ++s) {
            sc[s] = expf(sc[s] - mx);
            sum += sc[s];
        }
        const float inv = 1.f / sum;
        for (int s = 0; s < a.S; ++s) {
            const float p = sc[s] * inv;
            a.probs[w.item * a.S + s] = p;
            pd[s] = a.keep ? p * a.keep[w.item * a.S + s] : p;
        }
    }
    cp_async_wait_all();
    __syncwarp();
    if (w.active) {
        float ov[DM];
        _Pragma("unroll") for (int j = 0; j < d; ++j) ov[j] = 0.f;
        for (int s = 0; s < a.S; ++s) {
            const float* vs = w.base + (1 + s) * w.tile + w.slice;
            _Pragma("unroll") for (int j = 0; j < d; ++j) ov[j] = fmaf(pd[s], vs[j], ov[j]);
        }
        float* os = w.base + w.slice;   // over the q tile (each lane overwrites only its own slice)
        _Pragma("unroll") for (int j = 0; j < d; ++j) os[j] = ov[j];
    }
    __syncwarp();
    store_rows(w, a.o + w.b0 * a.ldo, a.ldo, w.base, lane);
}

template <int DT>
__global__ void __launch_bounds__(AW * 32) k_attn_bwd_st(const AttnArgs a) {
    extern __shared__ __align__(16) float attn_smem[];
    const int lane = threadIdx.x & 31;
    const int d = DT > 0 ? DT : a.d;
    constexpr int DM = DT > 0 ? DT : MAX_D;
    WarpTile w;
    if (!warp_tile<DT>(a, attn_smem, w)) return;
    // phase A: dL/do and the V parts -> dL/dp, the softmax Jacobian, dL/dV (in place over V)
    stage_rows(w, w.base, a.go + w.b0 * a.ldgo, a.ldgo, lane);
    for (int s = 0; s < a.S; ++s)
        stage_rows(w, w.base + (1 + s) * w.tile, a.kv + ((int64_t)s * a.N + w.b0) * a.ldkv + a.voff, a.ldkv, lane);
    float p[MAX_S], ds[MAX_S];
    if (w.active)
        for (int s = 0; s < a.S; ++s) {
            p[s] = a.probs[w.item * a.S + s];
            ds[s] = a.keep ? a.keep[w.item * a.S + s] : 1.f;   // (keep factor for now)
        }
    cp_async_wait_all();
    __syncwarp();
    const float scale = rsqrtf((float)d);
    if (w.active) {
        float g[DM];
        const float* gs = w.base + w.slice;
        _Pragma("unroll") for (int j = 0; j < d; ++j) g[j] = gs[j];
        float dot = 0.f;
        for (int s = 0; s < a.S; ++s) {
            float* vs = w.base + (1 + s) * w.tile + w.slice;
            float acc = 0.f;
            _Pragma("unroll") for (int j = 0; j < d; ++j) acc = fmaf(g[j], vs[j], acc);
            const float keep = ds[s], pdrop = p[s] * keep;
            _Pragma("unroll") for (int j = 0; j < d; ++j) vs[j] = pdrop * g[j];          // dL/dv[s]
            ds[s] = acc * keep;                                                           // dL/dp[s]
            dot = fmaf(p[s], ds[s], dot);
        }
        for (int s = 0; s < a.S; ++s) ds[s] = p[s] * (ds[s] - dot) * scale;               // dL/d(q.k_s)
    }
    __syncwarp();
    for (int s = 0; s < a.S; ++s)
        store_rows(w, a.gkv + ((int64_t)s * a.N + w.b0) * a.ldgkv + a.voff, a.ldgkv, w.base + (1 + s) * w.tile, lane);
    __syncwarp();   // the tiles have been read back: q and the K parts take their place
    // phase B: q and the K parts -> dL/dq, dL/dK (in place over K)
    stage_rows(w, w.base, a.q + w.b0 * a.ldq, a.ldq, lane);
    for (int s = 0; s < a.S; ++s)
        stage_rows(w, w.base + (1 + s) * w.tile, a.kv + ((int64_t)s * a.N + w.b0) * a.ldkv, a.ldkv, lane);
    cp_async_wait_all();
    __syncwarp();
    if (w.active) {
        float qv[DM], gq[DM];
        float* qs = w.base + w.slice;
        _Pragma("unroll") for (int j = 0; j < d; ++j) {
            qv[j] = qs[j];
            gq[j] = 0.f;
        }
        for (int s = 0; s < a.S; ++s) {
            float* ks = w.base + (1 + s) * w.tile + w.slice;
            _Pragma("unroll") for (int j = 0; j < d; ++j) {
                gq[j] = fmaf(ds[s], ks[j], gq[j]);
                ks[j] = ds[s] * qv[j];                                                    // dL/dk[s]
            }
        }
        _Pragma("unroll") for (int j = 0; j < d; ++j) qs[j] = gq[j];
    }
    __syncwarp();
    store_rows(w, a.gq + w.b0 * a.ldgq, a.ldgq, w.base, lane);
    for (int s = 0; s < a.S; ++s)
        store_rows(w, a.gkv + ((int64_t)s * a.N + w.b0) * a.ldgkv, a.ldgkv, w.base + (1 + s) * w.tile, lane);
}

// shared memory of one block of the staged kernels, or 0 when the shape does not fit them
size_t staged_smem(const AttnArgs& a, bool backward) {
    static const bool on = [] {
        const char* e = getenv("RGCN_B200_ATTN_STAGED");
        return !(e && e[0] == '0');
    }();
    if (!on || a.heads > 32) return 0;
    // whole quads are copied: every row must be 16-byte addressable and hold ceil4(heads * d) floats on both sides of
    // the K | V split
    const int ep = (a.heads * a.d + 3) & ~3;
    auto ok = [&](const void* p, int64_t ld) { return ((uintptr_t)p & 15) == 0 && ld % 4 == 0 && ld >= ep; };
    if (!ok(a.q, a.ldq) || !ok(a.kv, a.ldkv) || a.voff % 4 != 0 || a.voff < ep || a.ldkv < a.voff + ep) return 0;
    if (!backward && !ok(a.o, a.ldo)) return 0;
    if (backward && (!ok(a.go, a.ldgo) || !ok(a.gq, a.ldgq) || !ok(a.gkv, a.ldgkv) || a.ldgkv < a.voff + ep)) return 0;
    const size_t bytes = (size_t)AW * (1 + a.S) * (32 / a.heads) * (ep + 4) * sizeof(float);
    return bytes <= 96 * 1024 ? bytes : 0;
}

template <typename K>
int launch_staged(K kern, const AttnArgs& a, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) RGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t per_block = (int64_t)AW * (32 / a.heads);
    kern<<<(unsigned)((a.N + per_block - 1) / per_block), AW * 32, smem, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int check(const AttnArgs& a, const char* who) {
    if (!a.q || !a.kv || !a.probs || a.S <= 0 || a.S > MAX_S || a.heads <= 0 || a.d <= 0 || a.d > MAX_D || a.N < 0 ||
        a.ldq < a.heads * a.d || a.voff < a.heads * a.d || a.ldkv < a.voff + a.heads * a.d)
        return fail(RGCN_ERR_INVALID_ARG, std::string(who) + ": bad argument (1 <= S <= 8, head_dim <= 64)");
    return 0;
}

}  // namespace
}  // namespace rgcn

using namespace rgcn;

extern "C" int rgcn_attn_head_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, int64_t v_offset, int32_t num_sums,
                                  int64_t num_nodes, int32_t heads, int32_t head_dim, const float* keep_scaled, float* probs,
                                  float* o, int64_t ldo, void* stream) {
    AttnArgs a{};
    a.q = q; a.ldq = ldq; a.kv = kv; a.ldkv = ldkv; a.voff = v_offset; a.S = num_sums; a.heads = heads; a.d = head_dim; a.N = num_nodes;
    a.keep = keep_scaled; a.probs = probs; a.o = o; a.ldo = ldo;
    int rc = check(a, "rgcn_attn_head_fwd");
    if (rc) return rc;
    if (!o || ldo < heads * head_dim) return fail(RGCN_ERR_INVALID_ARG, "rgcn_attn_head_fwd: bad output");
    if (num_nodes == 0) return 0;
    const int64_t total = num_nodes * heads;
    note_launch(1);
    if (const size_t smem = staged_smem(a, false)) {
        if (head_dim == 21) return launch_staged(k_attn_fwd_st<21>, a, smem, (cudaStream_t)stream);   // emb 63, 3 summaries
        if (head_dim == 16) return launch_staged(k_attn_fwd_st<16>, a, smem, (cudaStream_t)stream);
        return launch_staged(k_attn_fwd_st<0>, a, smem, (cudaStream_t)stream);
    }
    const int grid = (int)((total + 255) / 256);
    if (head_dim == 21) k_attn_fwd<21><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else if (head_dim == 16) k_attn_fwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else k_attn_fwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int rgcn_attn_head_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, int64_t v_offset, int32_t num_sums,
                                  int64_t num_nodes, int32_t heads, int32_t head_dim, const float* keep_scaled,
                                  const float* probs, const float* go, int64_t ldgo, float* gq, int64_t ldgq, float* gkv,
                                  int64_t ldgkv, void* stream) {
    AttnArgs a{};
    a.q = q; a.ldq = ldq; a.kv = kv; a.ldkv = ldkv; a.voff = v_offset; a.S = num_sums; a.heads = heads; a.d = head_dim; a.N = num_nodes;
    a.keep = keep_scaled; a.probs = const_cast<float*>(probs);
    a.go = go; a.ldgo = ldgo; a.gq = gq; a.ldgq = ldgq; a.gkv = gkv; a.ldgkv = ldgkv;
    int rc = check(a, "rgcn_attn_head_bwd");
    if (rc) return rc;
    if (!go || !gq || !gkv || ldgo < heads * head_dim || ldgq < heads * head_dim || ldgkv < v_offset + heads * head_dim)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_attn_head_bwd: bad gradient buffers");
    if (num_nodes == 0) return 0;
    const int64_t total = num_nodes * heads;
    note_launch(1);
    if (const size_t smem = staged_smem(a, true)) {
        if (head_dim == 21) return launch_staged(k_attn_bwd_st<21>, a, smem, (cudaStream_t)stream);
        if (head_dim == 16) return launch_staged(k_attn_bwd_st<16>, a, smem, (cudaStream_t)stream);
        return launch_staged(k_attn_bwd_st<0>, a, smem, (cudaStream_t)stream);
    }
    const int grid = (int)((total + 255) / 256);
    if (head_dim == 21) k_attn_bwd<21><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else if (head_dim == 16) k_attn_bwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else k_attn_bwd<0><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
