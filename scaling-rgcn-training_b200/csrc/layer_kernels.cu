// layer_kernels.cu — K1/K3/K4: the R-GCN layer's forward, dL/dx and dL/dW passes for sm_100a.
//
// What they replace (reference model/layers.py:21,23 -> PyG RGCNConv loop path, SURVEY.md
// Appendix A): per relation a boolean mask over all E edges, an index_select gather, two
// scatter_add_ (sum and count), a divide over a dense [N,Fin] buffer and an SGEMM — R times —
// plus their autograd.  Here one pass walks the blocked relational CSR once:
//
//   warp <- batch of <=16 (relation, owner) segments of ONE relation
//     gather   coalesced row loads of the segment's sources (U rows in flight per warp),
//              weighted by 1/cnt(relation,dst), reduced in registers, mean rows staged in smem
//     mma      [16 x Kp] . B_rel[Kp x Np] on the tensor pipe, error-compensated 3xTF32
//              (hi*hi + lo*hi + hi*lo) so results stay within fp32 round-off of the reference
//     scatter  REDG.ADD.F32x4 of the 16 result rows into the owners' output rows
//
// The self loop (root) is relation R of the same structure; bias rides on it.
// Long segments were cut into chunk rows at build time; a pre-pass reduces those first.
#include <algorithm>

#include "common.cuh"

namespace rgcn {

namespace {

constexpr int TILE_WARPS = 8;
constexpr int WG_WARPS = 4;
constexpr int GATHER_U = 8;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void red_add_v4(float* p, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// Weighted sum of gathered rows over entries [ebeg,eend).  Lane l owns columns l and l+32.
// FLAGGED: entries carry LAST_FLAG; at each flag the running sums go to H[cs*ldh + col], cs++.
// !FLAGGED: a single sum is returned in (acc0, acc1).
template <int KP, bool FLAGGED>
__device__ __forceinline__ void gather_rows(const uint32_t* __restrict__ idx_arr, const float* __restrict__ w_arr,
                                            int ebeg, int eend, const float* __restrict__ feat, int64_t ldf, int kin,
                                            const float* __restrict__ aux, int64_t n_nodes, bool relu,
                                            float* __restrict__ H, int ldh, int lane, float& acc0, float& acc1) {
    acc0 = 0.f;
    acc1 = 0.f;
    int cs = 0;
    const bool c0 = lane < kin;
    const bool c1 = (KP > 32) && (lane + 32 < kin);
    for (int base = ebeg; base < eend; base += 32) {
        const int m = min(32, eend - base);
        uint32_t my_idx = 0;
        float my_w = 0.f;
        if (lane < m) {
            my_idx = idx_arr[base + lane];
            my_w = w_arr[base + lane];
        }
        for (int j0 = 0; j0 < m; j0 += GATHER_U) {
            float v0[GATHER_U], v1[GATHER_U];
#pragma unroll
            for (int u = 0; u < GATHER_U; ++u) {
                v0[u] = 0.f;
                v1[u] = 0.f;
                if (j0 + u < m) {
                    const uint32_t raw = __shfl_sync(FULL, my_idx, j0 + u);
                    const int64_t row = (int64_t)(raw & IDX_MASK);
                    const float* rp = row < n_nodes ? feat + row * ldf : aux + (row - n_nodes) * KP;
                    if (c0) v0[u] = __ldg(rp + lane);
                    if (KP > 32) {
                        if (c1) v1[u] = __ldg(rp + lane + 32);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < GATHER_U; ++u) {
                if (j0 + u < m) {
                    const uint32_t raw = __shfl_sync(FULL, my_idx, j0 + u);
                    const float w = __shfl_sync(FULL, my_w, j0 + u);
                    float x0 = v0[u], x1 = v1[u];
                    if (relu && (int64_t)(raw & IDX_MASK) < n_nodes) {
                        x0 = fmaxf(x0, 0.f);
                        x1 = fmaxf(x1, 0.f);
                    }
                    acc0 = fmaf(w, x0, acc0);
                    acc1 = fmaf(w, x1, acc1);
                    if (FLAGGED && (raw & LAST_FLAG)) {
                        if (lane < KP) H[cs * ldh + lane] = acc0;
                        if (KP > 32) H[cs * ldh + lane + 32] = acc1;
                        acc0 = 0.f;
                        acc1 = 0.f;
                        ++cs;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// chunk pre-pass: aux[c] = sum_{k in chunk c} raw_w[k] * feat[raw_idx[k]]   (one warp per chunk)
// ---------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_chunk_sum(const int32_t* __restrict__ raw_idx, const float* __restrict__ raw_w,
                                                   const int32_t* __restrict__ chunk_beg,
                                                   const int32_t* __restrict__ chunk_end, int num_chunks,
                                                   const float* __restrict__ feat, int64_t ldf, int kin,
                                                   int64_t n_nodes, int relu, float* __restrict__ aux) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= num_chunks) return;
    float a0, a1;
    gather_rows<KP, false>(reinterpret_cast<const uint32_t*>(raw_idx), raw_w, chunk_beg[c], chunk_end[c], feat, ldf,
                           kin, nullptr, n_nodes, relu != 0, nullptr, 0, lane, a0, a1);
    if (lane < KP) aux[(int64_t)c * KP + lane] = a0;
    if (KP > 32) aux[(int64_t)c * KP + lane + 32] = a1;
}

// ---------------------------------------------------------------------------------------------
// B-operand preparation: fragment-ordered, pre-split into tf32 hi/lo.
//   wfrag[((rel*KT + kt)*NT + nt)*32 + lane] = (b0.hi, b1.hi, b0.lo, b1.lo)
//   b0 = B[8kt + lane%4][8nt + lane/4], b1 = B[8kt + lane%4 + 4][8nt + lane/4]
//   B = W_rel (rows<fin, cols<fout) or its transpose; rel == R is the root matrix.
// ---------------------------------------------------------------------------------------------
__global__ void k_wprep(const float* __restrict__ weight, const float* __restrict__ root, int R, int fin, int fout,
                        int KT, int NT, int transpose, float4* __restrict__ wfrag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)(R + 1) * KT * NT * 32;
    if (i >= total) return;
    const int lane = (int)(i & 31);
    int64_t q = i >> 5;
    const int nt = (int)(q % NT);
    q /= NT;
    const int kt = (int)(q % KT);
    const int rel = (int)(q / KT);
    const float* W = rel < R ? weight + (int64_t)rel * fin * fout : root;
    const int g = lane >> 2, t = lane & 3;
    const int n = 8 * nt + g;
    float b[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = 8 * kt + t + 4 * h;
        float v = 0.f;
        if (W) {
            if (!transpose) {
                if (k < fin && n < fout) v = W[(int64_t)k * fout + n];
            } else {
                if (k < fout && n < fin) v = W[(int64_t)n * fout + k];
            }
        }
        b[h] = v;
    }
    uint32_t h0, l0, h1, l1;
    split_tf32(b[0], h0, l0);
    split_tf32(b[1], h1, l1);
    wfrag[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
}

// ---------------------------------------------------------------------------------------------
// K1 / K3: tile pass
// ---------------------------------------------------------------------------------------------
struct TileArgs {
    const uint32_t* e_idx;
    const float* e_w;
    const int32_t* seg_ptr;
    const int32_t* seg_own;
    const int32_t* bat_seg0;
    const int32_t* bat_info;
    int num_batches;
    const float* feat;
    int64_t ldf;
    int kin;
    const float* aux;
    int64_t n_nodes;
    const float4* wfrag;
    const float* bias;
    int nbias;
    int self_rel;
    float* out;
    int64_t ldo;
    int nout;
    int relu_in;
};

template <int KT, int NT>
__global__ void __launch_bounds__(TILE_WARPS * 32) k_tile(const TileArgs a) {
    constexpr int KP = KT * 8, LDH = KP + 4;
    constexpr bool BREG = (KT * NT <= 16);
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* H = smem + warp * (BS * LDH);
    const int g = lane >> 2, t = lane & 3;
    const int64_t gw = (int64_t)blockIdx.x * TILE_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * TILE_WARPS;
    const int per = (int)((a.num_batches + nw - 1) / nw);
    const int b_begin = (int)min((int64_t)a.num_batches, gw * per);
    const int b_end = (int)min((int64_t)a.num_batches, gw * per + per);
    float4 bfrag[BREG ? KT * NT : 1];
    int cur_rel = -1;
    for (int b = b_begin; b < b_end; ++b) {
        const int seg0 = a.bat_seg0[b];
        const int info = a.bat_info[b];
        const int nseg = info & 0xff, rel = info >> 8;
        const float4* wf = a.wfrag + (int64_t)rel * (KT * NT * 32) + lane;
        if constexpr (BREG) {
            if (rel != cur_rel) {
#pragma unroll
                for (int i = 0; i < KT * NT; ++i) bfrag[i] = __ldg(wf + i * 32);
                cur_rel = rel;
            }
        }
        const int my_own = lane < nseg ? a.seg_own[seg0 + lane] : -1;
        const int my_ptr = lane <= nseg ? a.seg_ptr[seg0 + lane] : 0;
        const int ebeg = __shfl_sync(FULL, my_ptr, 0), eend = __shfl_sync(FULL, my_ptr, nseg);
        if (nseg < BS) {
            for (int r = nseg; r < BS; ++r) {
                if (lane < KP) H[r * LDH + lane] = 0.f;
                if (KP > 32) H[r * LDH + lane + 32] = 0.f;
            }
        }
        float u0, u1;
        gather_rows<KP, true>(a.e_idx, a.e_w, ebeg, eend, a.feat, a.ldf, a.kin, a.aux, a.n_nodes, a.relu_in != 0, H, LDH,
                              lane, u0, u1);
        __syncwarp();
        float d[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) d[n][0] = d[n][1] = d[n][2] = d[n][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            uint32_t ah[4], al[4];
            split_tf32(H[g * LDH + 8 * kt + t], ah[0], al[0]);
            split_tf32(H[(g + 8) * LDH + 8 * kt + t], ah[1], al[1]);
            split_tf32(H[g * LDH + 8 * kt + t + 4], ah[2], al[2]);
            split_tf32(H[(g + 8) * LDH + 8 * kt + t + 4], ah[3], al[3]);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                float4 bf;
                if constexpr (BREG) bf = bfrag[kt * NT + n];
                else bf = __ldg(wf + (kt * NT + n) * 32);
                const uint32_t bh0 = __float_as_uint(bf.x), bh1 = __float_as_uint(bf.y);
                const uint32_t bl0 = __float_as_uint(bf.z), bl1 = __float_as_uint(bf.w);
                mma_tf32(d[n], al[0], al[1], al[2], al[3], bh0, bh1);
                mma_tf32(d[n], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                mma_tf32(d[n], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
            }
        }
        if (rel == a.self_rel && a.bias != nullptr) {
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int col = 8 * n + 2 * t;
                const float bx = col < a.nbias ? a.bias[col] : 0.f;
                const float by = col + 1 < a.nbias ? a.bias[col + 1] : 0.f;
                d[n][0] += bx;
                d[n][1] += by;
                d[n][2] += bx;
                d[n][3] += by;
            }
        }
        const int own_lo = __shfl_sync(FULL, my_own, g), own_hi = __shfl_sync(FULL, my_own, g + 8);
        const bool odd = (t & 1) != 0;
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) {
            const int col = odd ? 8 * (2 * j + 1) + 2 * (t - 1) : 8 * (2 * j) + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // h = 0: row g ; h = 1: row g + 8
                const float e0 = d[2 * j][2 * h], e1 = d[2 * j][2 * h + 1];
                const float f0 = d[2 * j + 1][2 * h], f1 = d[2 * j + 1][2 * h + 1];
                const float rx = __shfl_xor_sync(FULL, odd ? e0 : f0, 1);
                const float ry = __shfl_xor_sync(FULL, odd ? e1 : f1, 1);
                const int own = h ? own_hi : own_lo;
                if (own >= 0 && col < a.nout) {
                    float* p = a.out + (int64_t)own * a.ldo + col;
                    if (odd) red_add_v4(p, rx, ry, f0, f1);
                    else red_add_v4(p, e0, e1, rx, ry);
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// K4: dL/dW_rel = sum_seg mean_seg^T (x) gout[owner_seg]  (relation-major BRC), root and bias on
// the self-loop relation.  Per batch: H[16 x Kp] (re-gathered means), G[16 x Np]; D[Kp x Np] +=
// H^T . G on the tensor pipe (3xTF32), accumulated in registers across the batches of a relation
// and flushed with one atomic add per element per (warp, relation).
// ---------------------------------------------------------------------------------------------
struct WGradArgs {
    const uint32_t* e_idx;
    const float* e_w;
    const int32_t* seg_ptr;
    const int32_t* seg_own;
    const int32_t* bat_seg0;
    const int32_t* bat_info;
    int num_batches;
    const float* feat;
    int64_t ldf;
    int kin;
    const float* aux;
    int64_t n_nodes;
    const float* gout;
    int64_t ldg;
    int nout;
    int self_rel;
    float* gweight;
    float* groot;
    float* gbias;
    int relu_in;
};

template <int KT, int NT>
__global__ void __launch_bounds__(WG_WARPS * 32) k_wgrad(const WGradArgs a) {
    constexpr int KP = KT * 8, NP = NT * 8, MT = KP / 16;
    constexpr int LDH = KP + 8, LDG = NP + 8;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* H = smem + warp * (BS * (LDH + LDG));
    float* G = H + BS * LDH;
    const int g = lane >> 2, t = lane & 3;
    const int64_t gw = (int64_t)blockIdx.x * WG_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * WG_WARPS;
    const int per = (int)((a.num_batches + nw - 1) / nw);
    const int b_begin = (int)min((int64_t)a.num_batches, gw * per);
    const int b_end = (int)min((int64_t)a.num_batches, gw * per + per);
    float d[MT][NT][4];
    float bsum[(NP + 31) / 32];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) d[m][n][0] = d[m][n][1] = d[m][n][2] = d[m][n][3] = 0.f;
#pragma unroll
    for (int i = 0; i < (NP + 31) / 32; ++i) bsum[i] = 0.f;
    int cur_rel = -1;

    auto flush = [&](int rel) {
        float* dst = rel == a.self_rel ? a.groot : (a.gweight ? a.gweight + (int64_t)rel * a.kin * a.nout : nullptr);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = 16 * m + g + ((i & 2) ? 8 : 0);
                    const int col = 8 * n + 2 * t + (i & 1);
                    if (dst && row < a.kin && col < a.nout) atomicAdd(dst + (int64_t)row * a.nout + col, d[m][n][i]);
                    d[m][n][i] = 0.f;
                }
            }
    };

    for (int b = b_begin; b < b_end; ++b) {
        const int seg0 = a.bat_seg0[b];
        const int info = a.bat_info[b];
        const int nseg = info & 0xff, rel = info >> 8;
        if (rel != cur_rel) {
            if (cur_rel >= 0) flush(cur_rel);
            cur_rel = rel;
        }
        if (rel != a.self_rel && a.gweight == nullptr) continue;
        if (rel == a.self_rel && a.groot == nullptr && a.gbias == nullptr) continue;
        const int my_own = lane < nseg ? a.seg_own[seg0 + lane] : -1;
        const int my_ptr = lane <= nseg ? a.seg_ptr[seg0 + lane] : 0;
        const int ebeg = __shfl_sync(FULL, my_ptr, 0), eend = __shfl_sync(FULL, my_ptr, nseg);
        if (nseg < BS) {
            for (int r = nseg; r < BS; ++r) {
                if (lane < KP) H[r * LDH + lane] = 0.f;
                if (KP > 32) H[r * LDH + lane + 32] = 0.f;
            }
        }
        // G rows: gout[owner]
#pragma unroll 8
        for (int r = 0; r < BS; ++r) {
            const int own = __shfl_sync(FULL, my_own, r);
            float x0 = 0.f, x1 = 0.f;
            if (own >= 0) {
                const float* gp = a.gout + (int64_t)own * a.ldg;
                if (lane < a.nout) x0 = __ldg(gp + lane);
                if (NP > 32 && lane + 32 < a.nout) x1 = __ldg(gp + lane + 32);
            }
            if (lane < NP) G[r * LDG + lane] = x0;
            if (NP > 32) G[r * LDG + lane + 32] = x1;
        }
        float u0, u1;
        gather_rows<KP, true>(a.e_idx, a.e_w, ebeg, eend, a.feat, a.ldf, a.kin, a.aux, a.n_nodes, a.relu_in != 0, H, LDH,
                              lane, u0, u1);
        __syncwarp();
        if (rel == a.self_rel) {
#pragma unroll
            for (int i = 0; i < (NP + 31) / 32; ++i) {
                const int c = lane + 32 * i;
                if (c < NP) {
                    float s = 0.f;
#pragma unroll
                    for (int r = 0; r < BS; ++r) s += G[r * LDG + c];
                    bsum[i] += s;
                }
            }
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                split_tf32(G[(8 * ks + t) * LDG + 8 * n + g], bh[n][0], bl[n][0]);
                split_tf32(G[(8 * ks + t + 4) * LDG + 8 * n + g], bh[n][1], bl[n][1]);
            }
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                uint32_t ah[4], al[4];
                split_tf32(H[(8 * ks + t) * LDH + 16 * m + g], ah[0], al[0]);
                split_tf32(H[(8 * ks + t) * LDH + 16 * m + g + 8], ah[1], al[1]);
                split_tf32(H[(8 * ks + t + 4) * LDH + 16 * m + g], ah[2], al[2]);
                split_tf32(H[(8 * ks + t + 4) * LDH + 16 * m + g + 8], ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    mma_tf32(d[m][n], al[0], al[1], al[2], al[3], bh[n][0], bh[n][1]);
                    mma_tf32(d[m][n], ah[0], ah[1], ah[2], ah[3], bl[n][0], bl[n][1]);
                    mma_tf32(d[m][n], ah[0], ah[1], ah[2], ah[3], bh[n][0], bh[n][1]);
                }
            }
        }
        __syncwarp();
    }
    if (cur_rel >= 0) flush(cur_rel);
    if (a.gbias) {
#pragma unroll
        for (int i = 0; i < (NP + 31) / 32; ++i) {
            const int c = lane + 32 * i;
            if (c < a.nout && bsum[i] != 0.f) atomicAdd(a.gbias + c, bsum[i]);
        }
    }
}

__global__ void k_copy_cols(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t n,
                            int cols) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * cols) return;
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    dst[r * ldd + c] = src[r * lds + c];
}

__global__ void k_relu_mask(float* __restrict__ gr, int64_t ldg, const float* __restrict__ pre, int64_t ldp, int64_t n,
                            int cols) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * cols) return;
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    if (!(pre[r * ldp + c] > 0.f)) gr[r * ldg + c] = 0.f;
}

template <int KT, int NT>
int run_tile(const TileArgs& a, int num_sms, cudaStream_t st) {
    constexpr int LDH = KT * 8 + 4;
    const size_t smem = (size_t)TILE_WARPS * BS * LDH * sizeof(float);
    static bool configured = false;
    if (!configured) {
        RGCN_CUDA(cudaFuncSetAttribute(k_tile<KT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int per_sm = 1;
    RGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile<KT, NT>, TILE_WARPS * 32, smem));
    per_sm = std::max(per_sm, 1);
    int64_t want = ((int64_t)a.num_batches + TILE_WARPS - 1) / TILE_WARPS;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * per_sm));
    k_tile<KT, NT><<<grid, TILE_WARPS * 32, smem, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

template <int KT, int NT>
int run_wgrad(const WGradArgs& a, int num_sms, cudaStream_t st) {
    constexpr int LDH = KT * 8 + 8, LDG = NT * 8 + 8;
    const size_t smem = (size_t)WG_WARPS * BS * (LDH + LDG) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        RGCN_CUDA(cudaFuncSetAttribute(k_wgrad<KT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int per_sm = 1;
    RGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wgrad<KT, NT>, WG_WARPS * 32, smem));
    per_sm = std::max(per_sm, 1);
    int64_t want = ((int64_t)a.num_batches + WG_WARPS - 1) / WG_WARPS;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * per_sm));
    k_wgrad<KT, NT><<<grid, WG_WARPS * 32, smem, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

#define RGCN_DISPATCH_KN(FN, kp, np, ...)                                        \
    do {                                                                         \
        const int _k = (kp) / 8, _n = (np) / 8;                                  \
        if (_k == 2 && _n == 2) return FN<2, 2>(__VA_ARGS__);                    \
        if (_k == 2 && _n == 4) return FN<2, 4>(__VA_ARGS__);                    \
        if (_k == 2 && _n == 8) return FN<2, 8>(__VA_ARGS__);                    \
        if (_k == 4 && _n == 2) return FN<4, 2>(__VA_ARGS__);                    \
        if (_k == 4 && _n == 4) return FN<4, 4>(__VA_ARGS__);                    \
        if (_k == 4 && _n == 8) return FN<4, 8>(__VA_ARGS__);                    \
        if (_k == 8 && _n == 2) return FN<8, 2>(__VA_ARGS__);                    \
        if (_k == 8 && _n == 4) return FN<8, 4>(__VA_ARGS__);                    \
        if (_k == 8 && _n == 8) return FN<8, 8>(__VA_ARGS__);                    \
        return fail(RGCN_ERR_UNSUPPORTED, "no tile kernel for this padded shape"); \
    } while (0)

}  // namespace

int pad_dim(int f) {
    if (f <= 0) return 0;
    if (f <= 16) return 16;
    if (f <= 32) return 32;
    if (f <= 64) return 64;
    return 0;
}

int launch_chunk_prepass(const TilePass& p, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_chunks == 0) return 0;
    const int wpb = 8;
    const int grid = (b.num_chunks + wpb - 1) / wpb;
    const int relu = p.relu_in ? 1 : 0;
    ProfScope prof(TAG_PREPASS, p.kin, b.num_chunks, st);
    note_launch(1);
    switch (p.kp) {
        case 16:
            k_chunk_sum<16><<<grid, wpb * 32, 0, st>>>(b.raw_idx, b.raw_w, b.chunk_beg, b.chunk_end, b.num_chunks,
                                                       p.feat, p.ldf, p.kin, p.n_nodes, relu, p.aux);
            break;
        case 32:
            k_chunk_sum<32><<<grid, wpb * 32, 0, st>>>(b.raw_idx, b.raw_w, b.chunk_beg, b.chunk_end, b.num_chunks,
                                                       p.feat, p.ldf, p.kin, p.n_nodes, relu, p.aux);
            break;
        case 64:
            k_chunk_sum<64><<<grid, wpb * 32, 0, st>>>(b.raw_idx, b.raw_w, b.chunk_beg, b.chunk_end, b.num_chunks,
                                                       p.feat, p.ldf, p.kin, p.n_nodes, relu, p.aux);
            break;
        default: return fail(RGCN_ERR_UNSUPPORTED, "chunk pre-pass: unsupported padded width");
    }
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_wprep(const WPrep& p, cudaStream_t st) {
    const int KT = p.kp / 8, NT = p.np / 8;
    const int64_t total = (int64_t)(p.R + 1) * KT * NT * 32;
    const int tpb = 256;
    ProfScope prof(TAG_WPREP, p.fin, p.fout, st);
    note_launch(1);
    k_wprep<<<(int)((total + tpb - 1) / tpb), tpb, 0, st>>>(p.weight, p.root, p.R, p.fin, p.fout, KT, NT,
                                                            p.transpose ? 1 : 0, p.wfrag);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_tile_pass(const TilePass& p, int num_sms, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_batches == 0) return 0;
    TileArgs a;
    a.e_idx = b.e_idx;
    a.e_w = b.e_w;
    a.seg_ptr = b.seg_ptr;
    a.seg_own = b.seg_own;
    a.bat_seg0 = b.bat_seg0;
    a.bat_info = b.bat_info;
    a.num_batches = b.num_batches;
    a.feat = p.feat;
    a.ldf = p.ldf;
    a.kin = p.kin;
    a.aux = p.aux;
    a.n_nodes = p.n_nodes;
    a.wfrag = p.wfrag;
    a.bias = p.bias;
    a.nbias = p.nbias;
    a.self_rel = p.self_rel;
    a.out = p.out;
    a.ldo = p.ldo;
    a.nout = p.nout;
    a.relu_in = p.relu_in ? 1 : 0;
    ProfScope prof(p.transposed ? TAG_TILE_BWD : TAG_TILE_FWD, p.kin, p.nout, st);
    note_launch(1);
    RGCN_DISPATCH_KN(run_tile, p.kp, p.np, a, num_sms, st);
}

int launch_wgrad_pass(const WGradPass& p, int num_sms, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_batches == 0) return 0;
    WGradArgs a;
    a.e_idx = b.e_idx;
    a.e_w = b.e_w;
    a.seg_ptr = b.seg_ptr;
    a.seg_own = b.seg_own;
    a.bat_seg0 = b.bat_seg0;
    a.bat_info = b.bat_info;
    a.num_batches = b.num_batches;
    a.feat = p.feat;
    a.ldf = p.ldf;
    a.kin = p.kin;
    a.aux = p.aux;
    a.n_nodes = p.n_nodes;
    a.gout = p.gout;
    a.ldg = p.ldg;
    a.nout = p.nout;
    a.self_rel = p.self_rel;
    a.gweight = p.gweight;
    a.groot = p.groot;
    a.gbias = p.gbias;
    a.relu_in = p.relu_in ? 1 : 0;
    ProfScope prof(TAG_WGRAD, p.kin, p.nout, st);
    note_launch(1);
    RGCN_DISPATCH_KN(run_wgrad, p.kp, p.np, a, num_sms, st);
}

int launch_copy_cols(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n, int cols, cudaStream_t st) {
    const int64_t total = n * cols;
    if (total == 0) return 0;
    ProfScope prof(TAG_COPY, cols, 0, st);
    note_launch(1);
    k_copy_cols<<<(int)((total + 255) / 256), 256, 0, st>>>(src, lds, dst, ldd, n, cols);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_relu_mask(float* g, int64_t ldg, const float* pre, int64_t ldp, int64_t n, int cols, cudaStream_t st) {
    const int64_t total = n * cols;
    if (total == 0) return 0;
    ProfScope prof(TAG_MASK, cols, 0, st);
    note_launch(1);
    k_relu_mask<<<(int)((total + 255) / 256), 256, 0, st>>>(g, ldg, pre, ldp, n, cols);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rgcn
