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
#include <cstdlib>

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
// weights: round-to-nearest split, done once per call in k_wprep
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
// activations (inner loops): hi = x with the 13 low mantissa bits cleared, lo = x - hi (exact).
// mma .tf32 reads only the upper 19 bits of each operand, so lo needs no further rounding.
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 4-byte async copy global -> shared; src_bytes = 0 zero-fills (rows without an owner)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// ---- streaming gather ---------------------------------------------------------------------------
// A warp walks a CONTIGUOUS run of entries (its batches are consecutive segments, and entries are
// stored in segment order).  Entries are taken 32 at a time (one coalesced index/weight load,
// prefetched one block ahead), rows 8 at a time: 16 broadcasts, then 16 independent row loads
// issued back to back (lane l reads columns l and l+32), then 8 weighted accumulations.
// Slots past the end of the run carry index 0 / weight 0: they load row 0 and change nothing,
// so the inner loops carry only the two per-lane column predicates.
constexpr int RING = 32;   // smem ring of finished segment rows per warp (>= 16 + 8)

struct LanePtrs {   // what a lane needs to address its columns of a gathered row
    const float* feat;   // feature matrix (row r at feat + r * ldf)
    const float* aux;    // chunk rows (row r - n_rows at aux + (r - n_rows) * KP)
    uint32_t ldf, n_rows;
    uint32_t col;        // this lane's first column (clamped to the logical width)
    bool c1;             // lane also owns column lane + 32
};
template <int KP>
__device__ __forceinline__ LanePtrs lane_ptrs(const float* feat, int64_t ldf, int kin, const float* aux,
                                              int64_t n_rows, int lane) {
    LanePtrs q;
    q.feat = feat;
    q.aux = aux ? aux : feat;
    q.ldf = (uint32_t)ldf;
    q.n_rows = (uint32_t)n_rows;
    q.col = (uint32_t)min(lane, kin - 1);   // lanes past the width re-read the last column; their
    q.c1 = (KP > 32) && (lane + 32 < kin);  // sums land in pad columns that meet zero rows of B
    return q;
}

template <int KP, bool RELU>
struct RowGroup {
    float v0[GATHER_U], v1[GATHER_U];
    uint32_t raw[GATHER_U];
    // CH: the block contains chunk rows (index >= n_rows) -> per-entry select; rare.
    // Element offsets are 32-bit (the host checks rows * ld < 2^32): one IMAD + one LEA pair per
    // entry, shared by both column halves (the second half is an immediate +128 B).
    template <bool CH>
    __device__ __forceinline__ void load(uint32_t blk_idx, int j0, const LanePtrs& q) {
#pragma unroll
        for (int u = 0; u < GATHER_U; ++u) raw[u] = __shfl_sync(FULL, blk_idx, j0 + u);
        const float* p[GATHER_U];
#pragma unroll
        for (int u = 0; u < GATHER_U; ++u) {
            const uint32_t row = raw[u] & IDX_MASK;
            if (CH && row >= q.n_rows) p[u] = q.aux + ((row - q.n_rows) * (uint32_t)KP + q.col);
            else p[u] = q.feat + (row * q.ldf + q.col);
        }
#pragma unroll
        for (int u = 0; u < GATHER_U; ++u) {
            v0[u] = __ldg(p[u]);
            v1[u] = 0.f;
            if (KP > 32) {
                if (q.c1) v1[u] = __ldg(p[u] + 32);
            }
        }
        if (RELU) {
#pragma unroll
            for (int u = 0; u < GATHER_U; ++u) {
                if (!CH || (raw[u] & IDX_MASK) < q.n_rows) {   // chunk rows are sums of rectified rows already
                    v0[u] = fmaxf(v0[u], 0.f);
                    v1[u] = fmaxf(v1[u], 0.f);
                }
            }
        }
    }
    __device__ __forceinline__ void load_any(uint32_t blk_idx, int j0, const LanePtrs& q, bool has_chunk) {
        if (has_chunk) load<true>(blk_idx, j0, q);
        else load<false>(blk_idx, j0, q);
    }
};

// Entry stream of one warp: block prefetch + the ring of finished segment rows.
template <int KP, int LDH>
struct Stream {
    uint32_t nxt_idx, blk_idx;
    float nxt_w, blk_w;
    int ring_w;
    float acc0, acc1;
    bool has_chunk;
    __device__ __forceinline__ void start(const uint32_t* e_idx, const float* e_w, int E0, int E1, int lane) {
        nxt_idx = 0;
        nxt_w = 0.f;
        if (E0 + lane < E1) {
            nxt_idx = e_idx[E0 + lane];
            nxt_w = e_w[E0 + lane];
        }
        ring_w = 0;
        acc0 = acc1 = 0.f;
    }
    __device__ __forceinline__ void next_block(const uint32_t* e_idx, const float* e_w, int pos, int E1, int lane,
                                               uint32_t n_rows) {
        blk_idx = nxt_idx;
        blk_w = nxt_w;
        nxt_idx = 0;
        nxt_w = 0.f;
        if (pos + 32 + lane < E1) {
            nxt_idx = e_idx[pos + 32 + lane];
            nxt_w = e_w[pos + 32 + lane];
        }
        has_chunk = __any_sync(FULL, (blk_idx & IDX_MASK) >= n_rows);
    }
    // branch-free: the running sums are stored on every entry and kept or cleared by the LAST flag
    template <bool RELU>
    __device__ __forceinline__ void consume(const RowGroup<KP, RELU>& rg, float wreg, int j0, float* Hlane, int lane) {
        float w[GATHER_U];
#pragma unroll
        for (int u = 0; u < GATHER_U; ++u) w[u] = __shfl_sync(FULL, wreg, j0 + u);
#pragma unroll
        for (int u = 0; u < GATHER_U; ++u) {
            acc0 = fmaf(w[u], rg.v0[u], acc0);
            acc1 = fmaf(w[u], rg.v1[u], acc1);
            float* hr = Hlane + (ring_w & (RING - 1)) * LDH;   // Hlane = ring base + lane
            if (KP >= 32 || lane < KP) hr[0] = acc0;
            if (KP > 32) hr[32] = acc1;
            const uint32_t last = rg.raw[u] >> 31;
            ring_w += (int)last;
            acc0 = last ? 0.f : acc0;
            acc1 = last ? 0.f : acc1;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// chunk pre-pass: aux[c] = sum_{k in chunk c} raw_w[k] * feat[raw_idx[k]]   (one warp per chunk)
// ---------------------------------------------------------------------------------------------
template <int KP, bool RELU>
__global__ void __launch_bounds__(256) k_chunk_sum(const int32_t* __restrict__ raw_idx, const float* __restrict__ raw_w,
                                                   const int32_t* __restrict__ chunk_beg,
                                                   const int32_t* __restrict__ chunk_end, int num_chunks,
                                                   const float* __restrict__ feat, int64_t ldf, int kin,
                                                   int64_t n_rows, float* __restrict__ aux,
                                                   const int32_t* __restrict__ chunk_out) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= num_chunks) return;
    const int64_t row = chunk_out ? chunk_out[c] : c;   // FWD_REL writes FWD's numbering
    const LanePtrs q = lane_ptrs<KP>(feat, ldf, kin, nullptr, n_rows, lane);
    const int e0 = chunk_beg[c], e1 = chunk_end[c];
    float acc0 = 0.f, acc1 = 0.f;
    for (int pos = e0; pos < e1; pos += 32) {
        uint32_t bi = 0;
        float bw = 0.f;
        if (pos + lane < e1) {
            bi = (uint32_t)raw_idx[pos + lane];
            bw = raw_w[pos + lane];
        }
        const int m = min(32, e1 - pos);
        for (int j0 = 0; j0 < m; j0 += GATHER_U) {
            RowGroup<KP, RELU> rg;
            rg.template load<false>(bi, j0, q);
#pragma unroll
            for (int u = 0; u < GATHER_U; ++u) {
                const float w = __shfl_sync(FULL, bw, j0 + u);
                acc0 = fmaf(w, rg.v0[u], acc0);
                acc1 = fmaf(w, rg.v1[u], acc1);
            }
        }
    }
    if (lane < min(KP, kin)) aux[row * KP + lane] = acc0;   // pad columns stay zero
    else if (lane < KP) aux[row * KP + lane] = 0.f;
    if (KP > 32) aux[row * KP + lane + 32] = acc1;          // v1 is zero past the width
}

// 16-column rows that are 16-byte addressable: 4 lanes per row, 8 rows per load instruction (the
// column-per-lane kernel above leaves half the warp idle and loads one row per instruction)
template <bool RELU>
__global__ void __launch_bounds__(256) k_chunk_sum16(const int32_t* __restrict__ raw_idx, const float* __restrict__ raw_w,
                                                     const int32_t* __restrict__ chunk_beg,
                                                     const int32_t* __restrict__ chunk_end, int num_chunks,
                                                     const float* __restrict__ feat, int64_t ldf, int kin,
                                                     float* __restrict__ aux, const int32_t* __restrict__ chunk_out) {
    const int lane = threadIdx.x & 31, q = lane & 3, r = lane >> 2;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= num_chunks) return;
    const int64_t row = chunk_out ? chunk_out[c] : c;
    const int e0 = chunk_beg[c], e1 = chunk_end[c];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool col_ok = 4 * q < kin;   // kin is a multiple of 4 here (whole quads inside the row)
    for (int pos = e0 + r; pos < e1; pos += 8) {
        const uint32_t idx = (uint32_t)raw_idx[pos];
        const float w = raw_w[pos];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) v = __ldg(reinterpret_cast<const float4*>(feat + (uint64_t)idx * (uint64_t)ldf) + q);
        if (RELU) {
            v.x = fmaxf(v.x, 0.f);
            v.y = fmaxf(v.y, 0.f);
            v.z = fmaxf(v.z, 0.f);
            v.w = fmaxf(v.w, 0.f);
        }
        acc.x = fmaf(w, v.x, acc.x);
        acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z);
        acc.w = fmaf(w, v.w, acc.w);
    }
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
        acc.x += __shfl_xor_sync(FULL, acc.x, sh);
        acc.y += __shfl_xor_sync(FULL, acc.y, sh);
        acc.z += __shfl_xor_sync(FULL, acc.z, sh);
        acc.w += __shfl_xor_sync(FULL, acc.w, sh);
    }
    if (r == 0) reinterpret_cast<float4*>(aux + row * 16)[q] = acc;
}

// ---------------------------------------------------------------------------------------------
// B-operand preparation: fragment-ordered, pre-split into tf32 hi/lo.
//   wfrag[((rel*KT + kt)*NT + nt)*32 + lane] = (b0.hi, b1.hi, b0.lo, b1.lo)
//   b0 = B[8kt + lane%4][8nt + lane/4], b1 = B[8kt + lane%4 + 4][8nt + lane/4]
//   B = W_rel (rows<fin, cols<fout) or its transpose; rel == R is the root matrix.
// ---------------------------------------------------------------------------------------------
__global__ void k_wprep(const float* __restrict__ weight, const float* __restrict__ root, int R, int fin, int fout,
                        int KT, int NT, int transpose, int perm, float4* __restrict__ wfrag,
                        float2* __restrict__ wfrag2) {
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
        // K slot (t + 4h) of step kt: natural order, or the vector-load order 16(kt/2) + 4t + 2(kt%2) + h
        const int k = perm ? 16 * (kt >> 1) + 4 * t + 2 * (kt & 1) + h : 8 * kt + t + 4 * h;
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
    if (wfrag2) wfrag2[i] = make_float2(b[0], b[1]);
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
    int num_entries;
    const int4* units;
    int num_units;
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

template <int KT, int NT, bool RELU, bool BREG>
__global__ void __launch_bounds__(TILE_WARPS * 32, BREG ? 1 : ((KT == 8 || NT > 2) ? 2 : 3)) k_tile(const TileArgs a) {
    constexpr int KP = KT * 8, LDH = KP + 4;
    constexpr bool PIPE = (KT == 8);   // wide rows: double-buffered groups (more bytes in flight per warp)
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* H = smem + warp * (RING * LDH);
    const int g = lane >> 2, t = lane & 3;
    const int gw = blockIdx.x * TILE_WARPS + warp;
    const int nw = gridDim.x * TILE_WARPS;
    const uint32_t* __restrict__ e_idx = a.e_idx;
    const float* __restrict__ e_w = a.e_w;
    const LanePtrs q = lane_ptrs<KP>(a.feat, a.ldf, a.kin, a.aux, a.n_nodes, lane);
    float4 bfrag[BREG ? KT * NT : 1];
    int cur_rel = -1;

    for (int unit = gw; unit < a.num_units; unit += nw) {   // equal-cost units, round-robin over warps
        const int4 u0 = a.units[unit], u1 = a.units[unit + 1];
        int b = u0.x;
        const int b_end = u1.x;
        if (b >= b_end) continue;
        int seg0 = u0.y;
        const int E0 = u0.z, E1 = u1.z;
        int info = a.bat_info[b];
        int nseg = info & 0xff, rel = info >> 8;
        int my_own = lane < nseg ? a.seg_own[seg0 + lane] : -1;
        int info_n = b + 1 < b_end ? a.bat_info[b + 1] : 0;
        int own_n = (b + 1 < b_end && lane < (info_n & 0xff)) ? a.seg_own[seg0 + nseg + lane] : -1;
        int ring_base = 0;
        Stream<KP, LDH> sm;
        sm.start(e_idx, e_w, E0, E1, lane);
        sm.next_block(e_idx, e_w, E0, E1, lane, q.n_rows);
        // drain every batch whose rows are all in the ring
        auto drain = [&]() {
        while (b < b_end && sm.ring_w - ring_base >= nseg) {
            __syncwarp();
            const float4* wf = a.wfrag + (int64_t)rel * (KT * NT * 32) + lane;
            if constexpr (BREG) {
                if (rel != cur_rel) {
#pragma unroll
                    for (int i = 0; i < KT * NT; ++i) bfrag[i] = __ldg(wf + i * 32);
                    cur_rel = rel;
                }
            }
            const float* h_lo = H + ((ring_base + g) & (RING - 1)) * LDH;
            const float* h_hi = H + ((ring_base + g + 8) & (RING - 1)) * LDH;
            float d[NT][4];
#pragma unroll
            for (int n = 0; n < NT; ++n) d[n][0] = d[n][1] = d[n][2] = d[n][3] = 0.f;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                uint32_t ah[4], al[4];
                split_fast(h_lo[8 * kt + t], ah[0], al[0]);
                split_fast(h_hi[8 * kt + t], ah[1], al[1]);
                split_fast(h_lo[8 * kt + t + 4], ah[2], al[2]);
                split_fast(h_hi[8 * kt + t + 4], ah[3], al[3]);
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
            // rows past nseg belong to later batches: their owners read -1 and are skipped
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
            // next batch: metadata was prefetched one batch ago
            ring_base += nseg;
            seg0 += nseg;
            ++b;
            info = info_n;
            nseg = info & 0xff;
            rel = info >> 8;
            my_own = own_n;
            if (b < b_end) {
                info_n = b + 1 < b_end ? a.bat_info[b + 1] : 0;
                own_n = (b + 1 < b_end && lane < (info_n & 0xff)) ? a.seg_own[seg0 + nseg + lane] : -1;
            }
        }
        };
        // Software pipeline over groups of 8 entries: the loads of group k+1 are in flight while
        // group k is accumulated and its finished batches go through the tensor pipe.
        if constexpr (PIPE) {
            RowGroup<KP, RELU> A, B;
            A.load_any(sm.blk_idx, 0, q, sm.has_chunk);
            for (int pos = E0; pos < E1; pos += 32) {
                const int m = min(32, E1 - pos);
    #pragma unroll 1
                for (int h = 0; h < 2; ++h) {   // two (A, B) group pairs per 32-entry block
                    const int ja = 16 * h, jb = ja + 8;
                    if (jb < m) B.load_any(sm.blk_idx, jb, q, sm.has_chunk);
                    if (ja < m) {
                        sm.consume(A, sm.blk_w, ja, H + lane, lane);
                        drain();
                    }
                    const float w_cur = sm.blk_w;
                    if (h == 0) {
                        if (16 < m) A.load_any(sm.blk_idx, 16, q, sm.has_chunk);
                    } else if (pos + 32 < E1) {
                        sm.next_block(e_idx, e_w, pos + 32, E1, lane, q.n_rows);
                        A.load_any(sm.blk_idx, 0, q, sm.has_chunk);
                    }
                    if (jb < m) {
                        sm.consume(B, w_cur, jb, H + lane, lane);
                        drain();
                    }
                }
            }

        } else {
            for (int pos = E0; pos < E1; pos += 32) {
                if (pos > E0) sm.next_block(e_idx, e_w, pos, E1, lane, q.n_rows);
                const int m = min(32, E1 - pos);
                for (int j0 = 0; j0 < m; j0 += GATHER_U) {
                    RowGroup<KP, RELU> rg;
                    rg.load_any(sm.blk_idx, j0, q, sm.has_chunk);
                    sm.consume(rg, sm.blk_w, j0, H + lane, lane);
                    drain();
                }
            }
        }
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
    int num_entries;
    const int4* units;
    int num_units;
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

template <int KT, int NT, bool RELU>
__global__ void __launch_bounds__(WG_WARPS * 32, (KT * NT <= 16) ? 4 : ((KT * NT <= 32) ? 2 : 1)) k_wgrad(const WGradArgs a) {
    constexpr int KP = KT * 8, NP = NT * 8, MT = KP / 16;
    constexpr int LDH = KP + 8, LDG = NP + 8;
    constexpr int GR = BS * NP / 32;   // registers holding a batch's 16 gout rows
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* H = smem + warp * (RING * LDH + 2 * BS * LDG);
    float* Gbuf = H + RING * LDH;   // two buffers of 16 gout rows, filled by cp.async one batch ahead
    const int g = lane >> 2, t = lane & 3;
    const int gw = blockIdx.x * WG_WARPS + warp;
    const int nw = gridDim.x * WG_WARPS;
    // ring rows are the K dimension here: anything ever read must be finite
    for (int i = lane; i < RING * LDH; i += 32) H[i] = 0.f;
    __syncwarp();
    const uint32_t* __restrict__ e_idx = a.e_idx;
    const float* __restrict__ e_w = a.e_w;
    const float* __restrict__ gout = a.gout;
    const LanePtrs q = lane_ptrs<KP>(a.feat, a.ldf, a.kin, a.aux, a.n_nodes, lane);

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
    // the 16 gout rows of a batch, requested (async, straight into shared memory) one batch ahead
    auto fetch_g = [&](float* dstbuf, int own) {
#pragma unroll
        for (int i = 0; i < GR; ++i) {
            const int e = i * 32 + lane;
            const int row = e / NP, col = e % NP;
            const int o = __shfl_sync(FULL, own, row);
            const bool ok = o >= 0 && col < a.nout;
            cp_async4(dstbuf + row * LDG + col, ok ? gout + (int64_t)o * a.ldg + col : gout, ok ? 4 : 0);
        }
        cp_async_commit();
    };

    for (int unit = gw; unit < a.num_units; unit += nw) {
        const int4 u0 = a.units[unit], u1 = a.units[unit + 1];
        int b = u0.x;
        const int b_end = u1.x;
        if (b >= b_end) continue;
        int seg0 = u0.y;
        const int E0 = u0.z, E1 = u1.z;
        int info = a.bat_info[b];
        int nseg = info & 0xff, rel = info >> 8;
        int my_own = lane < nseg ? a.seg_own[seg0 + lane] : -1;
        int info_n = b + 1 < b_end ? a.bat_info[b + 1] : 0;
        int own_n = (b + 1 < b_end && lane < (info_n & 0xff)) ? a.seg_own[seg0 + nseg + lane] : -1;
        int par = 0;
        cp_async_wait<0>();   // nothing of the previous unit may still be landing in the G buffers
        __syncwarp();
        fetch_g(Gbuf, my_own);
        fetch_g(Gbuf + BS * LDG, own_n);
        int ring_base = 0;
        Stream<KP, LDH> sm;
        sm.start(e_idx, e_w, E0, E1, lane);
        sm.next_block(e_idx, e_w, E0, E1, lane, q.n_rows);
        auto drain = [&]() {
        while (b < b_end && sm.ring_w - ring_base >= nseg) {
            if (rel != cur_rel) {
                if (cur_rel >= 0) flush(cur_rel);
                cur_rel = rel;
            }
            const bool wanted =
                rel == a.self_rel ? (a.groot != nullptr || a.gbias != nullptr) : a.gweight != nullptr;
            cp_async_wait<1>();   // this batch's rows have landed; the next batch's may still fly
            __syncwarp();
            const float* G = Gbuf + par * (BS * LDG);
            if (wanted) {
                if (rel == a.self_rel) {
#pragma unroll
                    for (int i = 0; i < (NP + 31) / 32; ++i) {
                        const int c = lane + 32 * i;
                        if (c < NP) {
                            float sacc = 0.f;
#pragma unroll
                            for (int r = 0; r < BS; ++r) sacc += G[r * LDG + c];
                            bsum[i] += sacc;
                        }
                    }
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        split_fast(G[(8 * ks + t) * LDG + 8 * n + g], bh[n][0], bl[n][0]);
                        split_fast(G[(8 * ks + t + 4) * LDG + 8 * n + g], bh[n][1], bl[n][1]);
                    }
                    const float* r0 = H + ((ring_base + 8 * ks + t) & (RING - 1)) * LDH;
                    const float* r1 = H + ((ring_base + 8 * ks + t + 4) & (RING - 1)) * LDH;
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        uint32_t ah[4], al[4];
                        split_fast(r0[16 * m + g], ah[0], al[0]);
                        split_fast(r0[16 * m + g + 8], ah[1], al[1]);
                        split_fast(r1[16 * m + g], ah[2], al[2]);
                        split_fast(r1[16 * m + g + 8], ah[3], al[3]);
#pragma unroll
                        for (int n = 0; n < NT; ++n) {
                            mma_tf32(d[m][n], al[0], al[1], al[2], al[3], bh[n][0], bh[n][1]);
                            mma_tf32(d[m][n], ah[0], ah[1], ah[2], ah[3], bl[n][0], bl[n][1]);
                            mma_tf32(d[m][n], ah[0], ah[1], ah[2], ah[3], bh[n][0], bh[n][1]);
                        }
                    }
                }
            }
            __syncwarp();
            ring_base += nseg;
            seg0 += nseg;
            ++b;
            info = info_n;
            nseg = info & 0xff;
            rel = info >> 8;
            my_own = own_n;
            info_n = 0;
            own_n = -1;
            if (b + 1 < b_end) {
                info_n = a.bat_info[b + 1];
                own_n = lane < (info_n & 0xff) ? a.seg_own[seg0 + nseg + lane] : -1;
            }
            fetch_g(Gbuf + par * (BS * LDG), own_n);   // refill the buffer just consumed (zeros past the end)
            par ^= 1;
        }
        };
        for (int pos = E0; pos < E1; pos += 32) {
            if (pos > E0) sm.next_block(e_idx, e_w, pos, E1, lane, q.n_rows);
            const int m = min(32, E1 - pos);
            for (int j0 = 0; j0 < m; j0 += GATHER_U) {
                RowGroup<KP, RELU> rg;
                rg.load_any(sm.blk_idx, j0, q, sm.has_chunk);
                sm.consume(rg, sm.blk_w, j0, H + lane, lane);
                drain();
            }
        }
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

// dst[r][0:cols] = src[r][0:cols]  (one warp per row: both sides coalesced)
__global__ void __launch_bounds__(256) k_copy_cols(const float* __restrict__ src, int64_t lds, float* __restrict__ dst,
                                                   int64_t ldd, int64_t n, int cols) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    for (int64_t r = wid; r < n; r += nw) {
        const float* sp = src + r * lds;
        float* dp = dst + r * ldd;
        for (int c = lane; c < cols; c += 32) dp[c] = __ldg(sp + c);
    }
}

// dst[r][c] = c < cols ? src[r][c] : 0 for c < ldd  (16-byte aligned mirror of an odd-width matrix)
__global__ void __launch_bounds__(256) k_pad_rows(const float* __restrict__ src, int64_t lds, int cols,
                                                  float* __restrict__ dst, int64_t ldd, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    for (int64_t r0 = wid * 4; r0 < n; r0 += nw * 4) {
        float v[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                v[i][j] = (r0 + i < n && c < cols) ? __ldg(src + (r0 + i) * lds + c) : 0.f;
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                if (r0 + i < n && c < (int)ldd) dst[(r0 + i) * ldd + c] = v[i][j];
            }
        for (int c = lane + 128; c < (int)ldd; c += 32)
            for (int i = 0; i < 4; ++i)
                if (r0 + i < n) dst[(r0 + i) * ldd + c] = c < cols ? __ldg(src + (r0 + i) * lds + c) : 0.f;
    }
}

__global__ void k_relu_mask(float* __restrict__ gr, int64_t ldg, const float* __restrict__ pre, int64_t ldp, int64_t n,
                            int cols) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * cols) return;
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    if (!(pre[r * ldp + c] > 0.f)) gr[r * ldg + c] = 0.f;
}

// rows of whole aligned quads (the 16-wide hidden layer): one float4 of each matrix per thread
__global__ void __launch_bounds__(256) k_relu_mask4(float4* __restrict__ gr, int64_t ldg4, const float4* __restrict__ pre,
                                                    int64_t ldp4, int64_t n, int cols4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, total = n * cols4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / cols4;
        const int c = (int)(i - r * cols4);
        const float4 p = __ldg(pre + r * ldp4 + c);
        float4 g = gr[r * ldg4 + c];
        g.x = p.x > 0.f ? g.x : 0.f;
        g.y = p.y > 0.f ? g.y : 0.f;
        g.z = p.z > 0.f ? g.z : 0.f;
        g.w = p.w > 0.f ? g.w : 0.f;
        gr[r * ldg4 + c] = g;
    }
}

template <int KT, int NT, bool RELU, bool BREG>
int run_tile_v(const TileArgs& a, int num_sms, cudaStream_t st) {
    constexpr int LDH = KT * 8 + 4;
    const size_t smem = (size_t)TILE_WARPS * RING * LDH * sizeof(float);
    static bool configured = false;
    auto kern = k_tile<KT, NT, RELU, BREG>;
    if (!configured) {
        RGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int per_sm = 1;
    RGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILE_WARPS * 32, smem));
    per_sm = std::max(per_sm, 1);
    // a few batches per warp at least, so the streaming prefetch has something to stream
    int64_t want = ((int64_t)a.num_units + TILE_WARPS - 1) / TILE_WARPS;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * per_sm));
    kern<<<grid, TILE_WARPS * 32, smem, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

bool breg_enabled() {   // RGCN_B200_BREG=1 keeps the B fragments of the current relation in registers
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RGCN_B200_BREG");
        v = (e && e[0] == '1') ? 1 : 0;   // default off: the extra 64 registers cost more warps than they save loads
    }
    return v == 1;
}

template <int KT, int NT>
int run_tile(const TileArgs& a, int num_sms, cudaStream_t st) {
    const bool breg = (KT * NT <= 16) && breg_enabled();
    if (a.relu_in) {
        if constexpr (KT * NT <= 16) {
            if (breg) return run_tile_v<KT, NT, true, true>(a, num_sms, st);
        }
        return run_tile_v<KT, NT, true, false>(a, num_sms, st);
    }
    if constexpr (KT * NT <= 16) {
        if (breg) return run_tile_v<KT, NT, false, true>(a, num_sms, st);
    }
    return run_tile_v<KT, NT, false, false>(a, num_sms, st);
}

template <int KT, int NT, bool RELU>
int run_wgrad_v(const WGradArgs& a, int num_sms, cudaStream_t st) {
    constexpr int LDH = KT * 8 + 8, LDG = NT * 8 + 8;
    const size_t smem = (size_t)WG_WARPS * (RING * LDH + 2 * BS * LDG) * sizeof(float);
    static bool configured = false;
    auto kern = k_wgrad<KT, NT, RELU>;
    if (!configured) {
        RGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int per_sm = 1;
    RGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WG_WARPS * 32, smem));
    per_sm = std::max(per_sm, 1);
    int64_t want = ((int64_t)a.num_units + WG_WARPS - 1) / WG_WARPS;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * per_sm));
    kern<<<grid, WG_WARPS * 32, smem, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

template <int KT, int NT>
int run_wgrad(const WGradArgs& a, int num_sms, cudaStream_t st) {
    if (a.relu_in) return run_wgrad_v<KT, NT, true>(a, num_sms, st);
    return run_wgrad_v<KT, NT, false>(a, num_sms, st);
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

namespace {
template <int KP>
void run_chunk(const Brc& b, const TilePass& p, int grid, int wpb, cudaStream_t st) {
    if (p.relu_in)
        k_chunk_sum<KP, true><<<grid, wpb * 32, 0, st>>>(b.raw_idx, b.raw_w, b.chunk_beg, b.chunk_end, b.num_chunks, p.feat,
                                                         p.ldf, p.kin, p.n_nodes, p.aux, b.chunk_out);
    else
        k_chunk_sum<KP, false><<<grid, wpb * 32, 0, st>>>(b.raw_idx, b.raw_w, b.chunk_beg, b.chunk_end, b.num_chunks, p.feat,
                                                          p.ldf, p.kin, p.n_nodes, p.aux, b.chunk_out);
}
}  // namespace

int launch_chunk_prepass(const TilePass& p, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_chunks == 0) return 0;
    const int wpb = 8;
    const int grid = (b.num_chunks + wpb - 1) / wpb;
    ProfScope prof(TAG_PREPASS, p.kin, b.num_chunks, st);
    note_launch(1);
    // (kin < 16 with whole quads, e.g. the 12 used columns of a caller-padded 11-wide row, is fine too)
    const bool quad16 = p.kp == 16 && p.ldf % 4 == 0 && p.kin % 4 == 0 && ((uintptr_t)p.feat & 15) == 0 &&
                        ((uintptr_t)p.aux & 15) == 0;
    if (quad16) {
        if (p.relu_in)
            k_chunk_sum16<true><<<grid, wpb * 32, 0, st>>>(b.raw_idx, b.raw_w, b.chunk_beg, b.chunk_end, b.num_chunks, p.feat,
                                                           p.ldf, p.kin, p.aux, b.chunk_out);
        else
            k_chunk_sum16<false><<<grid, wpb * 32, 0, st>>>(b.raw_idx, b.raw_w, b.chunk_beg, b.chunk_end, b.num_chunks, p.feat,
                                                            p.ldf, p.kin, p.aux, b.chunk_out);
        RGCN_CUDA(cudaGetLastError());
        return 0;
    }
    switch (p.kp) {
        case 16: run_chunk<16>(b, p, grid, wpb, st); break;
        case 32: run_chunk<32>(b, p, grid, wpb, st); break;
        case 64: run_chunk<64>(b, p, grid, wpb, st); break;
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
                                                            p.transpose ? 1 : 0, p.perm ? 1 : 0, p.wfrag, p.wfrag2);
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
    a.num_entries = (int)b.num_entries;
    a.units = b.units;
    a.num_units = b.num_units;
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
    ProfScope prof(p.transposed ? TAG_TILE_BWD : TAG_TILE_FWD, p.kin, p.tag_out, st);
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
    a.num_entries = (int)b.num_entries;
    a.units = b.units;
    a.num_units = b.num_units;
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
    k_copy_cols<<<(int)std::min<int64_t>((n + 7) / 8, 148 * 16), 256, 0, st>>>(src, lds, dst, ldd, n, cols);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_pad_rows(const float* src, int64_t lds, int cols, float* dst, int64_t ldd, int64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    ProfScope prof(TAG_COPY, cols, (int)ldd, st);
    note_launch(1);
    k_pad_rows<<<(int)std::min<int64_t>((n + 31) / 32, 148 * 8), 256, 0, st>>>(src, lds, cols, dst, ldd, n);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_relu_mask(float* g, int64_t ldg, const float* pre, int64_t ldp, int64_t n, int cols, cudaStream_t st) {
    const int64_t total = n * cols;
    if (total == 0) return 0;
    ProfScope prof(TAG_MASK, cols, 0, st);
    note_launch(1);
    if (cols % 4 == 0 && ldg % 4 == 0 && ldp % 4 == 0 && (((uintptr_t)g | (uintptr_t)pre) & 15) == 0) {
        const int64_t quads = total / 4;
        const int grid = (int)std::min<int64_t>((quads + 255) / 256, 148 * 16);
        k_relu_mask4<<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(g), ldg / 4, reinterpret_cast<const float4*>(pre),
                                           ldp / 4, n, cols / 4);
        RGCN_CUDA(cudaGetLastError());
        return 0;
    }
    k_relu_mask<<<(int)((total + 255) / 256), 256, 0, st>>>(g, ldg, pre, ldp, n, cols);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rgcn
