// graph_build.cu — K0: one-time on-device build of the blocked relational CSRs (BRC).
//
// Consumes the reference's edge tensors as they are (graphs/graph.py:55-69: int64, strided
// views of one [E,3] buffer, inverse edges as relation 2r+1, multi-edges kept) and produces
// the relation-then-owner sorted structures every layer kernel reads.  Integer work only;
// checked bit-exactly against oracle/csr_oracle.py.  Device-wide sort/scan primitives are CUB
// (ships with the CUDA toolkit); everything else is hand-written.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace rgcn {

static thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

void Brc::release() { *this = Brc(); }   // (the arrays belong to the graph's arena)

namespace {

// arenas of the build in progress on this thread (set by graph_create_impl)
thread_local Arena* t_persist = nullptr;
thread_local Arena* t_scratch = nullptr;

template <typename T>
cudaError_t palloc(T** p, size_t bytes) {   // persistent array of the graph being built
    return t_persist->take(reinterpret_cast<void**>(p), bytes);
}

template <typename T>
struct Dev {   // scratch buffer from the build's scratch arena; released by the enclosing ScratchScope
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        n = count;
        return t_scratch->take(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(T));
    }
};
struct ScratchScope {   // every scratch buffer taken inside the scope is handed back at its end
    Arena::Mark m;
    ScratchScope() : m(t_scratch->mark()) {}
    ~ScratchScope() { t_scratch->rewind(m); }
};

constexpr int TPB = 256;
inline int blocks_for(int64_t n) { return (int)std::max<int64_t>(1, (n + TPB - 1) / TPB); }

// edges -> int32 (validated, global ids)
__global__ void k_edges(const int64_t* __restrict__ src, int64_t ss, const int64_t* __restrict__ dst, int64_t ds,
                        const int64_t* __restrict__ et, int64_t es, int64_t E, int64_t N, int R,
                        int32_t* __restrict__ src32, int32_t* __restrict__ dst32, int32_t* __restrict__ rel32,
                        int* __restrict__ err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= E) return;
    int64_t s = src[i * ss], d = dst[i * ds], r = et[i * es];
    if (s < 0 || s >= N || d < 0 || d >= N || r < 0 || r >= R) {
        atomicOr(err, 1);
        s = d = 0;
        r = 0;
    }
    src32[i] = (int32_t)s;
    dst32[i] = (int32_t)d;
    rel32[i] = (int32_t)r;
}

// per-edge mean normaliser 1/cnt(rel,dst): sort by rel*N+dst, run lengths, scatter back by edge id
__global__ void k_wkeys(const int32_t* __restrict__ dst, const int32_t* __restrict__ rel, int64_t E, int64_t N,
                        uint64_t* __restrict__ key, int32_t* __restrict__ eid) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= E) return;
    key[i] = (uint64_t)rel[i] * (uint64_t)N + (uint64_t)dst[i];
    eid[i] = (int32_t)i;
}
__global__ void k_run_starts(const int32_t* __restrict__ head, const int32_t* __restrict__ scan, int64_t n,
                             int32_t* __restrict__ start) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    start[scan[i] - 1] = (int32_t)i;
}
__global__ void k_edge_weights(const int32_t* __restrict__ perm, const int32_t* __restrict__ scan,
                               const int32_t* __restrict__ start, int64_t n, float* __restrict__ w_edge) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t s = scan[i] - 1;
    w_edge[perm[i]] = 1.0f / (float)(start[s + 1] - start[s]);
}

// entry list of one BRC of the partition [lo,hi): edges whose owner end lies in the range (kept in
// input order), then one self loop per owned node.  Owner ids local, gather ids global.
__global__ void k_flag_owned(const int32_t* __restrict__ own_g, int64_t E, int32_t lo, int32_t hi,
                             int32_t* __restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= E) return;
    flag[i] = (own_g[i] >= lo && own_g[i] < hi) ? 1 : 0;
}
__global__ void k_fill_entries(const int32_t* __restrict__ own_g, const int32_t* __restrict__ gat_g,
                               const int32_t* __restrict__ rel_g, const float* __restrict__ w_edge,
                               const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int64_t E, int64_t nsel,
                               int32_t lo, int32_t hi, int R, int push, int32_t* __restrict__ own,
                               int32_t* __restrict__ gat, int32_t* __restrict__ rel, float* __restrict__ w) {
    // pull (push == 0): the OWNER end is in [lo, hi) -> owner ids local, gather ids global
    // push            : the GATHER end is in [lo, hi) -> gather ids local, owner ids global
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < E) {
        if (flag[i]) {
            int32_t o = pos[i];
            own[o] = push ? own_g[i] : own_g[i] - lo;
            gat[o] = push ? gat_g[i] - lo : gat_g[i];
            rel[o] = rel_g[i];
            w[o] = w_edge[i];
        }
    } else if (i < E + (hi - lo)) {
        int32_t v = (int32_t)(i - E);
        int64_t o = nsel + v;
        own[o] = push ? lo + v : v;
        gat[o] = push ? v : lo + v;
        rel[o] = R;
        w[o] = 1.0f;
    }
}

// sort key: edges by (owner range, relation, owner); self loops (relation R) after every range, by owner
__global__ void k_keys(const int32_t* __restrict__ own, const int32_t* __restrict__ rel, int64_t n2, int R, int NR,
                       uint64_t self_base, uint64_t* __restrict__ key, int32_t* __restrict__ eid) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n2) return;
    uint64_t o = (uint64_t)own[i];
    key[i] = rel[i] == R ? self_base + o : (o / NR) * ((uint64_t)(R + 1) * NR) + (uint64_t)rel[i] * NR + (o % NR);
    eid[i] = (int32_t)i;
}

__global__ void k_heads(const uint64_t* __restrict__ skey, int64_t n2, int32_t* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n2) return;
    head[i] = (i == 0 || skey[i] != skey[i - 1]) ? 1 : 0;
}

// segid = inclusive_scan(head) - 1 ; at heads record the segment's start / owner / relation
__global__ void k_seg_fill(const uint64_t* __restrict__ skey, const int32_t* __restrict__ head,
                           const int32_t* __restrict__ scan, int64_t n2, int R, int NR, uint64_t self_base,
                           int32_t* __restrict__ seg_ptr0, int32_t* __restrict__ seg_own, int32_t* __restrict__ seg_rel) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n2 || !head[i]) return;
    int32_t s = scan[i] - 1;
    uint64_t k = skey[i];
    uint64_t span = (uint64_t)(R + 1) * NR;
    uint64_t rng = k / span, rem = k % span;
    seg_ptr0[s] = (int32_t)i;
    if (k >= self_base) {
        seg_rel[s] = R;
        seg_own[s] = (int32_t)(k - self_base);
    } else {
        seg_rel[s] = (int32_t)(rem / NR);
        seg_own[s] = (int32_t)(rng * NR + rem % NR);
    }
}

__global__ void k_raw(const int32_t* __restrict__ perm, const int32_t* __restrict__ gat,
                      const float* __restrict__ w_entry, int64_t n2, int32_t* __restrict__ raw_idx,
                      float* __restrict__ raw_w) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n2) return;
    int32_t e = perm[i];
    raw_idx[i] = gat[e];
    raw_w[i] = w_entry[e];
}

__global__ void k_seg_counts(const int32_t* __restrict__ seg_ptr0, int32_t S, int T, int CH,
                             int32_t* __restrict__ out_cnt, int32_t* __restrict__ nchunk) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    int32_t cnt = seg_ptr0[s + 1] - seg_ptr0[s];
    int32_t nc = cnt > T ? (cnt + CH - 1) / CH : 0;
    nchunk[s] = nc;
    out_cnt[s] = cnt > T ? nc : cnt;
}

__global__ void k_compact(const int32_t* __restrict__ scan, const int32_t* __restrict__ seg_ptr0,
                          const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ chunk_base,
                          const int32_t* __restrict__ raw_idx, const float* __restrict__ raw_w, int64_t n2, int64_t N,
                          int T, int CH, uint32_t* __restrict__ e_idx, float* __restrict__ e_w,
                          int32_t* __restrict__ chunk_beg, int32_t* __restrict__ chunk_end,
                          const int32_t* __restrict__ seg_own, int32_t* __restrict__ e_own) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n2) return;
    int32_t s = scan[i] - 1;
    int32_t a = seg_ptr0[s], b = seg_ptr0[s + 1];
    int32_t cnt = b - a, p = (int32_t)i - a;
    int32_t o = seg_ptr[s];
    if (cnt <= T) {
        e_idx[o + p] = (uint32_t)raw_idx[i] | (p == cnt - 1 ? LAST_FLAG : 0u);
        e_w[o + p] = raw_w[i];
        e_own[o + p] = seg_own[s];
    } else if (p % CH == 0) {
        int32_t j = p / CH, nc = (cnt + CH - 1) / CH;
        int32_t c = chunk_base[s] + j;
        chunk_beg[c] = (int32_t)i;
        chunk_end[c] = min((int32_t)i + CH, b);
        e_idx[o + j] = (uint32_t)(N + c) | (j == nc - 1 ? LAST_FLAG : 0u);
        e_w[o + j] = 1.0f;
        e_own[o + j] = seg_own[s];
    }
}

__global__ void k_group_heads(const int32_t* __restrict__ seg_own, const int32_t* __restrict__ seg_rel, int32_t S,
                              int NR, int32_t* __restrict__ ghead) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    int h = 1;
    if (s > 0) h = (seg_rel[s] != seg_rel[s - 1]) || (seg_own[s] / NR != seg_own[s - 1] / NR);
    ghead[s] = h;
}

__global__ void k_group_fill(const int32_t* __restrict__ ghead, const int32_t* __restrict__ gscan, int32_t S,
                             int32_t* __restrict__ grp_seg) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S || !ghead[s]) return;
    grp_seg[gscan[s] - 1] = (int32_t)s;
}

__global__ void k_group_nb(const int32_t* __restrict__ grp_seg, int32_t G, int32_t* __restrict__ nb) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    nb[g] = (grp_seg[g + 1] - grp_seg[g] + BS - 1) / BS;
}

__global__ void k_batch_fill(const int32_t* __restrict__ bat_base, const int32_t* __restrict__ grp_seg,
                             const int32_t* __restrict__ seg_rel, int32_t G, int32_t NB,
                             int32_t* __restrict__ bat_seg0, int32_t* __restrict__ bat_info) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= NB) return;
    // largest g with bat_base[g] <= b
    int lo = 0, hi = G - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (bat_base[mid] <= (int32_t)b) lo = mid;
        else hi = mid - 1;
    }
    int32_t s0 = grp_seg[lo] + ((int32_t)b - bat_base[lo]) * BS;
    int32_t ns = min(BS, grp_seg[lo + 1] - s0);
    bat_seg0[b] = s0;
    bat_info[b] = (seg_rel[s0] << 8) | ns;
}

__global__ void k_set_i32(int32_t* p, int32_t v) { *p = v; }

// entry tiles: each (range, relation) group's entries cut into runs of <= ET
__global__ void k_group_ntiles(const int32_t* __restrict__ grp_seg, const int32_t* __restrict__ seg_ptr, int32_t G,
                               int32_t* __restrict__ nt) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    nt[g] = (seg_ptr[grp_seg[g + 1]] - seg_ptr[grp_seg[g]] + ET - 1) / ET;
}
__global__ void k_tile_fill(const int32_t* __restrict__ tile_base, const int32_t* __restrict__ grp_seg,
                            const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_rel, int32_t G,
                            int32_t NT, int32_t* __restrict__ tile_e0, int32_t* __restrict__ tile_info) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= NT) return;
    int lo = 0, hi = G - 1;   // largest g with tile_base[g] <= i
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tile_base[mid] <= (int32_t)i) lo = mid;
        else hi = mid - 1;
    }
    const int32_t e_beg = seg_ptr[grp_seg[lo]], e_end = seg_ptr[grp_seg[lo + 1]];
    const int32_t e0 = e_beg + ((int32_t)i - tile_base[lo]) * ET;
    tile_e0[i] = e0;
    tile_info[i] = (seg_rel[grp_seg[lo]] << 8) | min(ET, e_end - e0);
}

// Work units: cut the batch list into NU pieces of equal cost c(b) = first_entry(b) + 24 b
// (entries + a per-batch overhead), so that warps taking units round-robin finish together.
__global__ void k_units(const int32_t* __restrict__ bat_seg0, const int32_t* __restrict__ seg_ptr, int32_t NB,
                        int32_t S, int32_t E3, int32_t NU, int4* __restrict__ units) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u > NU) return;
    if (u == NU) {
        units[u] = make_int4(NB, S, E3, 0);
        return;
    }
    const int64_t total = (int64_t)E3 + (int64_t)UNIT_BATCH_COST * NB;
    const int64_t target = total * u / NU;
    int lo = 0, hi = NB;   // first batch with c(b) >= target
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int64_t c = (int64_t)seg_ptr[bat_seg0[mid]] + (int64_t)UNIT_BATCH_COST * mid;
        if (c >= target) hi = mid;
        else lo = mid + 1;
    }
    if (lo >= NB) units[u] = make_int4(NB, S, E3, 0);
    else units[u] = make_int4(lo, bat_seg0[lo], seg_ptr[bat_seg0[lo]], 0);
}

template <typename T>
cudaError_t scan_inclusive(const T* in, T* out, int64_t n, cudaStream_t st) {
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, in, out, (int)n, st);
    if (e != cudaSuccess) return e;
    ScratchScope scratch_scope;
    Dev<char> tmp;
    if ((e = tmp.alloc(tmp_bytes)) != cudaSuccess) return e;
    e = cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, in, out, (int)n, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}
template <typename T>
cudaError_t scan_exclusive(const T* in, T* out, int64_t n, cudaStream_t st) {
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, st);
    if (e != cudaSuccess) return e;
    ScratchScope scratch_scope;
    Dev<char> tmp;
    if ((e = tmp.alloc(tmp_bytes)) != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, (int)n, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

// Super tiles for the tcgen05 kernel (M = 128 rows per MMA): runs of consecutive entry tiles of one relation, cut
// every ST_TILES tiles.  Entries of consecutive tiles are contiguous, so a super tile is [stile_e0[j], stile_e0[j+1]).
constexpr int ST_TILES = 8;   // 8 x 16 = 128 entries
__global__ void k_stile_runstart(const int32_t* __restrict__ tile_info, int32_t NT, int32_t* __restrict__ start) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= NT) return;
    start[i] = (i == 0 || (tile_info[i] >> 8) != (tile_info[i - 1] >> 8)) ? (int32_t)i : 0;
}
__global__ void k_stile_heads(const int32_t* __restrict__ runstart, int32_t NT, int32_t* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= NT) return;
    head[i] = (((int32_t)i - runstart[i]) % ST_TILES == 0) ? 1 : 0;
}
__global__ void k_stile_fill(const int32_t* __restrict__ head, const int32_t* __restrict__ pos,
                             const int32_t* __restrict__ tile_e0, const int32_t* __restrict__ tile_info, int32_t NT,
                             int32_t NS, int32_t* __restrict__ stile_e0, int32_t* __restrict__ stile_rel) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= NT) return;
    if (head[i]) {
        stile_e0[pos[i]] = tile_e0[i];
        stile_rel[pos[i]] = tile_info[i] >> 8;
    }
    if (i == NT - 1) stile_e0[NS] = tile_e0[i] + (tile_info[i] & 0xff);
}
struct MaxOp {
    __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};
int build_stiles(Brc& b, cudaStream_t st) {
    const int32_t NT = b.num_tiles_noself;
    b.num_stiles = 0;
    if (NT <= 0) return 0;
    ScratchScope scratch_scope;
    Dev<int32_t> runstart, head, pos;
    Dev<char> tmp;
    RGCN_CUDA(runstart.alloc(NT));
    RGCN_CUDA(head.alloc((size_t)NT + 1));
    RGCN_CUDA(pos.alloc((size_t)NT + 1));
    k_stile_runstart<<<blocks_for(NT), TPB, 0, st>>>(b.tile_info, NT, runstart.p);
    size_t tmp_bytes = 0;
    RGCN_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, runstart.p, runstart.p, MaxOp(), (int)NT, st));
    RGCN_CUDA(tmp.alloc(tmp_bytes));
    RGCN_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tmp_bytes, runstart.p, runstart.p, MaxOp(), (int)NT, st));
    RGCN_CUDA(cudaMemsetAsync(head.p, 0, ((size_t)NT + 1) * 4, st));
    k_stile_heads<<<blocks_for(NT), TPB, 0, st>>>(runstart.p, NT, head.p);
    RGCN_CUDA(scan_exclusive(head.p, pos.p, (int64_t)NT + 1, st));
    int32_t NS = 0;
    RGCN_CUDA(cudaMemcpy(&NS, pos.p + NT, 4, cudaMemcpyDeviceToHost));
    RGCN_CUDA(palloc(&b.stile_e0, ((size_t)NS + 2) * 4));
    RGCN_CUDA(palloc(&b.stile_rel, ((size_t)NS + 2) * 4));
    k_stile_fill<<<blocks_for(NT), TPB, 0, st>>>(head.p, pos.p, b.tile_e0, b.tile_info, NT, NS, b.stile_e0, b.stile_rel);
    RGCN_CUDA(cudaGetLastError());
    RGCN_CUDA(cudaStreamSynchronize(st));
    b.num_stiles = NS;
    return 0;
}

int bit_length(uint64_t v) {
    int b = 0;
    while (v) {
        ++b;
        v >>= 1;
    }
    return std::max(b, 1);
}

// ---- chunk numbering shared between FWD and FWD_REL -------------------------------------------
// Both structures chunk the same (relation, dst) segments into the same pieces (same T / CH, entries
// of a segment in input order), only the ORDER of the segments differs.  FWD_REL's chunk c is the
// c-th chunk in (relation, dst, piece) order; a stable sort of FWD's chunks by (relation, dst) gives
// FWD's id of that chunk.  FWD_REL's chunk entries are rewritten to FWD ids, so the chunk rows the
// forward pass computed serve dL/dW unchanged (spec: oracle/csr_oracle.py share_chunks).
__global__ void k_chunk_keys(const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_own,
                             const int32_t* __restrict__ seg_rel, const uint32_t* __restrict__ e_idx, int32_t S,
                             uint32_t n_gat, uint64_t n_own, uint64_t* __restrict__ key, int32_t* __restrict__ id) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int32_t a = seg_ptr[s], b = seg_ptr[s + 1];
    if (a >= b || (e_idx[a] & IDX_MASK) < n_gat) return;   // not a chunked segment
    for (int32_t o = a; o < b; ++o) {
        const uint32_t c = (e_idx[o] & IDX_MASK) - n_gat;
        key[c] = (uint64_t)seg_rel[s] * n_own + (uint64_t)seg_own[s];
        id[c] = (int32_t)c;
    }
}
__global__ void k_chunk_renumber(uint32_t* __restrict__ e_idx, int64_t E3, uint32_t n_gat,
                                 const int32_t* __restrict__ order) {
    int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= E3) return;
    const uint32_t v = e_idx[o], idx = v & IDX_MASK;
    if (idx >= n_gat) e_idx[o] = (n_gat + (uint32_t)order[idx - n_gat]) | (v & LAST_FLAG);
}

int share_chunks(const Brc& fwd, Brc& rel, int64_t n_own, int64_t n_gat, int R, cudaStream_t st) {
    const int32_t NC = fwd.num_chunks;
    if (NC != rel.num_chunks)   // both structures chunk the same (relation, dst) segments: a mismatch is a builder bug
        return fail(RGCN_ERR_INVALID_ARG, "share_chunks: FWD and FWD_REL disagree on the number of chunks");
    if (NC == 0) return 0;
    ScratchScope scratch_scope;
    Dev<uint64_t> key, skey;
    Dev<int32_t> id;
    Dev<char> tmp;
    RGCN_CUDA(key.alloc(NC));
    RGCN_CUDA(skey.alloc(NC));
    RGCN_CUDA(id.alloc(NC));
    RGCN_CUDA(palloc(&rel.chunk_out, (size_t)NC * 4));
    k_chunk_keys<<<blocks_for(fwd.num_seg), TPB, 0, st>>>(fwd.seg_ptr, fwd.seg_own, fwd.seg_rel, fwd.e_idx, fwd.num_seg,
                                                         (uint32_t)n_gat, (uint64_t)n_own, key.p, id.p);
    const int end_bit = bit_length((uint64_t)(R + 1) * (uint64_t)n_own);
    size_t tmp_bytes = 0;
    RGCN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, skey.p, id.p, rel.chunk_out, (int)NC, 0, end_bit, st));
    RGCN_CUDA(tmp.alloc(tmp_bytes));
    RGCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, skey.p, id.p, rel.chunk_out, (int)NC, 0, end_bit, st));
    k_chunk_renumber<<<blocks_for(rel.num_entries), TPB, 0, st>>>(rel.e_idx, rel.num_entries, (uint32_t)n_gat, rel.chunk_out);
    RGCN_CUDA(cudaGetLastError());
    RGCN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Build one BRC from an entry list (owner local id, gather global id, relation, weight).
// n_own owners (ranges of NR), chunk rows are numbered from n_gat (the gather row count).
// The self loops belong to the owners [self_lo, self_hi) (all owners unless the structure is source-partitioned).
int build_brc(const int32_t* own, const int32_t* gat, const int32_t* rel, const float* w_entry, int64_t n2,
              int64_t n_own, int64_t n_gat, int R, int NR, int T, int CH, cudaStream_t st, Brc* out,
              int64_t self_lo = 0, int64_t self_hi = -1) {
    if (self_hi < 0) self_hi = n_own;
    const int64_t N = n_own;
    Brc b;
    b.num_entries0 = n2;
    b.range_nodes = NR;
    ScratchScope scratch_scope;
    Dev<uint64_t> key, skey;
    Dev<int32_t> eid, head, scan;
    RGCN_CUDA(key.alloc(n2));
    RGCN_CUDA(skey.alloc(n2));
    RGCN_CUDA(eid.alloc(n2));
    RGCN_CUDA(head.alloc(n2));
    RGCN_CUDA(scan.alloc(n2));
    RGCN_CUDA(palloc(&b.perm, std::max<int64_t>(n2, 1) * 4));
    const uint64_t nranges = std::max<uint64_t>((uint64_t)((N + NR - 1) / NR), 1);
    const uint64_t self_base = nranges * (uint64_t)(R + 1) * (uint64_t)NR;
    k_keys<<<blocks_for(n2), TPB, 0, st>>>(own, rel, n2, R, NR, self_base, key.p, eid.p);
    {
        uint64_t max_key = self_base + (uint64_t)N;
        int end_bit = std::min(64, bit_length(max_key));
        size_t tmp_bytes = 0;
        RGCN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, skey.p, eid.p, b.perm, (int)n2, 0,
                                                  end_bit, st));
        Dev<char> tmp;
        RGCN_CUDA(tmp.alloc(tmp_bytes));
        RGCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, skey.p, eid.p, b.perm, (int)n2, 0, end_bit,
                                                  st));
        RGCN_CUDA(cudaStreamSynchronize(st));
    }
    k_heads<<<blocks_for(n2), TPB, 0, st>>>(skey.p, n2, head.p);
    RGCN_CUDA(scan_inclusive(head.p, scan.p, n2, st));
    int32_t S = 0;
    if (n2 > 0) RGCN_CUDA(cudaMemcpy(&S, scan.p + (n2 - 1), 4, cudaMemcpyDeviceToHost));
    b.num_seg = S;
    RGCN_CUDA(palloc(&b.seg_ptr0, (size_t)(S + 1) * 4));
    RGCN_CUDA(palloc(&b.seg_ptr, (size_t)(S + 1) * 4));
    RGCN_CUDA(palloc(&b.seg_own, std::max(S, 1) * 4));
    RGCN_CUDA(palloc(&b.seg_rel, std::max(S, 1) * 4));
    k_seg_fill<<<blocks_for(n2), TPB, 0, st>>>(skey.p, head.p, scan.p, n2, R, NR, self_base, b.seg_ptr0, b.seg_own,
                                               b.seg_rel);
    k_set_i32<<<1, 1, 0, st>>>(b.seg_ptr0 + S, (int32_t)n2);
    RGCN_CUDA(palloc(&b.raw_idx, std::max<int64_t>(n2, 1) * 4));
    RGCN_CUDA(palloc(&b.raw_w, std::max<int64_t>(n2, 1) * 4));
    k_raw<<<blocks_for(n2), TPB, 0, st>>>(b.perm, gat, w_entry, n2, b.raw_idx, b.raw_w);

    // chunking of long segments
    Dev<int32_t> out_cnt, nchunk, chunk_base;
    RGCN_CUDA(out_cnt.alloc(S + 1));
    RGCN_CUDA(nchunk.alloc(S + 1));
    RGCN_CUDA(chunk_base.alloc(S + 1));
    RGCN_CUDA(cudaMemsetAsync(out_cnt.p, 0, (size_t)(S + 1) * 4, st));
    RGCN_CUDA(cudaMemsetAsync(nchunk.p, 0, (size_t)(S + 1) * 4, st));
    k_seg_counts<<<blocks_for(S), TPB, 0, st>>>(b.seg_ptr0, S, T, CH, out_cnt.p, nchunk.p);
    RGCN_CUDA(scan_exclusive(out_cnt.p, b.seg_ptr, (int64_t)S + 1, st));
    RGCN_CUDA(scan_exclusive(nchunk.p, chunk_base.p, (int64_t)S + 1, st));
    int32_t E3 = 0, NC = 0;
    RGCN_CUDA(cudaMemcpy(&E3, b.seg_ptr + S, 4, cudaMemcpyDeviceToHost));
    RGCN_CUDA(cudaMemcpy(&NC, chunk_base.p + S, 4, cudaMemcpyDeviceToHost));
    b.num_entries = E3;
    b.num_chunks = NC;
    RGCN_CUDA(palloc(&b.e_idx, ((size_t)std::max(E3, 1) + 16) * 4));   // +16: whole 16-byte chunks are staged
    RGCN_CUDA(palloc(&b.e_w, ((size_t)std::max(E3, 1) + 16) * 4));
    RGCN_CUDA(palloc(&b.chunk_beg, std::max(NC, 1) * 4));
    RGCN_CUDA(palloc(&b.chunk_end, std::max(NC, 1) * 4));
    RGCN_CUDA(palloc(&b.e_own, ((size_t)std::max(E3, 1) + 16) * 4));
    k_compact<<<blocks_for(n2), TPB, 0, st>>>(scan.p, b.seg_ptr0, b.seg_ptr, chunk_base.p, b.raw_idx, b.raw_w, n2, n_gat, T,
                                              CH, b.e_idx, b.e_w, b.chunk_beg, b.chunk_end, b.seg_own, b.e_own);

    // groups and batches
    Dev<int32_t> ghead, gscan;
    RGCN_CUDA(ghead.alloc(S));
    RGCN_CUDA(gscan.alloc(S));
    int32_t G = 0;
    if (S > 0) {
        k_group_heads<<<blocks_for(S), TPB, 0, st>>>(b.seg_own, b.seg_rel, S, NR, ghead.p);
        RGCN_CUDA(scan_inclusive(ghead.p, gscan.p, S, st));
        RGCN_CUDA(cudaMemcpy(&G, gscan.p + (S - 1), 4, cudaMemcpyDeviceToHost));
    }
    b.num_groups = G;
    Dev<int32_t> grp_seg, nb, bat_base;
    RGCN_CUDA(grp_seg.alloc(G + 1));
    RGCN_CUDA(nb.alloc(G + 1));
    RGCN_CUDA(bat_base.alloc(G + 1));
    int32_t NB = 0;
    if (G > 0) {
        k_group_fill<<<blocks_for(S), TPB, 0, st>>>(ghead.p, gscan.p, S, grp_seg.p);
        k_set_i32<<<1, 1, 0, st>>>(grp_seg.p + G, S);
        RGCN_CUDA(cudaMemsetAsync(nb.p, 0, (size_t)(G + 1) * 4, st));
        k_group_nb<<<blocks_for(G), TPB, 0, st>>>(grp_seg.p, G, nb.p);
        RGCN_CUDA(scan_exclusive(nb.p, bat_base.p, (int64_t)G + 1, st));
        RGCN_CUDA(cudaMemcpy(&NB, bat_base.p + G, 4, cudaMemcpyDeviceToHost));
    }
    b.num_batches = NB;
    RGCN_CUDA(palloc(&b.bat_seg0, std::max(NB, 1) * 4));
    RGCN_CUDA(palloc(&b.bat_info, std::max(NB, 1) * 4));
    if (NB > 0)
        k_batch_fill<<<blocks_for(NB), TPB, 0, st>>>(bat_base.p, grp_seg.p, b.seg_rel, G, NB, b.bat_seg0, b.bat_info);
    {   // entry tiles
        Dev<int32_t> ntl, tile_base;
        RGCN_CUDA(ntl.alloc(G + 1));
        RGCN_CUDA(tile_base.alloc(G + 1));
        int32_t NTL = 0;
        if (G > 0) {
            RGCN_CUDA(cudaMemsetAsync(ntl.p, 0, (size_t)(G + 1) * 4, st));
            k_group_ntiles<<<blocks_for(G), TPB, 0, st>>>(grp_seg.p, b.seg_ptr, G, ntl.p);
            RGCN_CUDA(scan_exclusive(ntl.p, tile_base.p, (int64_t)G + 1, st));
            RGCN_CUDA(cudaMemcpy(&NTL, tile_base.p + G, 4, cudaMemcpyDeviceToHost));
        }
        b.num_tiles = NTL;
        // self loops sort last: their groups (one per owner range) are the tail of the group list
        b.num_tiles_noself = NTL;
        if (G > 0) {
            // (one self-loop group per owner range that holds self loops)
            const int64_t self_groups = self_hi > self_lo ? (self_hi - 1) / NR - self_lo / NR + 1 : 0;
            const int32_t first_self_group = G - (int32_t)self_groups;
            if (first_self_group >= 0)
                RGCN_CUDA(cudaMemcpy(&b.num_tiles_noself, tile_base.p + first_self_group, 4, cudaMemcpyDeviceToHost));
        }
        RGCN_CUDA(palloc(&b.tile_e0, ((size_t)std::max(NTL, 1) + 16) * 4));
        RGCN_CUDA(palloc(&b.tile_info, ((size_t)std::max(NTL, 1) + 16) * 4));
        if (NTL > 0)
            k_tile_fill<<<blocks_for(NTL), TPB, 0, st>>>(tile_base.p, grp_seg.p, b.seg_ptr, b.seg_rel, G, NTL, b.tile_e0,
                                                         b.tile_info);
    }
    {
        const int64_t total = (int64_t)E3 + (int64_t)UNIT_BATCH_COST * NB;
        int64_t nu = std::min<int64_t>(UNIT_MAX, std::max<int64_t>(1, total / UNIT_MIN_COST));
        b.num_units = (int32_t)nu;
        RGCN_CUDA(palloc(&b.units, (size_t)(nu + 1) * sizeof(int4)));
        k_units<<<blocks_for(nu + 1), TPB, 0, st>>>(b.bat_seg0, b.seg_ptr, NB, S, E3, (int32_t)nu, b.units);
    }
    RGCN_CUDA(cudaStreamSynchronize(st));
    RGCN_CUDA(cudaGetLastError());
    {
        int rc = build_stiles(b, st);
        if (rc) return rc;
    }
    b.bytes = n2 * 12 + (int64_t)S * 16 + (int64_t)E3 * 8 + (int64_t)NC * 8 + (int64_t)NB * 8;
    *out = b;
    return 0;
}

}  // namespace
}  // namespace rgcn

using namespace rgcn;

extern "C" const char* rgcn_last_error(void) { return rgcn::last_error_cstr(); }
extern "C" int rgcn_abi_version(void) { return RGCN_B200_ABI_VERSION; }

extern "C" void rgcn_graph_destroy(rgcn_graph* g) {
    if (!g) return;
    for (int i = 0; i < 3; ++i) {
        if (i == RGCN_BRC_FWD_REL && g->rel_is_fwd) continue;
        g->brc[i].release();
    }
    g->arena.free_all();
    if (g->side) cudaStreamDestroy(g->side);
    if (g->ev_fork) cudaEventDestroy(g->ev_fork);
    if (g->ev_join) cudaEventDestroy(g->ev_join);
    delete g;
}

namespace {

// w_edge[e] = 1 / |{e' : rel(e') = rel(e), dst(e') = dst(e)}|  over ALL edges of the graph
int compute_edge_weights(const int32_t* dst32, const int32_t* rel32, int64_t E, int64_t N, int R, float* w_edge,
                         cudaStream_t st) {
    if (E == 0) return 0;
    ScratchScope scratch_scope;
    Dev<uint64_t> key, skey;
    Dev<int32_t> eid, perm, head, scan, start;
    RGCN_CUDA(key.alloc(E));
    RGCN_CUDA(skey.alloc(E));
    RGCN_CUDA(eid.alloc(E));
    RGCN_CUDA(perm.alloc(E));
    RGCN_CUDA(head.alloc(E));
    RGCN_CUDA(scan.alloc(E));
    k_wkeys<<<blocks_for(E), TPB, 0, st>>>(dst32, rel32, E, N, key.p, eid.p);
    const int end_bit = std::min(64, bit_length((uint64_t)R * (uint64_t)N));
    size_t tmp_bytes = 0;
    RGCN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, skey.p, eid.p, perm.p, (int)E, 0, end_bit, st));
    Dev<char> tmp;
    RGCN_CUDA(tmp.alloc(tmp_bytes));
    RGCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, skey.p, eid.p, perm.p, (int)E, 0, end_bit, st));
    k_heads<<<blocks_for(E), TPB, 0, st>>>(skey.p, E, head.p);
    RGCN_CUDA(scan_inclusive(head.p, scan.p, E, st));
    int32_t S = 0;
    RGCN_CUDA(cudaMemcpy(&S, scan.p + (E - 1), 4, cudaMemcpyDeviceToHost));
    RGCN_CUDA(start.alloc((size_t)S + 1));
    k_run_starts<<<blocks_for(E), TPB, 0, st>>>(head.p, scan.p, E, start.p);
    k_set_i32<<<1, 1, 0, st>>>(start.p + S, (int32_t)E);
    k_edge_weights<<<blocks_for(E), TPB, 0, st>>>(perm.p, scan.p, start.p, E, w_edge);
    RGCN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Entry list (owner-local, gather-global, rel, weight) of the edges whose owner end is in [lo,hi),
// in input order, followed by the self loops of the owned nodes; then the BRC(s) over it.
// push: select on the GATHER end instead (source-partitioned forward structures: gathered rows are the
// rank's own, owner rows span the whole graph; num_all = nodes of the whole graph).
int build_side(const int32_t* own_g, const int32_t* gat_g, const int32_t* rel_g, const float* w_edge, int64_t E,
               int64_t num_all, int R, int64_t lo, int64_t hi, int NR, int T, int CH, cudaStream_t st, Brc* blocked,
               Brc* relmajor, bool push = false) {
    ScratchScope scratch_scope;
    Dev<int32_t> flag, pos, own, gat, rel;
    Dev<float> w;
    RGCN_CUDA(flag.alloc(E + 1));
    RGCN_CUDA(pos.alloc(E + 1));
    int32_t nsel = 0;
    if (E > 0) {
        RGCN_CUDA(cudaMemsetAsync(flag.p, 0, (size_t)(E + 1) * 4, st));
        k_flag_owned<<<blocks_for(E), TPB, 0, st>>>(push ? gat_g : own_g, E, (int32_t)lo, (int32_t)hi, flag.p);
        RGCN_CUDA(scan_exclusive(flag.p, pos.p, E + 1, st));
        RGCN_CUDA(cudaMemcpy(&nsel, pos.p + E, 4, cudaMemcpyDeviceToHost));
    }
    const int64_t n_sel = hi - lo;
    const int64_t n2 = (int64_t)nsel + n_sel;
    RGCN_CUDA(own.alloc(n2));
    RGCN_CUDA(gat.alloc(n2));
    RGCN_CUDA(rel.alloc(n2));
    RGCN_CUDA(w.alloc(n2));
    k_fill_entries<<<blocks_for(E + n_sel), TPB, 0, st>>>(own_g, gat_g, rel_g, w_edge, flag.p, pos.p, E, nsel,
                                                          (int32_t)lo, (int32_t)hi, R, push ? 1 : 0, own.p, gat.p, rel.p,
                                                          w.p);
    const int64_t n_own = push ? num_all : n_sel;   // owner id space
    const int64_t n_gat = push ? n_sel : num_all;   // gather id space (chunk rows are numbered from here)
    int rc;
    const int64_t self_lo = push ? lo : 0, self_hi = push ? hi : n_sel;
    if ((rc = build_brc(own.p, gat.p, rel.p, w.p, n2, n_own, n_gat, R, NR, T, CH, st, blocked, self_lo, self_hi))) return rc;
    if (relmajor && (rc = build_brc(own.p, gat.p, rel.p, w.p, n2, n_own, n_gat, R, (int)std::max<int64_t>(n_own, 1), T,
                                    CH, st, relmajor, self_lo, self_hi)))
        return rc;
    return 0;
}

}  // namespace

namespace {
int graph_create_impl(const int64_t* src, int64_t src_stride, const int64_t* dst, int64_t dst_stride,
                      const int64_t* etype, int64_t etype_stride, int64_t num_edges, int64_t num_nodes,
                      int32_t num_relations, int64_t own_lo, int64_t own_hi, int32_t range_nodes,
                      int32_t split_threshold, int32_t chunk_size, bool push, void* stream, rgcn_graph** out) {
    if (!out) return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_create: out is null");
    *out = nullptr;
    if (num_edges < 0 || num_nodes <= 0 || num_relations <= 0)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_create: need num_edges >= 0, num_nodes > 0, num_relations > 0");
    if (own_lo < 0 || own_hi > num_nodes || own_lo >= own_hi)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_create: need 0 <= own_lo < own_hi <= num_nodes");
    if (num_edges > 0 && (!src || !dst || !etype))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_create: null edge tensor");
    if (num_edges + num_nodes >= (int64_t)0x7fffffff - 65536 || num_relations >= (1 << 22))
        return fail(RGCN_ERR_UNSUPPORTED, "rgcn_graph_create: graph too large for int32 indexing");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(ce != cudaSuccess ? (int)ce : (int)cudaErrorNoDevice,
                    "rgcn_graph_create: no CUDA device (the engine has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    rgcn_graph* g = new rgcn_graph();
    g->N = num_nodes;
    g->E = num_edges;
    g->R = num_relations;
    g->own_lo = own_lo;
    g->n_own = own_hi - own_lo;
    g->push = push;
    cudaGetDevice(&g->device);
    cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, g->device);
    if (cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking) != cudaSuccess) g->side = nullptr;
    if (cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming) != cudaSuccess) g->ev_fork = nullptr;
    if (cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming) != cudaSuccess) g->ev_join = nullptr;
    g->split_threshold = split_threshold > 0 ? split_threshold : 16;
    g->chunk_size = chunk_size > 0 ? chunk_size : 64;
    int64_t nr = range_nodes > 0 ? range_nodes : 16384;
    // small graphs: one range (pure relation-major); large: blocked so accumulate targets stay L2-hot
    const int64_t fwd_owners = push ? num_nodes : g->n_own;   // owner space of the forward structures
    const int64_t nr_bwd = std::min<int64_t>(nr, g->n_own);
    if (nr >= fwd_owners) nr = fwd_owners;
    g->range_nodes = (int32_t)nr;

    const int64_t E = num_edges;
    // arenas: the graph's own (persistent arrays) and this build's scratch; slabs sized to the graph so that a build
    // makes a handful of device allocations instead of ~170
    Arena scratch;
    {
        const size_t items = (size_t)(num_edges + num_nodes), mb2 = (size_t)2 << 20;
        auto slab = [&](size_t per_item, size_t cap) { return std::min(cap, std::max(mb2, (items * per_item + mb2 - 1) & ~(mb2 - 1))); };
        g->arena.slab_bytes = slab(48, (size_t)512 << 20);
        scratch.slab_bytes = slab(96, (size_t)1024 << 20);
    }
    struct ArenaGuard {
        Arena* s;
        ~ArenaGuard() {
            s->free_all();
            t_scratch = nullptr;
            t_persist = nullptr;
        }
    } arena_guard{&scratch};
    t_persist = &g->arena;
    t_scratch = &scratch;
    ScratchScope scratch_scope;
    Dev<int32_t> src32, dst32, rel32;
    Dev<float> w_edge;
    Dev<int> err;
    int rc = 0;
    auto bail = [&](int code) {
        rgcn_graph_destroy(g);
        return code;
    };
#define RGCN_CUDA_G(call)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return bail(fail((int)_e, std::string(#call) + ": " + cudaGetErrorString(_e)));        \
    } while (0)
    RGCN_CUDA_G(src32.alloc(E));
    RGCN_CUDA_G(dst32.alloc(E));
    RGCN_CUDA_G(rel32.alloc(E));
    RGCN_CUDA_G(w_edge.alloc(E));
    RGCN_CUDA_G(err.alloc(1));
    RGCN_CUDA_G(cudaMemsetAsync(err.p, 0, sizeof(int), st));
    if (E > 0)
        k_edges<<<blocks_for(E), TPB, 0, st>>>(src, src_stride, dst, dst_stride, etype, etype_stride, E, num_nodes,
                                               num_relations, src32.p, dst32.p, rel32.p, err.p);
    int herr = 0;
    RGCN_CUDA_G(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    RGCN_CUDA_G(cudaStreamSynchronize(st));
    if (herr)
        return bail(fail(RGCN_ERR_INDEX_RANGE,
                         "rgcn_graph_create: edge_index outside [0,num_nodes) or edge_type outside [0,num_relations)"));
    if ((rc = compute_edge_weights(dst32.p, rel32.p, E, num_nodes, num_relations, w_edge.p, st))) return bail(rc);
    const int T = g->split_threshold, CH = g->chunk_size;
    const bool one_range = nr >= fwd_owners;
    // forward: owner = dst, gather = src (+ the same entries in pure relation-major order for dL/dW);
    // pull: the edges whose dst is owned; push: the edges whose SRC is owned
    if ((rc = build_side(dst32.p, src32.p, rel32.p, w_edge.p, E, num_nodes, num_relations, own_lo, own_hi, (int)nr, T,
                         CH, st, &g->brc[RGCN_BRC_FWD], one_range ? nullptr : &g->brc[RGCN_BRC_FWD_REL], push)))
        return bail(rc);
    if (one_range) {
        g->brc[RGCN_BRC_FWD_REL] = g->brc[RGCN_BRC_FWD];
        g->rel_is_fwd = true;
    } else if ((rc = share_chunks(g->brc[RGCN_BRC_FWD], g->brc[RGCN_BRC_FWD_REL], fwd_owners, push ? g->n_own : num_nodes,
                                  num_relations, st))) {
        return bail(rc);
    }
    // transposed: owner = src, gather = dst, same per-edge weights (the edges whose src is owned, either mode)
    if ((rc = build_side(src32.p, dst32.p, rel32.p, w_edge.p, E, num_nodes, num_relations, own_lo, own_hi, (int)nr_bwd, T,
                         CH, st, &g->brc[RGCN_BRC_BWD], nullptr)))
        return bail(rc);
#undef RGCN_CUDA_G
    *out = g;
    return 0;
}
}  // namespace

extern "C" int rgcn_graph_create_part(const int64_t* src, int64_t src_stride, const int64_t* dst, int64_t dst_stride,
                                      const int64_t* etype, int64_t etype_stride, int64_t num_edges,
                                      int64_t num_nodes, int32_t num_relations, int64_t own_lo, int64_t own_hi,
                                      int32_t range_nodes, int32_t split_threshold, int32_t chunk_size, void* stream,
                                      rgcn_graph** out) {
    return graph_create_impl(src, src_stride, dst, dst_stride, etype, etype_stride, num_edges, num_nodes, num_relations,
                             own_lo, own_hi, range_nodes, split_threshold, chunk_size, false, stream, out);
}

extern "C" int rgcn_graph_create_push(const int64_t* src, int64_t src_stride, const int64_t* dst, int64_t dst_stride,
                                      const int64_t* etype, int64_t etype_stride, int64_t num_edges,
                                      int64_t num_nodes, int32_t num_relations, int64_t own_lo, int64_t own_hi,
                                      int32_t range_nodes, int32_t split_threshold, int32_t chunk_size, void* stream,
                                      rgcn_graph** out) {
    return graph_create_impl(src, src_stride, dst, dst_stride, etype, etype_stride, num_edges, num_nodes, num_relations,
                             own_lo, own_hi, range_nodes, split_threshold, chunk_size, true, stream, out);
}

extern "C" int rgcn_graph_create(const int64_t* src, int64_t src_stride, const int64_t* dst, int64_t dst_stride,
                                 const int64_t* etype, int64_t etype_stride, int64_t num_edges, int64_t num_nodes,
                                 int32_t num_relations, int32_t range_nodes, int32_t split_threshold,
                                 int32_t chunk_size, void* stream, rgcn_graph** out) {
    return rgcn_graph_create_part(src, src_stride, dst, dst_stride, etype, etype_stride, num_edges, num_nodes,
                                  num_relations, 0, num_nodes, range_nodes, split_threshold, chunk_size, stream, out);
}

extern "C" int rgcn_graph_query(const rgcn_graph* g, int32_t brc, int32_t key, int64_t* out) {
    if (!g || !out || brc < 0 || brc > 2) return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_query: bad argument");
    const Brc& b = g->brc[brc];
    switch (key) {
        case RGCN_Q_NUM_NODES: *out = g->N; break;
        case RGCN_Q_NUM_EDGES: *out = g->E; break;
        case RGCN_Q_NUM_RELATIONS: *out = g->R; break;
        case RGCN_Q_NUM_SEGMENTS: *out = b.num_seg; break;
        case RGCN_Q_NUM_ENTRIES: *out = b.num_entries; break;
        case RGCN_Q_NUM_CHUNKS: *out = b.num_chunks; break;
        case RGCN_Q_NUM_GROUPS: *out = b.num_groups; break;
        case RGCN_Q_NUM_BATCHES: *out = b.num_batches; break;
        case RGCN_Q_RANGE_NODES: *out = b.range_nodes; break;
        case RGCN_Q_NUM_OWNED: *out = g->n_own; break;
        case RGCN_Q_OWN_LO: *out = g->own_lo; break;
        case RGCN_Q_PUSH: *out = g->push ? 1 : 0; break;
        case RGCN_Q_NUM_ENTRIES0: *out = b.num_entries0; break;
        case RGCN_Q_NUM_TILES: *out = b.num_tiles; break;
        case RGCN_Q_NUM_TILES_NOSELF: *out = b.num_tiles_noself; break;
        case RGCN_Q_DEVICE_BYTES:
            *out = g->brc[0].bytes + g->brc[1].bytes + (g->rel_is_fwd ? 0 : g->brc[2].bytes);
            break;
        default: return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_query: unknown key");
    }
    return 0;
}

extern "C" int rgcn_graph_export(const rgcn_graph* g, int32_t brc, int32_t array, void* host_dst, int64_t bytes,
                                 void* stream) {
    if (!g || !host_dst || brc < 0 || brc > 2) return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_export: bad argument");
    const Brc& b = g->brc[brc];
    const void* p = nullptr;
    int64_t n = 0;
    switch (array) {
        case RGCN_A_PERM: p = b.perm; n = b.num_entries0; break;
        case RGCN_A_SEG_PTR: p = b.seg_ptr; n = (int64_t)b.num_seg + 1; break;
        case RGCN_A_SEG_OWN: p = b.seg_own; n = b.num_seg; break;
        case RGCN_A_SEG_REL: p = b.seg_rel; n = b.num_seg; break;
        case RGCN_A_SEG_PTR0: p = b.seg_ptr0; n = (int64_t)b.num_seg + 1; break;
        case RGCN_A_E_IDX: p = b.e_idx; n = b.num_entries; break;
        case RGCN_A_E_W: p = b.e_w; n = b.num_entries; break;
        case RGCN_A_RAW_IDX: p = b.raw_idx; n = b.num_entries0; break;
        case RGCN_A_RAW_W: p = b.raw_w; n = b.num_entries0; break;
        case RGCN_A_CHUNK_BEG: p = b.chunk_beg; n = b.num_chunks; break;
        case RGCN_A_CHUNK_END: p = b.chunk_end; n = b.num_chunks; break;
        case RGCN_A_BAT_SEG0: p = b.bat_seg0; n = b.num_batches; break;
        case RGCN_A_BAT_INFO: p = b.bat_info; n = b.num_batches; break;
        case RGCN_A_E_OWN: p = b.e_own; n = b.num_entries; break;
        case RGCN_A_TILE_E0: p = b.tile_e0; n = b.num_tiles; break;
        case RGCN_A_TILE_INFO: p = b.tile_info; n = b.num_tiles; break;
        case RGCN_A_CHUNK_OUT: p = b.chunk_out; n = b.chunk_out ? b.num_chunks : 0; break;
        default: return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_export: unknown array");
    }
    if (bytes != n * 4) return fail(RGCN_ERR_INVALID_ARG, "rgcn_graph_export: byte count mismatch");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    RGCN_CUDA(cudaMemcpyAsync(host_dst, p, (size_t)bytes, cudaMemcpyDeviceToHost, st));
    RGCN_CUDA(cudaStreamSynchronize(st));
    return 0;
}
