// common.cuh — shared declarations of the sm_100a R-GCN engine (internal; the public C ABI is
// include/rgcn_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rgcn_b200.h"

namespace rgcn {

constexpr int BS = 16;                        // segments per batch (MMA M)
constexpr int UNIT_BATCH_COST = 24;           // work-unit cost model: entries + 24 per batch
constexpr int UNIT_MAX = 32768, UNIT_MIN_COST = 1024;
constexpr int ET = 16;                        // entries per entry tile (MMA M of the entry-tile kernels)
constexpr uint32_t LAST_FLAG = 0x80000000u;   // bit 31 of e_idx: last entry of its segment
constexpr uint32_t IDX_MASK = 0x7fffffffu;

int fail(int code, const std::string& msg);

// pass tags reported by rgcn_profile_collect (see include/rgcn_b200.h)
enum { TAG_WPREP = 1, TAG_PREPASS = 2, TAG_TILE_FWD = 3, TAG_TILE_BWD = 4, TAG_WGRAD = 5, TAG_COPY = 6, TAG_MASK = 7,
       TAG_SIMPLE = 8, TAG_MAP = 9, TAG_SELF = 10, TAG_NVL_STORE = 11, TAG_NVL_REDUCE = 12, TAG_GEMM = 13 };
void note_launch(int n);
void set_nvl_mode(int m);   // nvl_comm.cu
struct ProfScope {   // records a CUDA-event pair around a launch when profiling is enabled
    ProfScope(int tag, int d0, int d1, cudaStream_t st);
    ~ProfScope();
    cudaStream_t st_;
    int idx_;
    bool nvtx_;   // an NVTX range (RGCN_B200_NVTX=1) named after the pass, for timeline tools
};

#define RGCN_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            return ::rgcn::fail((int)_e, std::string(#call) + ": " + cudaGetErrorString(_e)); \
        }                                                                                    \
    } while (0)

// Device memory in a few large slabs.  The graph build makes ~170 allocations; as individual cudaMalloc / cudaFree
// calls they cost 160 ms of host time around 6 ms of kernels (torch profiler, AM shape).  Sub-allocations are 256-byte
// aligned and never freed singly: the persistent arrays of a graph live in the graph's arena (freed with the graph),
// scratch in a per-build arena that is rewound at scope exit (same stream, so reuse is ordered).
struct Arena {
    std::vector<std::pair<char*, size_t>> slabs;
    size_t cur = 0, off = 0;            // next free byte: slabs[cur] + off
    size_t slab_bytes = 2u << 20;       // size of a new slab (or the request, if larger)
    struct Mark {
        size_t cur, off;
    };
    cudaError_t take(void** p, size_t bytes) {
        bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
        for (; cur < slabs.size(); ++cur, off = 0)
            if (off + bytes <= slabs[cur].second) {
                *p = slabs[cur].first + off;
                off += bytes;
                return cudaSuccess;
            }
        char* s = nullptr;
        const size_t sz = std::max(bytes, slab_bytes);
        cudaError_t e = cudaMalloc(&s, sz);
        if (e != cudaSuccess) return e;
        slabs.emplace_back(s, sz);
        cur = slabs.size() - 1;
        off = bytes;
        *p = s;
        return cudaSuccess;
    }
    Mark mark() const { return Mark{cur, off}; }
    void rewind(Mark m) { cur = m.cur, off = m.off; }
    void free_all() {
        for (auto& s : slabs) cudaFree(s.first);
        slabs.clear();
        cur = off = 0;
    }
};

// One blocked relational CSR (spec: oracle/csr_oracle.py; built by graph_build.cu).
struct Brc {
    int64_t num_entries0 = 0;   // E + N (edges + self loops)
    int64_t num_entries = 0;    // after chunking (E3)
    int32_t num_seg = 0, num_chunks = 0, num_groups = 0, num_batches = 0;
    int32_t range_nodes = 0;
    int32_t* perm = nullptr;       // [E+N] entry id per sorted position
    int32_t* seg_ptr0 = nullptr;   // [S+1] into raw_* (before chunking)
    int32_t* seg_ptr = nullptr;    // [S+1] into e_*
    int32_t* seg_own = nullptr;    // [S]
    int32_t* seg_rel = nullptr;    // [S]
    uint32_t* e_idx = nullptr;     // [E3] gather row | LAST_FLAG ; rows >= N are chunk rows
    float* e_w = nullptr;          // [E3]
    int32_t* raw_idx = nullptr;    // [E+N]
    float* raw_w = nullptr;        // [E+N]
    int32_t* chunk_beg = nullptr;  // [NC] into raw_*
    int32_t* chunk_end = nullptr;  // [NC]
    int32_t* chunk_out = nullptr;  // [NC] row of the chunk matrix this chunk is written to (null: its own id).
                                   // FWD_REL shares FWD's numbering, so backward can reuse forward's chunk rows
    int32_t* bat_seg0 = nullptr;   // [NB]
    int32_t* bat_info = nullptr;   // [NB] rel << 8 | nseg
    int4* units = nullptr;         // [NU+1] equal-cost work units: (first batch, first segment, first entry, 0)
    int32_t num_units = 0;
    int32_t* e_own = nullptr;      // [E3] owner row of every entry
    int32_t* tile_e0 = nullptr;    // [NT] entry tiles: <= 16 consecutive entries of one (range, relation) group
    int32_t* tile_info = nullptr;  // [NT] rel << 8 | count
    int32_t num_tiles = 0;
    int32_t num_tiles_noself = 0;  // tiles before the self-loop tiles (which sort last)
    int32_t* stile_e0 = nullptr;   // [NS+1] super tiles (tcgen05 kernel): <= 128 consecutive entries of one relation,
    int32_t* stile_rel = nullptr;  // [NS]   edge tiles only; super tile j = entries [stile_e0[j], stile_e0[j+1])
    int32_t num_stiles = 0;
    int64_t bytes = 0;
    void release();
};

}  // namespace rgcn

struct rgcn_graph {
    int64_t N = 0, E = 0;      // global node / edge counts (gather rows are global ids)
    int64_t own_lo = 0, n_own = 0;   // owned node range [own_lo, own_lo + n_own) (whole graph unless partitioned)
    int32_t R = 0;
    int32_t range_nodes = 0, split_threshold = 0, chunk_size = 0;
    int device = 0;
    int num_sms = 148;
    rgcn::Brc brc[3];
    rgcn::Arena arena;         // owns every array of the three structures
    bool rel_is_fwd = false;   // FWD_REL aliases FWD (graph fits one range)
    bool push = false;         // source-partitioned forward structures (rgcn_graph_create_push)
    // forward structures: rows of x that are gathered / rows of out that are written
    int64_t fwd_gather_rows() const { return push ? n_own : N; }
    int64_t fwd_out_rows() const { return push ? N : n_own; }
    // fork/join plumbing for passes of one layer call that do not depend on each other (created with the
    // graph, so nothing is allocated inside a CUDA-graph capture): a side stream and two timing-free events
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace rgcn {

// ---- launchers implemented in layer_kernels.cu / simple_kernels.cu / map_gather.cu --------------
int pad_dim(int f);   // 16/32/64 or 0

// Tile pass: out[own] += sum over a BRC's segments of (weighted gathered rows) . B_rel
// (forward: B = W ; transposed: B = W^T), self-loop relation carries root (+ bias).
struct TilePass {
    const Brc* brc;
    int64_t n_nodes;
    int self_rel;
    const float* feat; int64_t ldf; int kin;       // gathered rows
    float* aux;                                     // [num_chunks, kp] chunk rows (written by pre-pass)
    const float4* wfrag;                            // [(R+1), kp/8, np/8, 32] fragment-ordered hi/lo split
    const float2* wfrag2;                           // same fragments, unsplit fp32 pairs (half the bytes)
    const float* bias; int nbias;                   // nullable, [nbias]
    float* out; int64_t ldo; int nout;              // accumulate target (zeroed by caller), nout % 4 == 0
    int kp, np;
    bool relu_in;
    bool transposed;                                // dL/dx pass (profiling tag only)
    bool vec4;                                      // entry-tile kernel: 16-byte row loads, wfrag prepared with perm
    int tag_out;                                    // logical output width (profiling tag)
    bool packed;                                    // out rows are tightly packed with an odd width (ldo == nout,
                                                    // not a multiple of 4): bulk-reduce scatter into shifted windows
    int64_t out_rows;                               // rows of `out` (packed mode: bound of the last window)
    int ctas_per_sm = 0;                            // > 0: cap of resident CTAs per SM (pass shares the SMs with a concurrent one)
};
bool etile_packed_ok(const float* out, int64_t ldo, int nout, int np);
int launch_chunk_prepass(const TilePass& p, cudaStream_t st);
int launch_tile_pass(const TilePass& p, int num_sms, cudaStream_t st);

struct WPrep {
    const float* weight;   // [R, fin, fout]
    const float* root;     // [fin, fout] nullable
    int R, fin, fout;
    int kp, np;            // padded dims of the pass (kp x np B-operand)
    bool transpose;        // B[k][n] = W[n][k]
    bool perm;             // K order of the vector-load entry-tile kernel (see etile_kernels.cu)
    float4* wfrag;
    float2* wfrag2;        // nullable: unsplit fp32 pairs in the same fragment order
};
int launch_wprep(const WPrep& p, cudaStream_t st);

struct WGradPass {
    const Brc* brc;        // relation-major forward BRC
    int64_t n_nodes;
    int self_rel;
    const float* feat; int64_t ldf; int kin;
    float* aux;            // forward chunk rows [num_chunks, kp]
    const float* gout; int64_t ldg; int nout;
    float* gweight;        // [R, kin, nout] nullable (zeroed by caller)
    float* groot;          // [kin, nout] nullable (zeroed)
    float* gbias;          // [nout] nullable (zeroed)
    int kp, np;
    bool relu_in;
    bool vec4;             // entry-tile kernel with 16-byte row loads
    int ctas_per_sm = 0;   // > 0: cap of resident CTAs per SM
};
int launch_wgrad_pass(const WGradPass& p, int num_sms, cudaStream_t st);
// entry-tile variants (etile_kernels.cu): rows gathered straight into MMA fragments
bool etile_vec4_ok(const float* feat, int64_t ldf, int kin, const float* aux);
int launch_etile_pass(const TilePass& p, int num_sms, cudaStream_t st);   // edge tiles only: run launch_selfloop_pass first
int launch_selfloop_pass(const TilePass& p, int64_t own_lo, int64_t n_own, int R, int num_sms, cudaStream_t st);
int launch_ewgrad_pass(const WGradPass& p, int num_sms, cudaStream_t st);
// dL/dW on the transposed structure (wide owner rows, narrow gathered rows): see k_ewgrad_t
bool ewgrad_t_ok(const WGradPass& p);
int launch_ewgrad_t_pass(const WGradPass& p, float* scratch, int num_sms, cudaStream_t st);
bool selfloop_pad_ok(int kp, int np);
int launch_selfloop_pad(const TilePass& p, const float* x_raw, int64_t ld_raw, float* mirror, int64_t ldm, int64_t n_own,
                        int R, int num_sms, cudaStream_t st);

// tcgen05 entry-tile pass (etile_tc.cu): 64 gathered columns x 64 produced columns, 16-byte addressable rows on
// both sides; wtc = operand images written by launch_wprep_tc
bool etile_tc_ok(const TilePass& p);
int64_t wprep_tc_floats(int R);
int launch_wprep_tc(const float* weight, const float* root, int R, int fin, int fout, bool transpose, float* wtc,
                    cudaStream_t st);
int launch_etile_tc(const TilePass& p, const float* wtc, int R, int num_sms, cudaStream_t st);

int launch_copy_cols(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n, int cols, cudaStream_t st);
int launch_pad_rows(const float* src, int64_t lds, int cols, float* dst, int64_t ldd, int64_t n, cudaStream_t st);
int launch_relu_mask(float* g, int64_t ldg, const float* pre, int64_t ldp, int64_t n, int cols, cudaStream_t st);

// generic scalar path (any fin/fout)
struct SimplePass {
    const Brc* brc; int64_t n_nodes; int self_rel;
    const float* feat; int64_t ldf; int kin;
    const float* weight; const float* root; const float* bias; bool transpose; int w_rows, w_cols;   // W_r is [w_rows, w_cols]
    float* out; int64_t ldo; int nout;
    bool relu_in;
};
int launch_simple_pass(const SimplePass& p, cudaStream_t st);
struct SimpleWGrad {
    const Brc* brc; int64_t n_nodes; int self_rel;
    const float* feat; int64_t ldf; int kin;
    const float* gout; int64_t ldg; int nout;
    float* gweight; float* groot; float* gbias;
    bool relu_in;
};
int launch_simple_wgrad(const SimpleWGrad& p, cudaStream_t st);

}  // namespace rgcn
