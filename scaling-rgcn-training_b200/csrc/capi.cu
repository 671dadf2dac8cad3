// capi.cu — extern "C" layer entry points (include/rgcn_b200.h): orchestration of the passes.
#include <algorithm>

#include "common.cuh"

using namespace rgcn;

namespace {

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct WsCarver {
    char* base;
    int64_t cap, off = 0;
    bool ok = true;
    WsCarver(void* p, int64_t bytes) : base((char*)p), cap(bytes) {}
    template <typename T>
    T* take(int64_t count) {
        const int64_t bytes = align_up(std::max<int64_t>(count, 1) * (int64_t)sizeof(T), 256);
        if (off + bytes > cap || base == nullptr) {
            ok = false;
            return nullptr;
        }
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};

inline int64_t ws_take(int64_t count, int64_t elem) { return align_up(std::max<int64_t>(count, 1) * elem, 256); }

// Which kernel family gathers a pass whose gathered rows are padded to `kp` columns:
// RGCN_B200_ETILE=1 / 0 forces the entry-tile / staged kernels, default = entry tiles for rows of
// up to 32 columns, staged (coalesced 128-byte row loads through shared memory) for wider rows.
bool etile_choice(int kp, bool vec4_ok) {
    static int mode = -2;
    if (mode == -2) {
        const char* e = getenv("RGCN_B200_ETILE");
        mode = e ? (e[0] == '0' ? 0 : 1) : -1;
    }
    if (mode >= 0) return mode == 1;
    return kp <= 32 || vec4_ok;
}

// run-time options (rgcn_set_option); defaults may be overridden by the environment once
struct Options {
    int overlap = 1;       // independent passes of one layer call run concurrently (fork/join on the graph's side stream)
    int wg_ctas = 0;       // > 0: resident CTAs per SM of the dL/dW pass while it shares the SMs with the dL/dx chain
    int dx_ctas = 0;       // same for the dL/dx tile pass (0 = no cap: measured best, the second pass fills the SMs
                           // as the first drains; half-occupancy co-residency was 20 % slower on every pair)
    Options() {
        if (const char* e = getenv("RGCN_B200_OVERLAP")) overlap = atoi(e);
        if (const char* e = getenv("RGCN_B200_OVL_WG")) wg_ctas = atoi(e);
        if (const char* e = getenv("RGCN_B200_OVL_DX")) dx_ctas = atoi(e);
    }
};
Options& opts() {
    static Options o;
    return o;
}

// fork: the side stream waits for everything issued so far on `st`; join: `st` waits for the side stream.
// Both are event record + stream wait, which a CUDA-graph capture of `st` turns into graph edges.
struct Fork {
    const rgcn_graph* g;
    cudaStream_t st;
    bool open = false;
    Fork(const rgcn_graph* g_, cudaStream_t st_) : g(g_), st(st_) {}
    bool available() const { return opts().overlap != 0 && g->side && g->ev_fork && g->ev_join; }
    int begin() {
        RGCN_CUDA(cudaEventRecord(g->ev_fork, st));
        RGCN_CUDA(cudaStreamWaitEvent(g->side, g->ev_fork, 0));
        open = true;
        return 0;
    }
    int end() {
        if (!open) return 0;
        open = false;
        RGCN_CUDA(cudaEventRecord(g->ev_join, g->side));
        RGCN_CUDA(cudaStreamWaitEvent(st, g->ev_join, 0));
        return 0;
    }
};
// (an early error return between begin() and end() leaves the side stream un-joined; every such path
// returns a failure code and the caller raises, so no result of that call is consumed)

bool direct_target(const void* p, int64_t ld, int width) {
    return ld == width && width % 4 == 0 && ((uintptr_t)p & 15) == 0;
}

int64_t fwd_ws(const rgcn_graph* g, int fin, int fout) {
    const int kp = pad_dim(fin), np = pad_dim(fout);
    if (!kp || !np) return 256;
    int64_t b = 0;
    b += ws_take((int64_t)(g->R + 1) * kp * np * 2, 4);                      // wfrag (hi,lo)
    b += ws_take((int64_t)(g->R + 1) * kp * np, 4);                          // wfrag2 (fp32 pairs)
    b += ws_take((int64_t)g->brc[RGCN_BRC_FWD].num_chunks * kp, 4);          // chunk rows
    b += ws_take((int64_t)g->fwd_out_rows() * np, 4);                        // padded accumulate target
    if (kp == 64 && np == 64) b += ws_take(wprep_tc_floats(g->R), 4);        // tcgen05 operand images
    return b;
}

int64_t bwd_ws(const rgcn_graph* g, int fin, int fout) {
    const int kp = pad_dim(fin), np = pad_dim(fout);
    if (!kp || !np) return 256;
    int64_t b = 0;
    b += ws_take((int64_t)g->brc[RGCN_BRC_FWD_REL].num_chunks * kp, 4);      // chunk rows of x (dW pass)
    b += ws_take((int64_t)(g->R + 1) * kp * np * 2, 4);                      // W^T frags
    b += ws_take((int64_t)(g->R + 1) * kp * np, 4);                          // W^T frags, fp32 pairs
    b += ws_take((int64_t)g->brc[RGCN_BRC_BWD].num_chunks * np, 4);          // chunk rows of gout (dx pass)
    b += ws_take((int64_t)g->n_own * kp, 4);                                 // padded dx target
    if (kp == 64 && np == 64) b += ws_take(wprep_tc_floats(g->R), 4);        // tcgen05 operand images
    if (np == 16 && kp > np) b += ws_take((int64_t)(g->R + 1) * 16 * kp, 4); // dW^T accumulators (k_ewgrad_t)
    return b;
}

}  // namespace

extern "C" int64_t rgcn_layer_workspace_bytes(const rgcn_graph* g, int32_t fin, int32_t fout, int32_t backward) {
    if (!g || fin <= 0 || fout <= 0) return -1;
    return backward ? bwd_ws(g, fin, fout) : fwd_ws(g, fin, fout);
}

extern "C" int64_t rgcn_layer_chunk_rows_bytes(const rgcn_graph* g, int32_t fin) {
    if (!g || fin <= 0) return -1;
    const int kp = pad_dim(fin);
    if (!kp) return 0;   // generic kernels: no chunk rows
    return ws_take((int64_t)g->brc[RGCN_BRC_FWD].num_chunks * kp, 4);
}

namespace {
int layer_fwd(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin, const float* weight, const float* root,
              const float* bias, float* out, int64_t ldo, int32_t fout, uint32_t flags, void* workspace,
              int64_t workspace_bytes, float* chunk_rows, float* x_mirror, int64_t ld_mirror, void* stream) {
    if (!g || !x || !weight || !out || fin <= 0 || fout <= 0 || ldx < fin || ldo < fout ||
        ((uintptr_t)chunk_rows & 15) != 0 ||
        (x_mirror && (((uintptr_t)x_mirror & 15) != 0 || ld_mirror % 4 != 0 || ld_mirror < ((fin + 3) & ~3))))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_layer_fwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const bool relu = (flags & RGCN_F_RELU_IN) != 0;
    const int kp = pad_dim(fin), np = pad_dim(fout);
    // tile kernels address gathered rows with 32-bit element offsets
    const int64_t n_gat = g->fwd_gather_rows(), n_out = g->fwd_out_rows();   // rows of x / rows of out
    const bool big = (uint64_t)n_gat * (uint64_t)ldx >= (1ull << 32) ||
                     (uint64_t)n_out * (uint64_t)std::max<int64_t>(ldo, np) >= (1ull << 32);
    if (g->push && ((flags & RGCN_F_FORCE_SIMPLE) || !kp || !np || big))
        return fail(RGCN_ERR_UNSUPPORTED, "rgcn_layer_fwd: a source-partitioned graph runs on the tile kernels only (widths <= 64)");
    if ((flags & RGCN_F_FORCE_SIMPLE) || !kp || !np || big) {
        // (the mirror is always written when one is given: the caller hands it to backward as x)
        if (x_mirror) {
            int prc = launch_pad_rows(x, ldx, fin, x_mirror, ld_mirror, n_gat, st);
            if (prc) return prc;
        }
        RGCN_CUDA(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)fout * 4, (size_t)g->n_own, st));
        SimplePass p{};
        p.brc = &g->brc[RGCN_BRC_FWD];
        p.n_nodes = g->N;
        p.self_rel = g->R;
        p.feat = x; p.ldf = ldx; p.kin = fin;
        p.weight = weight; p.root = root; p.bias = bias; p.transpose = false; p.w_rows = fin; p.w_cols = fout;
        p.out = out; p.ldo = ldo; p.nout = fout; p.relu_in = relu;
        return launch_simple_pass(p, st);
    }
    WsCarver ws(workspace, workspace_bytes);
    float4* wfrag = (float4*)ws.take<float>((int64_t)(g->R + 1) * kp * np * 2);
    float2* wfrag2 = (float2*)ws.take<float>((int64_t)(g->R + 1) * kp * np);
    float* aux = ws.take<float>((int64_t)g->brc[RGCN_BRC_FWD].num_chunks * kp);
    if (chunk_rows) aux = chunk_rows;   // kept by the caller for rgcn_layer_bwd_reuse
    // accumulate straight into `out` when its rows are 16-byte addressable and hold whole quads
    // (ldo == fout % 4 == 0, or a caller-padded row: ldo % 4 == 0 and ldo >= ceil4(fout); the pad
    // columns then receive zeros); otherwise through a padded buffer + column copy
    const int fout4 = (fout + 3) & ~3;
    const bool direct = ldo % 4 == 0 && ldo >= fout4 && ((uintptr_t)out & 15) == 0;
    float* target = direct ? out : ws.take<float>(n_out * np);
    float* wtc = (kp == 64 && np == 64) ? ws.take<float>(wprep_tc_floats(g->R)) : nullptr;
    if (!ws.ok) return fail(RGCN_ERR_WORKSPACE, "rgcn_layer_fwd: workspace too small (see rgcn_layer_workspace_bytes)");
    const int64_t tld = direct ? ldo : np;
    const int tn = direct ? fout4 : np;
    int rc = 0;
    // wide rows that are not 16-byte addressable (Fin = 63) go through the staged kernels; everything
    // else is gathered straight into MMA fragments
    // rows that are not 16-byte addressable (emb = 63) and a caller-provided mirror: the gathers read the
    // zero-padded mirror; it is written by the fused pad + self-loop pass when the shape allows, else by
    // a plain padding copy
    const float* x_raw = x;
    const int64_t ld_raw = ldx;
    bool fused_pad = false;
    if (x_mirror) {
        // the fused pass emits root fragments in the vector K order and replaces the self-loop pass of the
        // entry-tile family, so it is taken only when the MIRROR will be gathered by k_etile with vector loads
        const bool mirror_v4 = etile_vec4_ok(x_mirror, ld_mirror, fin, aux);
        // (x rows 0.. are the owned rows: the whole graph, or a source-partitioned rank's shard)
        fused_pad = !etile_vec4_ok(x, ldx, fin, aux) && mirror_v4 && etile_choice(kp, mirror_v4) &&
                    selfloop_pad_ok(kp, np) && (g->push || g->n_own == g->N);
        if (!fused_pad && (rc = launch_pad_rows(x, ldx, fin, x_mirror, ld_mirror, n_gat, st))) return rc;
        x = x_mirror;
        ldx = ld_mirror;
    }
    const bool v4ok = etile_vec4_ok(x, ldx, fin, aux);
    const bool et = etile_choice(kp, v4ok);
    const bool v4 = et && v4ok;
    WPrep wp{weight, root, g->R, fin, fout, kp, np, false, v4, wfrag, wfrag2};
    TilePass p{};
    p.brc = &g->brc[RGCN_BRC_FWD];
    p.n_nodes = n_gat;
    p.self_rel = g->R;
    p.feat = x; p.ldf = ldx; p.kin = fin;
    p.aux = aux;
    p.wfrag = wfrag;
    p.wfrag2 = wfrag2;
    p.bias = bias; p.nbias = fout;
    p.out = target; p.ldo = tld; p.nout = tn; p.tag_out = fout;
    p.kp = kp; p.np = np;
    p.relu_in = relu;
    p.vec4 = v4;
    p.out_rows = n_out;
    if (et) {
        // the chunk pre-pass (long segments -> chunk rows) and the root pass are independent: the pre-pass
        // runs on the side stream (reading the caller's rows when the mirror is still being written)
        Fork fk(g, st);
        const bool ovl = fk.available() && g->brc[RGCN_BRC_FWD].num_chunks > 0;
        TilePass pre = p;
        if (fused_pad) {
            pre.feat = x_raw;
            pre.ldf = ld_raw;
        }
        if (ovl) {
            if ((rc = fk.begin())) return rc;
            if ((rc = launch_chunk_prepass(pre, g->side))) return rc;
        }
        if ((rc = launch_wprep(wp, st))) return rc;
        // root + bias with plain stores (initialises the target; the fused variant also writes the mirror)
        TilePass ps = p;
        if (g->push) {
            // x holds the owned rows, the target the rows of ALL nodes: the rows this rank does not own start
            // from zero, the owned ones from root + bias
            RGCN_CUDA(cudaMemsetAsync(target, 0, (size_t)n_out * tld * 4, st));
            ps.out = target + g->own_lo * tld;
        }
        if (fused_pad) rc = launch_selfloop_pad(ps, x_raw, ld_raw, x_mirror, ld_mirror, g->n_own, g->R, g->num_sms, st);
        else rc = launch_selfloop_pass(ps, g->push ? 0 : g->own_lo, g->n_own, g->R, g->num_sms, st);
        if (rc) return rc;
        if (ovl) {
            if ((rc = fk.end())) return rc;
        } else if ((rc = launch_chunk_prepass(p, st))) {
            return rc;
        }
        // the edge tiles accumulate: 64 x 64 column passes on the tcgen05 kernel, the rest on the mma.sync one
        if (wtc && etile_tc_ok(p)) {
            if ((rc = launch_wprep_tc(weight, root, g->R, fin, fout, false, wtc, st))) return rc;
            if ((rc = launch_etile_tc(p, wtc, g->R, g->num_sms, st))) return rc;
        } else if ((rc = launch_etile_pass(p, g->num_sms, st))) {
            return rc;
        }
    } else {
        if ((rc = launch_wprep(wp, st))) return rc;
        if ((rc = launch_chunk_prepass(p, st))) return rc;
        if (g->push) return fail(RGCN_ERR_UNSUPPORTED, "rgcn_layer_fwd: source-partitioned graphs need the entry-tile kernels");
        RGCN_CUDA(cudaMemsetAsync(target, 0, (size_t)n_out * tld * 4, st));
        if ((rc = launch_tile_pass(p, g->num_sms, st))) return rc;
    }
    if (!direct) return launch_copy_cols(target, tld, out, ldo, n_out, fout, st);
    return 0;
}
}  // namespace

extern "C" int rgcn_layer_fwd(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin, const float* weight,
                              const float* root, const float* bias, float* out, int64_t ldo, int32_t fout,
                              uint32_t flags, void* workspace, int64_t workspace_bytes, void* stream) {
    return layer_fwd(g, x, ldx, fin, weight, root, bias, out, ldo, fout, flags, workspace, workspace_bytes, nullptr, nullptr,
                     0, stream);
}

extern "C" int rgcn_layer_fwd_keep(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin, const float* weight,
                                   const float* root, const float* bias, float* out, int64_t ldo, int32_t fout,
                                   uint32_t flags, void* workspace, int64_t workspace_bytes, float* chunk_rows,
                                   float* x_mirror, int64_t ld_mirror, void* stream) {
    return layer_fwd(g, x, ldx, fin, weight, root, bias, out, ldo, fout, flags, workspace, workspace_bytes, chunk_rows,
                     x_mirror, ld_mirror, stream);
}

namespace {
int layer_bwd(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin, const float* weight, const float* root,
              const float* gout, int64_t ldg, const float* gout_gather, int64_t ldgg, int32_t fout, float* gx,
              int64_t ldgx, float* gweight, float* groot, float* gbias, uint32_t flags, void* workspace,
              int64_t workspace_bytes, const float* x_chunk_rows, void* stream) {
    if (!g || !x || !weight || !gout || fin <= 0 || fout <= 0 || ldx < fin || ldg < fout || (gx && ldgx < fin) ||
        ((uintptr_t)x_chunk_rows & 15) != 0)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_layer_bwd: bad argument");
    if (!gout_gather) {
        if (gx && g->n_own != g->N)
            return fail(RGCN_ERR_INVALID_ARG, "rgcn_layer_bwd: a partitioned graph needs gout_gather (all nodes' rows) for dL/dx");
        gout_gather = gout;
        ldgg = ldg;
    } else if (ldgg < fout) {
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_layer_bwd: bad argument");
    }
    // rows of the owned nodes (ReLU mask): x holds all rows (pull) or just the owned ones (push)
    const float* x_own = g->push ? x : x + g->own_lo * ldx;
    const int64_t n_gat = g->fwd_gather_rows();
    if (g->push && gout_gather == gout && g->n_own != g->N)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_layer_bwd: a source-partitioned graph needs gout_gather (all nodes' rows)");
    cudaStream_t st = (cudaStream_t)stream;
    const bool relu = (flags & RGCN_F_RELU_IN) != 0;
    const bool mask = relu && !(flags & RGCN_F_NO_RELU_MASK);   // apply the ReLU mask to gx here
    const int kp = pad_dim(fin), np = pad_dim(fout);
    const bool big = (uint64_t)n_gat * (uint64_t)ldx >= (1ull << 32) ||
                     (uint64_t)g->N * (uint64_t)(gout_gather ? ldgg : ldg) >= (1ull << 32) ||
                     (uint64_t)g->n_own * (uint64_t)std::max<int64_t>(ldgx, kp) >= (1ull << 32);
    const bool simple = (flags & RGCN_F_FORCE_SIMPLE) || !kp || !np || big;
    const bool need_w = gweight || groot || gbias;
    if (g->push && simple)
        return fail(RGCN_ERR_UNSUPPORTED, "rgcn_layer_bwd: a source-partitioned graph runs on the tile kernels only (widths <= 64)");
    int rc;
    if (gweight) RGCN_CUDA(cudaMemsetAsync(gweight, 0, (size_t)g->R * fin * fout * 4, st));
    if (groot) RGCN_CUDA(cudaMemsetAsync(groot, 0, (size_t)fin * fout * 4, st));
    if (gbias) RGCN_CUDA(cudaMemsetAsync(gbias, 0, (size_t)fout * 4, st));
    if (simple) {
        if (need_w) {
            SimpleWGrad p{};
            p.brc = &g->brc[RGCN_BRC_FWD_REL];
            p.n_nodes = g->N; p.self_rel = g->R;
            p.feat = x; p.ldf = ldx; p.kin = fin;
            p.gout = gout; p.ldg = ldg; p.nout = fout;
            p.gweight = gweight; p.groot = groot; p.gbias = gbias; p.relu_in = relu;
            if ((rc = launch_simple_wgrad(p, st))) return rc;
        }
        if (gx) {
            RGCN_CUDA(cudaMemset2DAsync(gx, (size_t)ldgx * 4, 0, (size_t)fin * 4, (size_t)g->n_own, st));
            SimplePass p{};
            p.brc = &g->brc[RGCN_BRC_BWD];
            p.n_nodes = g->N; p.self_rel = g->R;
            p.feat = gout_gather; p.ldf = ldgg; p.kin = fout;
            p.weight = weight; p.root = root; p.bias = nullptr; p.transpose = true; p.w_rows = fin; p.w_cols = fout;
            p.out = gx; p.ldo = ldgx; p.nout = fin; p.relu_in = false;
            if ((rc = launch_simple_pass(p, st))) return rc;
            if (mask && (rc = launch_relu_mask(gx, ldgx, x_own, ldx, g->n_own, fin, st))) return rc;
        }
        return 0;
    }
    WsCarver ws(workspace, workspace_bytes);
    float* xaux = ws.take<float>((int64_t)g->brc[RGCN_BRC_FWD_REL].num_chunks * kp);
    float4* wtfrag = (float4*)ws.take<float>((int64_t)(g->R + 1) * kp * np * 2);
    float2* wtfrag2 = (float2*)ws.take<float>((int64_t)(g->R + 1) * kp * np);
    float* gaux = ws.take<float>((int64_t)g->brc[RGCN_BRC_BWD].num_chunks * np);
    // dx target: gx itself when its rows are whole aligned quads, or tightly packed odd-width rows that
    // the bulk-reduce scatter can address (Fin = 63); else a padded buffer + column copy
    const bool dx_v4ok = gx && etile_vec4_ok(gout_gather, ldgg, fout, nullptr);
    const bool packed = gx && !direct_target(gx, ldgx, fin) && dx_v4ok && etile_choice(np, dx_v4ok) &&
                        etile_packed_ok(gx, ldgx, fin, kp);
    // (a caller-padded gradient buffer — rows of ceil4(fin) or more floats, 16-byte aligned — is accumulated in place)
    const bool padded = gx && !packed && ldgx % 4 == 0 && ldgx >= ((fin + 3) & ~3) && ((uintptr_t)gx & 15) == 0;
    const bool direct = gx && (packed || padded || direct_target(gx, ldgx, fin));
    float* target = gx ? (direct ? gx : ws.take<float>((int64_t)g->n_own * kp)) : nullptr;
    float* wtc = (gx && kp == 64 && np == 64) ? ws.take<float>(wprep_tc_floats(g->R)) : nullptr;
    float* wt_scratch = (need_w && np == 16 && kp > np) ? ws.take<float>((int64_t)(g->R + 1) * 16 * kp) : nullptr;
    if (!ws.ok) return fail(RGCN_ERR_WORKSPACE, "rgcn_layer_bwd: workspace too small (see rgcn_layer_workspace_bytes)");
    // dL/dW of a layer that is wider on its input side (63 -> 16) runs on the TRANSPOSED structure: the narrow gout
    // rows are gathered (as dL/dx does), the wide x rows are read by owner id and stay in L2 (k_ewgrad_t)
    WGradPass wt{};
    wt.brc = &g->brc[RGCN_BRC_BWD];
    wt.n_nodes = g->N; wt.self_rel = g->R;
    wt.feat = x_own; wt.ldf = ldx; wt.kin = fin;
    wt.aux = gaux;
    wt.gout = gout_gather; wt.ldg = ldgg; wt.nout = fout;
    wt.gweight = gweight; wt.groot = groot; wt.gbias = gbias;
    wt.kp = kp; wt.np = np; wt.relu_in = relu;
    const bool wg_t = need_w && wt_scratch && etile_choice(kp, true) && ewgrad_t_ok(wt);
    bool bwd_pre_done = false;
    if (wg_t) {   // both sides read the chunk rows of gout: before the fork
        TilePass pre{};
        pre.brc = &g->brc[RGCN_BRC_BWD];
        pre.n_nodes = g->N;
        pre.feat = gout_gather; pre.ldf = ldgg; pre.kin = fout; pre.aux = gaux; pre.kp = np; pre.relu_in = false;
        if ((rc = launch_chunk_prepass(pre, st))) return rc;
        bwd_pre_done = true;
    }
    // dL/dW (+ its pre-pass) and the dL/dx chain are independent: with both requested the dL/dW pass runs
    // on the side stream, each side keeping to its share of every SM so that they are co-resident
    Fork fk(g, st);
    const bool ovl = need_w && gx && fk.available();
    cudaStream_t st_w = st;
    if (ovl) {
        if ((rc = fk.begin())) return rc;
        st_w = g->side;
    }
    if (wg_t) {
        wt.ctas_per_sm = ovl ? opts().wg_ctas : 0;
        if ((rc = launch_ewgrad_t_pass(wt, wt_scratch, g->num_sms, st_w))) return rc;
    } else if (need_w) {
        // chunk rows of x in the relation-major ordering, then dW / droot / dbias
        TilePass pre{};
        pre.brc = &g->brc[RGCN_BRC_FWD_REL];
        pre.n_nodes = n_gat;
        pre.feat = x; pre.ldf = ldx; pre.kin = fin; pre.aux = xaux; pre.kp = kp; pre.relu_in = relu;
        // FWD_REL shares FWD's chunk numbering: the rows the forward pass kept are these rows
        if (x_chunk_rows) xaux = const_cast<float*>(x_chunk_rows);
        else if ((rc = launch_chunk_prepass(pre, st_w))) return rc;
        WGradPass p{};
        p.brc = &g->brc[RGCN_BRC_FWD_REL];
        p.n_nodes = n_gat; p.self_rel = g->R;
        p.feat = x; p.ldf = ldx; p.kin = fin; p.aux = xaux;
        // gout rows are addressed by OWNER id: local (pull) or global (push: the all-gathered rows)
        p.gout = g->push ? gout_gather : gout;
        p.ldg = g->push ? ldgg : ldg;
        p.nout = fout;
        p.gweight = gweight; p.groot = groot; p.gbias = gbias;
        p.kp = kp; p.np = np; p.relu_in = relu;
        const bool v4ok = kp >= 32 && etile_vec4_ok(x, ldx, fin, xaux);
        p.vec4 = v4ok;
        p.ctas_per_sm = ovl ? opts().wg_ctas : 0;
        if ((rc = etile_choice(kp, v4ok) ? launch_ewgrad_pass(p, g->num_sms, st_w) : launch_wgrad_pass(p, g->num_sms, st_w))) return rc;
    }
    if (gx) {
        // dx: transposed structure, gathers gout rows (width fout), B = W^T : [np x kp]
        const bool v4ok = etile_vec4_ok(gout_gather, ldgg, fout, gaux);
        const bool et = etile_choice(np, v4ok);
        const bool v4 = et && v4ok;
        WPrep wp{weight, root, g->R, fin, fout, np, kp, true, v4, wtfrag, wtfrag2};
        if ((rc = launch_wprep(wp, st))) return rc;
        const int64_t tld = direct ? ldgx : kp;
        TilePass p{};
        p.brc = &g->brc[RGCN_BRC_BWD];
        p.n_nodes = g->N; p.self_rel = g->R;
        p.feat = gout_gather; p.ldf = ldgg; p.kin = fout;
        p.aux = gaux;
        p.wfrag = wtfrag;
        p.wfrag2 = wtfrag2;
        p.bias = nullptr; p.nbias = 0;
        p.out = target; p.ldo = tld; p.nout = direct ? (packed ? fin : ((fin + 3) & ~3)) : kp; p.tag_out = fin;
        p.kp = np; p.np = kp;
        p.relu_in = false;
        p.transposed = true;
        p.vec4 = v4;
        p.packed = packed && v4;
        p.out_rows = g->n_own;
        p.ctas_per_sm = ovl ? opts().dx_ctas : 0;
        if (!bwd_pre_done && (rc = launch_chunk_prepass(p, st))) return rc;
        if (et) {
            if ((rc = launch_selfloop_pass(p, g->own_lo, g->n_own, g->R, g->num_sms, st))) return rc;
            if (wtc && etile_tc_ok(p)) {
                if ((rc = launch_wprep_tc(weight, root, g->R, fin, fout, true, wtc, st))) return rc;
                if ((rc = launch_etile_tc(p, wtc, g->R, g->num_sms, st))) return rc;
            } else if ((rc = launch_etile_pass(p, g->num_sms, st))) {
                return rc;
            }
        } else {
            RGCN_CUDA(cudaMemsetAsync(target, 0, (size_t)g->n_own * tld * 4, st));
            if ((rc = launch_tile_pass(p, g->num_sms, st))) return rc;
        }
        if (!direct && (rc = launch_copy_cols(target, tld, gx, ldgx, g->n_own, fin, st))) return rc;
        if (mask && (rc = launch_relu_mask(gx, ldgx, x_own, ldx, g->n_own, fin, st))) return rc;
    }
    return fk.end();
}
}  // namespace

extern "C" int rgcn_set_option(int32_t option, int64_t value) {
    switch (option) {
        case RGCN_OPT_OVERLAP: opts().overlap = (int)value; return 0;
        case RGCN_OPT_OVERLAP_WGRAD_CTAS: opts().wg_ctas = (int)value; return 0;
        case RGCN_OPT_OVERLAP_DX_CTAS: opts().dx_ctas = (int)value; return 0;
        case RGCN_OPT_NVL_MODE: set_nvl_mode((int)value); return 0;
        default: return fail(RGCN_ERR_INVALID_ARG, "rgcn_set_option: unknown option");
    }
}

extern "C" int rgcn_layer_bwd(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin, const float* weight,
                              const float* root, const float* gout, int64_t ldg, const float* gout_gather,
                              int64_t ldgg, int32_t fout, float* gx, int64_t ldgx, float* gweight, float* groot,
                              float* gbias, uint32_t flags, void* workspace, int64_t workspace_bytes, void* stream) {
    return layer_bwd(g, x, ldx, fin, weight, root, gout, ldg, gout_gather, ldgg, fout, gx, ldgx, gweight, groot, gbias, flags,
                     workspace, workspace_bytes, nullptr, stream);
}

extern "C" int rgcn_layer_bwd_reuse(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin, const float* weight,
                                    const float* root, const float* gout, int64_t ldg, const float* gout_gather,
                                    int64_t ldgg, int32_t fout, float* gx, int64_t ldgx, float* gweight, float* groot,
                                    float* gbias, uint32_t flags, void* workspace, int64_t workspace_bytes,
                                    const float* x_chunk_rows, void* stream) {
    return layer_bwd(g, x, ldx, fin, weight, root, gout, ldg, gout_gather, ldgg, fout, gx, ldgx, gweight, groot, gbias, flags,
                     workspace, workspace_bytes, x_chunk_rows, stream);
}

extern "C" int rgcn_pad_rows(const float* src, int64_t lds, int32_t cols, float* dst, int64_t ldd, int64_t n,
                             void* stream) {
    if (!src || !dst || cols <= 0 || lds < cols || ldd < cols || n < 0)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_pad_rows: bad argument");
    return launch_pad_rows(src, lds, cols, dst, ldd, n, (cudaStream_t)stream);
}
