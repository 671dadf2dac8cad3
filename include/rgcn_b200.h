/* rgcn_b200.h — C ABI of the B200-native R-GCN message-passing engine (librgcn_b200.so).
 *
 * Drop-in boundary for ONE hot path of tiddoloos/Scaling-RGCN-training: the R-GCN layer
 * forward/backward that the reference reaches through torch_geometric.nn.RGCNConv
 * (reference model/layers.py:7,15-16 construct; :21,23,62,64,108,110 call; autograd backward
 * triggered at model/modelTrainer.py:66), the edge tensors that feed it
 * (graphs/graph.py:55-69) and the summary->original embedding map-gather that feeds its first
 * layer (model/embeddingTricks.py:8-49).  The reference is pure Python and has no FFI of its
 * own; these entry points are what a ctypes binding on the reference side would bind
 * (INTEGRATION.md shows that stub).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types, no exceptions cross the boundary.
 *   - every call returns int: 0 = ok, <0 = rgcn_status, >0 = cudaError_t.  rgcn_last_error()
 *     returns a thread-local message for the last non-zero return.
 *   - all data pointers are DEVICE pointers on the current CUDA device unless named host_*.
 *     Tensors are owned by the caller (PyTorch); the engine retains none of them past the call.
 *     The only engine-owned object is the opaque rgcn_graph.
 *   - `stream` is a cudaStream_t passed as void*; every launch goes to it; no hidden syncs in
 *     rgcn_layer_fwd/bwd/map_gather (rgcn_graph_create synchronises: it is a one-time build).
 *   - features/parameters are fp32, row-major, leading dimension given in elements.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef RGCN_B200_H
#define RGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGCN_B200_ABI_VERSION 6

typedef enum {
    RGCN_OK = 0,
    RGCN_ERR_INVALID_ARG = -1,   /* null pointer, negative size, bad enum */
    RGCN_ERR_INDEX_RANGE = -2,   /* src/dst outside [0,N) or type outside [0,R) */
    RGCN_ERR_UNSUPPORTED = -3,   /* shape not supported by any kernel */
    RGCN_ERR_WORKSPACE = -4      /* workspace too small */
} rgcn_status;

typedef struct rgcn_graph rgcn_graph;

/* Which blocked relational CSR (BRC) of a graph: owner = dst (forward), owner = src
 * (transposed, for dL/dx), or forward in pure relation-major order (for dL/dW). */
enum { RGCN_BRC_FWD = 0, RGCN_BRC_BWD = 1, RGCN_BRC_FWD_REL = 2 };

/* rgcn_graph_query keys */
enum {
    RGCN_Q_NUM_NODES = 0, RGCN_Q_NUM_EDGES = 1, RGCN_Q_NUM_RELATIONS = 2,
    RGCN_Q_NUM_SEGMENTS = 3, RGCN_Q_NUM_ENTRIES = 4, RGCN_Q_NUM_CHUNKS = 5,
    RGCN_Q_NUM_GROUPS = 6, RGCN_Q_NUM_BATCHES = 7, RGCN_Q_RANGE_NODES = 8,
    RGCN_Q_DEVICE_BYTES = 9, RGCN_Q_NUM_OWNED = 10, RGCN_Q_OWN_LO = 11, RGCN_Q_NUM_ENTRIES0 = 12,
    RGCN_Q_NUM_TILES = 13, RGCN_Q_NUM_TILES_NOSELF = 14, RGCN_Q_PUSH = 15
};

/* rgcn_graph_export array ids (element type in brackets) */
enum {
    RGCN_A_PERM = 0 /*i32[E+N]*/, RGCN_A_SEG_PTR = 1 /*i32[S+1]*/, RGCN_A_SEG_OWN = 2 /*i32[S]*/,
    RGCN_A_SEG_REL = 3 /*i32[S]*/, RGCN_A_SEG_PTR0 = 4 /*i32[S+1] before chunking; diff = multiplicity*/, RGCN_A_E_IDX = 5 /*u32[E3]*/,
    RGCN_A_E_W = 6 /*f32[E3]*/, RGCN_A_RAW_IDX = 7 /*i32[E+N]*/, RGCN_A_RAW_W = 8 /*f32[E+N]*/,
    RGCN_A_CHUNK_BEG = 9 /*i32[NC]*/, RGCN_A_CHUNK_END = 10 /*i32[NC]*/,
    RGCN_A_BAT_SEG0 = 11 /*i32[NB]*/, RGCN_A_BAT_INFO = 12 /*i32[NB]*/,
    RGCN_A_E_OWN = 13 /*i32[E3]*/, RGCN_A_TILE_E0 = 14 /*i32[NT]*/, RGCN_A_TILE_INFO = 15 /*i32[NT]*/,
    RGCN_A_CHUNK_OUT = 16 /*i32[NC]: chunk-matrix row of each chunk; FWD_REL only (it shares FWD's numbering), else empty*/
};

/* layer flags */
#define RGCN_F_RELU_IN     1u  /* apply max(0,.) to every gathered input row (fuses the reference's
                                  F.relu between the two layers, model/layers.py:22) */
#define RGCN_F_FORCE_SIMPLE 2u /* use the generic (any-shape, scalar-FMA) kernels */
#define RGCN_F_NO_RELU_MASK 4u /* backward with RGCN_F_RELU_IN: leave gx UNMASKED — the caller applies the ReLU
                                  mask itself (fused into the exchange of gx: rgcn_nvl_store_rows relu_pre) */

const char* rgcn_last_error(void);
int rgcn_abi_version(void);

/* Process-wide run-time options (no reference counterpart; the defaults are what bench.py measures).
 * RGCN_OPT_OVERLAP: 1 = passes of one layer call that do not depend on each other (dL/dW vs the dL/dx
 * chain, chunk pre-pass vs root pass) run concurrently on an engine-owned side stream, forked from and
 * joined back into the caller's stream inside the call (capturable in a CUDA graph); 0 = one stream.
 * RGCN_OPT_OVERLAP_*_CTAS: resident CTAs per SM each of the two concurrent backward passes keeps to.
 * RGCN_OPT_NVL_MODE: the NVLink exchange kernels use 1 = multimem.* with weak semantics (default; the rank
 * barrier orders them), 0 = relaxed.sys, 2 = plain peer pointers even when a multicast address is given. */
enum { RGCN_OPT_OVERLAP = 0, RGCN_OPT_OVERLAP_WGRAD_CTAS = 1, RGCN_OPT_OVERLAP_DX_CTAS = 2, RGCN_OPT_NVL_MODE = 3 };
int rgcn_set_option(int32_t option, int64_t value);

/* K0 — replaces nothing in the reference one-to-one: PyG re-derives `edge_type == r` masks on
 * every forward; the engine sorts once.  Inputs are the tensors of
 * graphs/graph.py:65-69 as they are: int64, possibly strided (element strides given).
 * range_nodes <= 0, split_threshold <= 0, chunk_size <= 0 pick defaults. */
int rgcn_graph_create(const int64_t* src, int64_t src_stride,
                      const int64_t* dst, int64_t dst_stride,
                      const int64_t* etype, int64_t etype_stride,
                      int64_t num_edges, int64_t num_nodes, int32_t num_relations,
                      int32_t range_nodes, int32_t split_threshold, int32_t chunk_size,
                      void* stream, rgcn_graph** out);
/* Destination-partitioned variant (multi-GPU, one process per GPU): this graph owns the nodes
 * [own_lo, own_hi).  Its forward structures hold the edges whose dst is owned, its transposed
 * structure the edges whose src is owned; owner ids are local (id - own_lo), gathered rows keep
 * GLOBAL ids, and the mean normalisers are those of the whole graph.  Layer calls then take x /
 * gout_gather with all N rows (the caller all-gathers them) and produce the owned rows only. */
int rgcn_graph_create_part(const int64_t* src, int64_t src_stride,
                           const int64_t* dst, int64_t dst_stride,
                           const int64_t* etype, int64_t etype_stride,
                           int64_t num_edges, int64_t num_nodes, int32_t num_relations,
                           int64_t own_lo, int64_t own_hi,
                           int32_t range_nodes, int32_t split_threshold, int32_t chunk_size,
                           void* stream, rgcn_graph** out);
/* Source-partitioned ("push") variant: the forward structures hold the edges whose SRC is owned, with
 * LOCAL gathered ids (src - own_lo) and GLOBAL owner ids (dst); the transposed structure is the one of
 * rgcn_graph_create_part.  rgcn_layer_fwd then takes x with the OWNED rows only and writes a PARTIAL output
 * with one row per node of the whole graph (root + bias on the owned rows, zero-started elsewhere): the caller
 * reduce-scatters the partials, so the wide layer-1 input never crosses the links.  rgcn_layer_bwd takes x with
 * the owned rows, gout_gather with all nodes' rows (used for dL/dW as well as dL/dx; `gout` is ignored) and
 * produces the owned rows of gx and partial parameter gradients.  Tile kernels only (widths <= 64). */
int rgcn_graph_create_push(const int64_t* src, int64_t src_stride,
                           const int64_t* dst, int64_t dst_stride,
                           const int64_t* etype, int64_t etype_stride,
                           int64_t num_edges, int64_t num_nodes, int32_t num_relations,
                           int64_t own_lo, int64_t own_hi,
                           int32_t range_nodes, int32_t split_threshold, int32_t chunk_size,
                           void* stream, rgcn_graph** out);
void rgcn_graph_destroy(rgcn_graph* g);
int rgcn_graph_query(const rgcn_graph* g, int32_t brc, int32_t key, int64_t* out);
int rgcn_graph_export(const rgcn_graph* g, int32_t brc, int32_t array, void* host_dst,
                      int64_t bytes, void* stream);

/* Workspace size (bytes) for one forward or backward call of an (fin -> fout) layer. */
int64_t rgcn_layer_workspace_bytes(const rgcn_graph* g, int32_t fin, int32_t fout, int32_t backward);

/* RGCNConv.forward (reference call sites model/layers.py:21,23):
 *   out[i] = sum_r mean_{e: type=r, dst=i} x[src_e] . W_r + x[i] . root + bias
 * weight [R, fin, fout] (basis form is expanded by the caller), root/bias nullable.
 * x has one row per node of the whole graph, out one row per OWNED node (all of them unless the
 * graph was created with rgcn_graph_create_part); out is fully overwritten. */
int rgcn_layer_fwd(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin,
                   const float* weight, const float* root, const float* bias,
                   float* out, int64_t ldo, int32_t fout, uint32_t flags,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* Autograd backward of the above (reference model/modelTrainer.py:66).  Any of
 * gx / gweight / groot / gbias may be null = not needed (-e_freeze / -w_grad False).
 * Non-null outputs are fully overwritten.  With RGCN_F_RELU_IN, x is the pre-activation and
 * gx is the gradient w.r.t. that pre-activation (ReLU mask applied).
 * gout holds the rows of the OWNED nodes; gout_gather the rows of ALL nodes (null = same tensor,
 * only valid for an unpartitioned graph).  gx / gweight / groot / gbias cover the owned nodes
 * (partial sums in the partitioned case: the caller all-reduces the parameter gradients). */
int rgcn_layer_bwd(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin,
                   const float* weight, const float* root,
                   const float* gout, int64_t ldg, const float* gout_gather, int64_t ldgg, int32_t fout,
                   float* gx, int64_t ldgx, float* gweight, float* groot, float* gbias,
                   uint32_t flags, void* workspace, int64_t workspace_bytes, void* stream);

/* Chunk-row reuse between forward and backward.  Long (relation, dst) segments are pre-reduced into
 * "chunk rows" ([num_chunks, pad(fin)] floats, pad = 16/32/64); dL/dW needs exactly the rows the
 * forward pass computed (same x, same flags).  rgcn_layer_fwd_keep writes them to the caller's
 * `chunk_rows` buffer (rgcn_layer_chunk_rows_bytes(g, fin) bytes, 16-byte aligned) instead of its
 * workspace; rgcn_layer_bwd_reuse takes them back (null = recompute, i.e. plain rgcn_layer_bwd).
 * x_mirror (nullable, [num_nodes, ld_mirror], ld_mirror % 4 == 0, 16-byte aligned): when the rows of x are
 * not 16-byte addressable (emb = 63) the engine gathers from this zero-padded copy instead, which it
 * writes itself — fused with the root/self-loop pass where the shape allows, so x is read once.  The
 * caller then passes the mirror (ldx = ld_mirror) as x to the backward call. */
int64_t rgcn_layer_chunk_rows_bytes(const rgcn_graph* g, int32_t fin);
int rgcn_layer_fwd_keep(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin,
                        const float* weight, const float* root, const float* bias,
                        float* out, int64_t ldo, int32_t fout, uint32_t flags,
                        void* workspace, int64_t workspace_bytes, float* chunk_rows,
                        float* x_mirror, int64_t ld_mirror, void* stream);
int rgcn_layer_bwd_reuse(const rgcn_graph* g, const float* x, int64_t ldx, int32_t fin,
                         const float* weight, const float* root,
                         const float* gout, int64_t ldg, const float* gout_gather, int64_t ldgg, int32_t fout,
                         float* gx, int64_t ldgx, float* gweight, float* groot, float* gbias,
                         uint32_t flags, void* workspace, int64_t workspace_bytes, const float* x_chunk_rows,
                         void* stream);

/* K5 — get_tensor_list + sum/concat/stack (reference model/embeddingTricks.py:8-49).
 * For summary s: row i of T_s = emb[s][idx[s][i]] if idx[s][i] >= 0 else fallback[s][i]
 * (the reference's torch.rand row).  host_* are HOST arrays of num_sums device pointers.
 * mode 0: out[N,F] = T_0 + T_1 + ... (that order);  1: out[N,S*F] concat;  2: out[S,N,F] stack. */
int rgcn_map_gather(const float* const* host_emb, const int32_t* const* host_idx,
                    const float* const* host_fallback, int32_t num_sums,
                    int64_t num_nodes, int32_t feat, int32_t mode, float* out, void* stream);

/* Device side of `evaluate` (reference model/evaluation.py:14-31).  For the n evaluated rows
 * pred[idx[r]] (width num_classes): mode 0 = one-hot of the arg-max (the CE path: softmax is
 * monotone), mode 1 = round() of the already-sigmoided output (BCE path).  Against the int64
 * multi-hot labels y [n, num_classes] it fills counts[0:C] = true positives per class,
 * counts[C:2C] = false positives, counts[2C:3C] = false negatives, counts[3C] = rows predicted
 * exactly — all sklearn's accuracy_score / f1_score(weighted, macro) need. */
int rgcn_eval_counts(const float* pred, int64_t ldp, int32_t num_classes, const int64_t* idx, int64_t n,
                     const int64_t* y, int32_t mode, int64_t* counts, void* stream);

/* One Adam step with L2 weight decay on one fp32 tensor, in one pass — the update rule of
 * torch.optim.Adam(lr, weight_decay) as the reference builds it (model/modelTrainer.py:44;
 * amsgrad / maximize off).  `step` counts from 1. */
int rgcn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream);

/* Same update with the step count read from DEVICE memory (*step_dev >= 1, incremented by the caller
 * on the same stream before the call): nothing step-dependent is baked into the launch, so a
 * training step that contains it can be captured once in a CUDA graph and replayed. */
int rgcn_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, const int64_t* step_dev,
                       void* stream);

/* Zero-padded, 16-byte addressable mirror of an odd-width feature matrix (e.g. the reference's
 * emb = 63): dst[r][c] = c < cols ? src[r][c] : 0 for c < ldd.  Passing the mirror (ldx = ldd) as x
 * to rgcn_layer_fwd / rgcn_layer_bwd lets the kernels use 128-bit row loads. */
int rgcn_pad_rows(const float* src, int64_t lds, int32_t cols, float* dst, int64_t ldd, int64_t num_rows,
                  void* stream);

/* Row exchanges of the partitioned layers as engine kernels over NVLink / NVSwitch peer memory (no reference
 * counterpart).  The buffers are symmetric allocations of all ranks (host side: torch.distributed.
 * _symmetric_memory); `*_multicast` is the NVSwitch multicast address of the buffer, or null with `host_peers`
 * = HOST array of the `num_peers` ranks' device pointers (peer-mapped) when the fabric has no multicast.
 * Ordering between ranks (store -> barrier -> read) is the caller's.
 * rgcn_nvl_store_rows: all-gather half — row r of src ([rows, cols], leading dimension lds) lands in row
 *   row0 + r of EVERY rank's buffer ([*, ldd], ldd % 4 == 0, columns >= cols zero) through one multimem.st
 *   per 16 bytes.  relu_pre != null fuses the ReLU backward of the inter-layer activation
 *   (reference model/layers.py:22): src[r][c] is zeroed where relu_pre[r][c] <= 0 first (and written back).
 * rgcn_nvl_reduce_rows: reduce-scatter / all-reduce half — dst[r][c] = sum over ranks of
 *   buffer_rank[row0 + r][c], c < width (width % 4 == 0), one multimem.ld_reduce.add.f32 per 16 bytes:
 *   the switch adds the ranks' copies in flight. */
int rgcn_nvl_store_rows(float* src, int64_t lds, int32_t cols, const float* relu_pre, int64_t ld_pre,
                        float* dst_multicast, void* const* host_peers, int32_t num_peers, int64_t ldd,
                        int64_t row0, int64_t rows, void* stream);
int rgcn_nvl_reduce_rows(const float* src_multicast, void* const* host_peers, int32_t num_peers, int64_t lds,
                         int64_t row0, int64_t rows, float* dst, int64_t ldd, int32_t width, void* stream);

/* Sparse forms of the two exchanges (peer pointers only).  Which ranks reach a row is a property of the partitioned
 * graph: a rank's partial output is non-zero only in the rows its own edges reach, and it reads gout only in those
 * rows.  peer_mask[r]: bit p = rank p reaches row row0 + r (this rank's owned rows; its own bit set).
 * reduce_rows_sparse: dst[r] = sum over the set bits p (ascending) of peer_p[(row0 + r) * lds ..] — bit-identical to
 * the dense peer-pointer sum when the skipped rows are zero.  store_rows_sparse: rgcn_nvl_store_rows, each row
 * written only into the copies of the ranks whose bit is set. */
int rgcn_nvl_reduce_rows_sparse(void* const* host_peers, int32_t num_peers, const uint32_t* peer_mask, int64_t lds,
                                int64_t row0, int64_t rows, float* dst, int64_t ldd, int32_t width, void* stream);
int rgcn_nvl_store_rows_sparse(float* src, int64_t lds, int32_t cols, const float* relu_pre, int64_t ld_pre,
                               void* const* host_peers, int32_t num_peers, const uint32_t* peer_mask, int64_t ldd,
                               int64_t row0, int64_t rows, void* stream);

/* Dense contractions of the path on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMA operand
 * loads, accumulators in tensor memory), fp32-faithful through the error-compensated 3xTF32 split:
 *   C[m, n_store] = act( A[m, k] . W[n, k]^T + bias )       act 0 = identity, 1 = tanh,
 *                                                            2 = multiply by (1 - aux^2), the tanh backward
 *                                                            (aux [m, ldaux]: the saved tanh output; else null)
 * — the two nn.Linear calls of the reference's MLP transfer head (model/layers.py:105-107).
 * rgcn_gemm_prepack splits W ([n, k], leading dimension ldw; transpose = 1: W is given as [k, n]) once per call
 * into hi / lo parts, zero-padded to [n_pad, k_pad] (n_pad % 16 == 0, <= 256; k_pad % 32 == 0).
 * rgcn_gemm3x_tf32: rows of A (lda % 4 == 0, 16-byte aligned base) and of C (ldc % 4 == 0) must be 16-byte
 * addressable; columns [n, n_store) of C receive act(0) = 0 (bias_padded: [n_pad], zero beyond n, or null),
 * so a 63-wide result can be written straight into the zero-padded 64-wide rows the first layer gathers. */
int rgcn_gemm_prepack(const float* w, int64_t ldw, int32_t n, int32_t k, int32_t n_pad, int32_t k_pad,
                      int32_t transpose, float* w_hi, float* w_lo, void* stream);
int rgcn_gemm3x_tf32(const float* a, int64_t lda, int64_t m, int32_t k, const float* w_hi, const float* w_lo,
                     int32_t n_pad, int32_t k_pad, const float* bias_padded, int32_t act, const float* aux,
                     int64_t ldaux, float* c, int64_t ldc, int32_t n_store, void* stream);

/* c[n1, n2] = a[rows, n1]^T . b[rows, n2] (fp32-faithful 3xTF32 on tcgen05, split over the rows, c zeroed inside):
 * the parameter-gradient reductions of the transfer heads (autograd of reference model/layers.py:59-61, 103-107).
 * Rows of a and b 16-byte addressable (ld % 4 == 0, aligned base); n2 <= 192; ldc >= n2.
 * a_colsum [n1] / b_colsum [n2] (nullable): the column sums of a / b over the rows — the bias gradients that go with
 * c — accumulated while the operands pass through shared memory. */
int rgcn_gram3x_tf32(const float* a, int64_t lda, int32_t n1, const float* b, int64_t ldb, int32_t n2, int64_t rows,
                     float* c, int64_t ldc, float* a_colsum, float* b_colsum, void* stream);

/* Per-node core of the attention transfer head (reference model/layers.py:59-61: nn.MultiheadAttention over the
 * num_sums stacked summary embeddings, q = k = v, only attn_output[0] kept).  q [N, heads*head_dim] = the projected
 * query rows of summary 0 (bias added, unscaled); kv [num_sums*N, >= v_offset + heads*head_dim] = projected keys
 * (columns 0..) and values (columns v_offset..), row s*N + b for summary s (v_offset = ceil4(heads*head_dim) and
 * 16-byte addressable rows everywhere select the shared-memory-staged kernels); keep_scaled (nullable) [N, heads, num_sums] = dropout keep mask times 1/(1-p).
 * fwd: probs [N, heads, num_sums] = softmax_s(q.k_s / sqrt(head_dim)) (before dropout), o [N, heads*head_dim].
 * bwd: gq, gkv from dL/do (go); every element of gq / gkv is written. */
int rgcn_attn_head_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, int64_t v_offset, int32_t num_sums,
                       int64_t num_nodes, int32_t heads, int32_t head_dim, const float* keep_scaled, float* probs,
                       float* o, int64_t ldo, void* stream);
int rgcn_attn_head_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, int64_t v_offset, int32_t num_sums,
                       int64_t num_nodes, int32_t heads, int32_t head_dim, const float* keep_scaled,
                       const float* probs, const float* go, int64_t ldgo, float* gq, int64_t ldgq, float* gkv,
                       int64_t ldgkv, void* stream);

/* Instrumentation (no reference counterpart).  rgcn_kernel_launch_count: engine kernels launched
 * by this process so far.  rgcn_profile_enable(1): every pass launch is bracketed by a CUDA-event
 * pair on its own stream; rgcn_profile_collect synchronises those events, returns up to
 * max_records (tag, dims[2], milliseconds) records and clears the log.  Tags: 1 weight-fragment
 * prep, 2 chunk pre-pass, 3 forward tile pass, 4 dL/dx tile pass, 5 dL/dW pass, 6 column copy,
 * 7 ReLU mask, 8 generic kernels, 9 map gather, 10 self-loop (root + bias) pass, 11 / 12 the NVLink
 * exchange kernels (multicast store / load-reduce), 13 the tcgen05 dense contraction.
 * dims = (gathered width, output width). */
int64_t rgcn_kernel_launch_count(void);
int rgcn_profile_enable(int32_t on);
int rgcn_profile_collect(int32_t* tags, int32_t* dims, float* ms, int32_t max_records, int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* RGCN_B200_H */
