// nvl_comm.cu — the partitioned layers' row exchanges as the engine's OWN kernels over NVLink / NVSwitch
// peer memory (no reference counterpart: the reference is single-process, model/modelTrainer.py:16).
//
// The buffers are symmetric allocations (same size on every rank, peer-mapped; torch.distributed.
// _symmetric_memory on the host side) with an NVSwitch MULTICAST address:
//   all-gather      each rank stores its own rows ONCE to the multicast address (multimem.st): the switch
//                   replicates them into every rank's copy — egress = the shard, not (P-1) x the shard
//   reduce-scatter  each rank's layer kernels accumulate their partial output straight into their own copy;
//                   after a barrier a rank reads ITS rows through the multicast address with
//                   multimem.ld_reduce.add.f32: the switch sums the P copies in flight (NVLS)
//   all-reduce      the same load-reduce over the whole (small) buffer of parameter gradients
// Without multicast the same entry points take the peer pointers and loop over them.
// Ordering between ranks (store -> barrier -> read) is the caller's: rgcn_b200/partition.py NvlComm.
#include "common.cuh"

namespace rgcn {
// rgcn_set_option(RGCN_OPT_NVL_MODE): 0 = multimem with relaxed.sys semantics, 1 = multimem weak (default),
// 2 = peer pointers even when a multicast address is given
int g_nvl_mode = 1;
namespace {

constexpr int MAX_PEERS = 16;
struct PeerPtrs {
    float* p[MAX_PEERS];
    int n;   // 0: use the multicast address
};

__device__ __forceinline__ void mc_store4(float* addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void mc_store4_weak(float* addr, float4 v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 mc_load_reduce4_weak(const float* addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ float4 mc_load_reduce4(const float* addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(addr)
                 : "memory");
    return v;
}

// dst[(row0 + r) * ldd + c] = c < cols ? f(src[r * lds + c]) : 0   for every rank's copy of dst (ldd % 4 == 0).
// MASK: f(v) = pre[r * ldp + c] > 0 ? v : 0 — the ReLU backward of the inter-layer activation fused into the
// exchange of its result (and written back to src, which is the caller's gradient tensor).
// peer_mask (peer-pointer form only, nullable): bit p of peer_mask[r] = rank p reads row row0 + r at all (a property
// of the partitioned graph, NvlComm.set_touched) — the row is stored only into those ranks' copies.
template <bool MASK, bool WEAK>
__global__ void __launch_bounds__(256) k_nvl_store_rows(float* __restrict__ src, int64_t lds, int cols,
                                                        const float* __restrict__ pre, int64_t ldp, float* dst_mc,
                                                        PeerPtrs peers, const uint32_t* __restrict__ peer_mask, int64_t ldd,
                                                        int64_t row0, int64_t rows) {
    const int qpr = (int)(ldd >> 2);   // quads per row
    const int64_t total = rows * qpr;
    const bool vec = (lds & 3) == 0 && (cols & 3) == 0 && ((uintptr_t)src & 15) == 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / qpr;
        const int c = (int)(i - r * qpr) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float* s = src + r * lds + c;
        if (vec && c < cols) {
            v = *reinterpret_cast<const float4*>(s);
        } else {
            if (c < cols) v.x = s[0];
            if (c + 1 < cols) v.y = s[1];
            if (c + 2 < cols) v.z = s[2];
            if (c + 3 < cols) v.w = s[3];
        }
        if (MASK) {
            const float* q = pre + r * ldp + c;
            if (c < cols && !(q[0] > 0.f)) v.x = 0.f;
            if (c + 1 < cols && !(q[1] > 0.f)) v.y = 0.f;
            if (c + 2 < cols && !(q[2] > 0.f)) v.z = 0.f;
            if (c + 3 < cols && !(q[3] > 0.f)) v.w = 0.f;
            if (c < cols) s[0] = v.x;
            if (c + 1 < cols) s[1] = v.y;
            if (c + 2 < cols) s[2] = v.z;
            if (c + 3 < cols) s[3] = v.w;
        }
        const int64_t off = (row0 + r) * ldd + c;
        if (peers.n == 0) {
            if (WEAK) mc_store4_weak(dst_mc + off, v);
            else mc_store4(dst_mc + off, v);
        } else {
            const uint32_t m = peer_mask ? __ldg(peer_mask + r) : 0xffffffffu;
            for (int p = 0; p < peers.n; ++p)
                if ((m >> p) & 1u) *reinterpret_cast<float4*>(peers.p[p] + off) = v;
        }
    }
    __threadfence_system();
}

// dst[r * ldd + c] = sum over ranks of src_rank[(row0 + r) * lds + c]   (width % 4 == 0)
// U independent load-reduces per thread are in flight before the first store (the round trip goes through the
// switch: latency, not issue, bounds a single request stream)
template <bool WEAK, int U>
__global__ void __launch_bounds__(256) k_nvl_reduce_rows(const float* src_mc, PeerPtrs peers, int64_t lds, int64_t row0,
                                                         int64_t rows, float* __restrict__ dst, int64_t ldd, int width) {
    const int qpr = width >> 2;
    const int64_t total = rows * qpr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
        float4 v[U];
        int64_t doff[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            doff[u] = -1;
            if (i >= total) continue;
            const int64_t r = i / qpr;
            const int c = (int)(i - r * qpr) * 4;
            const int64_t off = (row0 + r) * lds + c;
            doff[u] = r * ldd + c;
            if (peers.n == 0) {
                v[u] = WEAK ? mc_load_reduce4_weak(src_mc + off) : mc_load_reduce4(src_mc + off);
            } else {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = 0; p < peers.n; ++p) {
                    float4 t;   // (not through the read-only path: the peers' copies were written during this launch sequence)
                    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                 : "l"(peers.p[p] + off)
                                 : "memory");
                    v[u].x += t.x, v[u].y += t.y, v[u].z += t.z, v[u].w += t.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (doff[u] >= 0) *reinterpret_cast<float4*>(dst + doff[u]) = v[u];
    }
}

__device__ __forceinline__ float4 ld_sys4(const float* addr) {
    float4 t;   // (not through the read-only path: the peers' copies were written during this launch sequence)
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                 : "l"(addr)
                 : "memory");
    return t;
}

// Sparse reduce-scatter: a rank's partial output is non-zero only in the rows its own edges reach, and which ranks
// reach an owned row is a property of the graph (peer_mask[r], bit p = rank p touches row row0 + r; built once by
// NvlComm.set_touched).  Only those copies are read: dst[r] = sum over set bits p (ascending) of peer_p[row0 + r].
// All NP (predicated) loads of a quad are issued before the first add.
template <int NP>
__global__ void __launch_bounds__(256) k_nvl_reduce_sparse(PeerPtrs peers, const uint32_t* __restrict__ peer_mask, int64_t lds,
                                                           int64_t row0, int64_t rows, float* __restrict__ dst, int64_t ldd,
                                                           int width) {
    constexpr int U = NP <= 4 ? 4 : 2;
    const int qpr = width >> 2;
    const int64_t total = rows * qpr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
        float4 t[U][NP];
        int64_t doff[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            doff[u] = -1;
            if (i >= total) continue;
            const int64_t r = i / qpr;
            const int c = (int)(i - r * qpr) * 4;
            const int64_t off = (row0 + r) * lds + c;
            doff[u] = r * ldd + c;
            const uint32_t m = __ldg(peer_mask + r);
#pragma unroll
            for (int p = 0; p < NP; ++p)
                t[u][p] = (p < peers.n && ((m >> p) & 1u)) ? ld_sys4(peers.p[p] + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (doff[u] < 0) continue;
            float4 v = t[u][0];
#pragma unroll
            for (int p = 1; p < NP; ++p) v.x += t[u][p].x, v.y += t[u][p].y, v.z += t[u][p].z, v.w += t[u][p].w;
            *reinterpret_cast<float4*>(dst + doff[u]) = v;
        }
    }
}

int make_peers(void* const* host_peers, int32_t num_peers, PeerPtrs* out) {
    out->n = 0;
    if (!host_peers || num_peers <= 0) return 0;
    if (num_peers > MAX_PEERS) return fail(RGCN_ERR_UNSUPPORTED, "rgcn_nvl: more than 16 peers");
    for (int i = 0; i < num_peers; ++i) out->p[i] = (float*)host_peers[i];
    out->n = num_peers;
    return 0;
}

// rgcn_set_option(RGCN_OPT_NVL_MODE): 0 = multimem with relaxed.sys semantics, 1 = multimem weak (default),
// 2 = peer pointers even when a multicast address is given
int grid_for(int64_t items, int num_sms) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((items + 255) / 256, (int64_t)num_sms * 8));
}

}  // namespace
}  // namespace rgcn

using namespace rgcn;

namespace rgcn {
void set_nvl_mode(int m) { g_nvl_mode = m; }
}  // namespace rgcn

namespace rgcn {
namespace {
int store_rows_impl(float* src, int64_t lds, int32_t cols, const float* relu_pre, int64_t ld_pre, float* dst_multicast,
                    void* const* host_peers, int32_t num_peers, const uint32_t* peer_mask, int64_t ldd, int64_t row0,
                    int64_t rows, void* stream) {
    if (!src || cols <= 0 || lds < cols || ldd < cols || (ldd & 3) || rows < 0 || row0 < 0 ||
        (!dst_multicast && num_peers <= 0) || ((uintptr_t)dst_multicast & 15) || (relu_pre && ld_pre < cols))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_nvl_store_rows: bad argument");
    if (rows == 0) return 0;
    PeerPtrs pp;
    const bool use_mc = dst_multicast && !peer_mask && !(g_nvl_mode == 2 && host_peers && num_peers > 0);
    int rc = make_peers(host_peers, use_mc ? 0 : num_peers, &pp);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = grid_for(rows * (ldd >> 2), sms);
    ProfScope prof(TAG_NVL_STORE, cols, (int)ldd, (cudaStream_t)stream);
    note_launch(1);
    cudaStream_t st = (cudaStream_t)stream;
    const bool weak = g_nvl_mode != 0;
    if (relu_pre) {
        if (weak) k_nvl_store_rows<true, true><<<grid, 256, 0, st>>>(src, lds, cols, relu_pre, ld_pre, dst_multicast, pp, peer_mask, ldd, row0, rows);
        else k_nvl_store_rows<true, false><<<grid, 256, 0, st>>>(src, lds, cols, relu_pre, ld_pre, dst_multicast, pp, peer_mask, ldd, row0, rows);
    } else {
        if (weak) k_nvl_store_rows<false, true><<<grid, 256, 0, st>>>(src, lds, cols, nullptr, 0, dst_multicast, pp, peer_mask, ldd, row0, rows);
        else k_nvl_store_rows<false, false><<<grid, 256, 0, st>>>(src, lds, cols, nullptr, 0, dst_multicast, pp, peer_mask, ldd, row0, rows);
    }
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace
}  // namespace rgcn

extern "C" int rgcn_nvl_store_rows(float* src, int64_t lds, int32_t cols, const float* relu_pre, int64_t ld_pre,
                                   float* dst_multicast, void* const* host_peers, int32_t num_peers, int64_t ldd,
                                   int64_t row0, int64_t rows, void* stream) {
    return store_rows_impl(src, lds, cols, relu_pre, ld_pre, dst_multicast, host_peers, num_peers, nullptr, ldd, row0, rows,
                           stream);
}

extern "C" int rgcn_nvl_store_rows_sparse(float* src, int64_t lds, int32_t cols, const float* relu_pre, int64_t ld_pre,
                                          void* const* host_peers, int32_t num_peers, const uint32_t* peer_mask, int64_t ldd,
                                          int64_t row0, int64_t rows, void* stream) {
    if (!peer_mask || !host_peers || num_peers <= 0)
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_nvl_store_rows_sparse: needs peer pointers and a peer mask");
    return store_rows_impl(src, lds, cols, relu_pre, ld_pre, nullptr, host_peers, num_peers, peer_mask, ldd, row0, rows, stream);
}

extern "C" int rgcn_nvl_reduce_rows(const float* src_multicast, void* const* host_peers, int32_t num_peers, int64_t lds,
                                    int64_t row0, int64_t rows, float* dst, int64_t ldd, int32_t width, void* stream) {
    if (!dst || width <= 0 || (width & 3) || lds < width || ldd < width || (lds & 3) || (ldd & 3) || rows < 0 || row0 < 0 ||
        (!src_multicast && num_peers <= 0) || ((uintptr_t)src_multicast & 15) || ((uintptr_t)dst & 15))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_nvl_reduce_rows: bad argument");
    if (rows == 0) return 0;
    PeerPtrs pp;
    const bool use_mc = src_multicast && !(g_nvl_mode == 2 && host_peers && num_peers > 0);
    int rc = make_peers(host_peers, use_mc ? 0 : num_peers, &pp);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = grid_for((rows * (width >> 2) + 3) / 4, sms);
    ProfScope prof(TAG_NVL_REDUCE, width, (int)std::min<int64_t>(rows, 1 << 30), (cudaStream_t)stream);
    note_launch(1);
    if (g_nvl_mode != 0)
        k_nvl_reduce_rows<true, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(src_multicast, pp, lds, row0, rows, dst, ldd, width);
    else
        k_nvl_reduce_rows<false, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(src_multicast, pp, lds, row0, rows, dst, ldd, width);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int rgcn_nvl_reduce_rows_sparse(void* const* host_peers, int32_t num_peers, const uint32_t* peer_mask, int64_t lds,
                                           int64_t row0, int64_t rows, float* dst, int64_t ldd, int32_t width, void* stream) {
    if (!dst || !peer_mask || !host_peers || num_peers <= 0 || width <= 0 || (width & 3) || lds < width || ldd < width ||
        (lds & 3) || (ldd & 3) || rows < 0 || row0 < 0 || ((uintptr_t)dst & 15))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_nvl_reduce_rows_sparse: bad argument");
    if (rows == 0) return 0;
    PeerPtrs pp;
    int rc = make_peers(host_peers, num_peers, &pp);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(TAG_NVL_REDUCE, width, (int)std::min<int64_t>(rows, 1 << 30), st);
    note_launch(1);
    const int64_t quads = rows * (width >> 2);
    if (num_peers <= 2) k_nvl_reduce_sparse<2><<<grid_for((quads + 3) / 4, sms), 256, 0, st>>>(pp, peer_mask, lds, row0, rows, dst, ldd, width);
    else if (num_peers <= 4) k_nvl_reduce_sparse<4><<<grid_for((quads + 3) / 4, sms), 256, 0, st>>>(pp, peer_mask, lds, row0, rows, dst, ldd, width);
    else if (num_peers <= 8) k_nvl_reduce_sparse<8><<<grid_for((quads + 1) / 2, sms), 256, 0, st>>>(pp, peer_mask, lds, row0, rows, dst, ldd, width);
    else k_nvl_reduce_sparse<16><<<grid_for((quads + 1) / 2, sms), 256, 0, st>>>(pp, peer_mask, lds, row0, rows, dst, ldd, width);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
