// gram_tc.cu — the parameter-gradient reductions of the transfer heads on the 5th-generation tensor cores:
//
//   C[n1, n2] = A[M, n1]^T . B[M, n2]          fp32 in, fp32 out, fp32-faithful (3xTF32), M = all nodes
//
// (reference model/layers.py:59-61, 103-107 under autograd: dL/dW of lin1 / lin2 / in_proj / out_proj — one
// [137 x 189] or [126 x 63] result reduced over 1.67 M .. 5 M rows; torch runs them as cuBLAS SIMT SGEMM.)
// Both operands are row-major with the REDUCTION index as the row, i.e. MN-major in tensor-core terms.  For 32-bit
// MN-major operands the tensor core takes one shared-memory layout only, the 128-byte swizzle with a 32-byte base
// (descriptor layout type SWIZZLE_128B_BASE32B: atoms of 32 floats along M/N x 4 rows along K, the 32-byte chunks of a
// row XORed with the row index), and the TMA unit writes exactly that (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a box
// of 32 columns x KB rows (KB = 16 or 32 rows per stage) lands as KB / 4 such atoms 512 B apart (SBO), the boxes of
// one tile sit one box (LBO = 128 x KB bytes) apart.  (The plain 128-byte swizzle is accepted and returns zeros.)
// Split-K: the row range is dealt over the CTAs.  The tensor core TRUNCATES when it adds into a tensor-memory
// accumulator, which over thousands of rows becomes a bias (measured 3e-5 relative over 4 300 rows; cuBLAS fp32: 3e-6),
// so an accumulator only ever holds 256 rows: it is then drained with round-to-nearest fp32 adds into a [BN x 128]
// shared-memory tile while the MMAs continue in the second accumulator; the tile is added to C at the end.
//
//   warp 0      TMA producer   4 + BN/32 boxes per KB-row stage (SASS UTMALDG), mbarrier tx
//   warp 1      MMA issuer     per stage KB/8 K steps x 3 tcgen05.mma.kind::tf32 (A_hi.B_hi, A_lo.B_hi, A_hi.B_lo), both
//                              operand descriptors MN-major; tcgen05.commit frees the stage / publishes an accumulator
//   warp 2      TMEM allocator (2 x 256 columns)
//   warps 4-11  splitter       BOTH operands are activations: each tile is split in shared memory into
//                              hi = rn_tf32(v) (in place) and lo = rn_tf32(v - hi); the column sums of A and B (the
//                              bias gradients that go with C) are accumulated on the way
//   warps 12-15 drain          tcgen05.ld -> += shared-memory tile; at the end red.global.add.f32 into C
#include <cuda.h>

#include "common.cuh"
#include "device_utils.cuh"

namespace rgcn {
namespace {

constexpr int GM = 128;                    // A columns per CTA (the M dimension of the MMA)
// rows (reduction index) per stage: 16 for wide operands (the drain tile leaves room for three 40 KB stages), 32 for
// BN <= 96 (half as many TMA boxes and barrier round trips per byte)
template <int KB>
struct Cfg {
    static constexpr int BOX = 32 * KB * 4;            // one TMA box: 32 columns x KB rows
    static constexpr int A_TILE = (GM / 32) * BOX;
    static constexpr int DRAIN_STAGES = 256 / KB;      // stages (256 rows) accumulated in tensor memory between two drains
};
constexpr int NUM_THREADS = 512;
constexpr int ACC_COLS = 256, ACC_STAGES = 2;
constexpr int MAX_BN = 192;                // the shared-memory tile is BN x 128 floats

struct GramArgs {
    int64_t rows;          // M
    int64_t rows_per_cta;  // multiple of KB
    int n1, n2;            // result shape
    int BN;                // n2 rounded up to 32
    float* C;
    int64_t ldc;
    int stages;
    float* a_colsum;       // nullable [n1]: sum over the rows of every column of A (the bias gradient that goes with C)
    float* b_colsum;       // nullable [n2]
};

__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t a = smem_u32(b);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// MN-major 32-bit operand: atoms of 32 floats (M/N) x 4 rows (K); atoms along M/N are LBO apart (one TMA box each),
// 4-row groups along K are SBO = 512 B apart (one tf32 instruction, K = 8, consumes two groups)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, int box_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)(box_bytes >> 4) << 16;               // leading byte offset: next 32-column atom
    d |= (uint64_t)(512 >> 4) << 32;                     // stride byte offset: next 4-row group
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)1 << 61;                              // SWIZZLE_128B_BASE32B
    return d;
}
// kind::tf32, fp32 accumulate, A and B MN-major (bits 15 / 16), M = 128
__device__ __forceinline__ uint32_t umma_idesc_mn(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int KB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_gram3x(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GramArgs g) {
    constexpr int BOX = Cfg<KB>::BOX, A_TILE = Cfg<KB>::A_TILE, DRAIN_STAGES = Cfg<KB>::DRAIN_STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int b_tile = (g.BN / 32) * BOX;
    const int stage_bytes = 2 * A_TILE + 2 * b_tile;          // A_hi | A_lo | B_hi | B_lo
    float* acc_s = reinterpret_cast<float*>(sm + (size_t)g.stages * stage_bytes);   // [BN][128]: column-major partial of C
    uint64_t* bars = reinterpret_cast<uint64_t*>(acc_s + (size_t)g.BN * GM);
    uint64_t* full = bars;                       // TMA bytes landed          [stages]
    uint64_t* split = bars + g.stages;           // tiles split into hi / lo  [stages]
    uint64_t* empty = bars + 2 * g.stages;       // MMAs of the stage retired [stages]
    uint64_t* acc_full = bars + 3 * g.stages;    // accumulator complete      [ACC_STAGES]
    uint64_t* acc_empty = acc_full + ACC_STAGES; // accumulator drained       [ACC_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    float* colsum_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [GM + MAX_BN]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.y;                   // which 128 columns of A
    const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_cta;
    const int64_t r1 = min(g.rows, r0 + g.rows_per_cta);
    const int num_kb = r0 < r1 ? (int)((r1 - r0 + KB - 1) / KB) : 0;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(split + s, 8);             // one arrival per splitter warp
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(acc_full + a, 1);
            mbar_init(acc_empty + a, 4);         // one arrival per drain warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(ACC_COLS * ACC_STAGES)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < GM + MAX_BN; i += NUM_THREADS) colsum_s[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty + s, ph ^ 1);
                const uint32_t st = base + (uint32_t)(s * stage_bytes);
                const int row = (int)(r0 + (int64_t)kb * KB);
                mbar_expect_tx(full + s, (uint32_t)(A_TILE + b_tile));
                for (int j = 0; j < GM / 32; ++j) tma_load_2d(st + j * BOX, &map_a, mb * GM + 32 * j, row, full + s);
                for (int j = 0; j < g.BN / 32; ++j) tma_load_2d(st + 2 * A_TILE + j * BOX, &map_b, 32 * j, row, full + s);
                if (++s == g.stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_mn(g.BN);
        int s = 0, a = 0;
        uint32_t ph = 0, aph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int in_acc = kb % DRAIN_STAGES;       // stages already in this accumulator
            if (in_acc == 0) {
                mbar_wait(acc_empty + a, aph ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait(split + s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool last = in_acc == DRAIN_STAGES - 1 || kb == num_kb - 1;
            if (lane == 0) {
                const uint32_t tmem_d = tmem_base + (uint32_t)(a * ACC_COLS);
                const uint32_t st = base + (uint32_t)(s * stage_bytes);
                const uint32_t a_hi = st, a_lo = st + A_TILE, b_hi = st + 2 * A_TILE, b_lo = b_hi + b_tile;
#pragma unroll
                for (int k = 0; k < KB / 8; ++k) {       // tf32: 8 rows per instruction = two 512-byte groups
                    const uint32_t ko = (uint32_t)(k * 1024);
                    umma_tf32(tmem_d, umma_desc_mn(a_hi + ko, BOX), umma_desc_mn(b_hi + ko, BOX), idesc, (in_acc | k) ? 1u : 0u);
                    umma_tf32(tmem_d, umma_desc_mn(a_lo + ko, BOX), umma_desc_mn(b_hi + ko, BOX), idesc, 1u);
                    umma_tf32(tmem_d, umma_desc_mn(a_hi + ko, BOX), umma_desc_mn(b_lo + ko, BOX), idesc, 1u);
                }
                umma_commit(empty + s);
                if (last) umma_commit(acc_full + a);
            }
            __syncwarp();
            if (last && ++a == ACC_STAGES) {
                a = 0;
                aph ^= 1;
            }
            if (++s == g.stages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== splitter (8 warps): hi = rn_tf32(v) in place, lo = rn_tf32(v - hi), for both tiles =====
        const int tid = threadIdx.x - 128;       // 0..255
        // column sums ride on the split: chunk q = tid + 256 i of a tile is (box tid / 128 + 2 i, row (tid % 128) / 8,
        // 16-byte chunk tid % 8); the TMA swizzle XORs the 32-byte chunk index with row % 4, so the four floats of the
        // chunk are columns 32 box + 8 ((tid % 8) / 2 ^ (row & 3)) + 4 (tid & 1) .. + 3 — fixed per thread and i
        const bool want_a = g.a_colsum != nullptr, want_b = g.b_colsum != nullptr && mb == 0;
        constexpr int CPB = 8 * KB;                      // 16-byte chunks per box
        const int row_in_box = (tid % CPB) >> 3;
        const int col_in_box = 8 * (((tid & 7) >> 1) ^ (row_in_box & 3)) + 4 * (tid & 1);
        constexpr int A_IT = A_TILE / 16 / 256, B_IT = 3;   // B: at most 768 chunks (6 boxes of 2 KB or 3 boxes of 4 KB)
        float4 sa[A_IT], sb[B_IT];
#pragma unroll
        for (int i = 0; i < A_IT; ++i) sa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < B_IT; ++i) sb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(full + s, ph);
            uint8_t* st = sm + (size_t)s * stage_bytes;
            auto split_one = [&](float4* hi, float4* lo, int q, float4& acc) {
                float4 v = hi[q], h, l;
                h.x = rn_tf32(v.x), h.y = rn_tf32(v.y), h.z = rn_tf32(v.z), h.w = rn_tf32(v.w);
                l.x = rn_tf32(v.x - h.x), l.y = rn_tf32(v.y - h.y), l.z = rn_tf32(v.z - h.z), l.w = rn_tf32(v.w - h.w);
                hi[q] = h;
                lo[q] = l;
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            };
#pragma unroll
            for (int i = 0; i < A_IT; ++i)
                split_one(reinterpret_cast<float4*>(st), reinterpret_cast<float4*>(st + A_TILE), tid + 256 * i, sa[i]);
#pragma unroll
            for (int i = 0; i < B_IT; ++i)
                if (tid + 256 * i < b_tile / 16)
                    split_one(reinterpret_cast<float4*>(st + 2 * A_TILE), reinterpret_cast<float4*>(st + 2 * A_TILE + b_tile),
                              tid + 256 * i, sb[i]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(split + s);
            if (++s == g.stages) {
                s = 0;
                ph ^= 1;
            }
        }
        if (want_a) {
#pragma unroll
            for (int i = 0; i < A_IT; ++i) {
                float* p = colsum_s + 32 * ((tid + 256 * i) / CPB) + col_in_box;
                atomicAdd(p, sa[i].x), atomicAdd(p + 1, sa[i].y), atomicAdd(p + 2, sa[i].z), atomicAdd(p + 3, sa[i].w);
            }
        }
        if (want_b) {
#pragma unroll
            for (int i = 0; i < B_IT; ++i) {
                if (tid + 256 * i >= b_tile / 16) continue;
                float* p = colsum_s + GM + 32 * ((tid + 256 * i) / CPB) + col_in_box;
                atomicAdd(p, sb[i].x), atomicAdd(p + 1, sb[i].y), atomicAdd(p + 2, sb[i].z), atomicAdd(p + 3, sb[i].w);
            }
        }
    } else if (warp >= 12 && num_kb > 0) {
        // ===== drain: TMEM lane = column 128 mb + 32 (warp % 4) + lane of A = row of C =====
        const int ew = warp - 12;
        const int row = ew * 32 + lane;
        float* mine = acc_s + row;                // element (row, c) at acc_s[c * 128 + row]: conflict-free per column
        const int num_drains = (num_kb + DRAIN_STAGES - 1) / DRAIN_STAGES;
        int a = 0;
        uint32_t aph = 0;
        for (int dr = 0; dr < num_drains; ++dr) {
            mbar_wait(acc_full + a, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * ACC_COLS);
            for (int c0 = 0; c0 < g.BN; c0 += 16) {
                float v[16];
                tmem_ld16(taddr0 + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float* p = mine + (c0 + i) * GM;
                    *p = dr ? *p + v[i] : v[i];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + a);
            if (++a == ACC_STAGES) {
                a = 0;
                aph ^= 1;
            }
        }
        const int crow = mb * GM + row;
        if (crow < g.n1) {
            float* cptr = g.C + (int64_t)crow * g.ldc;
            for (int c = 0; c < g.n2; ++c) atomicAdd(cptr + c, mine[c * GM]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACC_COLS * ACC_STAGES) : "memory");
    }
    if (num_kb > 0) {   // this CTA's share of the column sums
        if (g.a_colsum)
            for (int i = threadIdx.x; i < GM; i += NUM_THREADS)
                if (mb * GM + i < g.n1 && colsum_s[i] != 0.f) atomicAdd(g.a_colsum + mb * GM + i, colsum_s[i]);
        if (g.b_colsum && mb == 0)
            for (int i = threadIdx.x; i < g.n2; i += NUM_THREADS)
                if (colsum_s[GM + i] != 0.f) atomicAdd(g.b_colsum + i, colsum_s[GM + i]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [rows, cols] row-major, box = 32 columns x KB rows, 128-byte swizzle of 32-byte chunks, out-of-range elements read as zero
int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int kb) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(RGCN_ERR_UNSUPPORTED, "rgcn_gram: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)kb};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RGCN_ERR_INVALID_ARG, "rgcn_gram: cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return 0;
}

}  // namespace
}  // namespace rgcn

using namespace rgcn;

extern "C" int rgcn_gram3x_tf32(const float* a, int64_t lda, int32_t n1, const float* b, int64_t ldb, int32_t n2, int64_t rows,
                                float* c, int64_t ldc, float* a_colsum, float* b_colsum, void* stream) {
    if (!a || !b || !c || n1 <= 0 || n2 <= 0 || n2 > MAX_BN || rows < 0 || lda < n1 || ldb < n2 || (lda % 4) || (ldb % 4) ||
        ((uintptr_t)a & 15) || ((uintptr_t)b & 15) || ldc < n2 || rows >= (1ll << 31))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_gram3x_tf32: bad argument (rows of A and B 16-byte addressable, n2 <= 192)");
    cudaStream_t st = (cudaStream_t)stream;
    RGCN_CUDA(cudaMemset2DAsync(c, (size_t)ldc * 4, 0, (size_t)n2 * 4, (size_t)n1, st));
    if (a_colsum) RGCN_CUDA(cudaMemsetAsync(a_colsum, 0, (size_t)n1 * 4, st));
    if (b_colsum) RGCN_CUDA(cudaMemsetAsync(b_colsum, 0, (size_t)n2 * 4, st));
    if (rows == 0) return 0;
    GramArgs g{};
    g.rows = rows;
    g.n1 = n1;
    g.n2 = n2;
    g.BN = (n2 + 31) / 32 * 32;
    g.C = c;
    g.ldc = ldc;
    g.a_colsum = a_colsum;
    g.b_colsum = b_colsum;
    static const bool kb32_on = [] {
        const char* e = getenv("RGCN_B200_GRAM_KB32");
        return !(e && e[0] == '0');
    }();
    const int KB = (kb32_on && g.BN <= 96) ? 32 : 16;
    CUtensorMap ma, mbm;
    int rc;
    if ((rc = make_map(&ma, a, rows, n1, lda, KB))) return rc;
    if ((rc = make_map(&mbm, b, rows, n2, ldb, KB))) return rc;
    const int box = 32 * KB * 4, a_tile = (GM / 32) * box;
    const int stage_bytes = 2 * a_tile + 2 * (g.BN / 32) * box;
    const int acc_bytes = g.BN * GM * 4;
    const int budget = 227 * 1024 - 1024 - 256 - 2048 /*column sums*/ - acc_bytes;
    g.stages = std::max(2, std::min(8, budget / stage_bytes));
    const int smem = g.stages * stage_bytes + acc_bytes + 1024 + 256 + 2048;
    if (smem > 227 * 1024) return fail(RGCN_ERR_UNSUPPORTED, "rgcn_gram3x_tf32: tile does not fit shared memory");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int mblocks = (n1 + GM - 1) / GM;
    const int64_t kblocks = (rows + KB - 1) / KB;
    const int splits = (int)std::max<int64_t>(1, std::min<int64_t>(kblocks, sms / mblocks));
    g.rows_per_cta = (kblocks + splits - 1) / splits * KB;
    ProfScope prof(TAG_GEMM, n1, n2, st);
    note_launch(1);
    if (KB == 32) {
        RGCN_CUDA(cudaFuncSetAttribute(k_gram3x<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_gram3x<32><<<dim3((unsigned)splits, (unsigned)mblocks), NUM_THREADS, smem, st>>>(ma, mbm, g);
    } else {
        RGCN_CUDA(cudaFuncSetAttribute(k_gram3x<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_gram3x<16><<<dim3((unsigned)splits, (unsigned)mblocks), NUM_THREADS, smem, st>>>(ma, mbm, g);
    }
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
