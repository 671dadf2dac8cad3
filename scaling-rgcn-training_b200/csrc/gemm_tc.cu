// gemm_tc.cu — K2/K6: the dense contractions of the path on the 5th-generation tensor cores.
//
//   C[M, N] = act( A[M, K] . W[N, K]^T + bias[N] )        fp32 in, fp32 out, fp32-faithful (3xTF32)
//
// replaces the dense contractions of the reference's transfer heads (model/layers.py:105-107: lin2(tanh(lin1(E_cat))),
// two nn.Linear calls = cuBLAS SGEMM there; :59-61: the in / out projections of nn.MultiheadAttention) and their
// backward (dL/dh with the tanh backward in the epilogue) — AM-shape: 1.67 M x 189 x 137 and x 137 x 63 per step,
// 5 M x 63 x 126 for the K | V projection.  One persistent CTA per SM, 16 warps:
//
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (SASS UTMALDG) of the A tile [128 x 32 fp32] and the
//                              pre-split W tiles (hi, lo) [BN x 32] into a 128B-swizzled ring, mbarrier tx (W is
//                              loaded once and stays resident when all its K blocks fit next to three A stages)
//   warp 1      MMA issuer     one elected lane: per 32-column K block 4 x 3 tcgen05.mma.kind::tf32
//                              (SASS UTCHMMA... UTC*MMA) — A_hi.W_hi, A_lo.W_hi, A_hi.W_lo — accumulating
//                              in TENSOR MEMORY (two accumulator stages of 256 columns); tcgen05.commit
//                              releases the smem stage / publishes the accumulator
//   warp 2      TMEM allocator (tcgen05.alloc / dealloc, 512 columns)
//   warps 4-7   splitter       the tensor core reads 19 bits of an fp32: the A tile is split IN SHARED MEMORY
//                              into hi = rn_tf32(a) (overwriting the TMA tile) and lo = rn_tf32(a - hi) (second
//                              tile), both exactly representable, so the product keeps ~22 mantissa bits
//                              (error-compensated 3xTF32, the dropped lo.lo term is ~2^-22, zero-mean)
//                              whatever rounding the hardware applies
//   warps 8-15  epilogue       tcgen05.ld 32x32b (SASS LDTM) -> + bias -> tanh / tanh backward (optional) -> a
//                              swizzled shared-memory slab -> global rows, 8 rows x 64 contiguous bytes per store
//
// W is split once per call by rgcn_gemm_prepack (tiny); rows of A must be 16-byte addressable (TMA), the
// K tail and the M tail are zero-filled by the TMA unit itself.
#include <cuda.h>

#include "common.cuh"

namespace rgcn {
namespace {

constexpr int BM = 128, BK = 32;            // tile rows, K block (32 fp32 = one 128-byte swizzle row)
constexpr int A_TILE = BM * BK * 4;         // 16 KB
constexpr int NUM_THREADS = 512;            // 16 warps: producer, MMA, TMEM allocator, (idle), 4 splitters, 8 epilogue
constexpr int EPI_WARPS = 8;
constexpr int ACC_COLS = 256, ACC_STAGES = 2;

struct GemmArgs {
    int64_t M;
    int K_blocks;       // ceil(K / 32)
    int BN;             // padded N (multiple of 16, <= 256)
    int n_store;        // columns written per row of C (<= ldc); columns >= N receive act(0 + 0) = 0 for both acts
    const float* bias;  // [BN] padded with zeros, or null
    int act;            // 0 identity, 1 tanh, 2 multiply by (1 - aux^2): the tanh backward, aux = the saved tanh output
    const float* aux;   // [M, >= N] (act 2)
    int64_t ldaux;
    int n_aux;          // readable columns of aux
    float* C;
    int64_t ldc;
    int stages;
    int w_resident;     // 1: the split W tiles of ALL K blocks stay in shared memory for the CTA's lifetime (loaded once);
                        //    the ring then holds A tiles only — more bytes of A in flight per SM
};

// round to nearest tf32 (10-bit mantissa), result as an fp32 bit pattern: both parts of a split are then exactly
// representable, so nothing depends on how the tensor core narrows its fp32 inputs, and the rounding errors of
// the dropped lo.lo term have no systematic sign
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// K-major operand tile, 128-byte swizzle: 8-row x 128-byte atoms stacked along M/N (stride 1024 B), one atom on K
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);          // start address
    d |= (uint64_t)0 << 16;                              // leading byte offset (unused: one atom on K)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, both operands K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_gemm3x(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_whi,
         const __grid_constant__ CUtensorMap map_wlo, const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    // carve: [stages x (A_hi, A_lo, W_hi, W_lo)] 1024-byte aligned, then barriers + the TMEM base address
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int w_tile = g.BN * BK * 4;
    const int stage_bytes = 2 * A_TILE + (g.w_resident ? 0 : 2 * w_tile);
    const uint32_t w_res = base + (uint32_t)(g.stages * stage_bytes);      // resident W: [K_blocks][hi | lo]
    const int w_res_bytes = g.w_resident ? g.K_blocks * 2 * w_tile : 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)g.stages * stage_bytes + w_res_bytes);
    uint64_t* full = bars;                       // TMA bytes landed            [stages]
    uint64_t* split = bars + g.stages;           // A split into hi / lo        [stages]
    uint64_t* empty = bars + 2 * g.stages;       // MMAs of the stage retired   [stages]
    uint64_t* acc_full = bars + 3 * g.stages;    // accumulator complete        [ACC_STAGES]
    uint64_t* acc_empty = acc_full + ACC_STAGES; // accumulator drained         [ACC_STAGES]
    uint64_t* w_full = acc_empty + ACC_STAGES;   // resident W landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    // epilogue staging: per warp a [32 rows x 16 columns] slab, 16-byte chunks XOR-swizzled by the row so that both the
    // per-row writes and the transposed reads are conflict-free
    float* epi_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = (g.M + BM - 1) / BM;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(split + s, 4);         // one arrival per splitter warp
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(acc_full + a, 1);
            mbar_init(acc_empty + a, EPI_WARPS);   // one arrival per epilogue warp
        }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(ACC_COLS * ACC_STAGES)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (g.w_resident) {
                uint8_t* wr = sm + (size_t)g.stages * stage_bytes;
                mbar_expect_tx(w_full, (uint32_t)w_res_bytes);
                for (int kb = 0; kb < g.K_blocks; ++kb) {
                    tma_load_2d(wr + (size_t)kb * 2 * w_tile, &map_whi, kb * BK, 0, w_full);
                    tma_load_2d(wr + (size_t)kb * 2 * w_tile + w_tile, &map_wlo, kb * BK, 0, w_full);
                }
            }
            int s = 0;
            uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                for (int kb = 0; kb < g.K_blocks; ++kb) {
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = sm + (size_t)s * stage_bytes;
                    mbar_expect_tx(full + s, (uint32_t)(g.w_resident ? A_TILE : A_TILE + 2 * w_tile));
                    tma_load_2d(st, &map_a, kb * BK, (int)(t * BM), full + s);
                    if (!g.w_resident) {
                        tma_load_2d(st + 2 * A_TILE, &map_whi, kb * BK, 0, full + s);
                        tma_load_2d(st + 2 * A_TILE + w_tile, &map_wlo, kb * BK, 0, full + s);
                    }
                    if (++s == g.stages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc(g.BN);
        int s = 0, a = 0;
        uint32_t ph = 0, aph = 0;
        if (g.w_resident && blockIdx.x < num_tiles) mbar_wait(w_full, 0);
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            mbar_wait(acc_empty + a, aph ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(a * ACC_COLS);
            for (int kb = 0; kb < g.K_blocks; ++kb) {
                mbar_wait(full + s, ph);
                mbar_wait(split + s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t st = base + (uint32_t)(s * stage_bytes);
                    const uint32_t a_hi = st, a_lo = st + A_TILE;
                    const uint32_t w_hi = g.w_resident ? w_res + (uint32_t)(kb * 2 * w_tile) : st + 2 * A_TILE, w_lo = w_hi + w_tile;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {       // tf32: K = 8 (32 bytes) per instruction
                        const uint32_t ko = (uint32_t)(k * 32);
                        umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc, (kb | k) ? 1u : 0u);
                        umma_tf32(tmem_d, umma_desc(a_lo + ko), umma_desc(w_hi + ko), idesc, 1u);
                        umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(w_lo + ko), idesc, 1u);
                    }
                    umma_commit(empty + s);                  // the stage is free once these MMAs have read it
                    if (kb == g.K_blocks - 1) umma_commit(acc_full + a);
                }
                __syncwarp();
                if (++s == g.stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (++a == ACC_STAGES) {
                a = 0;
                aph ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== splitter: hi = rn_tf32(a) (in place), lo = rn_tf32(a - hi) =====
        const int tid = threadIdx.x - 128;       // 0..127
        int s = 0;
        uint32_t ph = 0;
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            for (int kb = 0; kb < g.K_blocks; ++kb) {
                mbar_wait(full + s, ph);
                float4* hi = reinterpret_cast<float4*>(sm + (size_t)s * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(sm + (size_t)s * stage_bytes + A_TILE);
                // elementwise: the swizzle is a permutation of 16-byte chunks, identical for both tiles
#pragma unroll
                for (int i = 0; i < A_TILE / 16 / 128; ++i) {
                    const int q = tid + 128 * i;
                    float4 v = hi[q], h, l;
                    h.x = rn_tf32(v.x), h.y = rn_tf32(v.y), h.z = rn_tf32(v.z), h.w = rn_tf32(v.w);
                    l.x = rn_tf32(v.x - h.x), l.y = rn_tf32(v.y - h.y), l.z = rn_tf32(v.z - h.z), l.w = rn_tf32(v.w - h.w);
                    hi[q] = h;
                    lo[q] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) mbar_arrive(split + s);
                if (++s == g.stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp >= 8) {
        // ===== epilogue: TMEM -> registers -> bias / tanh -> C =====
        // two warps per TMEM lane quarter, each taking every other 16-column chunk: a single warp per scheduler
        // issues its dependent chain at ~5 cycles per instruction, which bound the kernel (0.46 us per chunk)
        const int ew = warp & 3;                 // TMEM lanes 32 ew .. 32 ew + 31 (warp % 4 == ew)
        const int half = (warp - 8) >> 2;        // chunk parity
        int a = 0;
        uint32_t aph = 0;
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            mbar_wait(acc_full + a, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row = t * BM + ew * 32 + lane;
            const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * ACC_COLS);
            // the four store rows of this lane (slab row 8 it + lane / 4), columns c0 + 4 (lane % 4) ..
            float* cp4[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int64_t grow = t * BM + ew * 32 + it * 8 + (lane >> 2);
                cp4[it] = grow < g.M ? g.C + grow * g.ldc + 4 * (lane & 3) : nullptr;
            }
            for (int c0 = 16 * half; c0 < g.BN; c0 += 32) {
                float v[16];
                tmem_ld16(taddr0 + (uint32_t)c0, v);
                if (row < g.M) {
                    // (bias as four 16-byte loads and the activation branch outside the element loop: element-wise
                    // `v[i] + (bias ? ldg : 0)` compiled to 16 dependent load -> add -> branch steps per chunk and made
                    // the epilogue the bound of the whole kernel, ~1.2 us per 16 columns)
                    if (g.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(g.bias + c0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 b = __ldg(bp + i);
                            v[4 * i] += b.x;
                            v[4 * i + 1] += b.y;
                            v[4 * i + 2] += b.z;
                            v[4 * i + 3] += b.w;
                        }
                    }
                    if (g.act == 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
                    }
                    if (g.act == 2) {   // tanh backward: times (1 - h^2), h read 16 bytes at a time (rows of aux are 16-byte addressable)
                        const float* hrow = g.aux + row * g.ldaux + c0;
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (c0 + i + 3 < g.n_aux) h = __ldg(reinterpret_cast<const float4*>(hrow + i));
                            v[i] *= 1.f - h.x * h.x;
                            v[i + 1] *= 1.f - h.y * h.y;
                            v[i + 2] *= 1.f - h.z * h.z;
                            v[i + 3] *= 1.f - h.w * h.w;
                        }
                    }
                }
                // a thread holds 16 columns of ITS row: stored from here a warp instruction would touch 32 rows with 16
                // bytes each.  Through the slab every store instruction writes 8 rows x 64 contiguous bytes instead.
                float4* slab = reinterpret_cast<float4*>(epi_stage + (warp - 8) * 512);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    slab[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = it * 8 + (lane >> 2), q = lane & 3;
                    const float4 o = slab[r * 4 + (q ^ ((r >> 1) & 3))];
                    const int c = c0 + 4 * q;
                    if (cp4[it]) {
                        float* cp = cp4[it] + c0;
                        if (c + 3 < g.n_store) {
                            *reinterpret_cast<float4*>(cp) = o;
                        } else {
                            if (c < g.n_store) cp[0] = o.x;
                            if (c + 1 < g.n_store) cp[1] = o.y;
                            if (c + 2 < g.n_store) cp[2] = o.z;
                        }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + a);
            if (++a == ACC_STAGES) {
                a = 0;
                aph ^= 1;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACC_COLS * ACC_STAGES) : "memory");
    }
}

// W [N, K] -> hi / lo [n_pad, k_pad] (zero padded): hi = rn_tf32(w), lo = rn_tf32(w - hi)
__global__ void k_gemm_prepack(const float* __restrict__ w, int64_t ldw, int n, int k, int n_pad, int k_pad, int transpose,
                               float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_pad * k_pad) return;
    const int r = (int)(i / k_pad), c = (int)(i % k_pad);
    float v = 0.f;
    if (r < n && c < k) v = transpose ? w[(int64_t)c * ldw + r] : w[(int64_t)r * ldw + c];
    const float h = rn_tf32(v);
    hi[i] = h;
    lo[i] = rn_tf32(v - h);
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
int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(RGCN_ERR_UNSUPPORTED, "rgcn_gemm: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RGCN_ERR_INVALID_ARG, "rgcn_gemm: cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return 0;
}

}  // namespace
}  // namespace rgcn

using namespace rgcn;

extern "C" int rgcn_gemm_prepack(const float* w, int64_t ldw, int32_t n, int32_t k, int32_t n_pad, int32_t k_pad,
                                 int32_t transpose, float* w_hi, float* w_lo, void* stream) {
    if (!w || !w_hi || !w_lo || n <= 0 || k <= 0 || n_pad < n || k_pad < k || (n_pad % 16) || (k_pad % 32) || n_pad > 256 ||
        ldw < (transpose ? n : k))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_gemm_prepack: bad argument (n_pad % 16 == 0 <= 256, k_pad % 32 == 0)");
    const int64_t total = (int64_t)n_pad * k_pad;
    note_launch(1);
    k_gemm_prepack<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, ldw, n, k, n_pad, k_pad, transpose, w_hi, w_lo);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int rgcn_gemm3x_tf32(const float* a, int64_t lda, int64_t m, int32_t k, const float* w_hi, const float* w_lo,
                                int32_t n_pad, int32_t k_pad, const float* bias_padded, int32_t act, const float* aux,
                                int64_t ldaux, float* c, int64_t ldc, int32_t n_store, void* stream) {
    if (!a || !w_hi || !w_lo || !c || m < 0 || k <= 0 || lda < k || (lda % 4) || ((uintptr_t)a & 15) || (n_pad % 16) ||
        n_pad <= 0 || n_pad > 256 || (k_pad % 32) || k_pad < k || n_store <= 0 || n_store > n_pad || ldc < n_store ||
        (ldc % 4) || ((uintptr_t)c & 15) || ((uintptr_t)bias_padded & 15) || act < 0 || act > 2 ||
        (act == 2 && (!aux || ldaux <= 0 || (ldaux % 4) || ((uintptr_t)aux & 15))))
        return fail(RGCN_ERR_INVALID_ARG, "rgcn_gemm3x_tf32: bad argument (rows of A and C 16-byte addressable, n_pad % 16 == 0 <= 256)");
    if (m == 0) return 0;
    CUtensorMap ma, mh, ml;
    int rc;
    // the K tail (columns >= k) of A is zero-filled by the TMA unit; W is stored padded
    if ((rc = make_map(&ma, a, m, k, lda, BM))) return rc;
    if ((rc = make_map(&mh, w_hi, n_pad, k_pad, k_pad, n_pad))) return rc;
    if ((rc = make_map(&ml, w_lo, n_pad, k_pad, k_pad, n_pad))) return rc;
    GemmArgs g{};
    g.M = m;
    g.K_blocks = k_pad / BK;
    g.BN = n_pad;
    g.n_store = n_store;
    g.bias = bias_padded;
    g.act = act;
    g.aux = aux;
    g.ldaux = ldaux;
    g.n_aux = (int)std::min<int64_t>(ldaux, n_pad);   // (whole quads: ldaux % 4 == 0)
    g.C = c;
    g.ldc = ldc;
    const int w_tile = n_pad * BK * 4;
    const int budget = 227 * 1024 - 1024 /*alignment*/ - 256 /*barriers*/ - 16384 /*epilogue slabs*/;
    // W resident (all K blocks, hi + lo) when at least 3 A stages still fit: the ring then carries A only.  Measured
    // (ncu, AM shape): with W in the ring 3 stages = 48 KB of A in flight per SM bound the kernel at 1.7 TB/s
    static const bool res_on = [] {
        const char* e = getenv("RGCN_B200_GEMM_WRES");
        return !(e && e[0] == '0');
    }();
    const int w_res_bytes = g.K_blocks * 2 * w_tile;
    g.w_resident = res_on && w_res_bytes + 3 * 2 * A_TILE <= budget;
    const int stage_bytes = 2 * A_TILE + (g.w_resident ? 0 : 2 * w_tile);
    const int ring_budget = budget - (g.w_resident ? w_res_bytes : 0);
    g.stages = std::max(2, std::min(8, ring_budget / stage_bytes));
    const int smem = g.stages * stage_bytes + (g.w_resident ? w_res_bytes : 0) + 1024 + 256 + 16384;
    if (smem > 227 * 1024) return fail(RGCN_ERR_UNSUPPORTED, "rgcn_gemm3x_tf32: tile does not fit shared memory");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    RGCN_CUDA(cudaFuncSetAttribute(k_gemm3x, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t tiles = (m + BM - 1) / BM;
    const int grid = (int)std::min<int64_t>(tiles, sms);
    ProfScope prof(TAG_GEMM, k, n_pad, (cudaStream_t)stream);
    note_launch(1);
    k_gemm3x<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(ma, mh, ml, g);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}
