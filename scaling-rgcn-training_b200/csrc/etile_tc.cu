// etile_tc.cu — K2: the layer's forward / dL/dx pass for 64 gathered x 64 produced columns on the
// 5th-generation tensor cores (hidden-64 / basis-decomposed configuration, BASELINE.json configs[4]).
//
//   out[owner_e] += (w_e * x[src_e]) . B_rel          for the <= 128 entries of a SUPER TILE (one relation)
//
// is a [128 x 64] . [64 x 64] product per super tile — a grouped GEMM whose groups are the (range, relation)
// runs of the blocked relational CSR, with gathered A rows and scattered D rows.  With the warp-level
// mma.sync path this shape is tensor-bound (3xTF32: 2.9 TFLOP per pass at 10x AM-shape, ~3x over its HBM
// time); here one elected thread issues tcgen05.mma.kind::tf32 with the accumulator in tensor memory.
// One persistent CTA per SM, 16 warps:
//
//   warp 0      operand loader: the relation's pre-split, pre-swizzled B image (hi | lo, 32 KB, written once
//               per call by k_wprep_tc) arrives with ONE cp.async.bulk (SASS UBLKCP) per relation change
//   warp 1      MMA issuer: per 32-column half of K 4 x 3 tcgen05.mma (A_hi.B_hi, A_lo.B_hi, A_hi.B_lo),
//               tcgen05.commit frees the stage / publishes the accumulator (2 stages x 64 TMEM columns)
//   warp 2      TMEM allocator
//   warps 4-11  gather + split: two threads per row of the tile — 16-byte cp.async of x[src_t] (zero-fill form
//               for missing rows) straight into the 128B-swizzled operand tile, one stage ahead, indices read
//               three tiles ahead and the rows prefetched to L2 two tiles ahead; then each thread splits its
//               own chunks in place into hi = rn_tf32(a), lo = rn_tf32(a - hi) (and applies the fused ReLU)
//   warps 12-15 epilogue: thread t = TMEM lane t reads its result row (tcgen05.ld), scales it by the entry's
//               1/cnt weight, parks it in shared memory and issues ONE cp.reduce.async.bulk (f32 add, 256 B)
//               to out[owner_t]
// Rows on both sides must be 16-byte addressable with >= 64 floats per row (x: the zero-padded mirror for
// emb = 63; dL/dx: a 64-wide padded gradient buffer).  The root / self-loop term is written before this pass
// by the self-loop kernel, exactly as for the mma.sync entry-tile kernels.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace rgcn {
namespace {

constexpr int TM = 128;              // entries per super tile (MMA M)
constexpr int A_HALF = TM * 128;     // one 32-column half of the A tile: 128 rows x 128 B
constexpr int STAGES = 3;            // (A_hi, A_lo) ring, one stage = one half of K
constexpr int W_IMG = 4 * 64 * 128;  // B image of a relation: hi half 0 | hi half 1 | lo half 0 | lo half 1 (32 KB)
constexpr int W_BUFS = 2;
constexpr int OUT_STRIDE = 272;      // bytes per staged result row (256 + 16: STS.128 conflict-free per 8 lanes)
constexpr int ACC_COLS = 64, ACC_STAGES = 2;
constexpr int THREADS = 512;
constexpr int CHUNK_TILES = 4;       // consecutive super tiles a CTA takes at a time (one relation most of the time)
constexpr int SMEM_BYTES = STAGES * 2 * A_HALF + W_BUFS * W_IMG + TM * OUT_STRIDE + 256 + 1024;

struct TcArgs {
    const uint32_t* e_idx;
    const float* e_w;
    const int32_t* e_own;
    const int32_t* stile_e0;
    const int32_t* stile_rel;
    int num_stiles;
    const float* feat;
    int64_t ldf;
    int kin;
    const float* aux;
    int64_t n_rows;
    const float* wimg;   // [(R+1)][W_IMG / 4] floats
    float* out;
    int64_t ldo;
    int relu;
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
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {   // K-major, 128-byte swizzle, 8-row groups 1024 B apart
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}

// this CTA's tile sequence: chunks of CHUNK_TILES consecutive super tiles, chunks dealt round-robin over the grid
struct TileSeq {
    int chunk, j, num;
    __device__ TileSeq(int num_stiles) : chunk(blockIdx.x), j(0), num(num_stiles) {}
    __device__ int cur() const {
        const int t = chunk * CHUNK_TILES + j;
        return t < num ? t : -1;
    }
    __device__ void next() {
        if (++j == CHUNK_TILES || chunk * CHUNK_TILES + j >= num) {
            j = 0;
            chunk += gridDim.x;
        }
    }
};

__global__ void __launch_bounds__(THREADS, 1) k_etile_tc(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a_base = base;                                   // STAGES x (hi 16 KB | lo 16 KB)
    const uint32_t w_base = base + STAGES * 2 * A_HALF;             // W_BUFS x 32 KB
    uint8_t* out_stage = sm + STAGES * 2 * A_HALF + W_BUFS * W_IMG; // TM x OUT_STRIDE
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + TM * OUT_STRIDE);
    uint64_t* afull = bars;                      // [STAGES] rows gathered and split
    uint64_t* aempty = bars + STAGES;            // [STAGES] MMAs of the stage retired
    uint64_t* wfull = bars + 2 * STAGES;         // [W_BUFS]
    uint64_t* wempty = wfull + W_BUFS;           // [W_BUFS]
    uint64_t* acc_full = wempty + W_BUFS;        // [ACC_STAGES]
    uint64_t* acc_empty = acc_full + ACC_STAGES; // [ACC_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(afull + s, 8);             // one arrival per gather warp
            mbar_init(aempty + s, 1);
        }
        for (int b = 0; b < W_BUFS; ++b) {
            mbar_init(wfull + b, 1);
            mbar_init(wempty + b, 1);
        }
        for (int s = 0; s < ACC_STAGES; ++s) {
            mbar_init(acc_full + s, 1);
            mbar_init(acc_empty + s, 4);         // one arrival per epilogue warp
        }
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
        // ===== operand loader: one 32 KB bulk copy per relation change =====
        if (lane == 0) {
            int b = 0, prev_rel = -1;
            uint32_t ph = 0;
            for (TileSeq ts(a.num_stiles); ts.cur() >= 0; ts.next()) {
                const int rel = a.stile_rel[ts.cur()];
                if (rel == prev_rel) continue;
                prev_rel = rel;
                mbar_wait(wempty + b, ph ^ 1);
                mbar_expect_tx(wfull + b, W_IMG);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 w_base + (uint32_t)(b * W_IMG)),
                             "l"(a.wimg + (int64_t)rel * (W_IMG / 4)), "r"((uint32_t)W_IMG), "r"(smem_u32(wfull + b))
                             : "memory");
                if (++b == W_BUFS) {
                    b = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        int s = 0, acc = 0, b = -1, prev_rel = -1;
        uint32_t ph = 0, aph = 0, wph = 0;
        TileSeq ts(a.num_stiles);
        int rel_next = ts.cur() >= 0 ? a.stile_rel[ts.cur()] : -2, rel_next2 = -2;
        {
            TileSeq t2 = ts;
            t2.next();
            rel_next2 = t2.cur() >= 0 ? a.stile_rel[t2.cur()] : -2;
        }
        while (ts.cur() >= 0) {
            const int rel = rel_next;
            ts.next();
            const int next_rel = rel_next2;          // (the relation ids are read two tiles ahead)
            {
                TileSeq t2 = ts;
                t2.next();
                rel_next = next_rel;
                rel_next2 = (ts.cur() >= 0 && t2.cur() >= 0) ? a.stile_rel[t2.cur()] : -2;
            }
            if (rel != prev_rel) {          // the loader's next buffer holds this relation
                if (b < 0) {
                    b = 0;
                } else if (++b == W_BUFS) {
                    b = 0;
                    wph ^= 1;
                }
                mbar_wait(wfull + b, wph);
                prev_rel = rel;
            }
            mbar_wait(acc_empty + acc, aph ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
            const uint32_t wb = w_base + (uint32_t)(b * W_IMG);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(afull + s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t a_hi = a_base + (uint32_t)(s * 2 * A_HALF), a_lo = a_hi + A_HALF;
                    const uint32_t w_hi = wb + (uint32_t)(half * 64 * 128), w_lo = w_hi + 2 * 64 * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t ko = (uint32_t)(k * 32);
                        umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(w_hi + ko), idesc, (half | k) ? 1u : 0u);
                        umma_tf32(tmem_d, umma_desc(a_lo + ko), umma_desc(w_hi + ko), idesc, 1u);
                        umma_tf32(tmem_d, umma_desc(a_hi + ko), umma_desc(w_lo + ko), idesc, 1u);
                    }
                    umma_commit(aempty + s);
                    if (half == 1) {
                        umma_commit(acc_full + acc);
                        if (next_rel != rel) umma_commit(wempty + b);   // last tile that reads this operand buffer
                    }
                }
                __syncwarp();
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (++acc == ACC_STAGES) {
                acc = 0;
                aph ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== gather + split: 256 threads, thread u <-> (row u / 2, four of the eight 16-byte chunks of a half row) =====
        const int u = threadIdx.x - 128, t = u >> 1, part = u & 1;
        const uint32_t n_rows = (uint32_t)a.n_rows;
        const uint32_t row_off = (uint32_t)(t * 128), sw = (uint32_t)(t & 7);
        int s_issue = 0, s_done = 0;
        uint32_t ph_issue = 0;
        // (the raw index word is kept as loaded: nothing is computed from it until the copies need the address, so
        // the load of tile i+3's indices never stalls the warp)
        struct Row {
            uint32_t raw;   // e_idx word as loaded (bit 31 = end-of-segment flag)
            bool valid;
        };
        // two-level look-ahead: the entry span of a tile is read one step before its index word
        int span_e0 = 0, span_e1 = 0;
        auto load_span = [&](int tile) {
            span_e0 = a.stile_e0[tile];
            span_e1 = a.stile_e0[tile + 1];
        };
        auto row_of_span = [&]() {
            Row r{0u, false};
            if (span_e0 + t < span_e1) {
                r.raw = a.e_idx[span_e0 + t];
                r.valid = true;
            }
            return r;
        };
        auto src_of = [&](const Row& r) -> const float* {
            if (!r.valid) return nullptr;
            const uint32_t idx = r.raw & IDX_MASK;
            return idx < n_rows ? a.feat + (int64_t)idx * a.ldf : a.aux + (int64_t)(idx - n_rows) * 64;
        };
        auto issue = [&](const float* src, int half) {
            mbar_wait(aempty + s_issue, ph_issue ^ 1);
            const uint32_t dst = a_base + (uint32_t)(s_issue * 2 * A_HALF) + row_off;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 4 * part + i, col = half * 32 + 4 * c;
                const bool ok = src != nullptr && col < a.kin;
                cp_async16_zfill(dst + (((uint32_t)c ^ sw) << 4), ok ? (const void*)(src + col) : (const void*)a.feat, ok ? 16 : 0);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (++s_issue == STAGES) {
                s_issue = 0;
                ph_issue ^= 1;
            }
        };
        auto finish = [&](bool relu_row) {   // split this thread's own four chunks of the oldest issued stage and publish it
            float4* hi = reinterpret_cast<float4*>(sm + (size_t)s_done * 2 * A_HALF + row_off);
            float4* lo = reinterpret_cast<float4*>(sm + (size_t)s_done * 2 * A_HALF + A_HALF + row_off);
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = hi[(4 * part + i) ^ (int)sw];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = (4 * part + i) ^ (int)sw;
                float4 x = v[i], h, l;
                if (relu_row) x.x = fmaxf(x.x, 0.f), x.y = fmaxf(x.y, 0.f), x.z = fmaxf(x.z, 0.f), x.w = fmaxf(x.w, 0.f);
                h.x = rn_tf32(x.x), h.y = rn_tf32(x.y), h.z = rn_tf32(x.z), h.w = rn_tf32(x.w);
                l.x = rn_tf32(x.x - h.x), l.y = rn_tf32(x.y - h.y), l.z = rn_tf32(x.z - h.z), l.w = rn_tf32(x.w - h.w);
                hi[c] = h;
                lo[c] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(afull + s_done);
            if (++s_done == STAGES) s_done = 0;
        };
        // Look-ahead, so that no dependent load sits in front of the copies: entry spans and index words are read
        // LOOK tiles ahead, a row is pulled towards L2 (one 128-byte line per thread) LOOK - 1 tiles ahead — at
        // 10x AM-shape every gathered row is a TLB miss on top of a DRAM access — and the copies of tile i+1 are
        // issued before tile i is split.
        auto prefetch_l2 = [&](const Row& r) {
            const float* p = src_of(r);
            if (p != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32 * part) : "memory");
        };
        TileSeq ts(a.num_stiles);           // cursor of the tile whose entry span is read next (4 ahead of `cur`)
        bool span_ok = false;
        auto advance_span = [&]() {         // returns whether the PREVIOUS span was a tile; loads the next one
            const bool had = span_ok;
            span_ok = ts.cur() >= 0;
            if (span_ok) {
                load_span(ts.cur());
                ts.next();
            }
            return had;
        };
        struct Fetched {
            Row first;
            bool second;
        };
        advance_span();
        auto fetch_next = [&]() {           // index word of the tile whose span was read one step ago
            Fetched f{Row{0u, false}, false};
            if (span_ok) {
                f.first = row_of_span();
                f.second = true;
            }
            advance_span();
            return f;
        };
        // queue of LOOK tiles: [0] = the tile being gathered, [1] = the next one (its copies are issued before [0] is
        // split), [LOOK - 1] = the farthest one, whose row is prefetched towards L2 as soon as its index arrives
        constexpr int LOOK = 6;
        Row q[LOOK];
        bool has[LOOK];
#pragma unroll
        for (int i = 0; i < LOOK; ++i) {
            const Fetched f = fetch_next();
            q[i] = f.first;
            has[i] = f.second;
        }
        if (has[0]) {
#pragma unroll
            for (int i = 1; i < LOOK - 1; ++i)
                if (has[i]) prefetch_l2(q[i]);
            const float* src = src_of(q[0]);
            issue(src, 0);
            while (true) {
                const Fetched f = fetch_next();                  // index of the tile LOOK ahead (consumed LOOK - 1 trips later)
                if (has[LOOK - 1]) prefetch_l2(q[LOOK - 1]);     // its predecessor's row -> L2
                const bool relu_row = a.relu && q[0].valid && (q[0].raw & IDX_MASK) < n_rows;
                issue(src, 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                finish(relu_row);                                // half 0 of the current tile
                const float* src1 = has[1] ? src_of(q[1]) : nullptr;
                if (has[1]) {
                    issue(src1, 0);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                finish(relu_row);                                // half 1
                if (!has[1]) break;
#pragma unroll
                for (int i = 0; i < LOOK - 1; ++i) {
                    q[i] = q[i + 1];
                    has[i] = has[i + 1];
                }
                q[LOOK - 1] = f.first;
                has[LOOK - 1] = f.second;
                src = src1;
            }
        }
    } else if (warp >= 12) {
        // ===== epilogue: thread t <-> TMEM lane t <-> row t =====
        const int ew = warp - 12, t = ew * 32 + lane;
        float* my_row = reinterpret_cast<float*>(out_stage + t * OUT_STRIDE);
        int acc = 0;
        uint32_t aph = 0;
        // the row's owner / weight are read two tiles ahead of their use (a DRAM round trip in front of every
        // tile's scatter would otherwise bound the whole pipeline)
        struct Meta {
            int own;
            float w;
        };
        TileSeq ahead(a.num_stiles);
        auto fetch_meta = [&]() {
            Meta m{-1, 0.f};
            if (ahead.cur() >= 0) {
                const int tile = ahead.cur();
                const int e0 = a.stile_e0[tile], cnt = a.stile_e0[tile + 1] - e0;
                if (t < cnt) {
                    m.own = a.e_own[e0 + t];
                    m.w = a.e_w[e0 + t];
                }
                ahead.next();
            }
            return m;
        };
        Meta m0 = fetch_meta(), m1 = fetch_meta();
        for (TileSeq ts(a.num_stiles); ts.cur() >= 0; ts.next()) {
            const Meta m2 = fetch_meta();
            const int own = m0.own;
            const float w = m0.w;
            m0 = m1, m1 = m2;
            mbar_wait(acc_full + acc, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * ACC_COLS);
            float v[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) tmem_ld16(taddr + (uint32_t)c0, v + c0);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + acc);     // the accumulator is in registers: the next tile may start
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // my previous row has left shared memory
#pragma unroll
            for (int c = 0; c < 64; c += 4)
                *reinterpret_cast<float4*>(my_row + c) = make_float4(v[c] * w, v[c + 1] * w, v[c + 2] * w, v[c + 3] * w);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (own >= 0) {
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 256;" ::"l"(
                                 a.out + (int64_t)own * a.ldo),
                             "r"(smem_u32(my_row))
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++acc == ACC_STAGES) {
                acc = 0;
                aph ^= 1;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACC_COLS * ACC_STAGES) : "memory");
    }
}

// B image of every relation (root = relation R): Bop[n][k] = B[k][n], split into rn_tf32 hi / lo and laid out
// exactly as the tensor core reads it (two 32-column halves of [64 rows x 128 B], 16-byte chunks XOR-swizzled by
// the row) so that one flat bulk copy brings a relation's operand in.
//   forward (transpose = 0): B[k][n] = W_rel[k][n]      dL/dx (transpose = 1): B[k][n] = W_rel[n][k]
__global__ void k_wprep_tc(const float* __restrict__ weight, const float* __restrict__ root, int R, int fin, int fout,
                           int transpose, float* __restrict__ wimg) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)(R + 1) * 64 * 64) return;
    const int k = (int)(i & 63), n = (int)((i >> 6) & 63), rel = (int)(i >> 12);
    const float* W = rel < R ? weight + (int64_t)rel * fin * fout : root;
    float v = 0.f;
    if (W) {
        if (!transpose) {
            if (k < fin && n < fout) v = W[(int64_t)k * fout + n];
        } else {
            if (n < fin && k < fout) v = W[(int64_t)n * fout + k];
        }
    }
    const float h = rn_tf32(v), l = rn_tf32(v - h);
    const int half = k >> 5, kk = k & 31, chunk = kk >> 2, j = kk & 3;
    const int off = half * (64 * 32) + n * 32 + ((chunk ^ (n & 7)) << 2) + j;   // floats within the hi (or lo) part
    float* img = wimg + (int64_t)rel * (W_IMG / 4);
    img[off] = h;
    img[2 * 64 * 32 + off] = l;
}

int tc_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RGCN_B200_TC");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}

}  // namespace

bool etile_tc_ok(const TilePass& p) {
    return tc_mode() == 1 && p.kp == 64 && p.np == 64 && p.vec4 && !p.packed && p.brc->stile_e0 != nullptr &&
           p.ldf % 4 == 0 && p.ldo % 4 == 0 && p.ldo >= 64 && ((uintptr_t)p.out & 15) == 0 && ((uintptr_t)p.feat & 15) == 0 &&
           (p.aux == nullptr || ((uintptr_t)p.aux & 15) == 0);
}

int64_t wprep_tc_floats(int R) { return (int64_t)(R + 1) * (W_IMG / 4); }

int launch_wprep_tc(const float* weight, const float* root, int R, int fin, int fout, bool transpose, float* wtc,
                    cudaStream_t st) {
    const int64_t total = (int64_t)(R + 1) * 64 * 64;
    ProfScope prof(TAG_WPREP, fin, fout, st);
    note_launch(1);
    k_wprep_tc<<<(int)((total + 255) / 256), 256, 0, st>>>(weight, root, R, fin, fout, transpose ? 1 : 0, wtc);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_etile_tc(const TilePass& p, const float* wtc, int R, int num_sms, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_stiles == 0) return 0;
    TcArgs a{};
    a.e_idx = b.e_idx;
    a.e_w = b.e_w;
    a.e_own = b.e_own;
    a.stile_e0 = b.stile_e0;
    a.stile_rel = b.stile_rel;
    a.num_stiles = b.num_stiles;
    a.feat = p.feat;
    a.ldf = p.ldf;
    a.kin = p.kin;
    a.aux = p.aux ? p.aux : p.feat;
    a.n_rows = p.n_nodes;
    a.wimg = wtc;
    a.out = p.out;
    a.ldo = p.ldo;
    a.relu = p.relu_in ? 1 : 0;
    RGCN_CUDA(cudaFuncSetAttribute(k_etile_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int chunks = (b.num_stiles + CHUNK_TILES - 1) / CHUNK_TILES;
    const int grid = std::max(1, std::min(chunks, num_sms));
    ProfScope prof(p.transposed ? TAG_TILE_BWD : TAG_TILE_FWD, p.kin, p.tag_out, st);
    note_launch(1);
    k_etile_tc<<<grid, THREADS, SMEM_BYTES, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rgcn
