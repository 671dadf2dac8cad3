// etile_kernels.cu — entry-tile kernels: the layer's forward / dL/dx (k_etile) and dL/dW (k_ewgrad)
// with the gathered rows loaded STRAIGHT INTO tensor-core fragments.
//
// A tile is <= 16 consecutive entries (edges, self loops or chunk rows) of one relation.  With the
// entries as the M dimension of mma.m16n8k8, lane (g, t) of a warp owns entries g and g+8 and the
// K positions t, t+4 of every 8-column step — so its A fragment is exactly
//     x[src(g)][8k+t], x[src(g+8)][8k+t], x[src(g)][8k+t+4], x[src(g+8)][8k+t+4]
// i.e. four plain 4-byte loads off two row pointers with immediate offsets; one load instruction
// covers 8 rows x 16 B (sector pairs shared by the t / t+4 halves).  The ROWS are never staged in
// shared memory, there is no per-entry accumulate chain and no segment bookkeeping in the kernel:
// the per-(relation, dst) mean is the sum of the entries' 1/cnt-weighted rows, and the sum happens in
// the scatter (REDG.ADD.F32x4 per entry row, or one bulk reduce per row for 64-column rows).  Only
// the INDEX streams of a work unit go through shared memory (cp.async, one unit ahead).
// dL/dW needs no segments at all:
//     dW_rel = sum_entries (w_e x[src_e])^T (x) gout[owner_e]   =  A^T . B over 16-entry K slices
// with A^T fragments again direct loads (lane (g,t): entries t, t+4; feature rows 16m+g, +8).
//
// Per 16 entries (63 -> 16): 32 row loads, 32 fmul, 64 split ops, 48 HMMA (3xTF32), 16 B-fragment
// loads, 4 vector atomics — about half the instructions of the staged version and no
// dependent chains longer than the MMA accumulate.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "device_utils.cuh"

namespace rgcn {
namespace {

using namespace dev;

constexpr int EW = 8;    // warps per CTA
constexpr int UT = 16;   // tiles per work unit (round-robin over warps)

struct ETileArgs {
    const uint32_t* e_idx;
    const float* e_w;
    const int32_t* e_own;
    const int32_t* tile_e0;
    const int32_t* tile_info;
    int num_tiles;
    const float* feat;
    int64_t ldf;
    int kin;
    const float* aux;
    int64_t n_rows;
    const float4* wfrag;
    const float2* wfrag2;
    const float* bias;
    int nbias;
    int self_rel;
    float* out;
    int64_t ldo;
    int nout;
    int64_t out_elems;   // packed mode: floats in `out`
    // dL/dW only
    const float* gout;
    int64_t ldg;
    float* gweight;
    float* groot;
    float* gbias;
};

__device__ __forceinline__ float4 ldg128_hint(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
// (row gathers keep their L1 allocation: with L1::no_allocate the two 64-byte halves of a 128-byte
// line become separate L2 requests and the 256-byte-row forward pass ran 40 % slower)
// round-to-nearest split (unbiased): used for ONE operand when both are activations, so that the
// dropped lo*lo term has no systematic sign over long sums
__device__ __forceinline__ void split_rn(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// Kernel parameters used inside the load sequences are pinned in registers: left alone, ptxas
// re-reads them from the constant bank (LDC) between the row loads, and an LDC that shares a
// scoreboard slot with an outstanding LDG waits for DRAM — the loads of one tile then go out in
// three or four round trips instead of one (ncu: half of all stall samples on those sites).
__device__ __forceinline__ const float* pin(const float* p) {
    asm volatile("" : "+l"(p));
    return p;
}
__device__ __forceinline__ uint32_t pin(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ int pin(int v) {
    asm volatile("" : "+r"(v));
    return v;
}

struct RowRef {   // one gathered row of a lane
    const float* p;   // row pointer + lane's k offset
    float w;
    int own;
    bool real;        // a feature row (ReLU applies), not a chunk row
};

// ---------------------------------------------------------------------------------------------
// forward / dL/dx :  out[owner_e] += (w_e * x[src_e]) . B_rel   for the 16 entries of a tile
// ---------------------------------------------------------------------------------------------
// V4: rows are 16-byte addressable -> one LDG.128 per row per 16 columns.  Lane (g,t) then holds
// columns 16j+4t..+3, which serve K slots (t, t+4) of steps 2j and 2j+1; the B fragments are
// prepared with the matching row permutation (k_wprep perm).  K is a contraction index, so any
// consistent permutation is exact.
//
// BULK: the scatter goes through shared memory and the bulk-copy engine instead of the LSU: the
// warp parks its 16 x (8 NT) result tile in a (double-buffered, bank-skewed) shared tile and 16
// lanes each issue ONE cp.reduce.async.bulk (f32 add) of a whole row to out[owner].  A vector RED
// costs ~1.3 LSU cycles per LANE whatever its width, so a 64-column row is 16 lane-slots; as a
// bulk reduce it is one request to the copy engine.
template <int NT>
struct BulkTile {
    static constexpr int ROWW = NT * 8 + 8;        // words per staged row: +8 skews rows over the banks
    static constexpr int NBUF = 1;                      // result-tile buffers per warp
    static constexpr int WARP_WORDS = NBUF * 16 * ROWW;
    static constexpr int CTA_BYTES = EW * WARP_WORDS * 4;
};

//
// BULK == 2 (packed odd-width rows, e.g. the [N, 63] embedding gradient): row i starts at float
// 63 i, which is 16-byte aligned only for i % 4 == 0.  The row is staged shifted by f = (63 i) % 4
// with zeros around it and the reduce covers the aligned window [63 i - f, +8 NT + 4): the
// neighbours receive +0.0f.  No padded target, no column copy afterwards.
// Metadata staging: a warp's unit is UTU consecutive tiles = one contiguous span of <= 16 UTU entries.
// The unit's e_idx / e_w / e_own slices and its tile list are copied global -> shared with 16-byte
// cp.async ONE UNIT AHEAD (double buffered), so the tile loop reads its indices from shared memory:
// no per-tile index loads from global, no exposed index latency, no deep register pipeline.
template <int BULK>
struct MetaStage {
    static constexpr int UTU = (BULK != 0) ? 8 : UT;   // bulk kernels also park result tiles in shared memory
    static constexpr int SPAN = UTU * 16 + 4;          // entries staged per unit (start rounded down to a quad)
    static constexpr int MW = 3 * SPAN + 2 * UTU;      // words per buffer: idx, w, own, tile_e0, tile_info
    static constexpr int WARP_WORDS = 2 * MW;
    static constexpr int CTA_BYTES = EW * WARP_WORDS * 4;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}

// STAGE (16-column gathered rows that are 16-byte addressable): the ROWS go through shared memory too.
// Each warp keeps a ring of RowStage::D tile slots (16 rows x 64 B) filled by 16-byte cp.async (zero-fill
// form for the unused rows / columns) RowStage::D tiles ahead of the tensor-pipe work, so the bytes in
// flight per warp no longer sit in registers: D KB instead of 1 KB per warp.  The index streams of a unit
// arrive by the bulk-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP), two units ahead, in three
// buffers — the row prefetch runs across unit boundaries and needs the next unit's indices early.
struct RowStage {
    static constexpr int D = 4;                       // tile slots per warp
    static constexpr int NB = 3;                      // metadata buffers per warp
    static constexpr int SLOT_WORDS = 16 * 16;        // 16 rows x 16 floats
    static constexpr int BAR_WORDS = EW * NB * 2;     // one 8-byte mbarrier per buffer per warp
    static constexpr int META_WORDS = EW * NB * MetaStage<1>::MW;
    static constexpr int ROW_WORDS = EW * D * SLOT_WORDS;
    static constexpr int CTA_BYTES = (BAR_WORDS + META_WORDS + ROW_WORDS) * 4;
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int KT, int NT, bool RELU, bool V4, int BULK = 0, int STAGE = 0>
__global__ void __launch_bounds__(EW * 32, (KT * NT >= 16) ? 2 : 3) k_etile(const ETileArgs a) {
    constexpr int KP = KT * 8;
    static_assert(STAGE == 0 || (KT == 2 && V4), "row staging is built for 16-column vector-addressable rows");
    using MS = MetaStage<(BULK != 0 || STAGE != 0) ? 1 : 0>;
    constexpr int UTU = MS::UTU, SPAN = MS::SPAN, MW = MS::MW;
    extern __shared__ __align__(16) uint32_t dyn_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int gw = blockIdx.x * EW + warp, nw = gridDim.x * EW;
    uint32_t* meta = dyn_smem + (STAGE ? RowStage::BAR_WORDS + warp * (RowStage::NB * MW) : warp * MS::WARP_WORDS);
    [[maybe_unused]] float* bulk_smem =
        reinterpret_cast<float*>(dyn_smem + (STAGE ? RowStage::CTA_BYTES / 4 : EW * MS::WARP_WORDS));
    const uint64_t pol_f =
        (uint64_t)a.n_rows * (uint64_t)a.ldf * 4ull <= (112ull << 20) ? policy_evict_last() : policy_evict_normal();
    constexpr bool BREG = (KT * NT <= 4);
    float4 bfrag[BREG ? KT * NT : 1];
    // 16 -> 64 dL/dx: the 16 unsplit fragments of the current relation stay in registers too (their
    // re-read from L1 every tile is 4 KB per warp per tile: a fifth of the kernel's cycles in L1 returns)
    constexpr bool BREG2 = (BULK != 0 && KT == 2 && NT == 8);
    float2 bfrag2[BREG2 ? KT * NT : 1];
    int breg_rel = -1;
    constexpr int KOFF = V4 ? 4 : 1;   // lane's column offset inside a row: 4t (vector) or t
    struct TileRef {   // the two rows of this lane in one tile
        RowRef g8, h8;
        int rel;
    };
    const int num_units = (a.num_tiles + UTU - 1) / UTU;

    // entry span [first, end) of a unit, read straight from the tile list (one unit ahead of its use)
    auto unit_span = [&](int u, int& first, int& end) {
        const int t0 = u * UTU, t1 = min(a.num_tiles, t0 + UTU);
        first = a.tile_e0[t0];
        end = a.tile_e0[t1 - 1] + (a.tile_info[t1 - 1] & 0xff);
    };
    auto stage_unit = [&](int u, int first, int end, uint32_t* m) {
        const int eb = first & ~3;
        const int nchunk = (end - eb + 3) >> 2;
        for (int c = lane; c < nchunk; c += 32) {
            cp_async16(m + 4 * c, a.e_idx + eb + 4 * c);
            cp_async16(m + SPAN + 4 * c, a.e_w + eb + 4 * c);
            cp_async16(m + 2 * SPAN + 4 * c, a.e_own + eb + 4 * c);
        }
        const int t0 = u * UTU;
        if (lane < UTU / 4) cp_async16(m + 3 * SPAN + 4 * lane, a.tile_e0 + t0 + 4 * lane);
        else if (lane < UTU / 2) cp_async16(m + 3 * SPAN + UTU + 4 * (lane - UTU / 4), a.tile_info + t0 + 4 * (lane - UTU / 4));
    };
    // lane bases: feature rows / chunk rows.  Pinned (see pin()) where registers allow: the wide-row
    // instances are at the 128-register cap and pinning them spills (measured slower)
    constexpr bool PIN = (KT <= 2);
    const float* const feat_t = PIN ? pin(a.feat + KOFF * t) : a.feat + KOFF * t;
    const float* const aux_t = PIN ? pin(a.aux + KOFF * t) : a.aux + KOFF * t;
    const uint32_t n_rows = PIN ? pin((uint32_t)a.n_rows) : (uint32_t)a.n_rows;
    const uint32_t ldf = PIN ? pin((uint32_t)a.ldf) : (uint32_t)a.ldf;
    const int kin = PIN ? pin(a.kin) : a.kin;
    auto make_row = [&](const uint32_t* m, int o, bool valid) {
        RowRef r;
        uint32_t idx = 0;
        r.w = 0.f;
        r.own = -1;
        if (valid) {
            idx = m[o] & IDX_MASK;
            r.w = __uint_as_float(m[SPAN + o]);
            r.own = (int)m[2 * SPAN + o];
        }
        r.real = idx < n_rows;
        const uint32_t off = r.real ? idx * ldf : (idx - n_rows) * (uint32_t)KP;   // < 2^32 elements (checked by the caller)
        r.p = (r.real ? feat_t : aux_t) + off;
        return r;
    };
    auto tile_ref = [&](const uint32_t* m, int i, int eb) {
        TileRef r;
        const int e0 = (int)m[3 * SPAN + i], info = (int)m[3 * SPAN + UTU + i];
        r.rel = info >> 8;
        const int o = e0 - eb + g, cnt = info & 0xff;
        r.g8 = make_row(m, o, g < cnt);
        r.h8 = make_row(m, o + 8, g + 8 < cnt);
        return r;
    };
    // A fragments: raw loads (all independent); the arithmetic happens one tile later
    auto load_rows = [&](const TileRef& r, float (&av)[KT][4]) {
        if constexpr (V4) {
#pragma unroll
            for (int j = 0; j < KT / 2; ++j) {
                float4 vg = make_float4(0.f, 0.f, 0.f, 0.f), vh = vg;
                if (16 * j + 4 * t < kin) {
                    vg = ldg128_hint(reinterpret_cast<const float4*>(r.g8.p + 16 * j), pol_f);
                    vh = ldg128_hint(reinterpret_cast<const float4*>(r.h8.p + 16 * j), pol_f);
                }
                av[2 * j][0] = vg.x;       // step 2j  : slot t   <- col 16j+4t
                av[2 * j][2] = vg.y;       //            slot t+4 <- col 16j+4t+1
                av[2 * j + 1][0] = vg.z;   // step 2j+1: slot t   <- col 16j+4t+2
                av[2 * j + 1][2] = vg.w;   //            slot t+4 <- col 16j+4t+3
                av[2 * j][1] = vh.x;
                av[2 * j][3] = vh.y;
                av[2 * j + 1][1] = vh.z;
                av[2 * j + 1][3] = vh.w;
            }
        } else {
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const bool c_lo = 8 * kt + t < kin, c_hi = 8 * kt + 4 + t < kin;
                av[kt][0] = c_lo ? ldg_hint(r.g8.p + 8 * kt, pol_f) : 0.f;
                av[kt][1] = c_lo ? ldg_hint(r.h8.p + 8 * kt, pol_f) : 0.f;
                av[kt][2] = c_hi ? ldg_hint(r.g8.p + 8 * kt + 4, pol_f) : 0.f;
                av[kt][3] = c_hi ? ldg_hint(r.h8.p + 8 * kt + 4, pol_f) : 0.f;
            }
        }
    };

    if (gw >= num_units) return;
    [[maybe_unused]] int bulk_par = 0;
    auto process = [&](const TileRef& cur, const float (&av)[KT][4]) {
        const int rel = cur.rel;
        const RowRef cg = cur.g8, ch = cur.h8;
        const float4* wf = a.wfrag + (int64_t)rel * (KT * NT * 32) + lane;
        const float2* wf2 = a.wfrag2 + (int64_t)rel * (KT * NT * 32) + lane;
        if constexpr (BREG) {   // few fragments: keep the current relation's in registers
            if (rel != breg_rel) {
#pragma unroll
                for (int q = 0; q < KT * NT; ++q) bfrag[q] = __ldg(wf + q * 32);
                breg_rel = rel;
            }
        }
        if constexpr (BREG2) {
            if (rel != breg_rel) {
#pragma unroll
                for (int q = 0; q < KT * NT; ++q) bfrag2[q] = __ldg(wf2 + q * 32);
                breg_rel = rel;
            }
        }
        float d[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) d[n][0] = d[n][1] = d[n][2] = d[n][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            float x0 = av[kt][0], x1 = av[kt][1], x2 = av[kt][2], x3 = av[kt][3];
            if (RELU) {
                if (cg.real) {
                    x0 = fmaxf(x0, 0.f);
                    x2 = fmaxf(x2, 0.f);
                }
                if (ch.real) {
                    x1 = fmaxf(x1, 0.f);
                    x3 = fmaxf(x3, 0.f);
                }
            }
            uint32_t ah[4], al[4];
            split_fast(x0 * cg.w, ah[0], al[0]);
            split_fast(x1 * ch.w, ah[1], al[1]);
            split_fast(x2 * cg.w, ah[2], al[2]);
            split_fast(x3 * ch.w, ah[3], al[3]);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t bh0, bh1, bl0, bl1;
                if constexpr (BREG) {
                    const float4 bf = bfrag[kt * NT + n];
                    bh0 = __float_as_uint(bf.x), bh1 = __float_as_uint(bf.y);
                    bl0 = __float_as_uint(bf.z), bl1 = __float_as_uint(bf.w);
                } else {   // fp32 pairs from L1 (half the bytes of the pre-split form), split here
                    const float2 b2 = BREG2 ? bfrag2[kt * NT + n] : __ldg(wf2 + (kt * NT + n) * 32);
                    split_rn(b2.x, bh0, bl0);
                    split_rn(b2.y, bh1, bl1);
                }
                mma_tf32(d[n], al[0], al[1], al[2], al[3], bh0, bh1);
                mma_tf32(d[n], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                mma_tf32(d[n], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
            }
        }
        if constexpr (BULK != 0) {
            using BT = BulkTile<NT>;
            float* bt = bulk_smem + warp * BT::WARP_WORDS + bulk_par * (16 * BT::ROWW);
            // the bulk reads of this buffer must be over before it is rewritten
            if constexpr (BT::NBUF == 2) {
                bulk_par ^= 1;
                if (lane < 16) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            } else {
                if (lane < 16) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
            if constexpr (BULK == 2) {
                const int fg = cg.own >= 0 ? (int)(((int64_t)cg.own * a.ldo) & 3) : 0;
                const int fh = ch.own >= 0 ? (int)(((int64_t)ch.own * a.ldo) & 3) : 0;
                float* rg = bt + g * BT::ROWW;
                float* rh = bt + (g + 8) * BT::ROWW;
                rg[t < fg ? t : NT * 8 + t] = 0.f;   // the 4 window cells outside [f, f + 8 NT)
                rh[t < fh ? t : NT * 8 + t] = 0.f;
                rg += fg + 2 * t;
                rh += fh + 2 * t;
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    rg[8 * n] = d[n][0];
                    rg[8 * n + 1] = d[n][1];
                    rh[8 * n] = d[n][2];
                    rh[8 * n + 1] = d[n][3];
                }
            } else {
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    *reinterpret_cast<float2*>(bt + g * BT::ROWW + 8 * n + 2 * t) = make_float2(d[n][0], d[n][1]);
                    *reinterpret_cast<float2*>(bt + (g + 8) * BT::ROWW + 8 * n + 2 * t) = make_float2(d[n][2], d[n][3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const int o1 = __shfl_sync(FULL, cg.own, (lane & 7) * 4);
            const int o2 = __shfl_sync(FULL, ch.own, (lane & 7) * 4);
            const int own = lane < 8 ? o1 : o2;
            if (lane < 16) {
                if (own >= 0) {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(bt + lane * BT::ROWW);
                    if constexpr (BULK == 2) {
                        const int64_t off = (int64_t)own * a.ldo, w0 = off & ~(int64_t)3;
                        if (w0 + NT * 8 + 4 <= a.out_elems) {
                            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                         ::"l"(a.out + w0), "r"(src), "n"(NT * 32 + 16) : "memory");
                        } else {   // the window of the very last rows would pass the end of the buffer
                            const float* srow = bt + lane * BT::ROWW + (int)(off & 3);
                            for (int c = 0; c < a.nout; ++c) atomicAdd(a.out + off + c, srow[c]);
                        }
                    } else {
                        float* dst = a.out + (int64_t)own * a.ldo;
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                     ::"l"(dst), "r"(src), "n"(NT * 32) : "memory");
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // lanes t, t^1 swap halves so that each holds 4 consecutive columns of one n-tile pair
            const bool odd = (t & 1) != 0;
            const uint32_t ldo = (uint32_t)a.ldo;
#pragma unroll
            for (int j = 0; j < NT / 2; ++j) {
                const int col = 16 * j + (odd ? 6 : 0) + 2 * t;   // odd: 8(2j+1) + 2(t-1)
#pragma unroll
                for (int h = 0; h < 2; ++h) {   // h = 0: entry g ; h = 1: entry g + 8
                    const float p0 = d[2 * j][2 * h], p1 = d[2 * j][2 * h + 1];
                    const float q0 = d[2 * j + 1][2 * h], q1 = d[2 * j + 1][2 * h + 1];
                    const float rx = __shfl_xor_sync(FULL, odd ? p0 : q0, 1);
                    const float ry = __shfl_xor_sync(FULL, odd ? p1 : q1, 1);
                    const int own = h ? ch.own : cg.own;
                    if (own >= 0 && col < a.nout)   // row offsets fit 32 bits (checked by the caller)
                        red_add_v4(a.out + ((uint32_t)own * ldo + (uint32_t)col), odd ? rx : p0, odd ? ry : p1,
                                   odd ? q0 : rx, odd ? q1 : ry);
                }
            }
        }
    };

    if constexpr (STAGE != 0) {
        constexpr int D = RowStage::D, NB = RowStage::NB;
        uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem) + warp * NB;
        float* rows = reinterpret_cast<float*>(dyn_smem + RowStage::BAR_WORDS + RowStage::META_WORDS) + warp * (D * RowStage::SLOT_WORDS);
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) mbar_init(bars + b, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        const int ns = (num_units - gw + nw - 1) / nw;   // units of this warp: gw + s * nw, s < ns
        // lane 0: index streams of unit s -> buffer s % NB (bulk copies, completion counted on the buffer's mbarrier)
        // (the entry span of the unit staged NEXT is read one staging step ahead, so its latency is hidden)
        int sp_first = 0, sp_end = 0;
        unit_span(gw, sp_first, sp_end);
        auto stage_meta = [&](int s_) {
            const int first = sp_first, end = sp_end;
            if (s_ + 1 < ns) unit_span(gw + (s_ + 1) * nw, sp_first, sp_end);
            if (lane == 0) {
                const int u = gw + s_ * nw;
                const int eb = first & ~3;
                const uint32_t bytes = (uint32_t)((end - eb + 3) >> 2) * 16u;
                uint32_t* m = meta + (s_ % NB) * MW;
                uint64_t* bar = bars + (s_ % NB);
                mbar_expect_tx(bar, 3u * bytes + 2u * UTU * 4u);
                bulk_g2s(m, a.e_idx + eb, bytes, bar);
                bulk_g2s(m + SPAN, a.e_w + eb, bytes, bar);
                bulk_g2s(m + 2 * SPAN, a.e_own + eb, bytes, bar);
                bulk_g2s(m + 3 * SPAN, a.tile_e0 + u * UTU, UTU * 4, bar);
                bulk_g2s(m + 3 * SPAN + UTU, a.tile_info + u * UTU, UTU * 4, bar);
            }
        };
        struct Cursor {   // position in this warp's tile stream
            int s, i, nt, eb;
            const uint32_t* m;
        };
        auto enter = [&](Cursor& c) {   // the unit's index streams have landed (idempotent per phase)
            if (c.s >= ns) return;
            mbar_wait(bars + (c.s % NB), (uint32_t)((c.s / NB) & 1));
            c.m = meta + (c.s % NB) * MW;
            c.eb = (int)c.m[3 * SPAN] & ~3;
            const int u = gw + c.s * nw;
            c.nt = min(a.num_tiles, u * UTU + UTU) - u * UTU;
            c.i = 0;
        };
        auto issue_rows = [&](const Cursor& c, int slot) {
            const int e0 = (int)c.m[3 * SPAN + c.i], cnt = (int)c.m[3 * SPAN + UTU + c.i] & 0xff;
            float* dst = rows + slot * RowStage::SLOT_WORDS + 4 * lane;   // row (lane >> 2) + 8 k, quad (lane & 3)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int r = 8 * k + g;
                const bool valid = r < cnt && 4 * t < kin;
                const uint32_t idx = valid ? (c.m[e0 - c.eb + r] & IDX_MASK) : 0u;
                const bool real = idx < n_rows;
                const float* src = (real ? feat_t : aux_t) + (real ? idx * ldf : (idx - n_rows) * (uint32_t)KP);
                cp_async16_zfill(dst + k * 128, src, valid ? 16 : 0);
            }
        };
        stage_meta(0);
        if (ns > 1) stage_meta(1);
        if (ns > 2) stage_meta(2);
        Cursor cc{0, 0, 0, 0, nullptr}, ic;
        enter(cc);
        ic = cc;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (ic.s < ns) {
                issue_rows(ic, d);
                if (++ic.i == ic.nt) {
                    ++ic.s;
                    enter(ic);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int k = 0; cc.s < ns; ++k) {
            asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
            __syncwarp();
            const int slot = k % D;
            TileRef cur;
            {
                const int e0 = (int)cc.m[3 * SPAN + cc.i], info = (int)cc.m[3 * SPAN + UTU + cc.i];
                cur.rel = info >> 8;
                const int o = e0 - cc.eb + g, cnt = info & 0xff;
                cur.g8 = make_row(cc.m, o, g < cnt);
                cur.h8 = make_row(cc.m, o + 8, g + 8 < cnt);
            }
            float av[KT][4];
            {
                const float4 vg = *reinterpret_cast<const float4*>(rows + slot * RowStage::SLOT_WORDS + g * 16 + 4 * t);
                const float4 vh = *reinterpret_cast<const float4*>(rows + slot * RowStage::SLOT_WORDS + (g + 8) * 16 + 4 * t);
                av[0][0] = vg.x, av[0][2] = vg.y, av[1][0] = vg.z, av[1][2] = vg.w;
                av[0][1] = vh.x, av[0][3] = vh.y, av[1][1] = vh.z, av[1][3] = vh.w;
            }
            __syncwarp();   // every lane has read the slot before it is refilled
            if (ic.s < ns) {
                issue_rows(ic, slot);
                if (++ic.i == ic.nt) {
                    ++ic.s;
                    enter(ic);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            process(cur, av);
            if (++cc.i == cc.nt) {   // unit done: its buffer takes the unit NB - 1 ahead of the next one
                ++cc.s;
                __syncwarp();
                if (cc.s + NB - 1 < ns) stage_meta(cc.s + NB - 1);
                enter(cc);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if constexpr (BULK != 0) {
            if (lane < 16) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    int cur_first, cur_end, nxt_first = 0, nxt_end = 0;
    unit_span(gw, cur_first, cur_end);
    stage_unit(gw, cur_first, cur_end, meta);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (gw + nw < num_units) unit_span(gw + nw, nxt_first, nxt_end);
    int buf = 0;
    for (int unit = gw; unit < num_units; unit += nw, buf ^= 1) {
        const uint32_t* m = meta + buf * MW;
        const int eb = cur_first & ~3;
        const int nt = min(a.num_tiles, unit * UTU + UTU) - unit * UTU;
        // next unit: its metadata goes to the other buffer while this one is processed
        if (unit + nw < num_units) stage_unit(unit + nw, nxt_first, nxt_end, meta + (buf ^ 1) * MW);
        asm volatile("cp.async.commit_group;" ::: "memory");
        cur_first = nxt_first;
        cur_end = nxt_end;
        if (unit + 2 * nw < num_units) unit_span(unit + 2 * nw, nxt_first, nxt_end);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();

        // rows one tile ahead of the tensor-pipe work; two tiles per trip with swapped roles, so the
        // pipeline registers are never copied
        TileRef rA = tile_ref(m, 0, eb), rB = rA;
        float aA[KT][4], aB[KT][4];
        load_rows(rA, aA);
        for (int i = 0; i < nt; i += 2) {
            if (i + 1 < nt) {
                rB = tile_ref(m, i + 1, eb);
                load_rows(rB, aB);
            }
            process(rA, aA);
            if (i + 1 >= nt) break;
            if (i + 2 < nt) {
                rA = tile_ref(m, i + 2, eb);
                load_rows(rA, aA);
            }
            process(rB, aB);
        }
        __syncwarp();   // every lane is done with this metadata buffer before it is staged again
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if constexpr (BULK != 0) {
        if (lane < 16) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// dL/dW :  D[Kp x Np] += sum over the tile's entries  (w_e x[src_e])^T (x) gout[owner_e]
// lane (g,t): K slots = entries t, t+4 (first 8) and 8+t, 12+t (second 8); M rows = features 16m+g, +8
// ---------------------------------------------------------------------------------------------
// V4 (rows 16-byte addressable, Kp >= 32): lane (g,t) reads columns 32b+4g..+3 of its entries with
// LDG.128 (8 lanes x 16 B = 128 contiguous bytes per row) and those four values serve M rows
// (m-tile 2b, rows g / g+8) and (m-tile 2b+1, rows g / g+8): feature(m, h) = 32(m>>1) + 4g + 2(m&1) + h.
// M is an output index here, so the permutation is undone when D is flushed.
template <int KT, int NT, bool RELU, bool V4>
__global__ void __launch_bounds__(EW * 32, (KT * NT >= 32) ? 1 : 2) k_ewgrad(const ETileArgs a) {
    constexpr int KP = KT * 8, MT = KP / 16;
    static_assert(!V4 || KP >= 32, "vector loads need at least 32 columns");
    using MS = MetaStage<(KT >= 4) ? 1 : 0>;   // wide rows: smaller staging units leave more L1 for the gout rows
    constexpr int UTU = MS::UTU, SPAN = MS::SPAN, MW = MS::MW;
    extern __shared__ __align__(16) uint32_t dyn_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int gw = blockIdx.x * EW + warp, nw = gridDim.x * EW;
    uint32_t* meta = dyn_smem + warp * MS::WARP_WORDS;
    const uint64_t pol_f =
        (uint64_t)a.n_rows * (uint64_t)a.ldf * 4ull <= (112ull << 20) ? policy_evict_last() : policy_evict_normal();
    float d[MT][NT][4];
    float bsum[NT];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) d[m][n][0] = d[m][n][1] = d[m][n][2] = d[m][n][3] = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) bsum[n] = 0.f;
    int cur_rel = -1;

    auto flush = [&](int rel) {
        float* dst = rel == a.self_rel ? a.groot : (a.gweight ? a.gweight + (int64_t)rel * a.kin * a.nout : nullptr);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int h = (i & 2) ? 1 : 0;
                    const int row = V4 ? 32 * (m >> 1) + 4 * g + 2 * (m & 1) + h : 16 * m + g + 8 * h;
                    const int col = 8 * n + 2 * t + (i & 1);
                    if (dst && row < a.kin && col < a.nout) atomicAdd(dst + (int64_t)row * a.nout + col, d[m][n][i]);
                    d[m][n][i] = 0.f;
                }
            }
    };

    constexpr int KOFF = V4 ? 4 : 1;
    const float* const feat_g = pin(a.feat + KOFF * g);   // lane bases: feature rows / chunk rows / gout rows
    const float* const aux_g = pin(a.aux + KOFF * g);
    const float* const gout_g = pin(a.gout + g);
    const uint32_t n_rows = pin((uint32_t)a.n_rows), ldf = pin((uint32_t)a.ldf), ldg = pin((uint32_t)a.ldg);
    const int kin = pin(a.kin), nout = pin(a.nout);
    // contiguous span of tiles per warp (tiles cost the same), a multiple of the staging unit: D stays
    // in registers across the whole span and is flushed only where the relation changes
    const int per = ((a.num_tiles + nw - 1) / nw + UTU - 1) / UTU * UTU;
    const int t0 = min(a.num_tiles, gw * per), t1 = min(a.num_tiles, gw * per + per);
    if (t0 >= t1) return;
    const int num_units = (t1 - t0 + UTU - 1) / UTU;

    auto unit_span = [&](int u, int& first, int& end) {
        const int u0 = t0 + u * UTU, u1 = min(t1, u0 + UTU);
        first = a.tile_e0[u0];
        end = a.tile_e0[u1 - 1] + (a.tile_info[u1 - 1] & 0xff);
    };
    auto stage_unit = [&](int u, int first, int end, uint32_t* m) {
        const int eb = first & ~3;
        const int nchunk = (end - eb + 3) >> 2;
        for (int c = lane; c < nchunk; c += 32) {
            cp_async16(m + 4 * c, a.e_idx + eb + 4 * c);
            cp_async16(m + SPAN + 4 * c, a.e_w + eb + 4 * c);
            cp_async16(m + 2 * SPAN + 4 * c, a.e_own + eb + 4 * c);
        }
        const int u0 = t0 + u * UTU;
        if (lane < UTU / 4) cp_async16(m + 3 * SPAN + 4 * lane, a.tile_e0 + u0 + 4 * lane);
        else if (lane < UTU / 2) cp_async16(m + 3 * SPAN + UTU + 4 * (lane - UTU / 4), a.tile_info + u0 + 4 * (lane - UTU / 4));
    };
    struct Tile {
        RowRef r[4];   // entries t, t+4, 8+t, 12+t ; feature offset of lane g folded in
        int rel;
    };
    auto tile_ref = [&](const uint32_t* m, int i, int eb) {
        Tile tl;
        const int e0 = (int)m[3 * SPAN + i], info = (int)m[3 * SPAN + UTU + i];
        tl.rel = info >> 8;
        const int cnt = info & 0xff;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            RowRef r;
            const int o = e0 - eb + t + 4 * q;
            uint32_t idx = 0;
            r.w = 0.f;
            r.own = -1;
            if (t + 4 * q < cnt) {
                idx = m[o] & IDX_MASK;
                r.w = __uint_as_float(m[SPAN + o]);
                r.own = (int)m[2 * SPAN + o];
            }
            r.real = idx < n_rows;
            const uint32_t off = r.real ? idx * ldf : (idx - n_rows) * (uint32_t)KP;
            r.p = (r.real ? feat_g : aux_g) + off;
            tl.r[q] = r;
        }
        return tl;
    };
    auto load_tile = [&](const Tile& tl, float (&av)[2][MT][4], float (&bv)[2][NT][2]) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const RowRef& ra = tl.r[2 * ks];       // K slot t
            const RowRef& rb = tl.r[2 * ks + 1];   // K slot t + 4
            if constexpr (V4) {
#pragma unroll
                for (int jb = 0; jb < MT / 2; ++jb) {
                    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                    if (32 * jb + 4 * g < kin) {
                        va = ldg128_hint(reinterpret_cast<const float4*>(ra.p + 32 * jb), pol_f);
                        vb = ldg128_hint(reinterpret_cast<const float4*>(rb.p + 32 * jb), pol_f);
                    }
                    av[ks][2 * jb][0] = va.x;       // m-tile 2jb  : row g   (K slot t)
                    av[ks][2 * jb][1] = va.y;       //               row g+8
                    av[ks][2 * jb + 1][0] = va.z;   // m-tile 2jb+1: row g
                    av[ks][2 * jb + 1][1] = va.w;   //               row g+8
                    av[ks][2 * jb][2] = vb.x;       // same rows, K slot t+4
                    av[ks][2 * jb][3] = vb.y;
                    av[ks][2 * jb + 1][2] = vb.z;
                    av[ks][2 * jb + 1][3] = vb.w;
                }
            } else {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const bool c_lo = 16 * m + g < kin, c_hi = 16 * m + 8 + g < kin;
                    av[ks][m][0] = c_lo ? ldg_hint(ra.p + 16 * m, pol_f) : 0.f;
                    av[ks][m][1] = c_hi ? ldg_hint(ra.p + 16 * m + 8, pol_f) : 0.f;
                    av[ks][m][2] = c_lo ? ldg_hint(rb.p + 16 * m, pol_f) : 0.f;
                    av[ks][m][3] = c_hi ? ldg_hint(rb.p + 16 * m + 8, pol_f) : 0.f;
                }
            }
            const float* ga = gout_g + (uint32_t)max(ra.own, 0) * ldg;   // row offsets fit 32 bits (caller)
            const float* gb = gout_g + (uint32_t)max(rb.own, 0) * ldg;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const bool cn = 8 * n + g < nout;
                bv[ks][n][0] = (cn && ra.own >= 0) ? __ldg(ga + 8 * n) : 0.f;
                bv[ks][n][1] = (cn && rb.own >= 0) ? __ldg(gb + 8 * n) : 0.f;
            }
        }
    };
    auto compute = [&](const Tile& tl, const float (&av)[2][MT][4], const float (&bv)[2][NT][2]) {
        const int rel = tl.rel;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const RowRef& ra = tl.r[2 * ks];
            const RowRef& rb = tl.r[2 * ks + 1];
            uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                split_rn(bv[ks][n][0], bh[n][0], bl[n][0]);
                split_rn(bv[ks][n][1], bh[n][1], bl[n][1]);
                if (rel == a.self_rel) bsum[n] += bv[ks][n][0] + bv[ks][n][1];
            }
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                float x0 = av[ks][m][0], x1 = av[ks][m][1], x2 = av[ks][m][2], x3 = av[ks][m][3];
                if (RELU) {
                    if (ra.real) {
                        x0 = fmaxf(x0, 0.f);
                        x1 = fmaxf(x1, 0.f);
                    }
                    if (rb.real) {
                        x2 = fmaxf(x2, 0.f);
                        x3 = fmaxf(x3, 0.f);
                    }
                }
                uint32_t ah[4], al[4];
                split_fast(x0 * ra.w, ah[0], al[0]);
                split_fast(x1 * ra.w, ah[1], al[1]);
                split_fast(x2 * rb.w, ah[2], al[2]);
                split_fast(x3 * rb.w, ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    mma_tf32(d[m][n], al[0], al[1], al[2], al[3], bh[n][0], bh[n][1]);
                    mma_tf32(d[m][n], ah[0], ah[1], ah[2], ah[3], bl[n][0], bl[n][1]);
                    mma_tf32(d[m][n], ah[0], ah[1], ah[2], ah[3], bh[n][0], bh[n][1]);
                }
            }
        }
    };
    auto wanted = [&](int rel) {
        return rel == a.self_rel ? (a.groot != nullptr || a.gbias != nullptr) : a.gweight != nullptr;
    };
    auto enter = [&](int rel) {   // D belongs to one relation at a time
        if (rel != cur_rel) {
            if (cur_rel >= 0) flush(cur_rel);
            cur_rel = rel;
        }
    };

    int cur_first, cur_end, nxt_first = 0, nxt_end = 0;
    unit_span(0, cur_first, cur_end);
    stage_unit(0, cur_first, cur_end, meta);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (1 < num_units) unit_span(1, nxt_first, nxt_end);
    for (int u = 0, buf = 0; u < num_units; ++u, buf ^= 1) {
        const uint32_t* m = meta + buf * MW;
        const int eb = cur_first & ~3;
        const int nt = min(t1, t0 + u * UTU + UTU) - (t0 + u * UTU);
        if (u + 1 < num_units) stage_unit(u + 1, nxt_first, nxt_end, meta + (buf ^ 1) * MW);
        asm volatile("cp.async.commit_group;" ::: "memory");
        cur_first = nxt_first;
        cur_end = nxt_end;
        if (u + 2 < num_units) unit_span(u + 2, nxt_first, nxt_end);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        if constexpr (MT <= 2) {   // narrow rows: the next tile's rows are loaded before this tile's MMAs
            Tile tA = tile_ref(m, 0, eb), tB = tA;
            float aA[2][MT][4], bA[2][NT][2], aB[2][MT][4], bB[2][NT][2];
            bool wA = wanted(tA.rel), wB = false;
            if (wA) load_tile(tA, aA, bA);
            for (int i = 0; i < nt; i += 2) {
                if (i + 1 < nt) {
                    tB = tile_ref(m, i + 1, eb);
                    wB = wanted(tB.rel);
                    if (wB) load_tile(tB, aB, bB);
                }
                enter(tA.rel);
                if (wA) compute(tA, aA, bA);
                if (i + 1 >= nt) break;
                if (i + 2 < nt) {
                    tA = tile_ref(m, i + 2, eb);
                    wA = wanted(tA.rel);
                    if (wA) load_tile(tA, aA, bA);
                }
                enter(tB.rel);
                if (wB) compute(tB, aB, bB);
            }
        } else {
            for (int i = 0; i < nt; ++i) {
                const Tile tl = tile_ref(m, i, eb);
                enter(tl.rel);
                if (!wanted(tl.rel)) continue;
                float av[2][MT][4], bv[2][NT][2];
                load_tile(tl, av, bv);
                compute(tl, av, bv);
            }
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (cur_rel >= 0) flush(cur_rel);
    if (a.gbias) {   // column 8n+g summed over this lane's K slots; fold the 4 t-lanes, then one atomic
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            float sacc = bsum[n];
            sacc += __shfl_xor_sync(FULL, sacc, 1);
            sacc += __shfl_xor_sync(FULL, sacc, 2);
            if (t == 0 && 8 * n + g < a.nout && sacc != 0.f) atomicAdd(a.gbias + 8 * n + g, sacc);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// dL/dW on the TRANSPOSED structure (BWD: owner = src, gathered = dst) for layers that are wider on the input
// side than on the output side (63 -> 16):
//     T_r[16 x 8NT] += sum over the tile's entries  (w_e gout[dst_e])^T (x) x[src_e]          ( = dW_r^T )
// The relation-major form above gathers the WIDE rows x[src] at random (256 B per entry); here the random
// gather is the narrow gout row (64 B, the same rows the dL/dx pass gathers) and the wide row is read by OWNER
// id, which ascends inside every (range, relation) group.  All warps walk the tile sequence together — a warp's
// work is `upw` consecutive units inside each window of gridDim * EW * upw units — so the x rows of the few
// node ranges in flight (4 MB each) are served from L2: x is read from HBM once, not once per entry.
// Lane (g,t): A = gathered rows (M = gout column g / g+8, K slot = entry t / t+4); B = owner rows, read with
// LDG.128: columns 32b + 4g .. +3 of entry t serve n-tiles 4b .. 4b+3 at column g, i.e. D column (n = 4b + j, c)
// is x column 32b + 4c + j — undone by k_ewgrad_t_finish, which also transposes T into dW / droot.
// ---------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(EW * 32, 2) k_ewgrad_t(const ETileArgs a, const int upw) {
    using MS = MetaStage<0>;
    constexpr int UTU = MS::UTU, SPAN = MS::SPAN, MW = MS::MW, NB = NT / 4, LDT = NT * 8;
    extern __shared__ __align__(16) uint32_t dyn_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int gw = blockIdx.x * EW + warp, nw = gridDim.x * EW;
    uint32_t* meta = dyn_smem + warp * MS::WARP_WORDS;
    const uint64_t pol_x = policy_evict_last();
    float d[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) d[n][0] = d[n][1] = d[n][2] = d[n][3] = 0.f;
    float asum0 = 0.f, asum1 = 0.f;
    int cur_rel = -1;

    // row g (even t) or row g + 8 (odd t) of n-tile n as ONE 16-byte reduction: the lane pair (t, t^1) swaps halves
    auto flush = [&](int rel) {
        float* dst = a.gweight + (int64_t)rel * 16 * LDT + (int64_t)((t & 1) ? g + 8 : g) * LDT + ((t & 1) ? 2 * (t - 1) : 2 * t);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const bool odd = t & 1;
            const float s0 = odd ? d[n][0] : d[n][2], s1 = odd ? d[n][1] : d[n][3];
            const float r0 = __shfl_xor_sync(FULL, s0, 1), r1 = __shfl_xor_sync(FULL, s1, 1);
            if (odd) red_add_v4(dst + 8 * n, r0, r1, d[n][2], d[n][3]);
            else red_add_v4(dst + 8 * n, d[n][0], d[n][1], r0, r1);
            d[n][0] = d[n][1] = d[n][2] = d[n][3] = 0.f;
        }
    };

    const float* const feat_g = pin(a.feat + g);
    const float* const aux_g = pin(a.aux + g);
    const float* const x_g = pin(a.gout + 4 * g);
    const uint32_t n_rows = pin((uint32_t)a.n_rows), ldf = pin((uint32_t)a.ldf), ldx = pin((uint32_t)a.ldg);
    const int kin = pin(a.kin);
    const int total_units = (a.num_tiles + UTU - 1) / UTU;
    // unit u of this warp -> global unit: window (u / upw), this warp's block of upw units inside it
    auto unit_of = [&](int u) { return ((u / upw) * nw + gw) * upw + (u % upw); };
    const int win_units = nw * upw;
    int num_units = 0;   // units of this warp that exist
    {
        const int full = total_units / win_units, rem = total_units - full * win_units;
        num_units = full * upw + max(0, min(upw, rem - gw * upw));
    }
    if (num_units == 0) return;

    auto unit_span = [&](int u, int& first, int& end) {
        const int u0 = unit_of(u) * UTU, u1 = min(a.num_tiles, u0 + UTU);
        first = a.tile_e0[u0];
        end = a.tile_e0[u1 - 1] + (a.tile_info[u1 - 1] & 0xff);
    };
    auto stage_unit = [&](int u, int first, int end, uint32_t* m) {
        const int eb = first & ~3;
        const int nchunk = (end - eb + 3) >> 2;
        for (int c = lane; c < nchunk; c += 32) {
            cp_async16(m + 4 * c, a.e_idx + eb + 4 * c);
            cp_async16(m + SPAN + 4 * c, a.e_w + eb + 4 * c);
            cp_async16(m + 2 * SPAN + 4 * c, a.e_own + eb + 4 * c);
        }
        const int u0 = unit_of(u) * UTU;
        if (lane < UTU / 4) cp_async16(m + 3 * SPAN + 4 * lane, a.tile_e0 + u0 + 4 * lane);
        else if (lane < UTU / 2) cp_async16(m + 3 * SPAN + UTU + 4 * (lane - UTU / 4), a.tile_info + u0 + 4 * (lane - UTU / 4));
    };
    // a tile is kept as (first entry relative to the staged span, entry count, relation): indices, owners and
    // weights are re-read from the staged metadata where they are used, so two tiles in flight cost 6 registers
    struct Tile {
        int o0, cnt, rel;
    };
    auto tile_ref = [&](const uint32_t* m, int i, int eb) {
        Tile tl;
        const int info = (int)m[3 * SPAN + UTU + i];
        tl.o0 = (int)m[3 * SPAN + i] - eb + t;
        tl.rel = info >> 8;
        tl.cnt = info & 0xff;
        return tl;
    };
    const bool c_lo = g < kin, c_hi = g + 8 < kin;
    // K step ks of a tile: entries t + 8 ks (slot t) and t + 4 + 8 ks (slot t + 4)
    auto load_half = [&](const uint32_t* m, const Tile& tl, int ks, float (&av)[4], float4 (&bv)[2][NB]) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int q = 2 * ks + r;
            const bool valid = t + 4 * q < tl.cnt;
            const int o = tl.o0 + 4 * q;
            const uint32_t idx = valid ? (m[o] & IDX_MASK) : 0u;
            const uint32_t own = valid ? m[2 * SPAN + o] : 0u;
            const float* p = idx < n_rows ? feat_g + idx * ldf : aux_g + (idx - n_rows) * 16u;
            av[2 * r] = (c_lo && valid) ? __ldg(p) : 0.f;
            av[2 * r + 1] = (c_hi && valid) ? __ldg(p + 8) : 0.f;
            const float* xp = x_g + own * ldx;
#pragma unroll
            for (int b = 0; b < NB; ++b)
                bv[r][b] = valid ? ldg128_hint(reinterpret_cast<const float4*>(xp + 32 * b), pol_x) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto compute_half = [&](const uint32_t* m, const Tile& tl, int ks, const float (&av)[4], const float4 (&bv)[2][NB]) {
        const int oa = tl.o0 + 8 * ks, ob = oa + 4;
        const float wa = (t + 8 * ks < tl.cnt) ? __uint_as_float(m[SPAN + oa]) : 0.f;
        const float wb = (t + 4 + 8 * ks < tl.cnt) ? __uint_as_float(m[SPAN + ob]) : 0.f;
        if (tl.rel == a.self_rel) {
            asum0 += av[0] + av[2];
            asum1 += av[1] + av[3];
        }
        uint32_t ah[4], al[4];
        split_fast(av[0] * wa, ah[0], al[0]);
        split_fast(av[1] * wa, ah[1], al[1]);
        split_fast(av[2] * wb, ah[2], al[2]);
        split_fast(av[3] * wb, ah[3], al[3]);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float xa[4] = {bv[0][b].x, bv[0][b].y, bv[0][b].z, bv[0][b].w};
            const float xb[4] = {bv[1][b].x, bv[1][b].y, bv[1][b].z, bv[1][b].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t bh0, bl0, bh1, bl1;
                split_rn(xa[j], bh0, bl0);
                split_rn(xb[j], bh1, bl1);
                float(&dd)[4] = d[4 * b + j];
                mma_tf32(dd, al[0], al[1], al[2], al[3], bh0, bh1);
                mma_tf32(dd, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                mma_tf32(dd, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
            }
        }
    };
    auto wanted = [&](int rel) { return rel == a.self_rel ? (a.groot != nullptr || a.gbias != nullptr) : a.out != nullptr; };
    auto enter = [&](int rel) {
        if (rel != cur_rel) {
            if (cur_rel >= 0 && wanted(cur_rel)) flush(cur_rel);
            cur_rel = rel;
        }
    };

    int cur_first, cur_end, nxt_first = 0, nxt_end = 0;
    unit_span(0, cur_first, cur_end);
    stage_unit(0, cur_first, cur_end, meta);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (1 < num_units) unit_span(1, nxt_first, nxt_end);
    for (int u = 0, buf = 0; u < num_units; ++u, buf ^= 1) {
        const uint32_t* m = meta + buf * MW;
        const int eb = cur_first & ~3;
        const int ut0 = unit_of(u) * UTU;
        const int nt = min(a.num_tiles, ut0 + UTU) - ut0;
        if (u + 1 < num_units) stage_unit(u + 1, nxt_first, nxt_end, meta + (buf ^ 1) * MW);
        asm volatile("cp.async.commit_group;" ::: "memory");
        cur_first = nxt_first;
        cur_end = nxt_end;
        if (u + 2 < num_units) unit_span(u + 2, nxt_first, nxt_end);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        // (a variant that kept the next tile's K step in flight under this tile's MMAs spilled at 128 registers and
        // ran 20 % slower; latency is hidden by the 16 resident warps instead)
        for (int i = 0; i < nt; ++i) {
            const Tile tl = tile_ref(m, i, eb);
            enter(tl.rel);
            if (!wanted(tl.rel)) continue;
            float av0[4], av1[4];
            float4 bv0[2][NB], bv1[2][NB];
            load_half(m, tl, 0, av0, bv0);
            load_half(m, tl, 1, av1, bv1);
            compute_half(m, tl, 0, av0, bv0);
            compute_half(m, tl, 1, av1, bv1);
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (cur_rel >= 0 && wanted(cur_rel)) flush(cur_rel);
    if (a.gbias) {   // gout column g / g+8 summed over this lane's K slots of the self-loop tiles
        asum0 += __shfl_xor_sync(FULL, asum0, 1);
        asum0 += __shfl_xor_sync(FULL, asum0, 2);
        asum1 += __shfl_xor_sync(FULL, asum1, 1);
        asum1 += __shfl_xor_sync(FULL, asum1, 2);
        if (t == 0 && g < kin && asum0 != 0.f) atomicAdd(a.gbias + g, asum0);
        if (t == 0 && g + 8 < kin && asum1 != 0.f) atomicAdd(a.gbias + g + 8, asum1);
    }
}

// T[r][c][v] (c = gout column, v = 8n + c' with x column 32(n/4) + 4c' + n%4)  ->  dW[r][x column][c] / droot
__global__ void __launch_bounds__(256) k_ewgrad_t_finish(const float* __restrict__ T, int ldt, int R, int fin, int fout,
                                                         float* __restrict__ gweight, float* __restrict__ groot) {
    const int per = fin * fout;
    const int64_t total = (int64_t)(R + 1) * per;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / per), e = (int)(i - (int64_t)r * per);
        const int xc = e / fout, c = e - xc * fout;
        const int v = 8 * (4 * (xc >> 5) + (xc & 3)) + ((xc & 31) >> 2);
        const float val = T[((int64_t)r * 16 + c) * ldt + v];
        if (r < R) {
            if (gweight) gweight[i] = val;
        } else if (groot) {
            groot[e] = val;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// self loop + bias:  out[i] = act(x[i]) . root + bias  with PLAIN vector stores.  Runs before the
// edge tiles of a pass, so it also replaces the zero fill of the accumulate target; the self-loop
// tiles (which sort last in every BRC) are then skipped by k_etile.
// ---------------------------------------------------------------------------------------------
struct SelfArgs {
    const float* x;      // row 0 = first OWNED node
    int64_t ldx;
    int kin;
    const float4* rfrag;   // root fragments (slot R of wfrag)
    const float* bias;
    int nbias;
    float* out;
    int64_t ldo;
    int nout;
    int64_t n_own;
    float* mirror;       // k_selfloop_pad: zero-padded 16-byte addressable copy of x, written on the way
    int64_t ldm;
};

// PACK: `out` rows are tightly packed with a width that is not a multiple of 4 (ldo == nout): the
// warp's 16 rows are one contiguous, 64-byte aligned span, staged in shared memory and written flat.
template <int KT, int NT, bool RELU, bool V4, bool PACK = false>
__global__ void __launch_bounds__(EW * 32, 2) k_selfloop(const SelfArgs a) {
    constexpr bool BREG = (KT * NT <= 16);
    __shared__ __align__(16) float pack_stage[PACK ? EW * 16 * NT * 8 : 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t gw = (int64_t)blockIdx.x * EW + warp, nw = (int64_t)gridDim.x * EW;
    const float4* wf = a.rfrag + lane;
    float4 bfrag[BREG ? KT * NT : 1];
    if constexpr (BREG) {
#pragma unroll
        for (int i = 0; i < KT * NT; ++i) bfrag[i] = __ldg(wf + i * 32);
    }
    float bx[NT], by[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const int col = 8 * n + 2 * t;
        bx[n] = (a.bias && col < a.nbias) ? a.bias[col] : 0.f;
        by[n] = (a.bias && col + 1 < a.nbias) ? a.bias[col + 1] : 0.f;
    }
    constexpr int KOFF = V4 ? 4 : 1;
    const int64_t n_tiles = (a.n_own + 15) / 16;
    for (int64_t tile = gw; tile < n_tiles; tile += nw) {
        const int64_t ig = tile * 16 + g, ih = ig + 8;
        const bool vg = ig < a.n_own, vh = ih < a.n_own;
        const float* pg = a.x + (vg ? ig : 0) * a.ldx + KOFF * t;
        const float* ph = a.x + (vh ? ih : 0) * a.ldx + KOFF * t;
        float av[KT][4];
        if constexpr (V4) {
#pragma unroll
            for (int j = 0; j < KT / 2; ++j) {
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (16 * j + 4 * t < a.kin) {
                    v0 = __ldg(reinterpret_cast<const float4*>(pg + 16 * j));
                    v1 = __ldg(reinterpret_cast<const float4*>(ph + 16 * j));
                }
                av[2 * j][0] = v0.x;
                av[2 * j][2] = v0.y;
                av[2 * j + 1][0] = v0.z;
                av[2 * j + 1][2] = v0.w;
                av[2 * j][1] = v1.x;
                av[2 * j][3] = v1.y;
                av[2 * j + 1][1] = v1.z;
                av[2 * j + 1][3] = v1.w;
            }
        } else {
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const bool c_lo = 8 * kt + t < a.kin, c_hi = 8 * kt + 4 + t < a.kin;
                av[kt][0] = c_lo ? __ldg(pg + 8 * kt) : 0.f;
                av[kt][1] = c_lo ? __ldg(ph + 8 * kt) : 0.f;
                av[kt][2] = c_hi ? __ldg(pg + 8 * kt + 4) : 0.f;
                av[kt][3] = c_hi ? __ldg(ph + 8 * kt + 4) : 0.f;
            }
        }
        float d[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            d[n][0] = bx[n];
            d[n][1] = by[n];
            d[n][2] = bx[n];
            d[n][3] = by[n];
        }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            float x0 = av[kt][0], x1 = av[kt][1], x2 = av[kt][2], x3 = av[kt][3];
            if (RELU) {
                x0 = fmaxf(x0, 0.f);
                x1 = fmaxf(x1, 0.f);
                x2 = fmaxf(x2, 0.f);
                x3 = fmaxf(x3, 0.f);
            }
            uint32_t ah[4], al[4];
            split_fast(x0, ah[0], al[0]);
            split_fast(x1, ah[1], al[1]);
            split_fast(x2, ah[2], al[2]);
            split_fast(x3, ah[3], al[3]);
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
        if constexpr (PACK) {
            // the tile's 16 rows are ONE contiguous span of `out` (16 w floats, 64-byte aligned): staged
            // packed in shared memory and written by a single bulk copy (cp.async.bulk shared -> global)
            float* sb = pack_stage + warp * (16 * NT * 8);
            const int w = a.nout;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile's copy has read sb
            __syncwarp();
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int col = 8 * n + 2 * t;
                if (col < w) {
                    sb[g * w + col] = d[n][0];
                    sb[(g + 8) * w + col] = d[n][2];
                }
                if (col + 1 < w) {
                    sb[g * w + col + 1] = d[n][1];
                    sb[(g + 8) * w + col + 1] = d[n][3];
                }
            }
            const int rows = (int)min((int64_t)16, a.n_own - tile * 16);
            float* base = a.out + tile * 16 * w;
            const int bytes = rows * w * 4;
            if ((bytes & 15) == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base),
                                 "r"((uint32_t)__cvta_generic_to_shared(sb)), "r"(bytes)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {   // a last tile whose byte count is not a multiple of 16
                __syncwarp();
                for (int i = lane; i < rows * w; i += 32) base[i] = sb[i];
                __syncwarp();
            }
        } else {
        const bool odd = (t & 1) != 0;
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) {
            const int col = odd ? 8 * (2 * j + 1) + 2 * (t - 1) : 8 * (2 * j) + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float p0 = d[2 * j][2 * h], p1 = d[2 * j][2 * h + 1];
                const float q0 = d[2 * j + 1][2 * h], q1 = d[2 * j + 1][2 * h + 1];
                const float rx = __shfl_xor_sync(FULL, odd ? p0 : q0, 1);
                const float ry = __shfl_xor_sync(FULL, odd ? p1 : q1, 1);
                const int64_t row = h ? ih : ig;
                if ((h ? vh : vg) && col < a.nout) {
                    float4* p = reinterpret_cast<float4*>(a.out + row * a.ldo + col);
                    *p = odd ? make_float4(rx, ry, q0, q1) : make_float4(p0, p1, rx, ry);
                }
            }
        }
        }
    }
    if constexpr (PACK) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// self loop fused with the padding copy (rows of 33..64 floats that are NOT 16-byte addressable, the
// reference's emb = 63): the pass that makes the zero-padded mirror reads every row of x anyway, so
// it also multiplies it by root.  Row-wise coalesced 4-byte loads -> shared tile -> (a) the mirror
// rows as 16-byte stores, (b) the A fragments.  Replaces k_pad_rows + k_selfloop<8, NT>: x is read
// once instead of x once and the mirror once.
// ---------------------------------------------------------------------------------------------
template <int NT, bool RELU>
__global__ void __launch_bounds__(EW * 32, 2) k_selfloop_pad(const SelfArgs a) {
    constexpr int KT = 8, SW = 80;   // staged row stride: the 8 lanes of a 128-bit phase hit distinct banks
    static_assert(KT * NT <= 16, "fragments are kept in registers");
    __shared__ __align__(16) float stage[EW * 16 * SW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t gw = (int64_t)blockIdx.x * EW + warp, nw = (int64_t)gridDim.x * EW;
    float* sb = stage + warp * (16 * SW);
    const float4* wf = a.rfrag + lane;
    float4 bfrag[KT * NT];
#pragma unroll
    for (int i = 0; i < KT * NT; ++i) bfrag[i] = __ldg(wf + i * 32);
    float bx[NT], by[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const int col = 8 * n + 2 * t;
        bx[n] = (a.bias && col < a.nbias) ? a.bias[col] : 0.f;
        by[n] = (a.bias && col + 1 < a.nbias) ? a.bias[col + 1] : 0.f;
    }
    const bool c0 = lane < a.kin, c1 = lane + 32 < a.kin;
    const int64_t n_tiles = (a.n_own + 15) / 16;
    for (int64_t tile = gw; tile < n_tiles; tile += nw) {
        const int rows = (int)min((int64_t)16, a.n_own - tile * 16);
        const float* src = a.x + tile * 16 * a.ldx + lane;
        float v0[16], v1[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            v0[r] = (r < rows && c0) ? __ldg(src + r * a.ldx) : 0.f;
            v1[r] = (r < rows && c1) ? __ldg(src + r * a.ldx + 32) : 0.f;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            sb[r * SW + lane] = v0[r];          // columns >= kin are the zero padding
            sb[r * SW + lane + 32] = v1[r];
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // the mirror: 16 rows x 16 quads, 512 contiguous bytes per instruction
            const int i = lane + 32 * k, r = i >> 4, q = i & 15;
            if (r < rows && 4 * q < a.ldm)
                *reinterpret_cast<float4*>(a.mirror + (tile * 16 + r) * a.ldm + 4 * q) =
                    *reinterpret_cast<const float4*>(sb + r * SW + 4 * q);
        }
        float d[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            d[n][0] = bx[n];
            d[n][1] = by[n];
            d[n][2] = bx[n];
            d[n][3] = by[n];
        }
#pragma unroll
        for (int j = 0; j < KT / 2; ++j) {
            const float4 vg = *reinterpret_cast<const float4*>(sb + g * SW + 16 * j + 4 * t);
            const float4 vh = *reinterpret_cast<const float4*>(sb + (g + 8) * SW + 16 * j + 4 * t);
            // K order of the vector-load kernels: step 2j <- cols (x, y), step 2j+1 <- (z, w)
            float xs[2][4] = {{vg.x, vh.x, vg.y, vh.y}, {vg.z, vh.z, vg.w, vh.w}};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t ah[4], al[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) split_fast(RELU ? fmaxf(xs[h][q], 0.f) : xs[h][q], ah[q], al[q]);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const float4 bf = bfrag[(2 * j + h) * NT + n];
                    const uint32_t bh0 = __float_as_uint(bf.x), bh1 = __float_as_uint(bf.y);
                    const uint32_t bl0 = __float_as_uint(bf.z), bl1 = __float_as_uint(bf.w);
                    mma_tf32(d[n], al[0], al[1], al[2], al[3], bh0, bh1);
                    mma_tf32(d[n], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                    mma_tf32(d[n], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                }
            }
        }
        const int64_t ig = tile * 16 + g, ih = ig + 8;
        const bool vgd = ig < a.n_own, vhd = ih < a.n_own;
        const bool odd = (t & 1) != 0;
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) {
            const int col = odd ? 8 * (2 * j + 1) + 2 * (t - 1) : 8 * (2 * j) + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float p0 = d[2 * j][2 * h], p1 = d[2 * j][2 * h + 1];
                const float q0 = d[2 * j + 1][2 * h], q1 = d[2 * j + 1][2 * h + 1];
                const float rx = __shfl_xor_sync(FULL, odd ? p0 : q0, 1);
                const float ry = __shfl_xor_sync(FULL, odd ? p1 : q1, 1);
                const int64_t row = h ? ih : ig;
                if ((h ? vhd : vgd) && col < a.nout) {
                    float4* p = reinterpret_cast<float4*>(a.out + row * a.ldo + col);
                    *p = odd ? make_float4(rx, ry, q0, q1) : make_float4(p0, p1, rx, ry);
                }
            }
        }
    }
}

template <int KT, int NT>
int run_selfloop(const SelfArgs& a, bool relu, bool v4, bool packed, int num_sms, cudaStream_t st) {
    auto launch = [&](auto kern) -> int {
        const int64_t tiles = (a.n_own + 15) / 16;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((tiles + EW - 1) / EW, (int64_t)num_sms * 2));
        kern<<<grid, EW * 32, 0, st>>>(a);
        RGCN_CUDA(cudaGetLastError());
        return 0;
    };
    if constexpr (NT == 8) {
        if (packed) {
            if (!v4) return fail(RGCN_ERR_UNSUPPORTED, "packed self-loop pass needs 16-byte addressable gathered rows");
            if (relu) return launch(k_selfloop<KT, NT, true, true, true>);
            return launch(k_selfloop<KT, NT, false, true, true>);
        }
    }
    if (v4) {
        if (relu) return launch(k_selfloop<KT, NT, true, true>);
        return launch(k_selfloop<KT, NT, false, true>);
    }
    if (relu) return launch(k_selfloop<KT, NT, true, false>);
    return launch(k_selfloop<KT, NT, false, false>);
}

// rows of at least this many 8-column steps are scattered by the bulk-copy engine (0 = never)
int bulk_min_nt() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RGCN_B200_BULK");
        v = e ? atoi(e) : 8;
    }
    return v;
}

// RGCN_B200_STAGE=0 keeps the gathered rows of the 16-column kernels in registers (round-1 path)
bool row_stage_on() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RGCN_B200_STAGE");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

template <int KT, int NT>
int run_etile(const ETileArgs& a, bool relu, bool v4, bool packed, int num_sms, int cap, cudaStream_t st) {
    auto launch = [&](auto kern, int bulk_bytes = 0, bool staged = false) -> int {
        int per_sm = 1;
        const bool small_units = bulk_bytes != 0 || staged;
        const int smem = (staged ? RowStage::CTA_BYTES : small_units ? MetaStage<1>::CTA_BYTES : MetaStage<0>::CTA_BYTES) + bulk_bytes;
        const int utu = small_units ? MetaStage<1>::UTU : MetaStage<0>::UTU;
        if (smem > 48 * 1024) RGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        RGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EW * 32, smem));
        per_sm = std::max(per_sm, 1);
        if (cap > 0) per_sm = std::min(per_sm, cap);
        const int64_t units = ((int64_t)a.num_tiles + utu - 1) / utu;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((units + EW - 1) / EW, (int64_t)num_sms * per_sm));
        kern<<<grid, EW * 32, smem, st>>>(a);
        RGCN_CUDA(cudaGetLastError());
        return 0;
    };
    // bulk scatter: whole padded rows (8 NT floats) are added, so the target row must hold them
    // 16-column vector-addressable rows: rows staged in shared memory, RowStage::D tiles ahead
    [[maybe_unused]] const bool staged = KT == 2 && v4 && row_stage_on();
    if constexpr (NT == 8) {
        if (packed) {
            if (!v4) return fail(RGCN_ERR_UNSUPPORTED, "packed scatter needs 16-byte addressable gathered rows");
            if constexpr (KT == 2) {
                if (staged) {
                    if (relu) return launch(k_etile<KT, NT, true, true, 2, 1>, BulkTile<NT>::CTA_BYTES, true);
                    return launch(k_etile<KT, NT, false, true, 2, 1>, BulkTile<NT>::CTA_BYTES, true);
                }
            }
            if (relu) return launch(k_etile<KT, NT, true, true, 2>, BulkTile<NT>::CTA_BYTES);
            return launch(k_etile<KT, NT, false, true, 2>, BulkTile<NT>::CTA_BYTES);
        }
    }
    if (packed) return fail(RGCN_ERR_UNSUPPORTED, "packed scatter is built for 64-column rows only");
    const bool bulk = v4 && bulk_min_nt() > 0 && NT >= bulk_min_nt() && a.ldo >= 8 * NT && a.ldo % 4 == 0 &&
                      ((uintptr_t)a.out & 15) == 0;
    if (bulk) {
        if constexpr (KT == 2) {
            if (staged) {
                if (relu) return launch(k_etile<KT, NT, true, true, 1, 1>, BulkTile<NT>::CTA_BYTES, true);
                return launch(k_etile<KT, NT, false, true, 1, 1>, BulkTile<NT>::CTA_BYTES, true);
            }
        }
        if (relu) return launch(k_etile<KT, NT, true, true, 1>, BulkTile<NT>::CTA_BYTES);
        return launch(k_etile<KT, NT, false, true, 1>, BulkTile<NT>::CTA_BYTES);
    }
    if (v4) {
        if constexpr (KT == 2) {
            if (staged) {
                if (relu) return launch(k_etile<KT, NT, true, true, 0, 1>, 0, true);
                return launch(k_etile<KT, NT, false, true, 0, 1>, 0, true);
            }
        }
        if (relu) return launch(k_etile<KT, NT, true, true>);
        return launch(k_etile<KT, NT, false, true>);
    }
    if (relu) return launch(k_etile<KT, NT, true, false>);
    return launch(k_etile<KT, NT, false, false>);
}

template <int KT, int NT>
int run_ewgrad(const ETileArgs& a, bool relu, bool v4, int num_sms, int cap, cudaStream_t st) {
    auto launch = [&](auto kern) -> int {
        int per_sm = 1;
        const int smem = (KT >= 4) ? MetaStage<1>::CTA_BYTES : MetaStage<0>::CTA_BYTES;
        if (smem > 48 * 1024) RGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        RGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EW * 32, smem));
        per_sm = std::max(per_sm, 1);
        if (cap > 0) per_sm = std::min(per_sm, cap);
        const int64_t units = ((int64_t)a.num_tiles + UT - 1) / UT;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((units + EW - 1) / EW, (int64_t)num_sms * per_sm));
        kern<<<grid, EW * 32, smem, st>>>(a);
        RGCN_CUDA(cudaGetLastError());
        return 0;
    };
    if constexpr (KT >= 4) {
        if (v4) {
            if (relu) return launch(k_ewgrad<KT, NT, true, true>);
            return launch(k_ewgrad<KT, NT, false, true>);
        }
    }
    if (relu) return launch(k_ewgrad<KT, NT, true, false>);
    return launch(k_ewgrad<KT, NT, false, false>);
}

#define RGCN_DISPATCH_E(FN, kp, np, ...)                                          \
    do {                                                                          \
        const int _k = (kp) / 8, _n = (np) / 8;                                   \
        if (_k == 2 && _n == 2) return FN<2, 2>(__VA_ARGS__);                     \
        if (_k == 2 && _n == 4) return FN<2, 4>(__VA_ARGS__);                     \
        if (_k == 2 && _n == 8) return FN<2, 8>(__VA_ARGS__);                     \
        if (_k == 4 && _n == 2) return FN<4, 2>(__VA_ARGS__);                     \
        if (_k == 4 && _n == 4) return FN<4, 4>(__VA_ARGS__);                     \
        if (_k == 4 && _n == 8) return FN<4, 8>(__VA_ARGS__);                     \
        if (_k == 8 && _n == 2) return FN<8, 2>(__VA_ARGS__);                     \
        if (_k == 8 && _n == 4) return FN<8, 4>(__VA_ARGS__);                     \
        if (_k == 8 && _n == 8) return FN<8, 8>(__VA_ARGS__);                     \
        return fail(RGCN_ERR_UNSUPPORTED, "no entry-tile kernel for this padded shape"); \
    } while (0)

ETileArgs base_args(const Brc& b) {
    ETileArgs a{};
    a.e_idx = b.e_idx;
    a.e_w = b.e_w;
    a.e_own = b.e_own;
    a.tile_e0 = b.tile_e0;
    a.tile_info = b.tile_info;
    a.num_tiles = b.num_tiles;
    return a;
}

}  // namespace

// one LDG.128 per row per 16 columns needs 16-byte addressable rows and a width in whole quads
bool etile_vec4_ok(const float* feat, int64_t ldf, int kin, const float* aux) {
    static int off = -1;
    if (off < 0) {
        const char* e = getenv("RGCN_B200_VEC4");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    // rows may be wider than the logical width (zero-padded mirror): whole quads must stay inside a row
    return !off && ldf % 4 == 0 && ((kin + 3) & ~3) <= ldf && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)aux & 15) == 0;
}

// tightly packed odd-width target rows (the [N, 63] gradient): bulk-reduce scatter into shifted windows
bool etile_packed_ok(const float* out, int64_t ldo, int nout, int np) {
    static int off = -1;
    if (off < 0) {
        const char* e = getenv("RGCN_B200_PACKED");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    return !off && bulk_min_nt() > 0 && bulk_min_nt() <= 8 && np == 64 && ldo == nout && nout % 4 != 0 && nout > 56 &&
           ((uintptr_t)out & 15) == 0;
}

int launch_etile_pass(const TilePass& p, int num_sms, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_tiles == 0) return 0;
    ETileArgs a = base_args(b);
    a.feat = p.feat;
    a.ldf = p.ldf;
    a.kin = p.kin;
    a.aux = p.aux ? p.aux : p.feat;
    a.n_rows = p.n_nodes;
    a.wfrag = p.wfrag;
    a.wfrag2 = p.wfrag2;
    a.bias = p.bias;
    a.nbias = p.nbias;
    a.self_rel = p.self_rel;
    a.out = p.out;
    a.ldo = p.ldo;
    a.nout = p.nout;
    a.num_tiles = b.num_tiles_noself;   // the self loops were written by launch_selfloop_pass
    a.out_elems = p.out_rows * p.ldo;
    if (a.num_tiles == 0) return 0;
    ProfScope prof(p.transposed ? TAG_TILE_BWD : TAG_TILE_FWD, p.kin, p.tag_out, st);
    note_launch(1);
    RGCN_DISPATCH_E(run_etile, p.kp, p.np, a, p.relu_in, p.vec4, p.packed, num_sms, p.ctas_per_sm, st);
}

// out[i] = act(x[own_lo + i]) . root + bias for every owned row (plain stores: also initialises `out`)
int launch_selfloop_pass(const TilePass& p, int64_t own_lo, int64_t n_own, int R, int num_sms, cudaStream_t st) {
    if (n_own == 0) return 0;
    SelfArgs a{};
    a.x = p.feat + own_lo * p.ldf;
    a.ldx = p.ldf;
    a.kin = p.kin;
    a.rfrag = p.wfrag + (int64_t)R * (p.kp / 8) * (p.np / 8) * 32;
    a.bias = p.bias;
    a.nbias = p.nbias;
    a.out = p.out;
    a.ldo = p.ldo;
    a.nout = p.nout;
    a.n_own = n_own;
    ProfScope prof(TAG_SELF, p.kin, p.tag_out, st);
    note_launch(1);
    RGCN_DISPATCH_E(run_selfloop, p.kp, p.np, a, p.relu_in, p.vec4, p.packed, num_sms, st);
}

// fused padding copy + self loop (see k_selfloop_pad): x = caller's rows (any leading dimension), mirror
// [n_own, ldm] receives the zero-padded copy; the B fragments must have been prepared with the vector K order
bool selfloop_pad_ok(int kp, int np) { return kp == 64 && np <= 16; }

int launch_selfloop_pad(const TilePass& p, const float* x_raw, int64_t ld_raw, float* mirror, int64_t ldm, int64_t n_own,
                        int R, int num_sms, cudaStream_t st) {
    if (n_own == 0) return 0;
    if (!selfloop_pad_ok(p.kp, p.np)) return fail(RGCN_ERR_UNSUPPORTED, "fused pad + self loop: unsupported shape");
    SelfArgs a{};
    a.x = x_raw;
    a.ldx = ld_raw;
    a.kin = p.kin;
    a.rfrag = p.wfrag + (int64_t)R * (p.kp / 8) * (p.np / 8) * 32;
    a.bias = p.bias;
    a.nbias = p.nbias;
    a.out = p.out;
    a.ldo = p.ldo;
    a.nout = p.nout;
    a.n_own = n_own;
    a.mirror = mirror;
    a.ldm = ldm;
    ProfScope prof(TAG_SELF, p.kin, p.tag_out, st);
    note_launch(1);
    const int64_t tiles = (n_own + 15) / 16;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((tiles + EW - 1) / EW, (int64_t)num_sms * 2));
    if (p.np == 16) {
        if (p.relu_in) k_selfloop_pad<2, true><<<grid, EW * 32, 0, st>>>(a);
        else k_selfloop_pad<2, false><<<grid, EW * 32, 0, st>>>(a);
    } else {   // np == 8 does not exist (pad_dim starts at 16)
        return fail(RGCN_ERR_UNSUPPORTED, "fused pad + self loop: unsupported output width");
    }
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_ewgrad_pass(const WGradPass& p, int num_sms, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_tiles == 0) return 0;
    ETileArgs a = base_args(b);
    a.feat = p.feat;
    a.ldf = p.ldf;
    a.kin = p.kin;
    a.aux = p.aux ? p.aux : p.feat;
    a.n_rows = p.n_nodes;
    a.self_rel = p.self_rel;
    a.gout = p.gout;
    a.ldg = p.ldg;
    a.nout = p.nout;
    a.gweight = p.gweight;
    a.groot = p.groot;
    a.gbias = p.gbias;
    ProfScope prof(TAG_WGRAD, p.kin, p.nout, st);
    note_launch(1);
    RGCN_DISPATCH_E(run_ewgrad, p.kp, p.np, a, p.relu_in, p.vec4, num_sms, p.ctas_per_sm, st);
}


bool ewgrad_t_ok(const WGradPass& p) {
    // opt-in (RGCN_B200_WGRAD_T=1): measured at AM shape it moves 1.94 GB of DRAM traffic instead of 3.0 GB but,
    // latency-bound at 16 warps per SM, takes 0.65 ms against 0.64 ms for the relation-major form
    static const bool on = [] {
        const char* e = getenv("RGCN_B200_WGRAD_T");
        return e && e[0] == '1';
    }();
    // gathered side exactly 16 padded columns, owner side 32 or 64 with 16-byte addressable rows that hold the
    // whole padded width; 32-bit element offsets on both
    return on && !p.relu_in && p.np == 16 && (p.kp == 32 || p.kp == 64) && p.kp > p.np && p.ldf % 4 == 0 && p.ldf >= p.kp &&
           ((uintptr_t)p.feat & 15) == 0;
}

// p.feat / ldf / kin = the OWNER-side rows (x of the owned nodes), p.gout / ldg / nout = the rows gathered through
// p.brc (the transposed structure: gout of all nodes), p.aux = that structure's chunk rows of gout [num_chunks, 16],
// scratch = (R + 1) * 16 * kp floats
int launch_ewgrad_t_pass(const WGradPass& p, float* scratch, int num_sms, cudaStream_t st) {
    const Brc& b = *p.brc;
    const int R = p.self_rel;
    RGCN_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(R + 1) * 16 * p.kp * 4, st));
    ProfScope prof(TAG_WGRAD, p.kin, p.nout, st);
    if (b.num_tiles > 0) {
        ETileArgs a = base_args(b);
        a.feat = p.gout;
        a.ldf = p.ldg;
        a.kin = p.nout;
        a.aux = p.aux ? p.aux : p.gout;
        a.n_rows = p.n_nodes;
        a.self_rel = R;
        a.gout = p.feat;
        a.ldg = p.ldf;
        a.nout = p.kin;
        a.gweight = scratch;
        a.out = p.gweight;        // only as the "relation tiles wanted" flag
        a.groot = p.groot;
        a.gbias = p.gbias;
        note_launch(1);
        auto launch = [&](auto kern) -> int {
            int per_sm = 1;
            const int smem = MetaStage<0>::CTA_BYTES;
            if (smem > 48 * 1024) RGCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            RGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EW * 32, smem));
            per_sm = std::max(per_sm, 1);
            if (p.ctas_per_sm > 0) per_sm = std::min(per_sm, p.ctas_per_sm);
            const int64_t units = ((int64_t)a.num_tiles + UT - 1) / UT;
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((units + EW - 1) / EW, (int64_t)num_sms * per_sm));
            static const int upw = [] {
                const char* e = getenv("RGCN_B200_WGRAD_T_UPW");
                return e ? std::min(64, std::max(1, atoi(e))) : 1;
            }();
            kern<<<grid, EW * 32, smem, st>>>(a, upw);
            RGCN_CUDA(cudaGetLastError());
            return 0;
        };
        int rc = p.kp == 64 ? launch(k_ewgrad_t<8>) : launch(k_ewgrad_t<4>);
        if (rc) return rc;
    }
    note_launch(1);
    const int64_t total = (int64_t)(R + 1) * p.kin * p.nout;
    k_ewgrad_t_finish<<<(int)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, st>>>(scratch, p.kp, R, p.kin, p.nout, p.gweight,
                                                                                       p.groot);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rgcn
