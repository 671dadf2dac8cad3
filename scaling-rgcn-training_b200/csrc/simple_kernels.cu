// simple_kernels.cu — generic scalar-FMA CUDA kernels over the same BRC structures, for shapes
// the tensor-pipe tile kernels do not cover (fin or fout > 64) and as an on-GPU cross-check
// (RGCN_F_FORCE_SIMPLE).  Same math as layer_kernels.cu, one warp per segment, fp32 FMA,
// scalar atomics.  Not a CPU fallback: these run on the device or not at all.
#include "common.cuh"

namespace rgcn {
namespace {

constexpr int SW = 4;   // warps per block

struct SimpleArgs {
    const int32_t* raw_idx; const float* raw_w; const int32_t* seg_ptr0; const int32_t* seg_own; const int32_t* seg_rel;
    int num_seg;
    const float* feat; int64_t ldf; int kin;
    const float* weight; const float* root; const float* bias; int transpose; int w_rows, w_cols;
    const float* gout; int64_t ldg;
    float* out; int64_t ldo; int nout;
    float* gweight; float* groot; float* gbias;
    int self_rel; int relu_in;
};

// h[kin] in dynamic smem per warp
__device__ __forceinline__ void seg_mean(const SimpleArgs& a, int s, float* h, int lane) {
    for (int c = lane; c < a.kin; c += 32) h[c] = 0.f;
    __syncwarp();
    for (int k = a.seg_ptr0[s]; k < a.seg_ptr0[s + 1]; ++k) {
        const float* rp = a.feat + (int64_t)a.raw_idx[k] * a.ldf;
        const float w = a.raw_w[k];
        for (int c = lane; c < a.kin; c += 32) {
            float v = rp[c];
            if (a.relu_in) v = fmaxf(v, 0.f);
            h[c] = fmaf(w, v, h[c]);
        }
    }
    __syncwarp();
}

__global__ void k_simple_pass(const SimpleArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* h = smem + warp * a.kin;
    const int s = blockIdx.x * SW + warp;
    if (s >= a.num_seg) return;
    seg_mean(a, s, h, lane);
    const int rel = a.seg_rel[s];
    const float* W = rel == a.self_rel ? a.root : a.weight + (int64_t)rel * a.w_rows * a.w_cols;
    float* op = a.out + (int64_t)a.seg_own[s] * a.ldo;
    for (int o = lane; o < a.nout; o += 32) {
        float acc = 0.f;
        if (W) {
            if (!a.transpose) for (int c = 0; c < a.kin; ++c) acc = fmaf(h[c], W[(int64_t)c * a.w_cols + o], acc);
            else for (int c = 0; c < a.kin; ++c) acc = fmaf(h[c], W[(int64_t)o * a.w_cols + c], acc);
        }
        if (rel == a.self_rel && a.bias) acc += a.bias[o];
        atomicAdd(op + o, acc);
    }
}

__global__ void k_simple_wgrad(const SimpleArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* h = smem + warp * a.kin;
    const int s = blockIdx.x * SW + warp;
    if (s >= a.num_seg) return;
    const int rel = a.seg_rel[s];
    float* dst = rel == a.self_rel ? a.groot : (a.gweight ? a.gweight + (int64_t)rel * a.kin * a.nout : nullptr);
    const float* gp = a.gout + (int64_t)a.seg_own[s] * a.ldg;
    if (rel == a.self_rel && a.gbias)
        for (int o = lane; o < a.nout; o += 32) atomicAdd(a.gbias + o, gp[o]);
    if (!dst) return;
    seg_mean(a, s, h, lane);
    const int total = a.kin * a.nout;
    for (int i = lane; i < total; i += 32) {
        const int c = i / a.nout, o = i % a.nout;
        atomicAdd(dst + i, h[c] * gp[o]);
    }
}

SimpleArgs base_args(const Brc& b) {
    SimpleArgs a{};
    a.raw_idx = b.raw_idx; a.raw_w = b.raw_w; a.seg_ptr0 = b.seg_ptr0; a.seg_own = b.seg_own; a.seg_rel = b.seg_rel;
    a.num_seg = b.num_seg;
    return a;
}

}  // namespace

int launch_simple_pass(const SimplePass& p, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_seg == 0) return 0;
    SimpleArgs a = base_args(b);
    a.feat = p.feat; a.ldf = p.ldf; a.kin = p.kin;
    a.weight = p.weight; a.root = p.root; a.bias = p.bias; a.transpose = p.transpose ? 1 : 0;
    a.w_rows = p.w_rows; a.w_cols = p.w_cols;
    a.out = p.out; a.ldo = p.ldo; a.nout = p.nout; a.self_rel = p.self_rel; a.relu_in = p.relu_in ? 1 : 0;
    const size_t smem = (size_t)SW * p.kin * sizeof(float);
    if (smem > 48 * 1024) return fail(RGCN_ERR_UNSUPPORTED, "simple pass: feature width too large");
    ProfScope prof(TAG_SIMPLE, p.kin, p.nout, st);
    note_launch(1);
    k_simple_pass<<<(b.num_seg + SW - 1) / SW, SW * 32, smem, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

int launch_simple_wgrad(const SimpleWGrad& p, cudaStream_t st) {
    const Brc& b = *p.brc;
    if (b.num_seg == 0) return 0;
    SimpleArgs a = base_args(b);
    a.feat = p.feat; a.ldf = p.ldf; a.kin = p.kin;
    a.gout = p.gout; a.ldg = p.ldg; a.nout = p.nout;
    a.gweight = p.gweight; a.groot = p.groot; a.gbias = p.gbias;
    a.self_rel = p.self_rel; a.relu_in = p.relu_in ? 1 : 0;
    const size_t smem = (size_t)SW * p.kin * sizeof(float);
    if (smem > 48 * 1024) return fail(RGCN_ERR_UNSUPPORTED, "simple wgrad: feature width too large");
    ProfScope prof(TAG_SIMPLE, p.kin, p.nout, st);
    note_launch(1);
    k_simple_wgrad<<<(b.num_seg + SW - 1) / SW, SW * 32, smem, st>>>(a);
    RGCN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rgcn
