// profile.cu — launch counting and optional per-pass CUDA-event timing (used by bench.py to time
// the dominant kernel live, on the stream it is launched on, inside the timed region).
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace rgcn {
namespace {
struct Rec {
    int tag, d0, d1;
    cudaEvent_t a, b;
};
std::mutex g_mu;
std::vector<Rec> g_recs;
// read on the autograd thread as well as the caller's: atomics
std::atomic<bool> g_enabled{false};
std::atomic<int64_t> g_launches{0};
}  // namespace

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {
const char* const kTagNames[] = {"pass", "wprep", "chunk_prepass", "tile_fwd", "tile_dx", "wgrad", "copy_cols", "relu_mask",
                                 "simple", "map_gather", "self_loop", "nvl_store", "nvl_reduce", "gemm"};   // TAG_* order
bool nvtx_on() {
    static const bool on = [] {
        const char* e = getenv("RGCN_B200_NVTX");
        return e && e[0] == '1';
    }();
    return on;
}
}  // namespace

ProfScope::ProfScope(int tag, int d0, int d1, cudaStream_t st) : st_(st), idx_(-1), nvtx_(false) {
    if (nvtx_on()) {
        char name[64];
        snprintf(name, sizeof(name), "rgcn:%s_%dx%d", tag >= 0 && tag < (int)(sizeof(kTagNames) / sizeof(kTagNames[0])) ? kTagNames[tag] : "pass",
                 d0, d1);
        nvtxRangePushA(name);
        nvtx_ = true;
    }
    if (!g_enabled.load(std::memory_order_relaxed)) return;
    Rec r{tag, d0, d1, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_mu);
    g_recs.push_back(r);
    idx_ = (int)g_recs.size() - 1;
}
ProfScope::~ProfScope() {
    if (nvtx_) nvtxRangePop();
    if (idx_ < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEventRecord(g_recs[idx_].b, st_);
}
}  // namespace rgcn

extern "C" int64_t rgcn_kernel_launch_count(void) { return rgcn::g_launches.load(); }

extern "C" int rgcn_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(rgcn::g_mu);
    rgcn::g_enabled.store(on != 0);
    return 0;
}

extern "C" int rgcn_profile_collect(int32_t* tags, int32_t* dims, float* ms, int32_t max_records, int32_t* n_out) {
    using namespace rgcn;
    if (!n_out) return fail(RGCN_ERR_INVALID_ARG, "rgcn_profile_collect: n_out is null");
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    for (auto& r : g_recs) {
        float t = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&t, r.a, r.b);
        if (n < max_records && tags && dims && ms) {
            tags[n] = r.tag;
            dims[2 * n] = r.d0;
            dims[2 * n + 1] = r.d1;
            ms[n] = t;
            ++n;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_recs.clear();
    *n_out = n;
    return 0;
}
