// Optional per-kernel timing with CUDA events on the launching stream (used by bench.py for the roofline line:
// the dominant kernel's average duration is measured inside the timed region, not under a profiler).
#include <vector>
#include "common.cuh"

namespace accbpg {

static bool g_prof_on = false;
struct EvPair { cudaEvent_t a, b; };
static std::vector<EvPair> g_prof_live[P_COUNT];
static std::vector<EvPair> g_prof_pool;

static const char* kProfNames[P_COUNT] = {
    "syrk_tma_kernel", "syrk_reduce_kernel", "chol_inv_step_kernel(all block columns)", "factor+gradient interval (chain with overlapped triangular GEMM)",
    "trmm_persistent_kernel", "grad_finalize_kernel", "burg_simplex_x_kernel", "matvec_kernel", "rmatvec_kernel",
    "fw_pass_kernel", "fw_iteration", "syrk_reduce_push_kernel", "gram_sum_received_kernel(wait+sum)",
    "burg_prepare_push_kernel", "peer_wait_kernel(gg)", "peer_sum_scalars_kernel", "fw_persistent_kernel(batch)"};

ProfScope::ProfScope(int id, cudaStream_t s) : id_(id), s_(s), active_(false) {
    if (!g_prof_on || id < 0 || id >= P_COUNT) return;
    EvPair p;
    if (!g_prof_pool.empty()) { p = g_prof_pool.back(); g_prof_pool.pop_back(); }
    else {
        if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
    }
    cudaEventRecord(p.a, s);
    g_prof_live[id].push_back(p);
    active_ = true;
}
ProfScope::~ProfScope() {
    if (active_) cudaEventRecord(g_prof_live[id_].back().b, s_);
}

}  // namespace accbpg

using namespace accbpg;

extern "C" {

int accbpg_prof_enable(int on) {
    g_prof_on = (on != 0);
    return ACCBPG_OK;
}

int accbpg_prof_count(void) { return P_COUNT; }

const char* accbpg_prof_name(int id) { return (id >= 0 && id < P_COUNT) ? kProfNames[id] : ""; }

int accbpg_prof_read(int id, double* total_ms, int64_t* count) {
    if (id < 0 || id >= P_COUNT || !total_ms || !count) return arg_err("prof_read: id or NULL pointer");
    double tot = 0.0;
    int64_t n = 0;
    for (EvPair& p : g_prof_live[id]) {
        ACCBPG_CUDA(cudaEventSynchronize(p.b));
        float ms = 0.f;
        ACCBPG_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
        tot += ms;
        ++n;
        g_prof_pool.push_back(p);
    }
    g_prof_live[id].clear();
    *total_ms = tot;
    *count = n;
    return ACCBPG_OK;
}

}  // extern "C"
