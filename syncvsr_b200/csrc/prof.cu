// Launch counter + CUDA-event profiler for the tensor-core kernels (used by bench.py to report the roofline of the
// dominant kernel from the timed region itself: events are recorded on the launching stream around each launch).
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace svsr {

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

struct ProfRec {
  cudaEvent_t e0, e1;
  int kind;
  double flops;
};
static std::vector<ProfRec> g_pool;
static size_t g_used = 0;
static bool g_on = false;
static std::mutex g_mu;

bool prof_enabled() { return g_on; }
void prof_begin(int kind, double flops, cudaStream_t s) {
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_used == g_pool.size()) {
    ProfRec r{};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    g_pool.push_back(r);
  }
  ProfRec& r = g_pool[g_used];
  r.kind = kind, r.flops = flops;
  cudaEventRecord(r.e0, s);
}
void prof_end(cudaStream_t s) {
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_pool[g_used].e1, s);
  ++g_used;
}

}  // namespace svsr

using namespace svsr;
extern "C" {
long long svsr_launch_count(void) { return launch_count(); }
int svsr_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  g_used = 0;
  return SVSR_OK;
}
// Call after synchronising the stream. kind: 0 = igemm_kernel (conv fprop/dgrad + linear), 1 = wgrad_kernel.
int svsr_prof_read(int kind, double* total_ms, double* total_flops, int* launches) {
  std::lock_guard<std::mutex> lk(g_mu);
  double ms = 0, fl = 0;
  int n = 0;
  for (size_t i = 0; i < g_used; ++i) {
    if (g_pool[i].kind != kind) continue;
    float t = 0;
    if (cudaEventElapsedTime(&t, g_pool[i].e0, g_pool[i].e1) != cudaSuccess) {
      set_last_error("prof_read: event not complete (synchronise first)");
      return SVSR_ERR_CUDA;
    }
    ms += t, fl += g_pool[i].flops, ++n;
  }
  *total_ms = ms, *total_flops = fl, *launches = n;
  return SVSR_OK;
}
}
