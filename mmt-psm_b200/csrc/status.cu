// Error text, version and launch accounting for libmmtpsm_b200.
#include <atomic>
#include <mutex>
#include <string.h>
#include <vector>

#include "common.cuh"

namespace mmt {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct ProfRec {
  cudaEvent_t t0, t1;
  int family;
  double flops, bytes;
};
static std::atomic<bool> g_prof{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_recs;
static thread_local cudaEvent_t g_open = nullptr;

bool prof_enabled() { return g_prof.load(std::memory_order_relaxed); }
void prof_begin(cudaStream_t stream) {
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  g_open = e;
}
void prof_end(int family, double flops, double bytes, cudaStream_t stream) {
  if (!g_open) return;
  ProfRec r;
  r.t0 = g_open;
  g_open = nullptr;
  if (cudaEventCreate(&r.t1) != cudaSuccess) { cudaEventDestroy(r.t0); return; }
  cudaEventRecord(r.t1, stream);
  r.family = family; r.flops = flops; r.bytes = bytes;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_recs.push_back(r);
}
}  // namespace mmt

extern "C" {
const char* mmt_last_error(void) { return mmt::g_err; }
const char* mmt_version(void) { return "mmtpsm_b200 0.1 (sm_100a)"; }
int64_t mmt_launch_count(void) { return mmt::g_launches.load(); }
void mmt_reset_launch_count(void) { mmt::g_launches.store(0); }
void mmt_add_launch_count(int64_t n) { mmt::g_launches.fetch_add(n, std::memory_order_relaxed); }

void mmt_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(mmt::g_prof_mu);
  for (auto& r : mmt::g_recs) { cudaEventDestroy(r.t0); cudaEventDestroy(r.t1); }
  mmt::g_recs.clear();
  mmt::g_prof.store(on != 0);
}
int mmt_prof_collect(double* ms, double* flops, double* bytes, int64_t* launches, int n_families) {
  MMT_REQUIRE(ms && flops && bytes && launches && n_families >= MMT_PROF_FAMILIES, "prof_collect: bad arguments");
  for (int i = 0; i < n_families; ++i) { ms[i] = flops[i] = bytes[i] = 0; launches[i] = 0; }
  std::lock_guard<std::mutex> lk(mmt::g_prof_mu);
  for (auto& r : mmt::g_recs) {
    MMT_CUDA(cudaEventSynchronize(r.t1));
    float t = 0.f;
    MMT_CUDA(cudaEventElapsedTime(&t, r.t0, r.t1));
    ms[r.family] += t; flops[r.family] += r.flops; bytes[r.family] += r.bytes; launches[r.family] += 1;
    cudaEventDestroy(r.t0); cudaEventDestroy(r.t1);
  }
  mmt::g_recs.clear();
  return MMT_OK;
}
}
