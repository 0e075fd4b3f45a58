// Error text, version and launch accounting for libmmtpsm_b200.
#include <atomic>
#include <string.h>

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
}  // namespace mmt

extern "C" {
const char* mmt_last_error(void) { return mmt::g_err; }
const char* mmt_version(void) { return "mmtpsm_b200 0.1 (sm_100a)"; }
int64_t mmt_launch_count(void) { return mmt::g_launches.load(); }
void mmt_reset_launch_count(void) { mmt::g_launches.store(0); }
}
