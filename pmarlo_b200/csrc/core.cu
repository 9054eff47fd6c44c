// core.cu -- error reporting, version, launch accounting for libpmb200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pmb {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace pmb

extern "C" {
const char* pmb_last_error(void) { return pmb::g_err; }
int pmb_version(void) { return 100; }
int64_t pmb_launch_count(void) { return pmb::g_launches.load(std::memory_order_relaxed); }
}
