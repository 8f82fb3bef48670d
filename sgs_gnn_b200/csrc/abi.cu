// Error reporting, version and launch accounting for libsgs_b200.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace sgs {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

}  // namespace sgs

extern "C" {

const char* sgs_last_error(void) { return sgs::g_err; }

int32_t sgs_version(void) { return 100; /* 0.1.0 */ }

int64_t sgs_launch_count(void) { return sgs::g_launches.load(std::memory_order_relaxed); }
}
