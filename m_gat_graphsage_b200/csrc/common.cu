// Library bookkeeping: version, thread-local error string, launch counter, SM count.
#include "common.cuh"

#include <cstdarg>
#include <mutex>

namespace mgs {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;  // immutable after first query; one GPU per process
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;  // B200
  }
  return cached;
}

}  // namespace mgs

extern "C" {
int mgs_version(void) { return MGS_VERSION; }
const char* mgs_last_error_string(void) { return mgs::t_error; }
uint64_t mgs_launch_count(void) { return mgs::g_launch_count.load(std::memory_order_relaxed); }
}
