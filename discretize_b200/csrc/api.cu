// Error plumbing and ABI version of libdzb.so.
#include <cstdio>
#include <cstring>
#include "tm_common.cuh"

namespace dzb {
static thread_local char g_err[512] = "";

void set_error(const char *msg) {
  std::strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[512];
    std::snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    set_error(buf);
    return -(int)e;
  }
  return 0;
}
}  // namespace dzb

extern "C" int dzb_version(void) { return DZB_ABI_VERSION; }
extern "C" const char *dzb_last_error(void) { return dzb::g_err; }

// experiment knob (not part of the stable ABI): the L2 fetch granularity hint (32 / 64 / 128 bytes).
// Random 8-byte gathers (interpolation, OcTree SpMV) pull a whole fetch unit per miss.
extern "C" int dzb_tune_l2_fetch(int bytes) {
  if (bytes > 0) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
    if (e != cudaSuccess) return -(int)e;
  }
  size_t v = 0;
  cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
  return (int)v;
}
