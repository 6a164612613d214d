// libaa_b200: version / error / device entry points shared by all translation units.
#include "aa_common.cuh"
#include <cstring>

namespace aa {
static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}
}  // namespace aa

extern "C" {
#pragma GCC visibility push(default)

int aa_version(void) { return 100; }  // 0.1.0

const char* aa_last_error(void) { return aa::t_error; }

int64_t aa_launch_count(void) { return aa::g_launch_count.load(); }

int aa_check_device(void) {
  int dev = 0;
  AA_CUDA(cudaGetDevice(&dev));
  int major = 0;
  AA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    aa::set_error("libaa_b200 ships sm_100a code only; device %d has compute capability %d.x", dev, major);
    return AA_ERR_ARCH;
  }
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
