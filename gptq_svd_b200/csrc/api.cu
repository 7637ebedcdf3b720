// Version, error string and device check of libtruncgptq.
#include <stdarg.h>

#include "common.cuh"

namespace tq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libtruncgptq has no CPU fallback)", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return TQ_ERR_CUDA;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return TQ_ERR_CUDA;
  }
  if (major != 10) {
    set_error("device compute capability %d.x is not sm_100 (B200): libtruncgptq is sm_100a only", major);
    return TQ_ERR_UNSUPPORTED;
  }
  return TQ_OK;
}

int num_sms() {
  static thread_local int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  cached = n;
  return n;
}

}  // namespace tq

extern "C" int tq_version(void) { return TQ_VERSION; }
extern "C" const char* tq_last_error(void) { return tq::g_err; }
