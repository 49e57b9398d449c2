// Library-level entry points: version, per-thread error message, device check.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace tpq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return TPQ_ERR_CUDA;
}

}  // namespace tpq

extern "C" int tpq_version(void) { return 1; }

extern "C" const char* tpq_last_error(void) { return tpq::g_err; }

extern "C" int tpq_device_supported(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { cudaGetLastError(); return 0; }
  return p.major == 10 && p.minor == 0;   // built for sm_100a only
}
