// Error plumbing and version of the C ABI (include/b2d.h).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/b2d.h"
#include "b2d_internal.h"

namespace b2d {
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B2D_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return B2D_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
  }
  return n;
}
}  // namespace b2d

extern "C" int b2d_version(void) { return B2D_VERSION; }
extern "C" const char* b2d_last_error(void) { return b2d::g_err; }

extern "C" int b2d_zero(void* p, int64_t bytes, void* stream) {
  if (!p || bytes < 0) return b2d::set_error(B2D_E_INVALID, "b2d_zero: bad argument");
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return b2d::set_error(B2D_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  return B2D_OK;
}
