// Internal helpers shared by the translation units of libb2d.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace b2d {
// Records a thread-local message (returned by b2d_last_error()) and returns `code`.
int set_error(int code, const char* fmt, ...);
// Checks cudaGetLastError() after a launch.
int check_launch(const char* what);
int num_sms();
// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency); nullptr if unavailable.
void* tensor_map_encode_fn();
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: opt in once per (kernel, device).
// `mask` is the caller's static per-kernel bit set (bit = device ordinal; devices >= 64 simply set it every time).
template <typename F>
inline cudaError_t smem_attr_once(F kernel, int bytes, unsigned long long& mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && ((mask >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) mask |= 1ull << dev;
  return e;
}
}  // namespace b2d
