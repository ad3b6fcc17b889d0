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
}  // namespace b2d
