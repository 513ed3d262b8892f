// Shared host-side helpers for libvp_b200 (error reporting, launch checks).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>

#include "vp_b200.h"

namespace vp {
int set_error(const char* fmt, ...);
void count_launch(int n);
inline int check_launch(const char* what) {
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}
inline cudaStream_t as_stream(vp_stream_t s) { return static_cast<cudaStream_t>(s); }
inline int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }
}  // namespace vp
