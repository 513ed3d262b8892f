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
// slab-mapped plane kernels (planes.cu): return 0 = launched, 1 = shape outside the fast path (caller falls back), -1 = error
int slab_inorm_act(const float* x, int xs, float* y, int ys, int n, int P, int C, const float* gamma, const float* beta, float eps, int act,
                   float alpha, float* stats, vp_stream_t stream);
int slab_inorm_act_bwd(const float* x, int xs, const float* const* dy, const int* dy_cs, int num_dy, float* dx, int dxs, int n, int P, int C,
                       const float* gamma, const float* beta, const float* stats, int act, float alpha, float* dgamma, float* dbeta,
                       vp_stream_t stream);
int slab_gates_fwd(const float* pre, int n, int P, int F, const float* c_prev, const float* g1, const float* b1, const float* g2,
                   const float* b2, float forget_bias, float eps, float* c_new, float* const* h_dst, const int* h_cs, int num_h, float* stats1,
                   float* stats2, vp_stream_t stream);
int slab_gates_bwd(const float* pre, int n, int P, int F, const float* c_prev, const float* g1, const float* b1, const float* g2,
                   const float* b2, const float* stats1, const float* stats2, float forget_bias, const float* const* dh, const int* dh_cs,
                   int num_dh, const float* dc_next, float* dpre, float* dc_prev, float* dg1, float* db1, float* dg2, float* db2,
                   vp_stream_t stream);
inline int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }
}  // namespace vp
