// Debug tool: ONE tcgen05.mma (kind::tf32, cta_group::1) on caller-supplied shared-memory images and descriptor fields, with
// the full accumulator returned.  Used by tests/gpu_probe_umma.py to read off which shared-memory word the tensor core
// fetches for operand element (row, k) under a given layout type / LBO / SBO -- the layouts of d0_layer.cu and of the halo
// engine were pinned this way, not from documentation.
#include <cstdint>

#include "common.h"
#include "ptx.cuh"

namespace vp {

struct UmmaProbeArgs {
  const uint32_t* a_img;
  const uint32_t* b_img;
  int a_words, b_words;
  uint32_t a_start, b_start;                 // byte offsets of the descriptor start addresses inside the images
  uint32_t a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout;
  uint32_t idesc;
  int n;
  float* out;                                // [128][n]
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const UmmaProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  uint32_t* sa = reinterpret_cast<uint32_t*>(smem);
  const int a_region = ((p.a_words * 4 + 1023) / 1024) * 1024;
  uint32_t* sb = reinterpret_cast<uint32_t*>(smem + a_region);
  for (int i = threadIdx.x; i < p.a_words; i += blockDim.x) sa[i] = p.a_img[i];
  for (int i = threadIdx.x; i < p.b_words; i += blockDim.x) sb[i] = p.b_img[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint64_t ad = make_smem_desc(smem_u32(sa) + p.a_start, p.a_lbo, p.a_sbo, 0, p.a_layout);
    const uint64_t bd = make_smem_desc(smem_u32(sb) + p.b_start, p.b_lbo, p.b_sbo, 0, p.b_layout);
    umma_tf32(tmem_base, ad, bd, p.idesc, 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int lane = threadIdx.x & 31;
  for (int cc = 0; cc < p.n; cc += 16) {
    float v[16];
    __syncwarp();
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + cc, v);
    for (int j = 0; j < 16 && cc + j < p.n; ++j) p.out[(warp * 32 + lane) * p.n + cc + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace vp

using namespace vp;

extern "C" int vp_debug_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, unsigned a_start, unsigned a_lbo,
                                   unsigned a_sbo, unsigned a_layout, int a_mn_major, unsigned b_start, unsigned b_lbo, unsigned b_sbo,
                                   unsigned b_layout, int b_mn_major, int n, float* out, vp_stream_t stream) {
  if (!a_img || !b_img || !out) return set_error("vp_debug_umma_probe: null pointer");
  if (n < 8 || n > 256 || n % 8) return set_error("vp_debug_umma_probe: n must be a multiple of 8 in 8..256");
  if (a_bytes % 4 || b_bytes % 4 || a_bytes + b_bytes > 200 * 1024) return set_error("vp_debug_umma_probe: bad image sizes");
  UmmaProbeArgs p;
  p.a_img = static_cast<const uint32_t*>(a_img); p.b_img = static_cast<const uint32_t*>(b_img);
  p.a_words = a_bytes / 4; p.b_words = b_bytes / 4;
  p.a_start = a_start; p.b_start = b_start;
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.a_layout = a_layout; p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_layout = b_layout;
  p.idesc = make_idesc_tf32(128, n, a_mn_major, b_mn_major);
  p.n = n; p.out = out;
  const size_t smem = static_cast<size_t>((a_bytes + 1023) / 1024 + (b_bytes + 1023) / 1024 + 1) * 1024;
  if (cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return set_error("cudaFuncSetAttribute(umma_probe_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
  umma_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(p);
  return check_launch("umma_probe_kernel");
}
