// Does the tensor pipe ramp up?  Lean single-thread issue loop (kind::tf32 128x128x8, commit every 4 MMAs), time stamps
// every 64 MMAs: prints cycles/MMA per segment for a cold start (after idle) and for back-to-back launches.
#include <cstdio>
#include <vector>
#include <unistd.h>
#include "ptx.cuh"
using namespace vp;

__global__ void __launch_bounds__(128, 1) ramp_kernel(int segments, long long* stamps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4], done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 3 * 32768 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + 0.001f * (i % 97);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tmem_base_smem, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t d = tmem_base_smem;
  if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 128, 0, 0);
      const uint64_t ad0 = make_smem_desc(smem_u32(smem), 16, 1024, 0);
      const uint32_t bar0 = smem_u32(&bar[0]);
      long long* st = stamps + static_cast<long long>(blockIdx.x) * (segments + 1);
      for (int seg = 0; seg < segments; ++seg) {
        st[seg] = clock64();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint64_t ad = ad0 + (j % 3) * 2048, bd = ad + 1024;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * (j % 3)) : "memory");
        }
        // keep the queue shallow so that the stamps follow execution, not issue: wait for this segment's last commit
        umma_commit(&done);
        mbar_wait(&done, seg & 1);
      }
      st[segments] = clock64();
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(d, 128);
}

int main() {
  const int segments = 40, ctas = 148;
  long long* d_st; cudaMalloc(&d_st, sizeof(long long) * ctas * (segments + 1));
  cudaFuncSetAttribute(ramp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  std::vector<long long> h(ctas * (segments + 1));
  for (int rep = 0; rep < 4; ++rep) {
    if (rep == 0 || rep == 2) usleep(200000);   // idle before reps 0 and 2; reps 1 and 3 follow immediately
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    ramp_kernel<<<ctas, 128, 100 * 1024>>>(segments, d_st);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h.data(), d_st, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("rep %d (%s): kernel %.1f us; cycles per MMA by 64-MMA segment (CTA 0):", rep, (rep == 0 || rep == 2) ? "after 200 ms idle" : "back-to-back", ms * 1e3);
    for (int s = 0; s < segments; ++s) printf(" %.0f", (h[s + 1] - h[s]) / 64.0);
    printf("\n");
  }
  return 0;
}
