// Lean issue loop: what is the hardware floor of tcgen05.mma kind::tf32 (SS mode) per MMA, without scalar overhead?
// One elected thread issues UNROLL MMAs per loop trip with immediate descriptor offsets; optional commit per 4 MMAs.
#include <cstdio>
#include <vector>
#include "ptx.cuh"
using namespace vp;

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}

template <int N, int KIND, int COMMIT, int SAMEADDR>
__global__ void __launch_bounds__(128, 1) floor_kernel(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4], done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  constexpr uint32_t stage_bytes = 16384u + N * 128u;
  for (uint32_t i = threadIdx.x; i < 3 * stage_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + 0.001f * (i % 97);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t d = tmem_base_smem;
  if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = KIND == 0 ? make_idesc_tf32(128, N, 0, 0)
                                           : ((1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24));
      const uint64_t ad0 = make_smem_desc(smem_u32(smem), 16, 1024, 0);
      const uint32_t bar0 = smem_u32(&bar[0]);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const uint64_t ad = ad0 + (SAMEADDR ? 0 : s * (stage_bytes >> 4)), bd = ad + 1024;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (KIND == 0) umma_tf32(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
            else umma_f16(d, ad + 2 * k, bd + 2 * k, idesc);
          }
          if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * s) : "memory");
        }
      }
      umma_commit(&done);
      mbar_wait(&done, 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(d, 256);
}

template <int N, int KIND, int COMMIT, int SAMEADDR>
void run(const char* name, long long* d_cycles, int ctas = 148) {
  const int iters = 700;
  const size_t smem = 3 * (16384 + N * 128) + 1024;
  cudaFuncSetAttribute(floor_kernel<N, KIND, COMMIT, SAMEADDR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  floor_kernel<N, KIND, COMMIT, SAMEADDR><<<ctas, 128, smem>>>(iters, d_cycles);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  floor_kernel<N, KIND, COMMIT, SAMEADDR><<<ctas, 128, smem>>>(iters, d_cycles);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); return; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), d_cycles, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
  const double mmas = iters * 12.0, kk = KIND == 0 ? 8 : 16;
  const double flops = 2.0 * 128 * N * kk * mmas * ctas;
  printf("%-40s cycles/MMA %.1f | %.1f us -> %.0f TFLOP/s | implied SM clock %.2f GHz\n", name, mx / mmas, ms * 1e3, flops / ms / 1e9, mx / (ms * 1e6));
}

int main() {
  long long* d_cycles; cudaMalloc(&d_cycles, 1024 * sizeof(long long));
  run<128, 0, 0, 0>("tf32 N=128 nocommit", d_cycles);
  run<128, 0, 1, 0>("tf32 N=128 commit/4", d_cycles);
  run<128, 0, 1, 1>("tf32 N=128 commit/4 same smem addr", d_cycles);
  run<256, 0, 0, 0>("tf32 N=256 nocommit", d_cycles);
  run<256, 0, 1, 0>("tf32 N=256 commit/4", d_cycles);
  run<64, 0, 1, 0>("tf32 N=64 commit/4", d_cycles);
  run<32, 0, 1, 0>("tf32 N=32 commit/4", d_cycles);
  run<192, 0, 1, 0>("tf32 N=192 commit/4", d_cycles);
  run<128, 1, 1, 0>("bf16 N=128 commit/4", d_cycles);
  run<256, 1, 1, 0>("bf16 N=256 commit/4", d_cycles);
  run<128, 0, 1, 0>("tf32 N=128 commit/4, 1 CTA", d_cycles, 1);
  run<256, 0, 1, 0>("tf32 N=256 commit/4, 1 CTA", d_cycles, 1);
  return 0;
}
