// Micro-benchmark: issue rate of tcgen05.mma (SS mode, operands in shared memory, no loads at all) for the tile
// shapes the conv engine uses.  Answers: what is the per-SM MMA floor for kind::tf32 M=128 N=128/256 (and kind::f16
// for comparison), with commits every 4 MMAs like the engine's ring?  Build: nvcc -arch=sm_100a -I../../video_prediction_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace vp;

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct P { int n, iters, commit_every, kind, nacc, mn_major; long long* cycles; };

__global__ void __launch_bounds__(128, 2) mma_rate_kernel(P p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8], done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  const int stages = 3;
  const uint32_t stage_bytes = 16384u + p.n * 128u;
  for (uint32_t i = threadIdx.x; i < stages * stage_bytes / 4; i += blockDim.x)
    reinterpret_cast<float*>(smem)[i] = 1.0f + 0.001f * (i % 97);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp == 1) {
    const uint32_t leader = elect_one_sync();
    const uint32_t idesc = p.kind == 0 ? make_idesc_tf32(128, p.n, p.mn_major, p.mn_major)
                                       : ((1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.n >> 3) << 17) | (8u << 24));
    long long t0 = 0;
    if (leader) {
      t0 = clock64();
      int s = 0;
      for (int it = 0; it < p.iters; ++it) {
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint64_t ad0 = p.mn_major ? make_smem_desc(a_addr, 8192, 512, 0, 1) : make_smem_desc(a_addr, 16, 1024, 0);
        const uint64_t bd0 = p.mn_major ? make_smem_desc(a_addr + 16384, 8192, 512, 0, 1) : make_smem_desc(a_addr + 16384, 16, 1024, 0);
        const uint32_t d = tmem_base + (it % p.nacc) * p.n;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t adv = p.mn_major ? k * 1024 : k * 32;
          if (p.kind == 0) umma_tf32(d, desc_advance(ad0, adv), desc_advance(bd0, adv), idesc, 1u);
          else umma_f16(d, desc_advance(ad0, adv), desc_advance(bd0, adv), idesc, 1u);
        }
        if (p.commit_every && (it % p.commit_every) == p.commit_every - 1) umma_commit(&bar[s]);
        if (++s == stages) s = 0;
      }
      umma_commit(&done);
    }
    __syncwarp();
    mbar_wait(&done, 0);
    if (leader) p.cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d_cycles;
  cudaMalloc(&d_cycles, 1024 * sizeof(long long));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { const char* name; int n, kind, commit_every, nacc, ctas, mn; };
  const Cfg cfgs[] = {
      {"tf32 N=128 commit/1 1cta/SM", 128, 0, 1, 1, 148, 0}, {"tf32 N=128 nocommit 1cta/SM", 128, 0, 0, 1, 148, 0},
      {"tf32 N=128 commit/1 2acc", 128, 0, 1, 2, 148, 0},      {"tf32 N=256 commit/1 1cta/SM", 256, 0, 1, 1, 148, 0},
      {"tf32 N=256 nocommit", 256, 0, 0, 1, 148, 0},           {"tf32 N=64 commit/1", 64, 0, 1, 1, 148, 0},
      {"tf32 N=128 commit/1 2cta/SM", 128, 0, 1, 1, 296, 0},   {"tf32 N=128 commit/1 1 CTA only", 128, 0, 1, 1, 1, 0},
      {"bf16 N=128 commit/1 1cta/SM", 128, 1, 1, 1, 148, 0},   {"bf16 N=256 commit/1 1cta/SM", 256, 1, 1, 1, 148, 0},
      {"tf32 N=128 commit/2", 128, 0, 2, 1, 148, 0},           {"tf32 N=128 commit/4", 128, 0, 4, 1, 148, 0},
      {"tf32 N=256 commit/2", 256, 0, 2, 1, 148, 0},           {"tf32 N=256 commit/4", 256, 0, 4, 1, 148, 0},
      {"tf32 N=256 commit/1 2acc", 256, 0, 1, 2, 148, 0},      {"tf32 N=192 commit/1", 192, 0, 1, 1, 148, 0},
      {"tf32 N=224 commit/1", 224, 0, 1, 1, 148, 0},           {"tf32 N=160 commit/1", 160, 0, 1, 1, 148, 0},
      {"tf32 MN-major N=128 commit/1", 128, 0, 1, 1, 148, 1},  {"tf32 MN-major N=96 commit/1", 96, 0, 1, 1, 148, 1},
      {"tf32 MN-major N=128 nocommit", 128, 0, 0, 1, 148, 1},
  };
  for (const Cfg& c : cfgs) {
    P p{c.n, 2000, c.commit_every, c.kind, c.nacc, c.mn, d_cycles};
    const size_t smem = 3 * (16384 + c.n * 128) + 1024 + (c.mn ? 48 * 1024 : 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mma_rate_kernel<<<c.ctas, 128, smem>>>(p);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    mma_rate_kernel<<<c.ctas, 128, smem>>>(p);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(err)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(c.ctas);
    cudaMemcpy(h.data(), d_cycles, c.ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60; for (long long v : h) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
    const double mmas = 2000.0 * 4;
    const double kk = c.kind == 0 ? 8 : 16;
    const double flops = 2.0 * 128 * c.n * kk * mmas * c.ctas;
    printf("%-34s cycles/MMA min %.1f max %.1f | kernel %.1f us -> %.0f TFLOP/s (peak-equiv %.0f%% of %s)\n", c.name, mn / mmas, mx / mmas,
           ms * 1e3, flops / ms / 1e9, 100.0 * flops / ms / 1e9 / (c.kind == 0 ? 1125.0 : 2250.0), c.kind == 0 ? "1125" : "2250");
  }
  return 0;
}
