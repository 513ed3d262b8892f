// Producer/consumer ring without data movement: how fast can ONE elected thread issue 4 MMAs per stage when it has to
// wait on a full barrier per stage and commit to an empty barrier, with a second elected thread playing the TMA producer
// (wait empty -> arrive full)?  Variants isolate the cost of each piece of the conv engine's issue loop.
#include <cstdio>
#include <vector>
#include "ptx.cuh"
#ifndef RANDOM_DATA
#define RANDOM_DATA 0
#endif
using namespace vp;

__device__ __forceinline__ uint32_t try_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void wait_addr(uint32_t addr, uint32_t parity) { while (!try_wait_addr(addr, parity)) {} }
__device__ __forceinline__ void commit_addr(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void arrive_addr(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void expect_tx_addr(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}

// MODE 0: runtime stage count, generic loop (like the engine).  MODE 1: stages unrolled at compile time.
template <int N, int STAGES, int MODE, int KSTEPS>
__global__ void __launch_bounds__(192, 1) ring_kernel(int iters, int rt_stages, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[8], empty_bar[8], done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  constexpr uint32_t stage_bytes = KSTEPS * (16384u + N * 128u) / 4;
  for (uint32_t i = threadIdx.x; i < STAGES * stage_bytes / 4; i += blockDim.x) { uint32_t hsh = (i + 1) * 2654435761u; hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13; reinterpret_cast<float*>(smem)[i] = RANDOM_DATA ? (static_cast<int>(hsh >> 8) - (1 << 23)) * (1.0f / (1 << 22)) : 1.0f + 0.001f * (i % 97); }
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); } mbar_init(&done, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 2) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t d = tmem_base_smem;
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  if (warp == 0) {
    if (elect_one_sync()) {
      if (MODE == 0) {
        int s = 0, ph = 0;
        for (int it = 0; it < iters * STAGES; ++it) {
          wait_addr(empty0 + 8 * s, ph ^ 1);
          arrive_addr(full0 + 8 * s);
          if (++s == rt_stages) { s = 0; ph ^= 1; }
        }
      } else if (MODE >= 2) {
        int s = 0, ph = 0;
        for (int it = 0; it < iters * STAGES; ++it) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive(&full_bar[s]);
          if (++s == rt_stages) { s = 0; ph ^= 1; }
        }
      } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int s = 0; s < STAGES; ++s) { wait_addr(empty0 + 8 * s, (it & 1) ^ 1); arrive_addr(full0 + 8 * s); }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
      const uint64_t ad0 = make_smem_desc(smem_u32(smem), 16, 1024, 0);
      const long long t0 = clock64();
      if (MODE >= 2) {
        int s = 0, ph = 0;
        uint64_t ad = ad0;
        for (int it = 0; it < iters * STAGES; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t bd = ad + 1024;
          umma_tf32(d, ad, bd, idesc, it > 0 ? 1u : 0u);
          umma_tf32(d, ad + 2, bd + 2, idesc, 1u);
          umma_tf32(d, ad + 4, bd + 4, idesc, 1u);
          umma_tf32(d, ad + 6, bd + 6, idesc, 1u);
          umma_commit(&empty_bar[s]);
          ad += stage_bytes >> 4;
          if (++s == rt_stages) { s = 0; ph ^= 1; ad = ad0; }
        }
      } else if (MODE == 0) {
        int s = 0, ph = 0;
        uint64_t ad = ad0;
        for (int it = 0; it < iters * STAGES; ++it) {
          wait_addr(full0 + 8 * s, ph);
          tc_fence_after();
          const uint64_t bd = ad + 1024;
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) umma_tf32(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
          commit_addr(empty0 + 8 * s);
          ad += stage_bytes >> 4;
          if (++s == rt_stages) { s = 0; ph ^= 1; ad = ad0; }
        }
      } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int s = 0; s < STAGES; ++s) {
            wait_addr(full0 + 8 * s, it & 1);
            tc_fence_after();
            const uint64_t ad = ad0 + s * (stage_bytes >> 4), bd = ad + 1024;
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) umma_tf32(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
            commit_addr(empty0 + 8 * s);
          }
        }
      }
      umma_commit(&done);
      mbar_wait(&done, 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  } else if (MODE == 3) {
    mbar_wait(&done, 0);   // epilogue warps of the engine: spin until the accumulator is complete
  } else if (MODE == 4) {
    wait_addr(smem_u32(&done), 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(d, 256);
}

template <int N, int STAGES, int MODE, int KSTEPS>
void run(const char* name, long long* d_cycles, int ctas = 148) {
  const int iters = 1200 / STAGES * 4 / KSTEPS;
  const size_t smem = STAGES * KSTEPS * (16384 + N * 128) / 4 + 1024;
  cudaFuncSetAttribute(ring_kernel<N, STAGES, MODE, KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  ring_kernel<N, STAGES, MODE, KSTEPS><<<ctas, 192, smem>>>(iters, STAGES, d_cycles);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  ring_kernel<N, STAGES, MODE, KSTEPS><<<ctas, 192, smem>>>(iters, STAGES, d_cycles);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); return; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), d_cycles, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
  const double mmas = static_cast<double>(iters) * STAGES * KSTEPS;
  printf("%-44s cycles/MMA %.1f (ideal %d) | %.1f us | implied clock %.2f GHz\n", name, mx / mmas, N / 2, ms * 1e3, mx / (ms * 1e6));
}

int main() {
  long long* d_cycles; cudaMalloc(&d_cycles, 1024 * sizeof(long long));
  run<128, 3, 2, 4>("N=128 S=3 engine helpers", d_cycles);
  run<128, 4, 2, 4>("N=128 S=4 engine helpers", d_cycles);
  run<128, 3, 3, 4>("N=128 S=3 engine helpers + spinning epilogue", d_cycles);
  run<128, 4, 3, 4>("N=128 S=4 engine helpers + spinning epilogue", d_cycles);
  run<128, 4, 4, 4>("N=128 S=4 engine helpers + plain-spin epilogue", d_cycles);
  run<128, 3, 0, 4>("N=128 S=3 generic loop, 4 MMA/stage", d_cycles);
  run<128, 4, 0, 4>("N=128 S=4 generic loop, 4 MMA/stage", d_cycles);
  run<128, 6, 0, 4>("N=128 S=6 generic loop, 4 MMA/stage", d_cycles);
  run<128, 3, 1, 4>("N=128 S=3 unrolled, 4 MMA/stage", d_cycles);
  run<128, 4, 1, 4>("N=128 S=4 unrolled, 4 MMA/stage", d_cycles);
  run<128, 6, 1, 4>("N=128 S=6 unrolled, 4 MMA/stage", d_cycles);
  run<128, 3, 1, 8>("N=128 S=3 unrolled, 8 MMA/stage", d_cycles);
  run<128, 2, 1, 8>("N=128 S=2 unrolled, 8 MMA/stage", d_cycles);
  run<256, 4, 0, 4>("N=256 S=4 generic loop, 4 MMA/stage", d_cycles);
  run<256, 4, 1, 4>("N=256 S=4 unrolled, 4 MMA/stage", d_cycles);
  run<64, 4, 1, 4>("N=64 S=4 unrolled, 4 MMA/stage", d_cycles);
  run<64, 4, 1, 8>("N=64 S=4 unrolled, 8 MMA/stage", d_cycles);
  return 0;
}
