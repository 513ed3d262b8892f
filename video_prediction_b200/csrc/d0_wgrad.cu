// Weight gradient of the FIRST video-discriminator layer (conv3d 3x3x3, stride 1, zero pad 1, <= 4 input channels
// stored as float4 voxels, 32 output channels; networks.py:83-84) on the tensor cores.
//
//   gw[tap][ci][co] += sum_v x[v + tap][ci] * dy[v][co]            GEMM-K = voxels, M = (tap, ci), N = co
//
// With 4 input channels the generic wgrad engine wastes 7/8 of every 32-channel operand box, and the CUDA-core kernel it
// replaces (conv3d_c4_wgrad_kernel, discrim.cu) is latency bound at ~0.9 ms per launch.  The only MN-major layout the
// tensor core reads for 32-bit operands is the 128B swizzle with 32-byte atoms (tests/gpu_probe_umma.py: the un-swizzled
// and 16-byte-atom layouts return zeros), whose fetch address is swizzle(start + (k%4)*128 + (k/4)*SBO + (m%32)*4 +
// (m/32)*LBO), keyed on the absolute address, for any 32-byte-aligned start.  A line of float4 voxels IS such an operand:
// the 32 M-elements of a 128-byte chunk are 8 consecutive voxels x 4 channels, and the chunk of K index k + 1 is the next
// 8 voxels.  So with K = "every 8th voxel":
//   * A(m = (g, j, ci), k) = halo_row[g][start + 8k + j][ci]: M-group g (LBO = row pitch) is one of four (dz, dy) halo rows,
//     j a voxel offset 0..7 of which three are the taps dx = -1..1 (the other lanes are never stored), K walks the 64-voxel
//     line segment in steps of 8.  A halo row is 10 voxel-octets from x0 - 8 (TMA box over a (32 floats = octet, W/8, H, D, N)
//     view; borders are zero fill = the convolution's padding);
//   * B(n = co, k) = dy[x0 + 8k + p][co] for phase p = 0..7: TMA delivers the dy tile phase-major ([p][k][co]) through a
//     tensor map whose dimensions are (co, octet, phase, line) with strides (4 B, 1024 B, 128 B, W * 128 B);
//   * phase p moves the A start by p voxels, but starts must be 32-byte aligned (a 16-byte start faults): even phases start
//     at voxel p + 6 and find tap dx in lane group j = dx + 2, odd phases start at p + 7 and find it in j = dx + 1; the two
//     parities accumulate in separate TMEM columns and are both added to gw by the epilogue.
// Per 64 voxels: 10 TMA requests (9 rows + dy) issued by 10 lanes of one warp instruction, 24 MMAs of M = 128, N = 32
// (8 phases x 3 groups of rows), accumulating in 2 x 96 TMEM columns over the CTA's whole voxel range.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "tensormap.h"

namespace vp {

constexpr int kD0Pix = 64;                                   // voxels (GEMM-K x phases) per stage
constexpr int kD0RowOct = 10;                                // halo row box: voxels x0 - 8 .. x0 + 71
constexpr int kD0RowBytes = kD0RowOct * 128;                 // = row pitch
constexpr int kD0DyBytes = kD0Pix * 128;
constexpr int kD0StageBytes = 20 * 1024;                     // dy tile (8 KB, 1024-aligned) + 9 rows (11520 B)
constexpr int kD0Stages = 8;
constexpr int kD0Tail = 8 * 1024;                            // the last row group's unused M-groups read past the last stage

struct D0WgradArgs {
  CUtensorMap xmap, dymap;
  float* gw;
  int tiles_w, H, D, N, CI;
  int total;                                                 // line segments = N * D * H * tiles_w
};

__global__ void __launch_bounds__(192, 1) d0_wgrad_kernel(const __grid_constant__ D0WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kD0Stages], empty_bar[kD0Stages], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it0 = static_cast<int>(static_cast<long long>(a.total) * blockIdx.x / gridDim.x);
  const int it1 = static_cast<int>(static_cast<long long>(a.total) * (blockIdx.x + 1) / gridDim.x);
  if (it1 <= it0) return;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kD0Stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t smem0 = smem_u32(smem);

  if (warp == 0) {
    // lanes 0..8: the nine halo rows, lane 9: the dy tile
    int mt = it0;
    int tw = mt % a.tiles_w; mt /= a.tiles_w;
    int line = mt;                                             // (n * D + z) * H + y
    int y = mt % a.H; mt /= a.H;
    int z = mt % a.D;
    int n = mt / a.D;
    const int rz = lane / 3 - 1, ry = lane % 3 - 1;
    uint32_t stage = 0, ph = 0;
#pragma unroll 1
    for (int it = it0; it < it1; ++it) {
      mbar_wait(&empty_bar[stage], ph ^ 1);
      const uint32_t dst = smem0 + stage * kD0StageBytes;
      const uint32_t fb = smem_u32(&full_bar[stage]);
      if (lane == 0) mbar_expect_tx(&full_bar[stage], 9 * kD0RowBytes + kD0DyBytes);
      __syncwarp();
      if (lane < 9) {
        tma_load_5d_addr(dst + kD0DyBytes + lane * kD0RowBytes, &a.xmap, fb, 0, tw * (kD0Pix / 8) - 1, y + ry, z + rz, n);
      } else if (lane == 9) {
        tma_load_4d_addr(dst, &a.dymap, fb, 0, tw * (kD0Pix / 8), 0, line);
      }
      if (++tw == a.tiles_w) {
        tw = 0; ++line;
        if (++y == a.H) { y = 0; if (++z == a.D) { z = 0; ++n; } }
      }
      if (++stage == kD0Stages) { stage = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_tf32(128, 32, 1, 1);
      const uint64_t a_base = make_smem_desc(smem0 + kD0DyBytes, kD0RowBytes, 512, 0, 1);     // LBO: next halo row; SBO: 4 octets on
      const uint64_t b_base = make_smem_desc(smem0, kD0DyBytes, 512, 0, 1);                    // one 32-channel group: LBO unused
      uint32_t stage = 0, ph = 0, first = 1;
#pragma unroll 1
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&full_bar[stage], ph);
        tc_fence_after();
        const uint32_t adv = stage * (kD0StageBytes >> 4);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          // row voxel index of (k, p, dx) is 8k + p + dx + 8; the start must be even: p + 6 (even p) or p + 7 (odd p)
          const uint32_t a_off = static_cast<uint32_t>((p & 1) ? p + 7 : p + 6);     // in 16-byte units
          const uint32_t acc = (first && p < 2) ? 0u : 1u;
#pragma unroll
          for (int rg = 0; rg < 3; ++rg)
            umma_tf32(tmem_base + (p & 1) * 96 + rg * 32, a_base + adv + a_off + rg * (4 * kD0RowBytes >> 4), b_base + adv + p * (1024 >> 4),
                      idesc, acc);
        }
        umma_commit(&empty_bar[stage]);
        first = 0;
        if (++stage == kD0Stages) { stage = 0; ph ^= 1; }
      }
      umma_commit(&tmem_full_bar);
    }
    __syncwarp();
  } else {
    // TMEM lane = 32 * g + 4 * j + ci: halo row 4 * rg + g, voxel offset j, input channel ci; the tap column is j - 1 in the
    // even-phase accumulators (columns 0..95) and j in the odd-phase ones (columns 96..191)
    const int g = warp & 3;
    const int j = lane >> 2, ci = lane & 3;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int par = 0; par < 2; ++par) {
      const int c = par ? j : j - 1;
#pragma unroll 1
      for (int rg = 0; rg < 3; ++rg) {
        const int r = 4 * rg + g;
        const bool valid = r < 9 && c >= 0 && c < 3 && ci < a.CI;
        float* o = a.gw + (static_cast<long long>(r * 3 + c) * a.CI + ci) * 32;
#pragma unroll 1
        for (int cc = 0; cc < 32; cc += 16) {
          float v[16];
          __syncwarp();
          tmem_ld16(tmem_base + (static_cast<uint32_t>(g * 32) << 16) + par * 96 + rg * 32 + cc, v);
          if (valid) {
#pragma unroll
            for (int q = 0; q < 16; q += 4) red_add_v4(o + cc + q, v[q], v[q + 1], v[q + 2], v[q + 3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

static int encode_map(CUtensorMap* m, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point not found");
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(first discriminator layer) failed with %d", static_cast<int>(r));
  return 0;
}

}  // namespace vp

using namespace vp;

extern "C" int vp_conv3d_c4_wgrad_tc(const float* x, const float* dy, float* gw, int n, int d, int h, int wd, int ci, vp_stream_t stream) {
  if (!x || !dy || !gw) return set_error("vp_conv3d_c4_wgrad_tc: null pointer");
  if (ci < 1 || ci > 4) return set_error("vp_conv3d_c4_wgrad_tc: ci must be 1..4");
  if (wd % kD0Pix) return set_error("vp_conv3d_c4_wgrad_tc: width %d is not a multiple of %d (use vp_conv3d_c4_wgrad)", wd, kD0Pix);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(gw)) & 15)
    return set_error("vp_conv3d_c4_wgrad_tc: pointers must be 16-byte aligned");
  const long long total = static_cast<long long>(n) * d * h * (wd / kD0Pix);
  if (total <= 0 || total > 0x7fffffffLL) return set_error("vp_conv3d_c4_wgrad_tc: bad dims");
  D0WgradArgs A;
  {
    // x: (octet of float4 voxels, w / 8, h, d, n), halo row boxes of 10 octets; dy: (channel, octet, phase, line) so that a 64-voxel box arrives
    // phase-major: [phase][octet][32 channels]
    const cuuint64_t W = wd, H = h, Dd = d;
    cuuint64_t xd[5] = {32, W / 8, H, Dd, static_cast<cuuint64_t>(n)};
    cuuint64_t xs[4] = {128, 16 * W, 16 * W * H, 16 * W * H * Dd};
    cuuint32_t xb[5] = {32, kD0RowOct, 1, 1, 1};
    if (encode_map(&A.xmap, x, 5, xd, xs, xb)) return -1;
    cuuint64_t yd[4] = {32, W / 8, 8, static_cast<cuuint64_t>(n) * Dd * H};
    cuuint64_t ys[3] = {1024, 128, 128 * W};
    cuuint32_t yb[4] = {32, 8, 8, 1};
    if (encode_map(&A.dymap, dy, 4, yd, ys, yb)) return -1;
  }
  A.gw = gw;
  A.tiles_w = wd / kD0Pix; A.H = h; A.D = d; A.N = n; A.CI = ci;
  A.total = static_cast<int>(total);
  const size_t smem = static_cast<size_t>(kD0Stages) * kD0StageBytes + kD0Tail + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(d0_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return set_error("cudaFuncSetAttribute(d0_wgrad_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
    configured = true;
  }
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = static_cast<int>(std::min<long long>(sms, total));
  d0_wgrad_kernel<<<grid, 192, smem, as_stream(stream)>>>(A);
  return check_launch("d0_wgrad_kernel");
}
