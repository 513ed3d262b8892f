// Halo-resident implicit GEMM for stride-1 convolutions (the ConvLSTM gate convolutions, rnn_ops.py:121).
//
// The box-mode engine (igemm.cu) re-fetches a 16 KB activation tile for every (tap, 32-channel chunk):
// 25 taps of a 5x5 kernel read the same pixels 25 times from L2, which caps tensor-pipe utilisation near 30 %.
// Here the input lives in HBM as a zero-padded, FLATTENED plane: row q = (n*Hp + y)*P + x with P = W + pad
// columns per line and Hp = H + pad lines per sample; the gap columns/lines are zeros shared by neighbouring
// lines/samples, so a filter tap (dy,dx) is the constant row offset dy*P + dx.  A CTA owns 256 consecutive rows:
// for each 32-channel chunk it TMA-loads ONE halo tile (256 + 4P + 4 rows) and every tap's A operand is the same
// tile addressed through a shifted UMMA descriptor start address (row granularity 128 B inside the 128B-swizzled
// tile).  Only the weights stream per tap, and each weight tile feeds two 128-row MMAs (M = 256 per CTA), so the
// L2 -> SM feed drops from ~128 B/cycle/SM to ~35 B/cycle/SM.
//
// Warp roles (224 threads): w0 weight producer, w1 MMA issuer + TMEM allocator, w2..5 epilogue, w6 halo producer.
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace vp {

constexpr int kFlatMaxTaps = 32;
constexpr int kFlatBStages = 6;

struct alignas(64) FlatArgs {
  CUtensorMap amap;  // 2-D {channels, rows}; box {32, box_rows}
  CUtensorMap bmap;  // packed weights 2-D {kc*32, slots*n_pad}; box {32, bn_tile}
  int32_t q_total, Hp, P, H, W;
  int32_t kc, n_pad, bn_tile, tmem_cols;
  int32_t ntaps, off_min, halo_rows, box_rows, nbox, halo_bytes;
  int32_t splits, desc_mode, b_stages;
  float* out;
  long long so_n, so_h, so_w;
  int32_t out_c, act, accumulate;
  float alpha;
  const float* bias;
  int32_t tap_off[kFlatMaxTaps];
  int32_t tap_wslot[kFlatMaxTaps];
};

__device__ __forceinline__ float flat_act(float v, int act, float alpha) {
  switch (act) {
    case VP_ACT_RELU: return fmaxf(v, 0.f);
    case VP_ACT_LRELU: return fmaxf(alpha * v, v);
    case VP_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case VP_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(224, 1) igemm_flat_kernel(const __grid_constant__ FlatArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t a_full[2], a_empty[2], b_full[kFlatBStages], b_empty[kFlatBStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int n0 = blockIdx.y * a.bn_tile;
  const int c_begin = static_cast<int>(static_cast<long long>(a.kc) * blockIdx.z / a.splits);
  const int c_end = static_cast<int>(static_cast<long long>(a.kc) * (blockIdx.z + 1) / a.splits);
  if (c_end <= c_begin) return;
  uint8_t* halo[2] = {smem, smem + a.halo_bytes};
  uint8_t* bring = smem + 2 * a.halo_bytes;
  const uint32_t b_bytes = static_cast<uint32_t>(a.bn_tile) * 128u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < a.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 6) {  // halo producer: one tile per 32-channel chunk, double buffered
    const uint32_t leader = elect_one_sync();
    for (int c = c_begin; c < c_end; ++c) {
      const int li = c - c_begin, buf = li & 1, ph = (li >> 1) & 1;
      mbar_wait(&a_empty[buf], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&a_full[buf], static_cast<uint32_t>(a.nbox * a.box_rows) * 128u);
        for (int b = 0; b < a.nbox; ++b)
          tma_load_2d(halo[buf] + static_cast<size_t>(b) * a.box_rows * 128, &a.amap, &a_full[buf], c * 32,
                      q0 + a.off_min + b * a.box_rows);
      }
      __syncwarp();
    }
  } else if (warp == 0) {  // weight producer: one [bn_tile x 32] tile per (chunk, tap)
    const uint32_t leader = elect_one_sync();
    int s = 0, ph = 0;
    for (int c = c_begin; c < c_end; ++c)
      for (int t = 0; t < a.ntaps; ++t) {
        mbar_wait(&b_empty[s], ph ^ 1);
        if (leader) {
          mbar_expect_tx(&b_full[s], b_bytes);
          tma_load_2d(bring + s * b_bytes, &a.bmap, &b_full[s], c * 32, a.tap_wslot[t] * a.n_pad + n0);
        }
        __syncwarp();
        if (++s == a.b_stages) { s = 0; ph ^= 1; }
      }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, a.bn_tile, 0, 0);
    const uint32_t leader = elect_one_sync();
    int s = 0, ph = 0;
    for (int c = c_begin; c < c_end; ++c) {
      const int li = c - c_begin, buf = li & 1, aph = (li >> 1) & 1;
      mbar_wait(&a_full[buf], aph);
      const uint32_t halo_addr = smem_u32(halo[buf]);
      for (int t = 0; t < a.ntaps; ++t) {
        mbar_wait(&b_full[s], ph);
        tc_fence_after();
        if (leader) {
          const uint64_t bd0 = make_smem_desc(smem_u32(bring + s * b_bytes), 16, 1024, 0);
          const uint32_t a0 = halo_addr + static_cast<uint32_t>(a.tap_off[t] - a.off_min) * 128u;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t a_addr = a0 + mt * (128u * 128u);
            const uint64_t ad0 = make_smem_desc(a_addr, 16, 1024, a.desc_mode ? ((a_addr >> 7) & 7u) : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tmem_base + mt * a.bn_tile, desc_advance(ad0, k * 32), desc_advance(bd0, k * 32), idesc,
                        (li > 0 || t > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&b_empty[s]);
        }
        __syncwarp();
        if (++s == a.b_stages) { s = 0; ph ^= 1; }
      }
      if (leader) umma_commit(&a_empty[buf]);
      __syncwarp();
    }
    if (leader) umma_commit(&tmem_full_bar);
    __syncwarp();
  } else {
    // epilogue: two passes of 128 rows; thread <-> one flattened position
    const int qd = warp & 3;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int plane = a.Hp * a.P;
    const bool add_bias = a.bias != nullptr && blockIdx.z == 0;
    for (int mt = 0; mt < 2; ++mt) {
      const int q = q0 + mt * 128 + qd * 32 + lane;
      const int n = q / plane, rem = q - n * plane;
      const int y = rem / a.P, x = rem - y * a.P;
      const bool rvalid = q < a.q_total && y < a.H && x < a.W;
      float* orow = a.out + n * a.so_n + y * a.so_h + x * a.so_w;
      for (int cc = 0; cc < a.bn_tile; cc += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + mt * a.bn_tile + cc, v);
        const int col0 = n0 + cc;
        if (rvalid && col0 < a.out_c) {
          if (add_bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (col0 + j < a.out_c) v[j] += __ldg(a.bias + col0 + j);
          }
          if (a.splits > 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (col0 + j < a.out_c) atomicAdd(orow + col0 + j, v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = flat_act(v[j], a.act, a.alpha);
            if (col0 + 16 <= a.out_c) {
              float4* o4 = reinterpret_cast<float4*>(orow + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (a.accumulate) { const float4 e = o4[j]; o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
                o4[j] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < a.out_c) orow[col0 + j] = a.accumulate ? orow[col0 + j] + v[j] : v[j];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 get_encode2() {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  return fn;
}

}  // namespace vp

using namespace vp;

// in: zero-padded flattened planes, dims (n, d=1, h=Hp, w=P); valid region H x W in the top-left corner of each plane;
// the gap (P - W columns, Hp - H lines) must be >= the filter reach on each side and hold zeros.
extern "C" int vp_conv_flat(const vp_tensor* in, int valid_h, int valid_w, const vp_conv_geom* g, const float* wpacked,
                            int n_pad, int kc, const vp_tensor* out, const float* bias, int act, float alpha, int split_k,
                            int accumulate, int desc_mode, vp_stream_t stream) {
  if (!in || !out || !g || !wpacked) return set_error("vp_conv_flat: null argument");
  if (g->kd != 1 || g->sd != 1 || g->sh != 1 || g->sw != 1) return set_error("vp_conv_flat: 2-D stride-1 convolutions only");
  if (g->kh * g->kw > kFlatMaxTaps) return set_error("vp_conv_flat: too many taps");
  if ((reinterpret_cast<uintptr_t>(in->ptr) & 15) || (in->cstride & 3)) return set_error("vp_conv_flat: misaligned input");
  if (n_pad % 16 || out->c > n_pad || kc * 32 < in->c) return set_error("vp_conv_flat: bad n_pad / kc");
  if (out->h != valid_h || out->w != valid_w || out->n != in->n) return set_error("vp_conv_flat: output dims mismatch");
  const int Hp = in->h, P = in->w;
  // reach of the filter: conv reads (y + r - ph, x + s - pw); transposed (stride 1) reads (y + ph - r, x + pw - s)
  const int up = g->transposed ? (g->kh - 1 - g->ph) : g->ph, down = g->transposed ? g->ph : (g->kh - 1 - g->ph);
  const int left = g->transposed ? (g->kw - 1 - g->pw) : g->pw, right = g->transposed ? g->pw : (g->kw - 1 - g->pw);
  if (Hp - valid_h < std::max(up, down) || P - valid_w < std::max(left, right))
    return set_error("vp_conv_flat: padding gap smaller than the filter reach");
  static thread_local FlatArgs A;
  std::memset(&A, 0, sizeof(A));
  A.q_total = in->n * Hp * P; A.Hp = Hp; A.P = P; A.H = valid_h; A.W = valid_w;
  A.kc = kc; A.n_pad = n_pad;
  A.bn_tile = n_pad <= 256 ? n_pad : (n_pad % 128 == 0 ? 128 : 0);
  if (n_pad > 128 && n_pad % 128 == 0) A.bn_tile = 128;
  if (const char* e = getenv("VP_FLAT_BN")) { const int v = atoi(e); if (v >= 16 && v <= 256 && v % 16 == 0 && n_pad % v == 0) A.bn_tile = v; }
  if (A.bn_tile == 0) return set_error("vp_conv_flat: n_pad must be <= 256 or a multiple of 128");
  int tc = 32; while (tc < 2 * A.bn_tile) tc *= 2;
  A.tmem_cols = tc;
  int omin = 1 << 30, omax = -(1 << 30);
  for (int r = 0; r < g->kh; ++r)
    for (int s = 0; s < g->kw; ++s) {
      const int dy = g->transposed ? (g->ph - r) : (r - g->ph), dx = g->transposed ? (g->pw - s) : (s - g->pw);
      const int off = dy * P + dx;
      A.tap_off[A.ntaps] = off; A.tap_wslot[A.ntaps] = r * g->kw + s; ++A.ntaps;
      omin = std::min(omin, off); omax = std::max(omax, off);
    }
  if (getenv("VP_FLAT_ALIGN8")) {   // timing experiment only (wrong results): are row-unaligned descriptor starts slower?
    omin = 1 << 30; omax = -(1 << 30);
    for (int t = 0; t < A.ntaps; ++t) { A.tap_off[t] = (A.tap_off[t] / 8) * 8; omin = std::min(omin, A.tap_off[t]); omax = std::max(omax, A.tap_off[t]); }
    omin = (omin / 8) * 8 - 8;
  }
  A.off_min = omin;
  A.halo_rows = 256 + (omax - omin);
  A.nbox = (A.halo_rows + 255) / 256;
  A.box_rows = ((A.halo_rows + A.nbox - 1) / A.nbox + 7) / 8 * 8;
  A.halo_bytes = (A.nbox * A.box_rows * 128 + 1023) / 1024 * 1024;
  A.splits = std::max(1, std::min(split_k, kc));
  if (A.splits > 1 && act != VP_ACT_NONE) return set_error("vp_conv_flat: split_k needs act NONE");
  A.desc_mode = desc_mode;
  A.out = out->ptr;
  A.so_w = out->cstride; A.so_h = A.so_w * out->w; A.so_n = A.so_h * out->h;
  A.out_c = out->c; A.act = act; A.alpha = alpha; A.accumulate = accumulate; A.bias = bias;
  EncodeTiledFn2 enc = get_encode2();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point not found");
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(in->c), static_cast<cuuint64_t>(A.q_total)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(in->cstride) * 4};
    cuuint32_t box[2] = {32, static_cast<cuuint32_t>(A.box_rows)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&A.amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, in->ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(flat activations) failed with %d", static_cast<int>(r));
  }
  {
    const int slots = g->kh * g->kw;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(kc) * 32, static_cast<cuuint64_t>(slots) * n_pad};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kc) * 128};
    cuuint32_t box[2] = {32, static_cast<cuuint32_t>(A.bn_tile)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&A.bmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wpacked), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(flat weights) failed with %d", static_cast<int>(r));
  }
  const size_t smem_max = 227 * 1024 - 1024;
  const long long b_budget = static_cast<long long>(smem_max) - 1024 - 2LL * A.halo_bytes;
  A.b_stages = static_cast<int>(std::min<long long>(kFlatBStages, b_budget / (A.bn_tile * 128)));
  if (A.b_stages < 2) return set_error("vp_conv_flat: halo of %d rows does not fit shared memory", A.halo_rows);
  const size_t smem = 2 * static_cast<size_t>(A.halo_bytes) + static_cast<size_t>(A.b_stages) * A.bn_tile * 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(igemm_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max)) != cudaSuccess)
      return set_error("cudaFuncSetAttribute(igemm_flat_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  dim3 grid((A.q_total + 255) / 256, n_pad / A.bn_tile, A.splits);
  igemm_flat_kernel<<<grid, 224, smem, as_stream(stream)>>>(A);
  return check_launch("igemm_flat_kernel");
}
