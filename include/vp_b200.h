/*
 * vp_b200.h -- C ABI of libvp_b200.so: the B200 (sm_100a) kernels behind the SAVP training path.
 *
 * The reference (alexlee-gk/video_prediction) has no FFI: its seam is the Python model registry
 * (video_prediction/models/__init__.py:12-25) whose classes call tf.nn.* primitives.  Each entry
 * point below replaces the TensorFlow primitive(s) named in its comment (file:line in the
 * reference) and is called only by the host layer in video_prediction_b200/, which keeps the
 * reference's registry / hparams / build_graph surface.  See INTEGRATION.md.
 *
 * Conventions
 *   - all tensors fp32, device pointers, NHWC / NDHWC ("channels last"), densely packed except
 *     for the channel stride (so a kernel can write into a slice of a wider concat buffer);
 *   - the caller allocates every buffer; the library never frees or retains a pointer;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), no hidden syncs;
 *   - return value 0 on success, <0 on error; vp_last_error() gives the message (thread-local).
 */
#ifndef VP_B200_H_
#define VP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vp_stream_t; /* cudaStream_t */

/* NDHWC view of a channel slice of an fp32 buffer.  2-D tensors use d = 1.
 * element (n,z,y,x,ch) lives at ptr[(((n*d + z)*h + y)*w + x)*cstride + ch], ch < c.
 * ptr must be 16-byte aligned and cstride a multiple of 4. */
typedef struct vp_tensor {
  float* ptr;
  int32_t n, d, h, w;
  int32_t c;       /* channels visible through this view            */
  int32_t cstride; /* channels of the underlying buffer (>= c)      */
} vp_tensor;

/* Convolution geometry.  transposed == 0:  out[o] = sum_r in[s*o + r - pad] * W[r]
 *                        transposed == 1:  out[i] = sum_{o,r : s*o + r - pad = i} in[o] * W[r]
 * Zero padding outside the input; W[r] is the "effective tap matrix" [cin][cout]. */
typedef struct vp_conv_geom {
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t pd, ph, pw; /* pad before */
  int32_t transposed;
} vp_conv_geom;

enum { VP_ACT_NONE = 0, VP_ACT_RELU = 1, VP_ACT_LRELU = 2, VP_ACT_SIGMOID = 3, VP_ACT_TANH = 4 };
enum { VP_WKIND_PLAIN = 0, VP_WKIND_POOLED = 1, VP_WKIND_UPSAMPLED = 2 };
enum { VP_WLAYOUT_FWD = 0, VP_WLAYOUT_DGRAD = 1 };

const char* vp_last_error(void);
int vp_version(void);

/* ---- tensor-core implicit-GEMM convolution (tcgen05 / TMEM / TMA) ------------------------------
 * Replaces tf.nn.conv2d (rnn_ops.py:121, ops.py:528), tf.nn.conv2d_transpose (ops.py:584),
 * tf.nn.conv3d (ops.py:773) and their gradients.
 * wpacked: [kd*kh*kw][n_pad][kc*32] floats (see vp_pack_weights); GEMM N = n_pad (multiple of 16),
 * GEMM K = kc 32-channel chunks of in->c per tap.  out->c columns are stored.
 * split_k > 1: partial sums are atomically added into `out` (caller zero-fills; act must be NONE;
 * bias is added by split 0). */
int vp_conv_igemm(const vp_tensor* in, const vp_conv_geom* g, const float* wpacked, int n_pad, int kc,
                  const vp_tensor* out, const float* bias, int act, float alpha, int split_k, vp_stream_t stream);

/* Weight gradient of the same convolution: dwpacked[tap][co][ci] += sum_o  dy.. * x..
 * (layout VP_WLAYOUT_FWD: rows = channels of dy (n_pad rows), cols = kc*32 channels of x).
 * Always accumulates atomically: caller zero-fills before the first call of a step. */
int vp_conv_wgrad(const vp_tensor* x, const vp_tensor* dy, const vp_conv_geom* g, float* dwpacked, int n_pad,
                  int kc, int split_k, vp_stream_t stream);

/* Weight packing: reference-layout kernel w [kd][kh][kw][ci_ref][co] (HWIO / DHWIO) ->
 * effective tap matrices in tensor-core layout, TF32-rounded.
 *  kind PLAIN:     taps = kd*kh*kw, Keff = w
 *  kind POOLED:    2-D only, taps = (kh+1)*(kw+1): avg-pool(2x2,s1,FULL) of the kernel (ops.py:838-842)
 *  kind UPSAMPLED: 2-D only, taps = (kh+3)*(kw+3): bilinear (x) kernel (ops.py:698-704)
 *  layout FWD:   wpacked[tap][co (n_pad)][ci_int (kc*32)]
 *  layout DGRAD: wpacked[tap][ci_int (n_pad)][co (kc*32)]
 * cmap[ci_int] = reference input channel feeding internal channel ci_int, or -1 (zero column);
 * cmap == NULL means identity.  inv_scale: optional device scalar; weights are divided by it
 * (spectral norm sigma, ops.py:1044). */
int vp_pack_weights(const float* w, int kd, int kh, int kw, int ci_ref, int co, int kind, int layout,
                    const int32_t* cmap, int ci_int, const float* inv_scale, float* wpacked, int n_pad, int kc,
                    vp_stream_t stream);
/* Adjoint of vp_pack_weights(layout FWD): dw[...] += L^T(dwpacked) (no scale applied). */
int vp_unpack_wgrad(const float* dwpacked, int kd, int kh, int kw, int ci_ref, int co, int kind,
                    const int32_t* cmap, int ci_int, float* dw, int n_pad, int kc, vp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VP_B200_H_ */
