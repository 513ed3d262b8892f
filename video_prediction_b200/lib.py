"""ctypes binding of libvp_b200.so (the C ABI declared in include/vp_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load, importing the
product path raises.  PyTorch is used only for device memory and streams."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

_LIB = None


class VpTensor(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('n', C.c_int32), ('d', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
                ('c', C.c_int32), ('cstride', C.c_int32)]


class VpConvGeom(C.Structure):
    _fields_ = [('kd', C.c_int32), ('kh', C.c_int32), ('kw', C.c_int32),
                ('sd', C.c_int32), ('sh', C.c_int32), ('sw', C.c_int32),
                ('pd', C.c_int32), ('ph', C.c_int32), ('pw', C.c_int32),
                ('transposed', C.c_int32)]


ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID, ACT_TANH = range(5)
WKIND_PLAIN, WKIND_POOLED, WKIND_UPSAMPLED = range(3)
WLAYOUT_FWD, WLAYOUT_DGRAD = range(2)


class VpError(RuntimeError):
    pass


def lib():
    """Loads (building first if the sources are newer) libvp_b200.so.  Raises if unavailable."""
    global _LIB
    if _LIB is None:
        path = _build.LIB_PATH
        if not os.path.exists(path):
            path = _build.build()
        _LIB = C.CDLL(path)
        _LIB.vp_last_error.restype = C.c_char_p
    return _LIB


def check(rc):
    if rc != 0:
        raise VpError(lib().vp_last_error().decode())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def tensor_view(t, c=None, c_off=0):
    """vp_tensor view of a contiguous channels-last torch tensor [N,H,W,C] or [N,D,H,W,C],
    optionally restricted to channels [c_off, c_off + c)."""
    assert t.is_contiguous() and t.dtype == torch.float32
    if t.dim() == 4:
        n, h, w, ct = t.shape
        d = 1
    else:
        n, d, h, w, ct = t.shape
    c = ct - c_off if c is None else c
    return VpTensor(t.data_ptr() + 4 * c_off, n, d, h, w, c, ct)


def geom(k, s=(1, 1, 1), p=(0, 0, 0), transposed=False):
    return VpConvGeom(k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2], int(transposed))


def conv_igemm(x_view, g, wpacked, n_pad, kc, out_view, bias=None, act=ACT_NONE, alpha=0.0, split_k=1):
    check(lib().vp_conv_igemm(C.byref(x_view), C.byref(g), ptr(wpacked), n_pad, kc, C.byref(out_view), ptr(bias),
                              act, C.c_float(alpha), split_k, stream_ptr()))


def conv_wgrad(x_view, dy_view, g, dwpacked, n_pad, kc, split_k=1):
    check(lib().vp_conv_wgrad(C.byref(x_view), C.byref(dy_view), C.byref(g), ptr(dwpacked), n_pad, kc, split_k,
                              stream_ptr()))


def eff_taps(k, kind):
    if kind == WKIND_POOLED:
        return (k[1] + 1) * (k[2] + 1)
    if kind == WKIND_UPSAMPLED:
        return (k[1] + 3) * (k[2] + 3)
    return k[0] * k[1] * k[2]


def pad_to(v, m):
    return (v + m - 1) // m * m


def pack_weights(w, k, ci_ref, co, kind, layout, ci_int=None, cmap=None, inv_scale=None, out=None):
    """Returns (wpacked, n_pad, kc)."""
    ci_int = ci_ref if ci_int is None else ci_int
    rows, cols = (co, ci_int) if layout == WLAYOUT_FWD else (ci_int, co)
    n_pad, kc = pad_to(rows, 16), pad_to(cols, 32) // 32
    if n_pad > 256:
        n_pad = pad_to(n_pad, 128)
    taps = eff_taps(k, kind)
    if out is None:
        out = torch.empty(taps * n_pad * kc * 32, device=w.device, dtype=torch.float32)
    check(lib().vp_pack_weights(ptr(w), k[0], k[1], k[2], ci_ref, co, kind, layout, ptr(cmap), ci_int,
                                ptr(inv_scale), ptr(out), n_pad, kc, stream_ptr()))
    return out, n_pad, kc


def unpack_wgrad(dwpacked, k, ci_ref, co, kind, dw, n_pad, kc, ci_int=None, cmap=None):
    ci_int = ci_ref if ci_int is None else ci_int
    check(lib().vp_unpack_wgrad(ptr(dwpacked), k[0], k[1], k[2], ci_ref, co, kind, ptr(cmap), ci_int, ptr(dw),
                                n_pad, kc, stream_ptr()))
