"""Per-CTA timeline of halo-mode launches (VP_FWD_TRACE=1): CTA lifetime, setup, and what the MMA-issuing thread waited for
(halo tiles / weight tiles / TMEM accumulators), in SM cycles."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from video_prediction_b200 import lib as L

NB = 32


def trace(name, xs, cin, cout, k, s, p, dgrad=False, act=0):
    cs = (cin + 3) // 4 * 4
    w = torch.randn(*(k if k[0] > 1 else k[1:]), cin, cout, device='cuda') * 0.03
    sp3 = (1,) + tuple(xs[1:]) if len(xs) == 3 else tuple(xs[1:])
    osp = tuple((d + 2 * pp - kq) // st + 1 for d, pp, kq, st in zip(sp3, p, k, s))
    oshape = (xs[0],) + (osp if len(xs) == 4 else osp[1:])
    if not dgrad:
        x = torch.randn(*xs, cs, device='cuda')
        wp, n_pad, kc = L.pack_weights(w, k, cin, cout, L.WKIND_PLAIN, L.WLAYOUT_FWD, ci_int=cs)
        out = torch.zeros(*oshape, cout, device='cuda')
        c_in, c_out, tr = cs, cout, False
    else:
        x = torch.randn(*oshape, cout, device='cuda')
        wp, n_pad, kc = L.pack_weights(w, k, cin, cout, L.WKIND_PLAIN, L.WLAYOUT_DGRAD, ci_int=cs)
        out = torch.zeros(*xs, cs, device='cuda')
        c_in, c_out, tr = cout, cs, True
    g = L.geom(k, s, p, tr)

    def call():
        L.conv_igemm(L.tensor_view(x, c_in), g, wp, n_pad, kc, L.tensor_view(out, c_out), None, act, 0.1, 0)
    os.environ.pop('VP_FWD_TRACE', None)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    os.environ['VP_FWD_TRACE'] = '1'
    call()
    torch.cuda.synchronize()
    os.environ.pop('VP_FWD_TRACE', None)
    n = 2048
    buf = (ctypes.c_ulonglong * (16 * n))()
    L.check(L.lib().vp_debug_read_trace(buf, n))
    t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 16).astype(np.int64)
    live = t[:, 0] > t[:, 0].max() - 5_000_000
    t = t[live]
    base = t[:, 0].min()
    life = (t[:, 2] - t[:, 0]) / 1e3
    med = lambda a: float(np.median(a))
    print('%-26s %3d CTAs on %3d SMs | kernel %.1f us (mean of 10) | CTA start spread %.1f us, lifetime med %.1f max %.1f us, setup %.1f us'
          % (name, len(t), len(set(t[:, 7].tolist())), e0.elapsed_time(e1) * 100, (t[:, 0].max() - base) / 1e3, med(life), life.max(),
             med(t[:, 1] - t[:, 0]) / 1e3))
    tot = np.maximum(t[:, 8], 1)
    print('   MMA thread: total %.0f cycles (%.2f GHz), waiting for halo %.0f%%, for weights %.0f%%, for TMEM %.0f%%, issuing %.0f%%' % (
        med(tot), med(tot / np.maximum((t[:, 2] - t[:, 1]), 1)), 100 * med(t[:, 9] / tot), 100 * med(t[:, 10] / tot), 100 * med(t[:, 11] / tot),
        100 * med((tot - t[:, 9] - t[:, 10] - t[:, 11]) / tot)))
    sys.stdout.flush()


if __name__ == '__main__':
    trace('lstm_h0 fwd', (NB, 32, 32), 72, 128, (1, 5, 5), (1, 1, 1), (0, 2, 2))
    trace('lstm_h0 dgrad', (NB, 32, 32), 72, 128, (1, 5, 5), (1, 1, 1), (0, 2, 2), dgrad=True)
    trace('lstm_h1 fwd', (NB, 16, 16), 136, 256, (1, 5, 5), (1, 1, 1), (0, 2, 2))
    trace('lstm_h2 fwd', (NB, 8, 8), 264, 512, (1, 5, 5), (1, 1, 1), (0, 2, 2))
    trace('D sn_conv1_0 k3 s1', (NB, 9, 32, 32), 64, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), act=L.ACT_LRELU)
    trace('D sn_conv0_1 k4 s(1,2,2)', (NB, 10, 64, 64), 32, 64, (4, 4, 4), (1, 2, 2), (1, 1, 1), act=L.ACT_LRELU)
    trace('D sn_conv2_0 k3 s1', (NB, 8, 16, 16), 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), act=L.ACT_LRELU)
