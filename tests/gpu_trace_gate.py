"""Per-CTA timeline of one gate-convolution launch (VP_FWD_TRACE=1): where does the time of a ~40 us kernel go?
Stamps (globaltimer ns, relative to the earliest CTA start): 0 CTA start, 1 MMA thread ready (setup done), 2 first operand
stage landed, 3 last MMA issued, 4 epilogue sees the accumulator, 5 epilogue stores done, 6 CTA exit, 7 = SM id."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['VP_FWD_TRACE'] = '1'
import numpy as np
import torch
from video_prediction_b200 import lib as L

SHAPES = [('lstm_h0', 32, 72, 128), ('lstm_h1', 16, 136, 256), ('lstm_h2', 8, 264, 512)]
NB = 32
for name, H, Cin, Cout in SHAPES:
    w = torch.randn(5, 5, Cin, Cout, device='cuda') * 0.03
    wp, n_pad, kc = L.pack_weights(w, (1, 5, 5), Cin, Cout, L.WKIND_PLAIN, L.WLAYOUT_FWD)
    g = L.geom((1, 5, 5), (1, 1, 1), (0, 2, 2), False)
    xs = [torch.randn(NB, H, H, Cin, device='cuda') for _ in range(4)]
    out = torch.zeros(NB, H, H, Cout, device='cuda')
    for t in range(4):
        L.conv_igemm(L.tensor_view(xs[t], Cin), g, wp, n_pad, kc, L.tensor_view(out, Cout), None, 0, 0.0, int(os.environ.get('SPLIT', '1')))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.conv_igemm(L.tensor_view(xs[0], Cin), g, wp, n_pad, kc, L.tensor_view(out, Cout), None, 0, 0.0, int(os.environ.get('SPLIT', '1')))
    e1.record()
    torch.cuda.synchronize()
    n = 2048
    buf = (ctypes.c_ulonglong * (16 * n))()
    L.check(L.lib().vp_debug_read_trace(buf, n))
    tr = np.frombuffer(buf, dtype=np.uint64).reshape(n, 16).astype(np.int64)
    # CTAs of this launch: stamp 0 within the last launch window
    t_hi = tr[:, 0].max()
    live = tr[:, 0] > t_hi - 2_000_000
    tr = tr[live]
    base = tr[:, 0].min()
    cyc = tr[:, 1].copy()
    print('   MMA thread: cycles waiting on full barriers median %.0f, issuing (4 MMA + commit) median %.0f, total %.0f' % (np.median(tr[:, 8]), np.median(tr[:, 9]), np.median(cyc)))
    rel = (tr[:, :7] - base) / 1e3
    rel[:, 1] = rel[:, 0]
    print('%s: %d CTAs traced, event time %.1f us, SMs used %d' % (name, len(tr), e0.elapsed_time(e1) * 1e3, len(set(tr[:, 7].tolist()))))
    names = ['start', 'mma_ready', 'first_stage', 'last_mma_issued', 'epi_begin', 'epi_end', 'exit']
    for i, nm in enumerate(names):
        col = rel[:, i]
        col = col[col >= 0]
        if col.size == 0:
            continue
        print('   %-16s min %7.2f  median %7.2f  max %7.2f us' % (nm, col.min(), np.median(col), col.max()))
    d = rel[:, 3] - rel[:, 2]
    print('   mma phase (first_stage -> last issue): min %.2f median %.2f max %.2f us' % (d.min(), np.median(d), d.max()))
    d = rel[:, 3] - rel[:, 2]
    print('   mma phase in SM cycles: median %.0f -> implied SM clock %.2f GHz' % (np.median(cyc), np.median(cyc / np.maximum(d, 1e-3)) / 1e3))
    d = rel[:, 5] - rel[:, 4]
    print('   epilogue: median %.2f max %.2f us;  setup (start->mma_ready): median %.2f;  fill (ready->first_stage): median %.2f' % (
        np.median(d), d.max(), np.median(rel[:, 1] - rel[:, 0]), np.median(rel[:, 2] - rel[:, 1])))
