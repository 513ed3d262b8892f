"""Weight-gradient GEMMs of the tap-group kernel (strided / 3-D / 3x3 convolutions): merged wide-N MMAs (default) against
one MMA per tap (VP_WGRAD_MERGE=0): agreement and CUDA-event timings."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from video_prediction_b200 import lib as L

NB = 32
CASES = [('D sn_conv0_1 k4 s(1,2,2)', (NB, 10, 64, 64), 32, 64, (4, 4, 4), (1, 2, 2), (1, 1, 1)),
         ('D sn_conv1_0 k3 s1', (NB, 9, 32, 32), 64, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
         ('D sn_conv1_1 k4 s(1,2,2)', (NB, 9, 32, 32), 64, 128, (4, 4, 4), (1, 2, 2), (1, 1, 1)),
         ('D sn_conv2_0 k3 s1', (NB, 8, 16, 16), 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
         ('D sn_conv2_1 k4 s2', (NB, 8, 16, 16), 128, 256, (4, 4, 4), (2, 2, 2), (1, 1, 1)),
         ('D sn_conv3_0 k3 s1', (NB, 4, 8, 8), 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
         ('3x3 head 32->32 (11 steps)', (NB * 11, 1, 64, 64), 32, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
         ('masks 3x3 60->8 (11 steps)', (NB * 11, 1, 64, 64), 60, 8, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
         ('h0 pooled 6x6 s2 16->32 (11 steps)', (NB * 11, 1, 64, 64), 16, 32, (1, 6, 6), (1, 2, 2), (0, 2, 2)),
         ('h1 pooled 4x4 s2 48->64 (11 steps)', (NB * 11, 1, 32, 32), 48, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1))]
for name, xs, cin, cout, k, s, p in CASES:
    osp = tuple((d + 2 * pp - kq) // st + 1 for d, pp, kq, st in zip(xs[1:], p, k, s))
    x = torch.randn(*xs, cin, device='cuda')
    dy = torch.randn(xs[0], *osp, cout, device='cuda') * 0.1
    g = L.geom(k, s, p, False)
    n_pad, kc = L.pad_to(cout, 16), L.pad_to(cin, 32) // 32
    taps = k[0] * k[1] * k[2]
    res, tms = {}, {}
    for mode in ('0', '1'):
        os.environ['VP_WGRAD_MERGE'] = mode
        dwp = torch.zeros(taps * n_pad * kc * 32, device='cuda')
        L.conv_wgrad(L.tensor_view(x, cin), L.tensor_view(dy, cout), g, dwp, n_pad, kc, 0)
        torch.cuda.synchronize()
        res[mode] = dwp.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.conv_wgrad(L.tensor_view(x, cin), L.tensor_view(dy, cout), g, dwp, n_pad, kc, 0)
        e1.record()
        torch.cuda.synchronize()
        tms[mode] = e0.elapsed_time(e1) / 5
    fl = 2.0 * dy.numel() * taps * cin
    err = (res['0'] - res['1']).abs().max().item() / (res['0'].abs().max().item() + 1e-30)
    print('%-36s one MMA per tap %7.1f us %5.0f TF/s | merged %7.1f us %5.0f TF/s | rel diff %.1e %s' % (
        name, tms['0'] * 1e3, fl / tms['0'] / 1e9, tms['1'] * 1e3, fl / tms['1'] / 1e9, err, 'OK' if err < 1e-4 else 'MISMATCH'))
