"""ConvLSTM weight-gradient shapes (cfg2, all 11 time steps batched, NB = 32): row-mode kernel vs tap-group kernel.
Checks that the two agree (and a small case against autograd) and prints TFLOP/s (algorithmic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from video_prediction_b200 import lib as L

SHAPES = [('lstm_h0', 32, 72, 128, 5), ('lstm_h1', 16, 136, 256, 5), ('lstm_h2', 8, 264, 512, 5), ('lstm_h4', 32, 104, 128, 5),
          ('conv3x3_64', 64, 40, 32, 3)]
N = int(os.environ.get('WG_N', '352'))
for name, H, Cin, Cout, k in SHAPES:
    n = N if H <= 32 else N // 4
    x = torch.randn(n, H, H, Cin, device='cuda')
    dy = torch.randn(n, H, H, Cout, device='cuda') * 0.1
    g = L.geom((1, k, k), (1, 1, 1), (0, k // 2, k // 2), False)
    n_pad, kc = L.pad_to(Cout, 16), L.pad_to(Cin, 32) // 32
    res = {}
    for mode in ('0', '1'):
        os.environ['VP_WGRAD_ROW'] = mode
        dwp = torch.zeros(k * k * n_pad * kc * 32, device='cuda')
        L.conv_wgrad(L.tensor_view(x, Cin), L.tensor_view(dy, Cout), g, dwp, n_pad, kc, int(os.environ.get('WG_SPLIT', '0')))
        torch.cuda.synchronize()
        res[mode] = dwp.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.conv_wgrad(L.tensor_view(x, Cin), L.tensor_view(dy, Cout), g, dwp, n_pad, kc, int(os.environ.get('WG_SPLIT', '0')))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * n * H * H * Cout * k * k * Cin
        print('%-10s row=%s : %8.1f us  %6.1f TFLOP/s' % (name, mode, ms * 1e3, fl / ms / 1e9))
    err = (res['0'] - res['1']).abs().max().item()
    ref = res['0'].abs().max().item()
    print('   row vs tap-group: max abs diff %.3e (max |dw| %.3e) %s' % (err, ref, 'OK' if err <= 2e-4 * ref + 1e-6 else 'MISMATCH'))
