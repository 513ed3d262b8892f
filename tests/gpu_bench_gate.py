"""Micro-benchmark of the ConvLSTM gate convolution shapes (cfg2, NB = 32): box-mode engine vs halo-resident flat
kernel, with tile-width overrides (VP_FWD_BN / VP_FLAT_BN).  Prints TFLOP/s (algorithmic FLOPs, cycling over 11
distinct input/output buffers so that consecutive launches do not hit in L2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from video_prediction_b200 import lib as L

SHAPES = [('lstm_h0', 32, 72, 128), ('lstm_h1', 16, 136, 256), ('lstm_h2', 8, 264, 512)]
NB, S = 32, 11
mode = sys.argv[1] if len(sys.argv) > 1 else 'box'
for name, H, Cin, Cout in SHAPES:
    w = torch.randn(5, 5, Cin, Cout, device='cuda') * 0.03
    wp, n_pad, kc = L.pack_weights(w, (1, 5, 5), Cin, Cout, L.WKIND_PLAIN, L.WLAYOUT_FWD)
    g = L.geom((1, 5, 5), (1, 1, 1), (0, 2, 2), False)
    outs = [torch.zeros(NB, H, H, Cout, device='cuda') for _ in range(S)]
    if mode == 'box':
        xs = [torch.randn(NB, H, H, Cin, device='cuda') for _ in range(S)]
        def run(t):
            L.conv_igemm(L.tensor_view(xs[t], Cin), g, wp, n_pad, kc, L.tensor_view(outs[t], Cout), None, 0, 0.0, int(os.environ.get('SPLIT', '1')))
    else:
        xs = [torch.zeros(NB, H + 2, H + 2, Cin, device='cuda') for _ in range(S)]
        for x in xs:
            x[:, :H, :H] = torch.randn(NB, H, H, Cin, device='cuda')
        def run(t):
            L.conv_flat(L.tensor_view(xs[t], Cin), H, H, g, wp, n_pad, kc, L.tensor_view(outs[t], Cout), None, 0, 0.0,
                        int(os.environ.get('SPLIT', '1')), 0, 0)
    for t in range(S):
        run(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for t in range(S):
            run(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (3 * S)
    fl = 2.0 * NB * H * H * Cout * 25 * Cin
    print('%s %-8s FWD_BN=%s FLAT_BN=%s SPLIT=%s : %7.1f us  %6.1f TFLOP/s' % (mode, name, os.environ.get('VP_FWD_BN'), os.environ.get('VP_FLAT_BN'),
                                                                     os.environ.get('SPLIT'), ms * 1e3, fl / ms / 1e9))
