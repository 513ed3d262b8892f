"""GPU probe: accuracy of ONE convolution against an fp64 reference on the same fp32 inputs, in the two arithmetic modes
(TF32 product mode / fp32-exact 3xTF32 mode), as a function of the reduction length K.  Separates operand rounding from the
tensor core's fp32 ACCUMULATION behaviour (a truncating accumulator shows up as an error that grows with K in exact mode)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from video_prediction_b200 import lib as L
from video_prediction_b200.models.savp_model import conv3x


def rnd(*s, seed=0, scale=1.0, pos=False):
    g = torch.Generator(device='cpu').manual_seed(seed)
    t = torch.randn(*s, generator=g) * scale
    return (t.abs() if pos else t).cuda()


for pos in (False, True):
    for cin, k in ((32, 1), (256, 1), (256, 3), (256, 5)):
        x = rnd(4, 16, 16, cin, pos=pos)
        w = rnd(k, k, cin, 64, seed=1, scale=0.05, pos=pos)
        ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
        ref32 = F.conv2d(x.cpu().permute(0, 3, 1, 2), w.cpu().permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1).cuda()
        out = {}
        for mode in ('tf32', 'exact'):
            o = torch.zeros(4, 16, 16, 64, device='cuda')
            wp, n_pad, kc = L.pack_weights(w, (1, k, k), cin, 64, L.WKIND_PLAIN, L.WLAYOUT_FWD)
            g = L.geom((1, k, k), (1, 1, 1), (0, k // 2, k // 2))
            if mode == 'tf32':
                L.conv_igemm(L.tensor_view(x, cin), g, wp, n_pad, kc, L.tensor_view(o, 64), None, 0, 0.0, 1)
            else:
                wl, _, _ = L.pack_weights(w, (1, k, k), cin, 64, L.WKIND_PLAIN, L.WLAYOUT_FWD | L.WLAYOUT_RESIDUAL)
                conv3x(x, cin, g, wp, wl, n_pad, kc, L.tensor_view(o, 64), None, 0, 0.0)
            torch.cuda.synchronize()
            out[mode] = ((o.double() - ref).norm() / ref.norm()).item(), ((o.double() - ref).abs().max() / ref.abs().max()).item()
        e32 = ((ref32.double() - ref).norm() / ref.norm()).item()
        print('K = %5d %s operands: rel-L2 error vs fp64: tf32 mode %.2e | exact mode %.2e (max %.2e) | CPU fp32 conv %.2e'
              % (k * k * cin, 'positive' if pos else 'signed  ', out['tf32'][0], out['exact'][0], out['exact'][1], e32))
