"""One-hot diagnosis of vp_conv3d_c4_wgrad_tc: x = delta at one voxel / channel, dy = delta at one voxel / channel; prints where
the product lands in gw[dz][dy][dx][ci][co] and where it should."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_prediction_b200 import lib as L  # noqa: E402

N, D, H, W, C = 1, 3, 4, 64, 3


def run(xv, dyv, ci, co):
    x = torch.zeros(N, D, H, W, 4, device='cuda')
    dy = torch.zeros(N, D, H, W, 32, device='cuda')
    x[0, xv[0], xv[1], xv[2], ci] = 1.0
    dy[0, dyv[0], dyv[1], dyv[2], co] = 1.0
    g = torch.zeros(27 * C * 32, device='cuda')
    L.check(L.lib().vp_conv3d_c4_wgrad_tc(L.ptr(x), L.ptr(dy), L.ptr(g), N, D, H, W, C, L.stream_ptr()))
    torch.cuda.synchronize()
    g = g.view(3, 3, 3, C, 32).cpu()
    nz = [(tuple(i.tolist()), float(g[tuple(i.tolist())])) for i in g.nonzero()]
    off = tuple(a - b + 1 for a, b in zip(xv, dyv))
    print('x@%s ci=%d  dy@%s co=%d  expect gw[%s][%d][%d]=1  got %s' % (xv, ci, dyv, co, off, ci, co, nz))


for xv, dyv in [((1, 1, 10), (1, 1, 10)), ((1, 1, 11), (1, 1, 10)), ((1, 1, 9), (1, 1, 10)), ((1, 2, 10), (1, 1, 10)),
                ((2, 1, 10), (1, 1, 10)), ((0, 0, 9), (1, 1, 10)), ((1, 1, 13), (1, 1, 13)), ((1, 1, 14), (1, 1, 13)),
                ((1, 1, 0), (1, 1, 0)), ((1, 1, 63), (1, 1, 63)), ((1, 1, 17), (1, 1, 16)), ((1, 1, 40), (1, 1, 41))]:
    run(xv, dyv, 1, 5)
