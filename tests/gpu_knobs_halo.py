"""GPU experiment: what bounds the halo kernel?  Times lstm_h0 / D sn_conv1_0 with the weight ring depth, the N tile and the
operand loads varied one at a time (VP_HALO_SKIP runs produce wrong results by design)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from video_prediction_b200 import lib as L

NB = 32


def setup(xs, cin, cout, k, s, p):
    cs = (cin + 3) // 4 * 4
    w = torch.randn(*(k if k[0] > 1 else k[1:]), cin, cout, device='cuda') * 0.03
    sp3 = (1,) + tuple(xs[1:]) if len(xs) == 3 else tuple(xs[1:])
    osp = tuple((d + 2 * pp - kq) // st + 1 for d, pp, kq, st in zip(sp3, p, k, s))
    oshape = (xs[0],) + (osp if len(xs) == 4 else osp[1:])
    x = torch.randn(*xs, cs, device='cuda')
    wp, n_pad, kc = L.pack_weights(w, k, cin, cout, L.WKIND_PLAIN, L.WLAYOUT_FWD, ci_int=cs)
    out = torch.zeros(*oshape, cout, device='cuda')
    g = L.geom(k, s, p, False)
    return lambda: L.conv_igemm(L.tensor_view(x, cs), g, wp, n_pad, kc, L.tensor_view(out, cout), None, 0, 0.1, 0)


def timeit(call, iters=20):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def burn(ms=30):
    """keep the GPU busy so that the SM clock is at its sustained level when the measurement starts"""
    a = torch.randn(4096, 4096, device='cuda')
    t = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while True:
        a = (a @ a).clamp_(-1, 1)
        e1.record()
        torch.cuda.synchronize()
        if e0.elapsed_time(e1) > ms:
            break


for name, args in (('lstm_h0', ((NB, 32, 32), 72, 128, (1, 5, 5), (1, 1, 1), (0, 2, 2))),
                   ('lstm_h1', ((NB, 16, 16), 136, 256, (1, 5, 5), (1, 1, 1), (0, 2, 2))),
                   ('D sn_conv1_0', ((NB, 9, 32, 32), 64, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1)))):
    call = setup(*args)
    base = dict(VP_HALO='1')
    for label, env in (('box mode', dict(VP_HALO='0')), ('halo default', {}), ('halo after 30 ms burn', dict(BURN='1')),
                       ('b_stages 3', dict(VP_HALO_BSTAGES='3')), ('b_stages 5', dict(VP_HALO_BSTAGES='5')),
                       ('N tile 64', dict(VP_HALO_BN='64')), ('N tile 32', dict(VP_HALO_BN='32')),
                       ('no halo loads', dict(VP_HALO_SKIP='1')), ('no weight loads', dict(VP_HALO_SKIP='2')),
                       ('no loads at all', dict(VP_HALO_SKIP='3'))):
        for k in ('VP_HALO_BSTAGES', 'VP_HALO_BN', 'VP_HALO_SKIP', 'VP_HALO'):
            os.environ.pop(k, None)
        os.environ.update(base)
        os.environ.update({k: v for k, v in env.items() if k != 'BURN'})
        if 'BURN' in env:
            burn()
        print('%-14s %-24s %7.1f us' % (name, label, timeit(call)))
        sys.stdout.flush()
