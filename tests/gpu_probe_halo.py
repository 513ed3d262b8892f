"""GPU probe: halo-mode forward / dgrad engine against the box-mode engine on every convolution geometry of the SAVP step
(same packed weights, same inputs: the two must agree to fp32 summation-order noise), with CUDA-event timings of both."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from video_prediction_b200 import lib as L


def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).cuda()


def ke(k, kind):
    return k if kind == L.WKIND_PLAIN else ((1, k[1] + 1, k[2] + 1) if kind == L.WKIND_POOLED else (1, k[1] + 3, k[2] + 3))


def run(name, xs, cin, cout, k, s, p, kind=0, transposed=False, dgrad=False, act=0, bias=False, iters=20, accumulate=False,
        actgrad=False, cs=None):
    """xs: input dims (n, [d,] h, w).  For dgrad the roles flip: the 'input' is dy with cout channels."""
    cs = cs or (cin + 3) // 4 * 4
    w = rnd(*(k if k[0] > 1 else k[1:]), cin, cout, seed=1, scale=0.05)
    kk = ke(k, kind)
    sp = xs[1:]
    sp3 = (1,) + tuple(sp) if len(sp) == 2 else tuple(sp)
    if not transposed:
        osp = tuple((d + 2 * pp - kq) // st + 1 for d, pp, kq, st in zip(sp3, p, kk, s))
    else:
        osp = tuple(d * st for d, st in zip(sp3, s))
    if not dgrad:
        x = rnd(*xs, cs)
        wp, n_pad, kc = L.pack_weights(w.contiguous(), k, cin, cout, kind, L.WLAYOUT_FWD, ci_int=cs)
        oshape = (xs[0],) + (osp if len(sp) == 3 else osp[1:]) + (cout,)
        xin, c_in, c_out, tr = x, cs, cout, transposed
    else:
        x = rnd(*((xs[0],) + (osp if len(sp) == 3 else osp[1:]) + (cout,)), seed=3)    # dy
        wp, n_pad, kc = L.pack_weights(w.contiguous(), k, cin, cout, kind, L.WLAYOUT_DGRAD, ci_int=cs)
        oshape = tuple(xs) + (cs,)
        xin, c_in, c_out, tr = x, cout, cs, not transposed
    b = rnd(c_out, seed=2) if bias else None
    geom = L.geom(kk, s, p, tr)
    outs, times = {}, {}
    aux = rnd(*oshape, seed=5) if actgrad else None
    for mode in ('0', '1', '1s0'):
        os.environ['VP_HALO'] = mode[0]
        os.environ['VP_HALO_STAGGER'] = '0' if mode.endswith('s0') else '1'
        out = torch.zeros(*oshape, device='cuda') if not accumulate else torch.ones(*oshape, device='cuda')
        def call():
            if actgrad:
                L.conv_igemm_actgrad(L.tensor_view(xin, c_in), geom, wp, n_pad, kc, L.tensor_view(out, c_out), aux.data_ptr(), 0, L.ACT_LRELU, 0.1)
            else:
                L.conv_igemm(L.tensor_view(xin, c_in), geom, wp, n_pad, kc, L.tensor_view(out, c_out), b, act, 0.1, 0, 1 if accumulate else 0)
        try:
            call()
            torch.cuda.synchronize()
        except Exception as ex:      # noqa: BLE001
            print('   %s mode %s: ERROR %s' % (name, mode, ex))
            outs[mode] = None
            continue
        outs[mode] = out.clone()
        if not accumulate:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                call()
            e1.record()
            torch.cuda.synchronize()
            times[mode] = e0.elapsed_time(e1) / iters * 1e3
    ref = outs['0']
    sc = ref.abs().max().item() + 1e-30
    m = 1
    for d in (oshape[:-1]):
        m *= d
    ntaps = kk[0] * kk[1] * kk[2]
    fl = 2.0 * m * cout * ntaps * cin if not dgrad else 2.0 * x.numel() / cout * cout * ntaps * cin
    if transposed != dgrad and any(st > 1 for st in s):
        fl /= (s[0] * s[1] * s[2])
    res = []
    for mode in ('1', '1s0'):
        if outs[mode] is None:
            res.append('ERR')
            continue
        err = (outs[mode] - ref).abs().max().item() / sc
        res.append('%s rel %.1e' % ('ok ' if err < 2e-5 else 'BAD', err))
    print('%-34s box %7.1f us %6.0f TF/s | halo %7.1f us %6.0f TF/s (%s) | no stagger %7.1f us (%s)' % (
        name, times.get('0', 0), fl / max(times.get('0', 1), 1e-9) / 1e6, times.get('1', 0), fl / max(times.get('1', 1), 1e-9) / 1e6, res[0],
        times.get('1s0', 0), res[1]))
    sys.stdout.flush()


if __name__ == '__main__':
    NB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    torch.cuda.set_device(0)
    t0 = time.time()
    # ConvLSTM gate convolutions (rnn_ops.py:121) at cfg2: forward and dgrad
    for nm, h, ci, co in (('lstm_h0', 32, 72, 128), ('lstm_h1', 16, 136, 256), ('lstm_h2', 8, 264, 512)):
        run(nm + ' fwd', (NB, h, h), ci, co, (1, 5, 5), (1, 1, 1), (0, 2, 2))
        run(nm + ' dgrad', (NB, h, h), ci, co, (1, 5, 5), (1, 1, 1), (0, 2, 2), dgrad=True)
    # encoder / decoder convs (pooled-kernel stride 2, bilinear-composed transposed), heads
    run('h0 conv_pool2d 5x5->6x6 s2', (NB, 64, 64), 14, 32, (1, 5, 5), (1, 2, 2), (0, 2, 2), kind=L.WKIND_POOLED, bias=True)
    run('h1 conv_pool2d 3x3->4x4 s2', (NB, 32, 32), 40, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1), kind=L.WKIND_POOLED, bias=True)
    run('h1 conv_pool2d dgrad', (NB, 32, 32), 40, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1), kind=L.WKIND_POOLED, dgrad=True)
    run('h2 conv_pool2d', (NB, 16, 16), 72, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), kind=L.WKIND_POOLED, bias=True)
    run('h3 upsample_conv2d', (NB, 8, 8), 136, 64, (1, 3, 3), (1, 2, 2), (0, 2, 2), kind=L.WKIND_UPSAMPLED, transposed=True, bias=True)
    run('h5 upsample_conv2d', (NB, 32, 32), 72, 32, (1, 3, 3), (1, 2, 2), (0, 2, 2), kind=L.WKIND_UPSAMPLED, transposed=True, bias=True)
    run('h5 upsample_conv2d dgrad', (NB, 32, 32), 72, 32, (1, 3, 3), (1, 2, 2), (0, 2, 2), kind=L.WKIND_UPSAMPLED, transposed=True, dgrad=True)
    run('h6 3x3 head', (NB, 64, 64), 32, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1), bias=True)
    run('masks 3x3 53->8', (NB, 64, 64), 53, 8, (1, 3, 3), (1, 1, 1), (0, 1, 1), bias=True, cs=60)
    run('masks dgrad', (NB, 64, 64), 53, 8, (1, 3, 3), (1, 1, 1), (0, 1, 1), dgrad=True, cs=60)
    run('posterior 4x4 s2 64->128', (NB * 5, 32, 32), 64, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), bias=True)
    # video discriminator (networks.py:83-102), one tower pass = 2B clips
    run('D sn_conv0_0 k3 s1 3->32', (NB, 10, 64, 64), 3, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), bias=True, act=L.ACT_LRELU, iters=5)
    run('D sn_conv0_0 dgrad', (NB, 10, 64, 64), 3, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), dgrad=True, iters=5)
    if True:
        x = rnd(NB, 10, 64, 64, 4)
        w = rnd(3, 3, 3, 3, 32, seed=1, scale=0.05)
        o = torch.zeros(NB, 10, 64, 64, 32, device='cuda')
        one = torch.ones(1, device='cuda')
        b = rnd(32, seed=2)
        for _ in range(2):
            L.conv3d_c4_fwd(x, w, one, b, o, NB, 10, 64, 64, 3, 0.1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.conv3d_c4_fwd(x, w, one, b, o, NB, 10, 64, 64, 3, 0.1)
        e1.record()
        torch.cuda.synchronize()
        print('D sn_conv0_0 CUDA-core kernel (conv3d_c4_fwd): %.1f us' % (e0.elapsed_time(e1) / 5 * 1e3))
    run('D sn_conv0_1 k4 s(1,2,2)', (NB, 10, 64, 64), 32, 64, (4, 4, 4), (1, 2, 2), (1, 1, 1), bias=True, act=L.ACT_LRELU, iters=5)
    run('D sn_conv0_1 dgrad', (NB, 10, 64, 64), 32, 64, (4, 4, 4), (1, 2, 2), (1, 1, 1), dgrad=True, iters=5)
    run('D sn_conv1_0 k3 s1', (NB, 9, 32, 32), 64, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), bias=True, act=L.ACT_LRELU, iters=5)
    run('D sn_conv1_0 dgrad+actgrad', (NB, 9, 32, 32), 64, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), dgrad=True, actgrad=True, iters=5)
    run('D sn_conv1_1 k4 s(1,2,2)', (NB, 9, 32, 32), 64, 128, (4, 4, 4), (1, 2, 2), (1, 1, 1), bias=True, act=L.ACT_LRELU, iters=5)
    run('D sn_conv1_1 dgrad', (NB, 9, 32, 32), 64, 128, (4, 4, 4), (1, 2, 2), (1, 1, 1), dgrad=True, iters=5)
    run('D sn_conv2_0 k3 s1', (NB, 8, 16, 16), 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), bias=True, act=L.ACT_LRELU, iters=5)
    run('D sn_conv2_1 k4 s2', (NB, 8, 16, 16), 128, 256, (4, 4, 4), (2, 2, 2), (1, 1, 1), bias=True, act=L.ACT_LRELU, iters=5)
    run('D sn_conv2_1 dgrad', (NB, 8, 16, 16), 128, 256, (4, 4, 4), (2, 2, 2), (1, 1, 1), dgrad=True, iters=5)
    run('D sn_conv3_0 k3 s1', (NB, 4, 8, 8), 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1), bias=True, act=L.ACT_LRELU, iters=5)
    run('D sn_conv3_0 dgrad', (NB, 4, 8, 8), 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1), dgrad=True, iters=5)
    run('accumulate lstm_h1 dgrad', (4, 16, 16), 136, 256, (1, 5, 5), (1, 1, 1), (0, 2, 2), dgrad=True, accumulate=True)
    print('done in %.1fs' % (time.time() - t0))
