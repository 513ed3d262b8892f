"""GPU probe for the tcgen05 implicit-GEMM engine: every case runs in its own subprocess (a device
trap in one case must not poison the CUDA context of the others) and is compared with the oracle's
fp64 restatement evaluated on the GPU.  Usage: python tests/gpu_probe_igemm.py [--case NAME]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ['conv3x3', 'conv5x5_lstm0', 'conv5x5_lstm2', 'pooled6x6', 'enc4x4s2', 'upsample', 'conv3d_k3',
         'conv3d_k4s122', 'conv3d_k4s222', 'slice_bias_act', 'splitk', 'dgrad_s1', 'dgrad_s2', 'dgrad_up',
         'wgrad_s1', 'wgrad_pooled', 'wgrad_up', 'wgrad_3d', 'smallN']


def run_case(name):
    import torch
    import torch.nn.functional as F
    from oracle import savp_oracle as O
    from video_prediction_b200 import lib as L
    torch.manual_seed(0)
    dev = 'cuda'

    def rnd(*s):
        return torch.randn(*s, device=dev, dtype=torch.float32)

    def tf32(x):  # round-to-nearest-even emulation of tf32 inputs
        xi = x.contiguous().view(torch.int32)
        xi = (xi + 0x0FFF + ((xi >> 13) & 1)) & ~0x1FFF
        return xi.view(torch.float32)

    def report(y, ref, scale_ref=None):
        err = (y.double() - ref).abs().max().item()
        sc = ref.abs().max().item()
        print('RESULT %s max_abs_err %.3e ref_max %.3e rel %.3e' % (name, err, sc, err / max(sc, 1e-30)))
        return err / max(sc, 1e-30)

    def conv_ref(x, w, strides, pads, transposed=False, out_hw=None):
        """fp64 reference; x NHWC/NDHWC [..., C]; w [k..., ci, co]; explicit zero pad-before `pads`."""
        nd = x.dim() - 2
        xd, wd = x.double(), w.double()
        if nd == 2:
            xn = xd.permute(0, 3, 1, 2)
            if not transposed:
                oh, ow = out_hw
                kh, kw = w.shape[:2]
                pa_h = max((oh - 1) * strides[0] + kh - x.shape[1] - pads[0], 0)
                pa_w = max((ow - 1) * strides[1] + kw - x.shape[2] - pads[1], 0)
                xn = F.pad(xn, (pads[1], pa_w, pads[0], pa_h))
                y = F.conv2d(xn, wd.permute(3, 2, 0, 1), stride=strides)[:, :, :oh, :ow]
            else:
                y = F.conv_transpose2d(xn, wd.permute(2, 3, 0, 1), stride=strides, padding=pads)
            return y.permute(0, 2, 3, 1)
        xn = xd.permute(0, 4, 1, 2, 3)
        od, oh, ow = out_hw
        k = w.shape[:3]
        pa = [max((o - 1) * s + kk - i - p, 0) for o, s, kk, i, p in zip((od, oh, ow), strides, k, x.shape[1:4], pads)]
        xn = F.pad(xn, (pads[2], pa[2], pads[1], pa[1], pads[0], pa[0]))
        y = F.conv3d(xn, wd.permute(4, 3, 0, 1, 2), stride=strides)[:, :, :od, :oh, :ow]
        return y.permute(0, 2, 3, 4, 1)

    def run_fwd(x, c_used, w_ref, k, s, p, out_shape, kind=L.WKIND_PLAIN, transposed=False, cmap=None, bias=None,
                act=L.ACT_NONE, alpha=0.0, split_k=1, out=None, out_off=0):
        ci_ref, co = w_ref.shape[-2], w_ref.shape[-1]
        cm = None if cmap is None else torch.tensor(cmap, device=dev, dtype=torch.int32)
        wp, n_pad, kc = L.pack_weights(w_ref.contiguous(), k, ci_ref, co, kind, L.WLAYOUT_FWD, ci_int=c_used, cmap=cm)
        if out is None:
            out = torch.zeros(*out_shape, co, device=dev)
        ke = k if kind == L.WKIND_PLAIN else ((1, k[1] + 1, k[2] + 1) if kind == L.WKIND_POOLED else (1, k[1] + 3, k[2] + 3))
        g = L.geom(ke, s, p, transposed)
        L.conv_igemm(L.tensor_view(x, c_used), g, wp, n_pad, kc, L.tensor_view(out, co, out_off), bias, act, alpha, split_k)
        torch.cuda.synchronize()
        return out

    if name == 'conv3x3':
        x, w = rnd(2, 32, 32, 32), rnd(3, 3, 32, 32) * 0.1
        y = run_fwd(x, 32, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 32, 32))
        return report(y, conv_ref(tf32(x), tf32(w), (1, 1), (1, 1), out_hw=(32, 32)))
    if name == 'conv5x5_lstm0':
        x, w = rnd(4, 32, 32, 72), rnd(5, 5, 72, 128) * 0.05
        y = run_fwd(x, 72, w, (1, 5, 5), (1, 1, 1), (0, 2, 2), (4, 32, 32))
        return report(y, conv_ref(tf32(x), tf32(w), (1, 1), (2, 2), out_hw=(32, 32)))
    if name == 'conv5x5_lstm2':
        x, w = rnd(6, 8, 8, 264), rnd(5, 5, 264, 512) * 0.02
        y = run_fwd(x, 264, w, (1, 5, 5), (1, 1, 1), (0, 2, 2), (6, 8, 8))
        return report(y, conv_ref(tf32(x), tf32(w), (1, 1), (2, 2), out_hw=(8, 8)))
    if name == 'pooled6x6':
        x, w = rnd(2, 64, 64, 16), rnd(5, 5, 14, 32) * 0.1
        x[..., 14:] = 7.0  # garbage in padding channels must be ignored (c_used = 14)
        y = run_fwd(x, 14, w, (1, 5, 5), (1, 2, 2), (0, 2, 2), (2, 32, 32), kind=L.WKIND_POOLED)
        ref = O.conv_pool2d(tf32(x)[..., :14].double(), w.double(), torch.zeros(32, device=dev, dtype=torch.float64))
        return report(y, ref)
    if name == 'enc4x4s2':
        x, w = rnd(4, 64, 64, 8), rnd(4, 4, 6, 64) * 0.1
        y = run_fwd(x, 6, w, (1, 4, 4), (1, 2, 2), (0, 1, 1), (4, 32, 32))
        return report(y, conv_ref(tf32(x)[..., :6], tf32(w), (2, 2), (1, 1), out_hw=(32, 32)))
    if name == 'upsample':
        x, w = rnd(2, 8, 8, 136), rnd(3, 3, 136, 64) * 0.05
        y = run_fwd(x, 136, w, (1, 3, 3), (1, 2, 2), (0, 2, 2), (2, 16, 16), kind=L.WKIND_UPSAMPLED, transposed=True)
        ref = O.upsample_conv2d(tf32(x).double(), w.double(), torch.zeros(64, device=dev, dtype=torch.float64))
        return report(y, ref)
    if name == 'conv3d_k3':
        x, w = rnd(2, 10, 64, 64, 4), rnd(3, 3, 3, 3, 32) * 0.1
        y = run_fwd(x, 3, w, (3, 3, 3), (1, 1, 1), (1, 1, 1), (2, 10, 64, 64))
        return report(y, conv_ref(tf32(x)[..., :3], tf32(w), (1, 1, 1), (1, 1, 1), out_hw=(10, 64, 64)))
    if name == 'conv3d_k4s122':
        x, w = rnd(2, 10, 32, 32, 32), rnd(4, 4, 4, 32, 64) * 0.05
        y = run_fwd(x, 32, w, (4, 4, 4), (1, 2, 2), (1, 1, 1), (2, 9, 16, 16))
        return report(y, conv_ref(tf32(x), tf32(w), (1, 2, 2), (1, 1, 1), out_hw=(9, 16, 16)))
    if name == 'conv3d_k4s222':
        x, w = rnd(2, 8, 16, 16, 128), rnd(4, 4, 4, 128, 256) * 0.02
        y = run_fwd(x, 128, w, (4, 4, 4), (2, 2, 2), (1, 1, 1), (2, 4, 8, 8))
        return report(y, conv_ref(tf32(x), tf32(w), (2, 2, 2), (1, 1, 1), out_hw=(4, 8, 8)))
    if name == 'slice_bias_act':
        x, w = rnd(2, 16, 16, 40), rnd(3, 3, 40, 32) * 0.1
        b = rnd(32)
        out = torch.full((2, 16, 16, 72), 5.0, device=dev)
        run_fwd(x, 40, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), None, bias=b, act=L.ACT_LRELU, alpha=0.2, out=out, out_off=8)
        ref = conv_ref(tf32(x), tf32(w), (1, 1), (1, 1), out_hw=(16, 16)) + b.double()
        ref = torch.maximum(0.2 * ref, ref)
        ok_rest = bool((out[..., :8] == 5.0).all() and (out[..., 40:] == 5.0).all())
        print('INFO untouched-channels-intact', ok_rest)
        r = report(out[..., 8:40], ref)
        return r if ok_rest else 1.0
    if name == 'splitk':
        x, w = rnd(6, 8, 8, 264), rnd(5, 5, 264, 512) * 0.02
        b = rnd(512)
        y = run_fwd(x, 264, w, (1, 5, 5), (1, 1, 1), (0, 2, 2), (6, 8, 8), bias=b, split_k=5)
        return report(y, conv_ref(tf32(x), tf32(w), (1, 1), (2, 2), out_hw=(8, 8)) + b.double())
    if name == 'smallN':
        x, w = rnd(2, 64, 64, 56), rnd(3, 3, 53, 7) * 0.1
        y = run_fwd(x, 53, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), None, out=torch.zeros(2, 64, 64, 8, device=dev))
        return report(y[..., :7], conv_ref(tf32(x)[..., :53], tf32(w), (1, 1), (1, 1), out_hw=(64, 64)))

    # ---------------- dgrad: dx = engine(dy, W packed DGRAD, geometry transposed-flag flipped)
    def run_dgrad(dy, w_ref, k, s, p, x_shape, c_int, kind=L.WKIND_PLAIN, fwd_transposed=False):
        ci_ref, co = w_ref.shape[-2], w_ref.shape[-1]
        wp, n_pad, kc = L.pack_weights(w_ref.contiguous(), k, ci_ref, co, kind, L.WLAYOUT_DGRAD, ci_int=c_int)
        dx = torch.zeros(*x_shape, c_int, device=dev)
        ke = k if kind == L.WKIND_PLAIN else ((1, k[1] + 1, k[2] + 1) if kind == L.WKIND_POOLED else (1, k[1] + 3, k[2] + 3))
        g = L.geom(ke, s, p, not fwd_transposed)
        L.conv_igemm(L.tensor_view(dy, co), g, wp, n_pad, kc, L.tensor_view(dx, c_int))
        torch.cuda.synchronize()
        return dx
    if name == 'dgrad_s1':
        x = rnd(2, 16, 16, 72).double().requires_grad_(True)
        w, dy = rnd(5, 5, 72, 128) * 0.05, rnd(2, 16, 16, 128)
        y = O.conv2d_tf(x, tf32(w).double(), padding='SAME')
        (gx,) = torch.autograd.grad(y, x, tf32(dy).double())
        dx = run_dgrad(dy, w, (1, 5, 5), (1, 1, 1), (0, 2, 2), (2, 16, 16), 72)
        return report(dx, gx)
    if name == 'dgrad_s2':
        x = rnd(2, 32, 32, 40).double().requires_grad_(True)
        w, dy = rnd(3, 3, 40, 64) * 0.05, rnd(2, 16, 16, 64)
        y = O.conv_pool2d(x, w.double(), torch.zeros(64, device=dev, dtype=torch.float64))
        (gx,) = torch.autograd.grad(y, x, tf32(dy).double())
        dx = run_dgrad(dy, w, (1, 3, 3), (1, 2, 2), (0, 1, 1), (2, 32, 32), 40, kind=L.WKIND_POOLED)
        return report(dx, gx)
    if name == 'dgrad_up':
        x = rnd(2, 8, 8, 136).double().requires_grad_(True)
        w, dy = rnd(3, 3, 136, 64) * 0.05, rnd(2, 16, 16, 64)
        y = O.upsample_conv2d(x, w.double(), torch.zeros(64, device=dev, dtype=torch.float64))
        (gx,) = torch.autograd.grad(y, x, tf32(dy).double())
        dx = run_dgrad(dy, w, (1, 3, 3), (1, 2, 2), (0, 2, 2), (2, 8, 8), 136, kind=L.WKIND_UPSAMPLED, fwd_transposed=True)
        return report(dx, gx)

    # ---------------- wgrad
    def run_wgrad(x, c_int, dy, w_shape, k, s, p, kind=L.WKIND_PLAIN, transposed=False, split_k=4):
        ci_ref, co = w_shape[-2], w_shape[-1]
        n_pad, kc = L.pad_to(co, 16), L.pad_to(c_int, 32) // 32
        ke = k if kind == L.WKIND_PLAIN else ((1, k[1] + 1, k[2] + 1) if kind == L.WKIND_POOLED else (1, k[1] + 3, k[2] + 3))
        taps = ke[0] * ke[1] * ke[2]
        dwp = torch.zeros(taps * n_pad * kc * 32, device=dev)
        g = L.geom(ke, s, p, transposed)
        L.conv_wgrad(L.tensor_view(x, c_int), L.tensor_view(dy, co), g, dwp, n_pad, kc, split_k)
        dw = torch.zeros(*w_shape, device=dev)
        L.unpack_wgrad(dwp, k, ci_ref, co, kind, dw, n_pad, kc, ci_int=c_int)
        torch.cuda.synchronize()
        return dw
    if name == 'wgrad_s1':
        x, dy = rnd(4, 16, 16, 72), rnd(4, 16, 16, 128)
        w = (rnd(5, 5, 72, 128) * 0.05).double().requires_grad_(True)
        y = O.conv2d_tf(tf32(x).double(), w, padding='SAME')
        (gw,) = torch.autograd.grad(y, w, tf32(dy).double())
        dw = run_wgrad(x, 72, dy, (5, 5, 72, 128), (1, 5, 5), (1, 1, 1), (0, 2, 2))
        return report(dw, gw)
    if name == 'wgrad_pooled':
        x, dy = rnd(2, 32, 32, 40), rnd(2, 16, 16, 64)
        w = (rnd(3, 3, 40, 64) * 0.05).double().requires_grad_(True)
        y = O.conv_pool2d(tf32(x).double(), w, torch.zeros(64, device=dev, dtype=torch.float64))
        (gw,) = torch.autograd.grad(y, w, tf32(dy).double())
        dw = run_wgrad(x, 40, dy, (3, 3, 40, 64), (1, 3, 3), (1, 2, 2), (0, 1, 1), kind=L.WKIND_POOLED)
        return report(dw, gw)
    if name == 'wgrad_up':
        x, dy = rnd(2, 8, 8, 136), rnd(2, 16, 16, 64)
        w = (rnd(3, 3, 136, 64) * 0.05).double().requires_grad_(True)
        y = O.upsample_conv2d(tf32(x).double(), w, torch.zeros(64, device=dev, dtype=torch.float64))
        (gw,) = torch.autograd.grad(y, w, tf32(dy).double())
        dw = run_wgrad(x, 136, dy, (3, 3, 136, 64), (1, 3, 3), (1, 2, 2), (0, 2, 2), kind=L.WKIND_UPSAMPLED, transposed=True)
        return report(dw, gw)
    if name == 'wgrad_3d':
        x, dy = rnd(2, 10, 32, 32, 32), rnd(2, 9, 16, 16, 64)
        w = (rnd(4, 4, 4, 32, 64) * 0.05).double().requires_grad_(True)
        xp = F.pad(tf32(x).double(), (0, 0, 1, 1, 1, 1, 1, 1))
        y = O.conv3d_tf_valid(xp, w, (1, 2, 2))
        (gw,) = torch.autograd.grad(y, w, tf32(dy).double())
        dw = run_wgrad(x, 32, dy, (4, 4, 4, 32, 64), (4, 4, 4), (1, 2, 2), (1, 1, 1))
        return report(dw, gw)
    raise SystemExit('unknown case ' + name)


def main():
    if '--case' in sys.argv:
        name = sys.argv[sys.argv.index('--case') + 1]
        rel = run_case(name)
        print('VERDICT %s %s' % (name, 'PASS' if rel < 2e-3 else 'FAIL'))
        return
    results = {}
    cases = CASES
    if '--only' in sys.argv:
        cases = sys.argv[sys.argv.index('--only') + 1].split(',')
    for name in cases:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', name], capture_output=True,
                               text=True, timeout=180)
            out = r.stdout + r.stderr
        except subprocess.TimeoutExpired:
            out = 'TIMEOUT'
        lines = [l for l in out.splitlines() if l.startswith(('RESULT', 'VERDICT', 'INFO'))]
        verdict = 'PASS' if any(l.startswith('VERDICT') and l.endswith('PASS') for l in lines) else 'FAIL'
        results[name] = verdict
        print('== %s: %s (%.1fs)' % (name, verdict, time.time() - t0))
        for l in lines:
            print('   ' + l)
        if verdict == 'FAIL':
            print('\n'.join('   | ' + l for l in out.splitlines()[-12:]))
        sys.stdout.flush()
    print('SUMMARY', results)


if __name__ == '__main__':
    main()
