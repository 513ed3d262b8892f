"""GPU probe: one full SAVP training step (D then G) of the CUDA path vs the CPU oracle (autograd)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import savp_oracle as O
from video_prediction_b200.models import SAVPVideoPredictionModel


def rel(a, b):
    a = a.detach().cpu().double().reshape(-1)
    b = b.detach().cpu().double().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), b.norm().item()


def run(hk, B, HW=64, C=3, A=0, step=5, tag=''):
    hp = O.make_hparams(**hk)
    params, trainable = O.init_params(hp, (HW, HW, C), action_dim=A, seed=0)
    inputs, noise = O.make_synthetic_inputs(hp, B, (HW, HW, C), action_dim=A)
    opt = dict(m={k: torch.zeros_like(v) for k, v in params.items()}, v={k: torch.zeros_like(v) for k, v in params.items()}, t=0)
    t0 = time.time()
    res = O.train_step(params, opt, hp, inputs, noise, step=step)
    t_cpu = time.time() - t0
    model = SAVPVideoPredictionModel(mode='train', hparams_dict=hk)
    model.set_params(params)
    binp = {'images': inputs['images'].permute(1, 0, 2, 3, 4)}
    if A:
        binp['actions'] = inputs['actions'].permute(1, 0, 2)
    model.build_graph(binp)
    model.global_step = step
    torch.cuda.synchronize()
    t0 = time.time()
    model.train_step(binp, noise)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    print('== %s B=%d: oracle CPU step %.1fs, CUDA path (eager, first call) %.3fs' % (tag, B, t_cpu, t_gpu))
    lv = model.losses()
    ref_l = dict(res['g_losses'])
    ref_l.update(res.get('d_losses', {}))
    worst = 0.0
    for k, v in ref_l.items():
        print('  loss %-44s cuda %.6f  oracle %.6f' % (k, lv[k], v))
        worst = max(worst, abs(lv[k] - v) / (abs(v) + 1e-6))
    bad = []
    for kind in ('g_grads', 'd_grads'):
        if kind not in res:
            continue
        errs = []
        for k, g in res[kind].items():
            if g is None:
                continue
            r, n = rel(model.grads[k], g)
            errs.append((r, k, n))
        gmax = max(e[2] for e in errs)
        noise = [e for e in errs if e[2] < 1e-5 * gmax]     # e.g. conv biases in front of an instance norm: exactly 0 in theory
        errs = [e for e in errs if e[2] >= 1e-5 * gmax]
        errs.sort(reverse=True)
        print('  %s: %d tensors (+%d with |ref| < 1e-5 of the largest, ignored), relative L2 errors:' % (kind, len(errs), len(noise)))
        for r, k, n in errs:
            print('     %.3e  |ref|=%.3e  %s' % (r, n, k))
        bad += [e for e in errs if e[0] > 5e-2]
    perr = []
    allg = dict(res['g_grads'])
    allg.update(res.get('d_grads', {}))
    for k, v in res['params'].items():
        if k in model.params:
            d = (model.params[k].detach().cpu() - v).abs()
            if k in allg and allg[k] is not None:   # Adam's first step is +-lr*sign(g): only compare where g is not noise
                d = d[allg[k].abs() > 1e-3 * allg[k].abs().max()]
            if d.numel():
                perr.append((d.max().item(), k))
    perr.sort(reverse=True)
    print('  params after Adam: worst max-abs diffs', ['%.2e %s' % e for e in perr[:3]])
    ok = worst < 2e-2 and not bad
    print('VERDICT %s %s (loss rel err %.2e, %d grad tensors over 5e-2)' % (tag, 'PASS' if ok else 'FAIL', worst, len(bad)))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'det'):
        run(dict(context_frames=2, sequence_length=12, nz=0, l1_weight=1.0, lr=1e-3), 2, tag='deterministic_l1')
    if which in ('all', 'vae'):
        run(dict(context_frames=2, sequence_length=12, nz=8, l1_weight=1.0, kl_weight=1e-3, kl_anneal_steps=(0, 10), lr=1e-3), 2,
            tag='vae_l1')
    if which in ('all', 'savp'):
        run(dict(context_frames=2, sequence_length=12, lr=2e-4, beta1=0.5, l1_weight=100., kl_weight=1.0,
                 video_sn_vae_gan_weight=0.1, video_sn_gan_weight=0.1, vae_gan_feature_cdist_weight=10.0,
                 kl_anneal_steps=(0, 10)), 2, tag='savp')
