"""Generates tests/golden/savp_b16_step.npz: the CPU oracle's full SAVP training step at BASELINE configs[1] size
(B=16, 64x64x3, 2+10, ours_savp hparams) -- the configuration bench.py times -- so that the -m gpu parity test compares
the code paths that produce the benchmark number (auto split-K, persistent tiles, row-mode wgrad) without paying several
CPU-minutes per test run.

One oracle run in plain fp32 (the reference's arithmetic).  The CUDA path is compared with it twice: in its product mode
(TF32 tensor-core operands, 5e-2 relative L2 per gradient tensor) and in the fp32-exact 3xTF32 mode (VP_EXACT=1, 2e-3).

Gradients (17.6 M floats) are stored as COUNT SKETCHES: every tensor is multiplied by a seeded +-1 sign vector and summed
in 512 contiguous buckets; E|sketch(a) - sketch(b)|^2 = |a - b|^2, so relative L2 errors are estimated to ~6 % from
512 numbers per tensor.  Outputs are stored as strided samples.

    python tests/golden/make_golden_b16.py            (about one minute on 8 cores)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import savp_oracle as O  # noqa: E402

HK = dict(context_frames=2, sequence_length=12, lr=2e-4, beta1=0.5, l1_weight=100., kl_weight=1.0, video_sn_vae_gan_weight=0.1,
          video_sn_gan_weight=0.1, vae_gan_feature_cdist_weight=10.0, kl_anneal_steps=(0, 10))
B, STEP, SHAPE = 16, 5, (64, 64, 3)
BUCKETS = 512
SAMPLE_STRIDE = 61


def sampling_mask():
    g = torch.Generator().manual_seed(1234)
    return torch.rand(HK['sequence_length'] - 1 - HK['context_frames'], B, generator=g) < 0.5


def sketch(name, t, buckets=BUCKETS):
    """Count sketch of a tensor: seeded signs (seed from the name), contiguous buckets."""
    v = t.detach().reshape(-1).double()
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b'\0'), 'little') % (2 ** 31) + len(name)
    g = torch.Generator().manual_seed(seed)
    sign = torch.randint(0, 2, (v.numel(),), generator=g, dtype=torch.int8).double() * 2 - 1
    v = v * sign
    nb = min(buckets, v.numel())
    idx = (torch.arange(v.numel(), dtype=torch.int64) * nb) // v.numel()
    return torch.zeros(nb, dtype=torch.float64).index_add_(0, idx, v)


def case():
    hp = O.make_hparams(**HK)
    params, _ = O.init_params(hp, SHAPE, seed=0)
    inputs, noise = O.make_synthetic_inputs(hp, B, SHAPE, seed=0)
    return hp, params, inputs, noise


def main():
    hp, params, inputs, noise = case()
    out = {}
    key = 'fp32'
    opt = dict(m={k: torch.zeros_like(v) for k, v in params.items()}, v={k: torch.zeros_like(v) for k, v in params.items()}, t=0)
    t0 = time.time()
    res = O.train_step(params, opt, hp, inputs, noise, step=STEP, sampling=sampling_mask())
    print('%s: oracle step %.1fs  g_loss %.6f d_loss %.6f' % (key, time.time() - t0, res['g_loss'], res['d_loss']), flush=True)
    out['%s/loss/g_loss' % key], out['%s/loss/d_loss' % key] = np.float64(res['g_loss']), np.float64(res['d_loss'])
    for k, v in list(res['g_losses'].items()) + list(res['d_losses'].items()):
        out['%s/loss/%s' % (key, k)] = np.float64(v)
    for kind in ('g_grads', 'd_grads'):
        for k, g in res[kind].items():
            if g is None:
                continue
            out['%s/sketch/%s' % (key, k)] = sketch(k, g).numpy()
            out['%s/norm/%s' % (key, k)] = np.float64(g.double().norm().item())
    for k in ('gen_images', 'gen_images_enc', 'zs_mu_enc'):
        out['%s/out/%s' % (key, k)] = res['outputs'][k].detach().reshape(-1)[::SAMPLE_STRIDE].numpy().astype(np.float32)
    for k in ('discriminator/video/sn_conv3_0/conv3d/u', 'discriminator/encoder/video/sn_conv0_1/conv3d/u'):
        out['%s/u/%s' % (key, k)] = res['params'][k].numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'savp_b16_step.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
