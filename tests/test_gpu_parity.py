"""GPU parity tests (-m gpu): the CUDA path through the reference-facing model API against the CPU oracle on
identical parameters, inputs and noise.  Tolerances: generator outputs 1e-3 max-abs (BASELINE.json north_star);
losses 1e-2 relative; gradients: TF32 tensor-core convolutions through an 11-step BPTT -> per-tensor cosine
similarity >= 0.99 and norm ratio within 10% for every tensor whose reference norm is not numerical noise."""
import os

import numpy as np
import pytest
import torch

from oracle import savp_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def Model():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from video_prediction_b200.models import get_model_class
    return get_model_class('savp')


def _run_forward(Model, hk, B, shape, A=0, seed=0):
    hp = O.make_hparams(**hk)
    params, _ = O.init_params(hp, shape, action_dim=A, seed=seed)
    inputs, noise = O.make_synthetic_inputs(hp, B, shape, action_dim=A, seed=seed)
    with torch.no_grad():
        ref = O.generator(O.Vars(params), hp, inputs, noise, O.ground_truth_mask(hp, B))
    model = Model(mode='test', hparams_dict=hk)
    model.set_params(params)
    binp = {'images': inputs['images'].permute(1, 0, 2, 3, 4)}
    if A:
        binp['actions'] = inputs['actions'].permute(1, 0, 2)
    model.build_graph(binp)
    model.set_inputs(binp, noise)
    model.generator_forward()
    torch.cuda.synchronize()
    return model, ref, inputs


def test_cfg1_deterministic_generator_matches_oracle(Model):
    # BASELINE configs[0]: deterministic generator, 64x64x3, 2 context + 10 predicted, batch 4
    model, ref, _ = _run_forward(Model, dict(context_frames=2, sequence_length=12, nz=0), 4, (64, 64, 3))
    err = (model.outputs['gen_images'].cpu() - ref['gen_images'].permute(1, 0, 2, 3, 4)).abs().max().item()
    assert err <= 1e-3, err
    assert model.outputs['gen_images'].shape == (4, 11, 64, 64, 3)          # batch-major (base_model.py:457-458)


def test_savp_generator_with_actions_matches_oracle(Model):
    model, ref, _ = _run_forward(Model, dict(context_frames=2, sequence_length=8, nz=8), 2, (64, 64, 3), A=4)
    for k in ('gen_images', 'gen_images_enc'):
        err = (model.outputs[k].cpu() - ref[k].permute(1, 0, 2, 3, 4)).abs().max().item()
        assert err <= 1e-3, (k, err)
    assert (model.outputs['zs_mu_enc'].cpu() - ref['zs_mu_enc'].permute(1, 0, 2)).abs().max() <= 1e-3
    m = model.outputs['masks']
    assert (m.sum(dim=-1) - 1).abs().max() < 1e-5                            # appendix C.3
    g = model.outputs['gen_images']
    assert g.min() >= -1e-6 and g.max() <= 1 + 1e-6


def test_kth_shape_single_channel_nz32(Model):
    # BASELINE configs[4] shape: 64x64x1, nz=32 (KTH hparams), shortened sequence
    model, ref, _ = _run_forward(Model, dict(context_frames=3, sequence_length=7, nz=32), 2, (64, 64, 1))
    err = (model.outputs['gen_images_enc'].cpu() - ref['gen_images_enc'].permute(1, 0, 2, 3, 4)).abs().max().item()
    assert err <= 1e-3, err


def test_cfg4_shape_128x128_generator_matches_oracle(Model):
    # BASELINE configs[3] shape: 128x128x3 (4 encoder / 4 decoder levels, savp_model.py:198-210), shortened sequence
    model, ref, _ = _run_forward(Model, dict(context_frames=2, sequence_length=5, nz=8), 1, (128, 128, 3))
    for k in ('gen_images', 'gen_images_enc'):
        err = (model.outputs[k].cpu() - ref[k].permute(1, 0, 2, 3, 4)).abs().max().item()
        assert err <= 1e-3, (k, err)


def test_golden_vectors_small_config(Model):
    g = np.load(os.path.join(GOLD, 'oracle_small.npz'))
    hk = dict(context_frames=2, sequence_length=5, nz=4, ngf=8, nef=8, ndf=8, clip_length=3)
    model, ref, _ = _run_forward(Model, hk, 2, (32, 32, 3), seed=3)
    out = model.outputs['gen_images'].cpu().numpy().transpose(1, 0, 2, 3, 4)
    assert np.abs(out - g['gen_images']).max() <= 1e-3
    out = model.outputs['gen_images_enc'].cpu().numpy().transpose(1, 0, 2, 3, 4)
    assert np.abs(out - g['gen_images_enc']).max() <= 1e-3


def test_full_size_batch_independence(Model):
    # BASELINE configs[1] size (batch 16): every sample's output depends only on that sample (DP correctness)
    hk = dict(context_frames=2, sequence_length=12, nz=8)
    hp = O.make_hparams(**hk)
    inputs, noise = O.make_synthetic_inputs(hp, 16, (64, 64, 3), seed=5)
    model = Model(mode='test', hparams_dict=hk)
    binp = {'images': inputs['images'].permute(1, 0, 2, 3, 4)}
    model.build_graph(binp)
    model.set_inputs(binp, noise)
    model.generator_forward()
    a = model.outputs['gen_images'].clone()
    perm = torch.randperm(16, generator=torch.Generator().manual_seed(0))
    binp2 = {'images': binp['images'][perm]}
    noise2 = {k: v[:, perm] for k, v in noise.items() if torch.is_tensor(v)}
    model.set_inputs(binp2, noise2)
    model.generator_forward()
    b = model.outputs['gen_images']
    assert (b - a[perm.cuda()]).abs().max().item() <= 1e-4   # split-K atomics reorder fp32 sums run to run
    assert a.min() >= -1e-6 and a.max() <= 1 + 1e-6


def _cos(a, b):
    a, b = a.double().reshape(-1).cpu(), b.double().reshape(-1).cpu()
    return (a @ b / (a.norm() * b.norm() + 1e-300)).item(), (a.norm() / (b.norm() + 1e-300)).item()


def test_savp_training_step_matches_oracle(Model):
    hk = dict(context_frames=2, sequence_length=12, lr=2e-4, beta1=0.5, l1_weight=100., kl_weight=1.0, video_sn_vae_gan_weight=0.1,
              video_sn_gan_weight=0.1, vae_gan_feature_cdist_weight=10.0, kl_anneal_steps=(0, 10))
    hp = O.make_hparams(**hk)
    B, step = 2, 5
    params, _ = O.init_params(hp, (64, 64, 3), seed=0)
    inputs, noise = O.make_synthetic_inputs(hp, B, (64, 64, 3))
    opt = dict(m={k: torch.zeros_like(v) for k, v in params.items()}, v={k: torch.zeros_like(v) for k, v in params.items()}, t=0)
    res = O.train_step(params, opt, hp, inputs, noise, step=step)
    model = Model(mode='train', hparams_dict=hk)
    model.set_params(params)
    binp = {'images': inputs['images'].permute(1, 0, 2, 3, 4)}
    model.build_graph(binp)
    model.global_step = step
    model.train_step(binp, noise, sampling=False)
    torch.cuda.synchronize()
    lv = model.losses()
    ref_l = dict(res['g_losses'])
    ref_l.update(res['d_losses'])
    for k, v in ref_l.items():
        assert abs(lv[k] - v) <= 1e-2 * abs(v) + 1e-6, (k, lv[k], v)
    for kind in ('g_grads', 'd_grads'):
        gmax = max(g.norm().item() for g in res[kind].values())
        for k, g in res[kind].items():
            if g.norm().item() < 1e-3 * gmax:
                continue            # e.g. conv biases feeding an instance norm (exactly zero in exact arithmetic), tiny encoder grads
            c, r = _cos(model.grads[k], g)
            assert c >= 0.99 and 0.9 <= r <= 1.1, (k, c, r)
    # spectral-norm u <- u' (ops.py:1046-1048) and TF-Adam update of a well-conditioned tensor
    k = 'discriminator/video/sn_conv3_0/conv3d/u'
    assert (model.params[k].cpu() - res['params'][k]).abs().max() <= 1e-4
    k = 'generator/rnn/savp_cell/masks/conv2d/kernel'
    sig = res['g_grads'][k].abs() > 1e-2 * res['g_grads'][k].abs().max()
    assert (model.params[k].cpu() - res['params'][k])[sig].abs().max() <= 0.2 * hp.lr
    assert model.global_step == step + 1
