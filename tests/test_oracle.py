"""CPU tests that pin the oracle: the reference's two docstring identities (ops.py:652-679, 799-817), an
independent direct-loop numpy restatement of the TF ops, the appendix-C invariants, fp32 vs fp64 agreement
and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import savp_oracle as O

D = torch.float64
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def t(*s, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g, dtype=D)


def test_conv_pool2d_docstring_identity():
    # ops.py:799-817: conv_pool2d(x,k,b,strides=2) == pool2d(conv2d(x,k,b), 2, 2), atol 1e-5, shapes [4,16,16,32]->64
    x, k, b = t(4, 16, 16, 32), t(3, 3, 32, 64, seed=1), t(64, seed=2)
    a = O.conv_pool2d(x, k, b)
    c = O.conv2d_tf(x, k, padding='SAME', bias=b)
    c = F.avg_pool2d(c.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert np.allclose(a.numpy(), c.numpy(), atol=1e-5)


def test_upsample_conv2d_docstring_identity():
    # ops.py:652-679: upsample_conv2d == upsample2d(VALID) -> conv2d(FULL) -> crop, atol 1e-5, shapes [4,8,8,64]->32
    x, k, b = t(4, 8, 8, 64), t(3, 3, 64, 32, seed=1), t(32, seed=2)
    a = O.upsample_conv2d(x, k, b)
    b2 = torch.tensor(O.bilinear_kernel_2x(), dtype=D)
    up = F.conv_transpose2d(x.permute(0, 3, 1, 2), b2[None, None].repeat(64, 1, 1, 1), stride=2, groups=64).permute(0, 2, 3, 1)
    full = O.conv2d_tf(up, k, padding='FULL', bias=b)
    ct = 1 + 2 - 1          # crop_top = (2 - 0)//2 + full_pad_end - same_pad_end
    full = full[:, ct:ct + 16, ct:ct + 16]
    assert np.allclose(a.numpy(), full.numpy(), atol=1e-5)


def _conv2d_loops(x, w, stride, pads):
    """Direct-loop numpy restatement of tf.nn.conv2d (cross-correlation, explicit pad-before/after)."""
    n, h, wd, ci = x.shape
    kh, kw, _, co = w.shape
    (pt, pb), (pl, pr) = pads
    xp = np.zeros((n, h + pt + pb, wd + pl + pr, ci))
    xp[:, pt:pt + h, pl:pl + wd] = x
    oh = (xp.shape[1] - kh) // stride + 1
    ow = (xp.shape[2] - kw) // stride + 1
    y = np.zeros((n, oh, ow, co))
    for i in range(oh):
        for j in range(ow):
            patch = xp[:, i * stride:i * stride + kh, j * stride:j * stride + kw]
            y[:, i, j] = np.tensordot(patch, w, axes=([1, 2, 3], [0, 1, 2]))
    return y


@pytest.mark.parametrize('size,k,s', [(8, 5, 1), (8, 3, 1), (8, 6, 2), (8, 4, 2), (7, 3, 2)])
def test_conv2d_same_padding_against_loops(size, k, s):
    x, w = t(2, size, size, 3), t(k, k, 3, 4, seed=1)
    pads = (O.same_pads(size, k, s), O.same_pads(size, k, s))
    ref = _conv2d_loops(x.numpy(), w.numpy(), s, pads)
    got = O.conv2d_tf(x, w, strides=(s, s), padding='SAME').numpy()
    assert got.shape == ref.shape and np.allclose(got, ref, atol=1e-10)


def test_conv2d_transpose_is_gradient_of_strided_conv():
    # upsample_conv2d's deconv == d/dx of the stride-2 SAME conv with kernel_up (ops.py:698-708)
    k = t(3, 3, 5, 4, seed=1)
    kup = O.upsampled_kernel(k)                        # [6,6,co=4,ci=5]
    x = t(2, 4, 4, 5)
    y = O.upsample_conv2d(x, k, torch.zeros(4, dtype=D))
    big = t(2, 8, 8, 4, seed=3).requires_grad_(True)
    fwd = O.conv2d_tf(big, kup, strides=(2, 2), padding='SAME')    # maps 8x8x4 -> 4x4x5
    (g,) = torch.autograd.grad(fwd, big, x)
    assert np.allclose(g.numpy(), y.numpy(), atol=1e-10)


def test_instance_norm_invariants():
    x = t(3, 6, 5, 4) * 3 + 2
    g, b = t(4, seed=1), t(4, seed=2)
    y = O.instance_norm(x, g, b)
    m = y.mean(dim=(1, 2))
    v = ((y - m[:, None, None]) ** 2).mean(dim=(1, 2))
    xv = ((x - x.mean(dim=(1, 2), keepdim=True)) ** 2).mean(dim=(1, 2))
    assert np.allclose(m.numpy(), b.expand(3, 4).numpy(), atol=1e-10)
    assert np.allclose(v.numpy(), (g ** 2 * xv / (xv + 1e-6)).numpy(), atol=1e-8)


def test_cdna_against_loops_and_fixed_point():
    img = torch.rand(2, 6, 7, 3, dtype=D)
    raw = t(2, 5, 5, 4, seed=1) * 0.3
    k = torch.relu(raw + torch.tensor(O.identity_kernel((5, 5)))[None, :, :, None] - 1e-12) + 1e-12
    k = k / k.sum(dim=(1, 2), keepdim=True)
    outs = O.apply_cdna_kernels(img, k)
    pad = np.pad(img.numpy(), ((0, 0), (2, 2), (2, 2), (0, 0)), mode='symmetric')
    for kk in range(4):
        ref = np.zeros((2, 6, 7, 3))
        for i in range(5):
            for j in range(5):
                ref += pad[:, i:i + 6, j:j + 7] * k[:, i, j, kk].numpy()[:, None, None, None]
        assert np.allclose(outs[kk].numpy(), ref, atol=1e-12)
    const = torch.full((2, 6, 7, 3), 0.37, dtype=D)
    for o in O.apply_cdna_kernels(const, k):
        assert np.allclose(o.numpy(), 0.37, atol=1e-12)
    assert O.identity_kernel((5, 5))[2, 2] == 1.0 and O.identity_kernel((5, 5)).sum() == 1.0
    assert np.allclose(O.identity_kernel((4, 4))[1:3, 1:3], 0.25)


def test_image_warp_zero_flow_is_identity_and_integer_shift():
    im = torch.rand(2, 5, 6, 3, dtype=D)
    assert np.allclose(O.image_warp(im, torch.zeros(2, 5, 6, 2, dtype=D)).numpy(), im.numpy())
    flow = torch.zeros(2, 5, 6, 2, dtype=D)
    flow[..., 0] = 1.0     # x + 1, clipped at the border
    w = O.image_warp(im, flow)
    assert np.allclose(w[:, :, :-1].numpy(), im[:, :, 1:].numpy())
    assert np.allclose(w[:, :, -1].numpy(), im[:, :, -1].numpy())


def test_loss_zero_points_and_schedules():
    z = torch.zeros(3, 2, 8, dtype=D)
    assert float(O.kl_loss(z, z)) == 0.0
    assert float(O.gan_loss(torch.ones(4, 1), 1.0, 'LSGAN')) == 0.0
    x = t(2, 3, 4, 5)
    assert abs(float(O.cosine_distance(x, x))) < 1e-12
    hp = O.make_hparams(context_frames=2, sequence_length=12, kl_weight=1.0)
    assert O.kl_weight(hp, 0) == 0.0 and O.kl_weight(hp, 75000) == 0.5 and O.kl_weight(hp, 200000) == 1.0
    assert O.learning_rate(hp, 0) == hp.lr and abs(O.learning_rate(hp, 250000) - hp.lr / 2) < 1e-12


def test_spectral_norm_matches_power_iteration_definition():
    W, u = t(3, 3, 3, 4, 6), t(1, 6, seed=1)
    Wb, u1 = O.spectral_normed_weight(W, u)
    Wr = W.reshape(-1, 6)
    v = (u @ Wr.t())
    v = v / v.norm()
    un = v @ Wr
    sigma = un.norm()
    assert np.allclose(Wb.numpy(), (W / sigma).numpy(), atol=1e-10)
    assert np.allclose(u1.numpy(), (un / sigma).numpy(), atol=1e-10)


def test_parameter_counts_match_survey():
    hp = O.make_hparams(context_frames=2, sequence_length=12, video_sn_gan_weight=0.1, video_sn_vae_gan_weight=0.1)
    params, trainable = O.init_params(hp, (64, 64, 3))
    g = sum(p.numel() for k, p in params.items() if k.startswith('generator/'))
    d = sum(p.numel() for k, p in params.items() if k.startswith('discriminator/') and trainable[k])
    assert g == 7349993 and d == 10288002        # 7.35 M + 10.29 M
    hp0 = O.make_hparams(context_frames=2, sequence_length=12, nz=0)
    p0, _ = O.init_params(hp0, (64, 64, 3))
    assert sum(p.numel() for p in p0.values()) == 6397177


def _small_case(dtype):
    hp = O.make_hparams(context_frames=2, sequence_length=5, nz=4, ngf=8, nef=8, ndf=8, clip_length=3)
    params, _ = O.init_params(hp, (32, 32, 3), seed=3, dtype=dtype)
    inputs, noise = O.make_synthetic_inputs(hp, 2, (32, 32, 3), seed=3, dtype=dtype)
    with torch.no_grad():
        out = O.generator(O.Vars(params, dtype=dtype), hp, inputs, noise, O.ground_truth_mask(hp, 2))
    return hp, params, inputs, noise, out


def test_generator_fp32_fp64_agree_and_batch_independent():
    _, _, _, _, o32 = _small_case(torch.float32)
    hp, params, inputs, noise, o64 = _small_case(torch.float64)
    assert (o32['gen_images'].double() - o64['gen_images']).abs().max() < 1e-4
    assert o64['masks'].sum(dim=-1).sub(1).abs().max() < 1e-12                       # masks sum to one
    assert o64['gen_images'].min() >= 0 and o64['gen_images'].max() <= 1              # convex combination
    perm = torch.tensor([1, 0])
    inp2 = {k: v[:, perm] for k, v in inputs.items()}
    noi2 = {k: v[:, perm] for k, v in noise.items() if torch.is_tensor(v)}
    with torch.no_grad():
        o2 = O.generator(O.Vars(params, dtype=torch.float64), hp, inp2, noi2, O.ground_truth_mask(hp, 2))
    assert (o2['gen_images'][:, perm] - o64['gen_images']).abs().max() < 1e-12       # DP correctness


def test_golden_vectors():
    path = os.path.join(GOLD, 'oracle_small.npz')
    if not os.path.exists(path):
        pytest.skip('golden file missing (run tests/golden/make_golden.py)')
    g = np.load(path)
    _, _, _, _, out = _small_case(torch.float32)
    assert np.allclose(out['gen_images'].numpy(), g['gen_images'], atol=1e-5)
    assert np.allclose(out['zs_mu_enc'].numpy(), g['zs_mu_enc'], atol=1e-5)


def test_train_step_runs_and_decreases_nothing_weird():
    hp = O.make_hparams(context_frames=2, sequence_length=6, nz=4, ngf=8, nef=8, ndf=8, clip_length=4, l1_weight=100.0,
                        kl_weight=1.0, kl_anneal_steps=(0, 10), video_sn_gan_weight=0.1, video_sn_vae_gan_weight=0.1,
                        vae_gan_feature_cdist_weight=10.0, lr=2e-4, beta1=0.5)
    params, _ = O.init_params(hp, (32, 32, 3), seed=1)
    inputs, noise = O.make_synthetic_inputs(hp, 2, (32, 32, 3), seed=1)
    opt = dict(m={k: torch.zeros_like(v) for k, v in params.items()}, v={k: torch.zeros_like(v) for k, v in params.items()}, t=0)
    res = O.train_step(params, opt, hp, inputs, noise, step=5)
    assert np.isfinite(res['g_loss']) and np.isfinite(res['d_loss'])
    assert all(g is not None for g in res['g_grads'].values())
    # u is replaced by u', trainable D weights moved by at most lr (Adam's first step is +-lr)
    k = 'discriminator/video/sn_conv0_0/conv3d/kernel'
    assert (res['params'][k] - params[k]).abs().max() <= hp.lr * 1.001
    assert not torch.equal(res['params']['discriminator/video/sn_conv0_0/conv3d/u'], params['discriminator/video/sn_conv0_0/conv3d/u'])
