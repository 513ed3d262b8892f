"""A SECOND, independent restatement of the composite generator path (SAVPCell.call, models/savp_model.py:393-646) against which
the oracle is pinned: plain numpy, float64, written from the TF semantics of SURVEY.md appendix A with formulations that
differ from the oracle's on purpose --
  * convolutions as sums of shifted slices (no F.conv2d),
  * conv_pool2d as a stride-1 SAME convolution followed by 2x2 average pooling (the identity of the reference's docstring,
    ops.py:799-817) instead of a stride-2 convolution with the pooled kernel,
  * upsample_conv2d as a scatter (every input pixel adds its 6x6 bilinear (x) kernel patch; conv2d_transpose, ops.py:584)
    instead of F.conv_transpose2d,
  * CDNA as np.pad(mode='symmetric') + 25 shifted multiply-adds per sample.
Two unrolled steps on a 32x32 image with one context frame, so the second step consumes the first step's OUTPUT (the
recurrence through gen_image and the ConvLSTM states is covered, not only one cell evaluation)."""
import numpy as np
import torch

from oracle import savp_oracle as O


def same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv_same(x, w, stride=1):
    B, H, W, _ = x.shape
    kh, kw, _, co = w.shape
    (pt, pb), (pl, pr) = same_pads(H, kh, stride), same_pads(W, kw, stride)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    Ho, Wo = -(-H // stride), -(-W // stride)
    out = np.zeros((B, Ho, Wo, co))
    for i in range(kh):
        for j in range(kw):
            out += xp[:, i:i + (Ho - 1) * stride + 1:stride, j:j + (Wo - 1) * stride + 1:stride, :] @ w[i, j]
    return out


def conv_pool2d(x, w, b):
    y = conv_same(x, w) + b
    B, H, W, C = y.shape
    return y.reshape(B, H // 2, 2, W // 2, 2, C).mean(axis=(2, 4))


def upsample_conv2d(x, w, b):
    kh, kw, ci, co = w.shape
    b1 = np.array([0.25, 0.75, 0.75, 0.25])
    B2 = np.pad(np.outer(b1, b1), kh - 1)                                   # FULL correlation: zero pad by k - 1
    P = B2.shape[0] - kh + 1                                                 # 6
    kup = np.zeros((P, P, ci, co))
    for p in range(P):
        for q in range(P):
            for i in range(kh):
                for j in range(kw):
                    kup[p, q] += B2[p + i, q + j] * w[i, j]
    Bn, H, W, _ = x.shape
    pad = same_pads(2 * H, P, 2)[0]                                          # of the stride-2 SAME conv this is the transpose of
    out = np.zeros((Bn, 2 * H + P, 2 * W + P, co))
    contrib = np.einsum('bhwi,pqio->bhwpqo', x, kup)
    for y in range(H):
        for xx in range(W):
            out[:, 2 * y:2 * y + P, 2 * xx:2 * xx + P] += contrib[:, y, xx]
    return out[:, pad:pad + 2 * H, pad:pad + 2 * W] + b


def inorm(x, g, b):
    m = x.mean(axis=(1, 2), keepdims=True)
    v = ((x - m) ** 2).mean(axis=(1, 2), keepdims=True)
    return (x - m) / np.sqrt(v + 1e-6) * g + b


def sigm(x):
    return 1.0 / (1.0 + np.exp(-x))


def conv_lstm(P, scope, x, state, F_):
    c, h = state
    gates = inorm(conv_same(np.concatenate([x, h], -1), P[scope + '/kernel']),
                  P[scope + '/input_transform_forget_output/gamma'], P[scope + '/input_transform_forget_output/beta'])
    i, j, f, o = (gates[..., k * F_:(k + 1) * F_] for k in range(4))
    c2 = inorm(c * sigm(f + 1.0) + sigm(i) * np.tanh(j), P[scope + '/state/gamma'], P[scope + '/state/beta'])
    h2 = np.tanh(c2) * sigm(o)
    return h2, (c2, h2)


def cdna(image, kernels):
    B, H, W, C = image.shape
    _, kh, kw, K = kernels.shape
    xp = np.pad(image, ((0, 0), same_pads(H, kh, 1), same_pads(W, kw, 1), (0, 0)), mode='symmetric')
    outs = []
    for k in range(K):
        o = np.zeros_like(image)
        for i in range(kh):
            for j in range(kw):
                o += xp[:, i:i + H, j:j + W, :] * kernels[:, i, j, k][:, None, None, None]
        outs.append(o)
    return outs


def cell_step(P, hp, sc, image_in, first, states, use_gt):
    image = image_in if use_gt else states['gen']
    enc, dec = O.layer_specs(hp, image.shape[1], image.shape[2])
    layers, new_rnn, ri = [], [], 0
    for i, (oc, use) in enumerate(enc):
        h = np.concatenate([image, first], -1) if i == 0 else layers[-1][-1]
        h = np.maximum(inorm(conv_pool2d(h, P['%s/h%d/conv_pool2d/kernel' % (sc, i)], P['%s/h%d/conv_pool2d/bias' % (sc, i)]),
                             P['%s/h%d/InstanceNorm/gamma' % (sc, i)], P['%s/h%d/InstanceNorm/beta' % (sc, i)]), 0)
        if use:
            rh, st = conv_lstm(P, '%s/lstm_h%d/basic_conv2dlstm_cell' % (sc, i), h, states['rnn'][ri], oc)
            ri += 1
            new_rnn.append(st)
            layers.append((h, rh))
        else:
            layers.append((h,))
    n_enc = len(layers)
    for i, (oc, use) in enumerate(dec):
        li = len(layers)
        h = layers[-1][-1] if i == 0 else np.concatenate([layers[-1][-1], layers[n_enc - i - 1][-1]], -1)
        h = np.maximum(inorm(upsample_conv2d(h, P['%s/h%d/upsample_conv2d/kernel' % (sc, li)], P['%s/h%d/upsample_conv2d/bias' % (sc, li)]),
                             P['%s/h%d/InstanceNorm/gamma' % (sc, li)], P['%s/h%d/InstanceNorm/beta' % (sc, li)]), 0)
        if use:
            rh, st = conv_lstm(P, '%s/lstm_h%d/basic_conv2dlstm_cell' % (sc, li), h, states['rnn'][ri], oc)
            ri += 1
            new_rnn.append(st)
            layers.append((h, rh))
        else:
            layers.append((h,))
    nl, top = len(layers), layers[-1][-1]

    def head(name, x, norm):
        y = conv_same(x, P['%s/%s/conv2d/kernel' % (sc, name)]) + P['%s/%s/conv2d/bias' % (sc, name)]
        return np.maximum(inorm(y, P['%s/%s/InstanceNorm/gamma' % (sc, name)], P['%s/%s/InstanceNorm/beta' % (sc, name)]), 0) if norm else y
    kh, kw = hp.kernel_size
    nk = hp.last_frames * hp.num_transformed_images
    flat = layers[n_enc - 1][-1].reshape(image.shape[0], -1)
    kern = (flat @ P[sc + '/cdna_kernels/dense/kernel'] + P[sc + '/cdna_kernels/dense/bias']).reshape(-1, kh, kw, nk)
    ident = np.zeros((kh, kw))
    ident[kh // 2, kw // 2] = 1.0                                            # odd kernel: the centre tap
    kern = np.maximum(kern + ident[None, :, :, None] - 1e-12, 0) + 1e-12
    kern = kern / kern.sum(axis=(1, 2), keepdims=True)
    scratch = sigm(head('scratch_image', head('h%d_scratch' % nl, top, True), False))
    transformed = cdna(image, kern) + [image, first, scratch]
    logits = head('masks', np.concatenate([head('h%d_masks' % nl, top, True)] + transformed, -1), False)
    e = np.exp(logits - logits.max(-1, keepdims=True))
    masks = e / e.sum(-1, keepdims=True)
    gen = sum(t * masks[..., k:k + 1] for k, t in enumerate(transformed))
    return gen, dict(gen=gen, rnn=new_rnn), masks


def test_two_unrolled_steps_match_the_oracle():
    hk = dict(context_frames=1, sequence_length=3, nz=0, ngf=4)
    hp = O.make_hparams(**hk)
    B, H, W, C = 2, 32, 32, 3
    params, _ = O.init_params(hp, (H, W, C), seed=3, dtype=torch.float64)
    g = torch.Generator().manual_seed(5)
    params = {k: (v + 0.05 * torch.randn(v.shape, generator=g, dtype=torch.float64) if ('bias' in k or 'beta' in k) else v) for k, v in params.items()}
    inputs, _ = O.make_synthetic_inputs(hp, B, (H, W, C), seed=1, dtype=torch.float64)
    gt = O.ground_truth_mask(hp, B)
    with torch.no_grad():
        ref = O.generator_given_z(O.Vars(params), hp, {'images': inputs['images']}, gt)
    P = {k: v.numpy() for k, v in params.items()}
    images = inputs['images'].numpy()
    sc = 'generator/rnn/savp_cell'
    enc, dec = O.layer_specs(hp, H, W)
    rnn, hh = [], H
    for oc, use in enc:
        hh //= 2
        if use:
            rnn.append((np.zeros((B, hh, hh, oc)), np.zeros((B, hh, hh, oc))))
    for oc, use in dec:
        hh *= 2
        if use:
            rnn.append((np.zeros((B, hh, hh, oc)), np.zeros((B, hh, hh, oc))))
    states = dict(gen=np.zeros((B, H, W, C)), rnn=rnn)
    for t in range(2):
        gen, states, masks = cell_step(P, hp, sc, images[t], images[0], states, use_gt=bool(gt[t][0]))
        err = np.abs(gen - ref['gen_images'][t].numpy()).max()
        assert err < 1e-9, 'step %d: max-abs %g' % (t, err)
        merr = np.abs(masks - ref['masks'][t].numpy()[..., 0, :]).max()
        assert merr < 1e-9, 'step %d masks: %g' % (t, merr)
    assert not bool(gt[1][0])                                                # the second step consumed the first step's output


# ---------------------------------------------------------------------------------------------------------------------------
# video discriminator (networks.video_sn_discriminator, networks.py:72-108): zero pad 1 on T, H, W, VALID conv3d, leaky relu
# 0.1, spectrally normalised kernels (one power iteration from the stored u, ops.py:1020-1049), dense logit; LSGAN loss.
def conv3d_valid(x, w, strides):
    kd, kh, kw, _, co = w.shape
    sd, sh, sw = strides
    B, D, H, W, _ = x.shape
    Do, Ho, Wo = (D - kd) // sd + 1, (H - kh) // sh + 1, (W - kw) // sw + 1
    out = np.zeros((B, Do, Ho, Wo, co))
    for a in range(kd):
        for i in range(kh):
            for j in range(kw):
                out += x[:, a:a + (Do - 1) * sd + 1:sd, i:i + (Ho - 1) * sh + 1:sh, j:j + (Wo - 1) * sw + 1:sw, :] @ w[a, i, j]
    return out


def sn_weight(W, u):
    Wr = W.reshape(-1, W.shape[-1])
    v = u @ Wr.T
    v = v / (np.linalg.norm(v) + 1e-12)
    u1 = v @ Wr
    u1 = u1 / (np.linalg.norm(u1) + 1e-12)
    return W / (v @ Wr @ u1.T).item(), u1


def test_video_discriminator_and_lsgan_loss_match_the_oracle():
    ndf, scope = 4, 'discriminator/video/encoder'
    T, B, H, W, C = 5, 2, 16, 16, 3
    clips = torch.rand(T, B, H, W, C, dtype=torch.float64, generator=torch.Generator().manual_seed(11))
    V = O.Vars(rng=np.random.RandomState(7), dtype=torch.float64)          # variables created by the oracle call (reference initialisers)
    u_out = {}
    with torch.no_grad():
        O.video_sn_discriminator(V, scope, clips, ndf)                      # creates the variables
        for k in V.params:                                                  # non-trivial biases
            if k.endswith('/bias'):
                V.params[k] = V.params[k] + 0.1 * torch.randn(V.params[k].shape, dtype=torch.float64, generator=torch.Generator().manual_seed(len(k)))
        feats, logits = O.video_sn_discriminator(V, scope, clips, ndf, u_out=u_out)
    P = {k: v.detach().numpy() for k, v in V.params.items()}
    x = clips.numpy().transpose(1, 0, 2, 3, 4)
    for li, (name, mult, k, strides) in enumerate(O.VIDEO_D_LAYERS):
        Wb, u1 = sn_weight(P['%s/%s/conv3d/kernel' % (scope, name)], P['%s/%s/conv3d/u' % (scope, name)])
        assert np.abs(u1 - u_out['%s/%s/conv3d/u' % (scope, name)].numpy()).max() < 1e-12
        y = conv3d_valid(np.pad(x, ((0, 0), (1, 1), (1, 1), (1, 1), (0, 0))), Wb, strides) + P['%s/%s/conv3d/bias' % (scope, name)]
        x = np.maximum(0.1 * y, y)
        assert x.shape == tuple(feats[li].shape) and np.abs(x - feats[li].numpy()).max() < 1e-9, name
    Wb, _ = sn_weight(P['%s/sn_fc4/dense/kernel' % scope], P['%s/sn_fc4/dense/u' % scope])
    lg = x.reshape(B, -1) @ Wb + P['%s/sn_fc4/dense/bias' % scope]
    assert np.abs(lg - logits.numpy()).max() < 1e-9
    # losses.gan_loss LSGAN (losses.py:29-54): mean (logit - label)^2
    assert abs(float(O.gan_loss(logits, 1.0, 'LSGAN')) - float(((lg - 1.0) ** 2).mean())) < 1e-12
