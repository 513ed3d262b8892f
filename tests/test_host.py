"""CPU tests of the host layer: hparams surface, registry, constructor errors, parameter specs vs the oracle,
C-ABI exports, and the data-parallel plumbing on gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hparams_surface():
    from video_prediction_b200.hparams import HParams
    hp = HParams(a=1, b=0.5, c='x', d=(1, 2), e=True)
    hp.parse('a=3,b=1e-3,c=hello,d=[5,6],e=false')
    assert hp.a == 3 and hp.b == 1e-3 and hp.c == 'hello' and hp.d == (5, 6) and hp.e is False
    hp.override_from_dict({'a': 7})
    assert hp.values()['a'] == 7
    with pytest.raises(ValueError):
        hp.set_hparam('nope', 1)
    with pytest.raises(ValueError):
        hp.parse('zzz=1')


def test_registry_and_constructor_errors():
    from video_prediction_b200.models import get_model_class, SAVPVideoPredictionModel
    assert get_model_class('savp') is SAVPVideoPredictionModel
    assert get_model_class('SAVPVideoPredictionModel') is SAVPVideoPredictionModel
    with pytest.raises(ValueError, match='Invalid model'):
        get_model_class('bogus')
    with pytest.raises(NotImplementedError):
        get_model_class('sv2p')
    with pytest.raises(ValueError):
        SAVPVideoPredictionModel(mode='train')                      # context_frames / sequence_length missing
    with pytest.raises(ValueError):
        SAVPVideoPredictionModel(mode='bad', hparams_dict=dict(context_frames=2, sequence_length=12))
    with pytest.raises(ValueError):
        SAVPVideoPredictionModel(hparams_dict=dict(context_frames=2, sequence_length=12, not_a_key=1))
    m = SAVPVideoPredictionModel(mode='train', hparams_dict=dict(context_frames=2, sequence_length=12, gan_weight=1.0),
                                 hparams='lr=0.01,kernel_size=[5,5]')           # deprecated key silently dropped
    assert m.hparams.lr == 0.01 and m.hparams.nz == 8 and m.hparams.l1_weight == 1.0 and m.deterministic is False


def test_shipped_hparams_files_parse_and_schedules():
    from video_prediction_b200.models import SAVPVideoPredictionModel
    ours_savp = dict(batch_size=16, lr=0.0002, beta1=0.5, beta2=0.999, l1_weight=100.0, l2_weight=0.0, kl_weight=1.0,
                     video_sn_vae_gan_weight=0.1, video_sn_gan_weight=0.1, vae_gan_feature_cdist_weight=10.0,
                     gan_feature_cdist_weight=0.0, state_weight=0.0)
    m = SAVPVideoPredictionModel(hparams_dict=dict(ours_savp, context_frames=2, sequence_length=12))
    assert m.kl_weight_at(0) == 0.0 and m.kl_weight_at(75000) == 0.5 and m.kl_weight_at(10 ** 6) == 1.0
    assert m.learning_rate_at(0) == 0.0002 and abs(m.learning_rate_at(250000) - 0.0001) < 1e-12
    assert m.learning_rate_at(300000) == 0.0


def test_param_specs_match_oracle_names_and_shapes():
    from oracle import savp_oracle as O
    from video_prediction_b200.models import SAVPVideoPredictionModel
    for nz, A, gan in ((0, 0, 0.0), (8, 0, 0.1), (8, 4, 0.1)):
        hk = dict(context_frames=2, sequence_length=12, nz=nz, video_sn_gan_weight=gan, video_sn_vae_gan_weight=gan if nz else 0.0)
        hp = O.make_hparams(**hk)
        ref, _ = O.init_params(hp, (64, 64, 3), action_dim=A)
        m = SAVPVideoPredictionModel(mode='train', hparams_dict=hk)
        m.B, m.T, m.H, m.W, m.C, m.A = 1, 12, 64, 64, 3, A
        m.S, m.Zc = 11, A + nz
        m.enc_specs, m.dec_specs = m.layer_specs(m.hparams.ngf, 64, 64)
        specs = m._generator_param_specs()
        specs.update(m._discriminator_param_specs())
        assert set(specs) == set(ref), sorted(set(specs) ^ set(ref))[:5]
        for k, (shape, _) in specs.items():
            assert tuple(shape) == tuple(ref[k].shape), k


def test_loss_totals_are_weighted_sums_like_the_reference():
    # base_model.py:461: total = accumulate_n(loss * weight); the per-term dicts stay unweighted.  Checked against the
    # oracle's generator_losses / total_loss on the same (unweighted) values, with the annealed KL weight at step 75000.
    from oracle import savp_oracle as O
    from video_prediction_b200.models import SAVPVideoPredictionModel
    hk = dict(context_frames=2, sequence_length=12, l1_weight=100.0, l2_weight=0.5, kl_weight=1.0, video_sn_gan_weight=0.1,
              video_sn_vae_gan_weight=0.1, vae_gan_feature_cdist_weight=10.0)
    m = SAVPVideoPredictionModel(mode='train', hparams_dict=hk)
    vals = {'gen_l1_loss': 0.02, 'gen_l2_loss': 0.3, 'gen_kl_loss': 0.5, 'gen_video_sn_gan_loss': 0.9,
            'gen_video_sn_vae_gan_loss': 0.8, 'gen_video_sn_vae_gan_feature_cdist_loss': 0.07,
            'gen_video_sn_gan_feature_cdist_loss': 0.0, 'discrim_video_sn_gan_loss': 0.25, 'discrim_video_sn_vae_gan_loss': 0.75}
    w = m.loss_weights(step=75000)
    g, d, gl, dl = m.split_losses(vals, w)
    assert 'gen_video_sn_gan_feature_cdist_loss' not in g          # weight 0 -> the reference never creates the term
    assert g['gen_l1_loss'] == 0.02 and d['discrim_video_sn_gan_loss'] == 0.25
    hp = O.make_hparams(**hk)
    ref_g = 100.0 * 0.02 + 0.5 * 0.3 + O.kl_weight(hp, 75000) * 0.5 + 0.1 * 0.9 + 0.1 * 0.8 + 10.0 * 0.07
    assert abs(gl - ref_g) < 1e-9 and abs(dl - (0.1 * 0.25 + 0.1 * 0.75)) < 1e-12
    assert m.loss_weights(step=0)['gen_kl_loss'] == 0.0            # KL weight is annealed from 0 (base_model.py:312-315)


def test_scheduled_sampling_schedule_and_mask():
    # savp_model.py:309-334: inverse sigmoid k / (k + exp((step - start) / k)), 1.0 before start, all-False below 1e-3,
    # nothing in test mode or with schedule 'none'; linear: 1 - (clip(step) - start) / (end - start)
    import math
    from video_prediction_b200.models import SAVPVideoPredictionModel
    hk = dict(context_frames=2, sequence_length=12)
    m = SAVPVideoPredictionModel(mode='train', hparams_dict=hk)
    assert abs(m.schedule_sampling_prob(0) - 900.0 / 901.0) < 1e-12
    assert abs(m.schedule_sampling_prob(9000) - 900.0 / (900.0 + math.exp(10.0))) < 1e-12
    m.B, m.NB, m.S = 4, 8, 11
    m.global_step = 0
    mask = m.draw_scheduled_sampling()
    assert tuple(mask.shape) == (9, 8) and mask.float().mean() > 0.9
    m.global_step = 20000                                         # prob = 900 / (900 + e^22.2) < 1e-3 -> deterministic
    assert m.draw_scheduled_sampling() is None
    m.global_step = 0
    a, m.rank = m.draw_scheduled_sampling(), 1
    m.hparams.set_hparam('schedule_sampling_k', 1.0)              # prob(0) = 0.5: masks of different ranks must differ
    m.rank = 0
    r0 = m.draw_scheduled_sampling()
    m.rank = 1
    assert not torch.equal(r0, m.draw_scheduled_sampling())
    t = SAVPVideoPredictionModel(mode='test', hparams_dict=hk)
    assert t.schedule_sampling_prob(0) == 0.0
    lin = SAVPVideoPredictionModel(mode='train', hparams_dict=dict(hk, schedule_sampling='linear', schedule_sampling_steps=(100, 300)))
    assert lin.schedule_sampling_prob(0) == 1.0 and lin.schedule_sampling_prob(200) == 0.5 and lin.schedule_sampling_prob(999) == 0.0
    none = SAVPVideoPredictionModel(mode='train', hparams_dict=dict(hk, schedule_sampling='none'))
    assert none.schedule_sampling_prob(0) == 0.0


def test_accepted_but_unbuilt_hparams_are_refused():
    from video_prediction_b200.models import SAVPVideoPredictionModel
    base = dict(context_frames=2, sequence_length=12)
    for bad in (dict(joint_gan_optimization=True), dict(state_weight=1.0), dict(tv_weight=0.1), dict(z_l1_weight=1.0),
                dict(use_same_discriminator=True), dict(dilation_rate=(2, 2)), dict(gan_feature_l2_weight=1.0),
                dict(conv_rnn='gru'), dict(transformation='dna'), dict(learn_prior=True)):
        m = SAVPVideoPredictionModel(mode='train', hparams_dict=dict(base, **bad))
        with pytest.raises(NotImplementedError):
            m._check_supported()
    with pytest.raises(ValueError):
        SAVPVideoPredictionModel(mode='train', hparams_dict=dict(base, gan_loss_type='WGAN'))._check_supported()
    for ok in (dict(gan_loss_type='GAN'), dict(gan_loss_type='SNGAN'), dict(l1_weight=1.0, l2_weight=1.0),
               dict(schedule_sampling='linear')):
        SAVPVideoPredictionModel(mode='train', hparams_dict=dict(base, **ok))._check_supported()


def test_concat_spec_channel_maps():
    from video_prediction_b200.models.savp_model import ConcatSpec
    sp = ConcatSpec([('image', 3), ('first', 3), ('z', 8)])
    assert sp.cstride == 16 and sp.ref_channels == 14
    assert sp.cmap == [0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, 9, 10, 11, 12, 13]
    assert sp.off('first') == 4 and sp.off('z') == 8


def test_c_abi_library_exports_every_declared_symbol():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    path = ge.build()
    header = open(os.path.join(ROOT, 'include', 'vp_b200.h')).read()
    declared = set(re.findall(r'\b(vp_[a-z0-9_]+)\s*\(', header))
    out = subprocess.run(['nm', '-D', '--defined-only', path], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (vp_[a-z0-9_]+)', out))
    assert declared and declared <= exported, sorted(declared - exported)
    import ctypes
    lib = ctypes.CDLL(path)
    assert lib.vp_version() >= 100


def _dp_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from video_prediction_b200 import dp
    r, l, w = dp.init_from_env('gloo')
    flat = torch.full((10,), float(rank + 1))
    dp.broadcast_state([flat])                     # -> all ones (rank 0's value)
    grad = torch.arange(6, dtype=torch.float32) * (rank + 1)
    ar = dp.make_allreduce()
    ar(grad)                                       # sum over ranks
    losses = dp.mean_scalars(torch.tensor([float(rank)]))
    sl = dp.shard_batch(32, r, w)
    q.put((rank, flat.tolist(), grad.tolist(), losses.item(), (sl.start, sl.stop)))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_data_parallel_plumbing_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    for rank, flat, grad, loss, sl in res:
        assert flat == [1.0] * 10
        assert grad == [0.0, 3.0, 6.0, 9.0, 12.0, 15.0]
        assert abs(loss - 0.5) < 1e-6
        assert sl == (rank * 16, rank * 16 + 16)
    from video_prediction_b200 import dp
    with pytest.raises(ValueError):
        dp.shard_batch(10, 0, 4)


def test_no_undefined_global_names_in_the_host_layer():
    """A cheap static check (no GPU needed to hit a NameError on a rarely taken path): every name loaded in the package's
    modules, the scripts and bench.py is a builtin, an import, an argument or assigned somewhere in the same file."""
    import ast
    import builtins
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, 'video_prediction_b200', '**', '*.py'), recursive=True) + \
        glob.glob(os.path.join(root, 'scripts', '*.py')) + [os.path.join(root, 'bench.py'), os.path.join(root, '__graft_entry__.py')]
    bad = []
    for path in files:
        tree = ast.parse(open(path).read())
        defined = set(dir(builtins)) | {'__file__', '__name__', '__doc__'}
        for n in ast.walk(tree):
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                defined.add(n.name)
            elif isinstance(n, ast.Import):
                defined.update((a.asname or a.name).split('.')[0] for a in n.names)
            elif isinstance(n, ast.ImportFrom):
                defined.update(a.asname or a.name for a in n.names)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                defined.add(n.id)
            elif isinstance(n, ast.arg):
                defined.add(n.arg)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                defined.add(n.name)
        bad += ['%s:%d %s' % (os.path.relpath(path, root), n.lineno, n.id) for n in ast.walk(tree)
                if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in defined]
    assert not bad, bad
