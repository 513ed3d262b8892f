"""GPU probe: generator forward of the CUDA path vs the CPU oracle (fp32) on identical params/inputs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import savp_oracle as O
from video_prediction_b200.models import SAVPVideoPredictionModel


def run(nz, B, T=12, ctx=2, HW=64, C=3, A=0, debug=True):
    hk = dict(context_frames=ctx, sequence_length=T, nz=nz)
    hp = O.make_hparams(**hk)
    params, _ = O.init_params(hp, (HW, HW, C), action_dim=A, seed=0)
    inputs, noise = O.make_synthetic_inputs(hp, B, (HW, HW, C), action_dim=A)
    V = O.Vars(params)
    taps = []
    t0 = time.time()
    with torch.no_grad():
        ref = O.generator(V, hp, inputs, noise, O.ground_truth_mask(hp, B), taps=taps)
    t_cpu = time.time() - t0
    model = SAVPVideoPredictionModel(mode='test', hparams_dict=hk)
    model.set_params({k: v for k, v in params.items() if k.startswith('generator/')})
    binp = {'images': inputs['images'].permute(1, 0, 2, 3, 4)}
    if A:
        binp['actions'] = inputs['actions'].permute(1, 0, 2)
    model.build_graph(binp)
    model.set_inputs(binp, noise)
    torch.cuda.synchronize()
    t0 = time.time()
    model.generator_forward()
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    out = model.outputs

    def cmp(name, a, b):
        a = a.detach().cpu().float()
        b = b.detach().cpu().float()
        print('  %-28s max_abs %.3e (ref max %.3e)' % (name, (a - b).abs().max().item(), b.abs().max().item()))
        return (a - b).abs().max().item()
    print('== nz=%d B=%d A=%d: oracle CPU %.2fs, CUDA path (eager launches) %.3fs' % (nz, B, A, t_cpu, t_gpu))
    Bf = model.Bf
    if debug:
        tp = taps[0]
        off = B if nz else 0   # taps are from the prior unroll (second half of the batch)
        cmp('t0 image', Bf['img'][0][off:off + B, ..., :C], tp['image'])
        for li, lay in enumerate(tp['layers']):
            d = model.gl[li]
            if d['use']:
                sp = d['rin_spec']
                cmp('t0 h%d (IN+relu)' % li, Bf['rin%d' % li][0][off:off + B, ..., sp.off('x'):sp.off('x') + d['oc']], lay[0])
                cmp('t0 lstm_h%d out' % li, Bf['rin%d' % li][1][off:off + B, ..., sp.off('h'):sp.off('h') + d['oc']], lay[1])
            else:
                cmp('t0 h%d (IN+relu)' % li, Bf['out%d' % li][0][off:off + B], lay[0])
        cmp('t0 cdna kernels', Bf['kern'][0][off:off + B].view(B, 5, 5, 4), tp['kernels'])
        msp = model.mk_spec
        so = msp.off('l6')
        cmp('t0 scratch', Bf['mk'][0][off:off + B, ..., so:so + C], tp['scratch'])
        cmp('t0 mask logits', Bf['mlog'][0][off:off + B, ..., :7], tp['mask_logits'])
    e = cmp('gen_images', out['gen_images'], ref['gen_images'].permute(1, 0, 2, 3, 4))
    if nz:
        e = max(e, cmp('gen_images_enc', out['gen_images_enc'], ref['gen_images_enc'].permute(1, 0, 2, 3, 4)))
        cmp('zs_mu_enc', out['zs_mu_enc'], ref['zs_mu_enc'].permute(1, 0, 2))
        cmp('zs_log_sigma_sq_enc', out['zs_log_sigma_sq_enc'], ref['zs_log_sigma_sq_enc'].permute(1, 0, 2))
    cmp('masks', out['masks'], ref['masks'].permute(1, 0, 2, 3, 4, 5))
    cmp('transformed_images', out['transformed_images'], ref['transformed_images'].permute(1, 0, 2, 3, 4, 5))
    print('VERDICT nz=%d %s (gen_images max_abs %.3e, tolerance 1e-3)' % (nz, 'PASS' if e <= 1e-3 else 'FAIL', e))


if __name__ == '__main__':
    run(0, 2)
    run(8, 2)
    run(8, 2, A=4, debug=False)
