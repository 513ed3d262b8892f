"""Profiling harness: builds the cfg2 model, warms up, then brackets ONE eager training step (or N gate-conv
launches) with cudaProfilerStart/Stop.  Use under `ncu --profile-from-start off ...`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from video_prediction_b200.models import SAVPVideoPredictionModel

mode = sys.argv[1] if len(sys.argv) > 1 else 'step'
model = SAVPVideoPredictionModel(mode='train', hparams_dict=dict(bench.SAVP_HPARAMS), num_gpus=1)
batch = bench.synthetic_batch(bench.PER_GPU_BATCH, 0)
model.build_graph(batch)
for _ in range(2):
    model.train_step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if mode == 'step':
    model.train_step(batch)
else:
    from video_prediction_b200 import lib as L
    for d in model.gl:
        if d['use']:
            li = d['li']
            d['rconv'].fwd(model.Bf['rin%d' % li][3], model.Bf['gpre%d' % li][3])
    # the first discriminator layer's own tensor-core kernels (csrc/d0_layer.cu) at the step's shape: 32 clips x 10 frames
    n, dd, h, w = 2 * bench.PER_GPU_BATCH, 10, bench.IMAGE[0], bench.IMAGE[1]
    x = torch.rand(n, dd, h, w, 4, device='cuda')
    wt, b = torch.randn(3, 3, 3, 3, 32, device='cuda') * 0.1, torch.zeros(32, device='cuda')
    sig = torch.ones(1, device='cuda')
    y, gw = torch.empty(n, dd, h, w, 32, device='cuda'), torch.zeros(27 * 3 * 32, device='cuda')
    if L.conv3d_c4_fwd_tc_ok(h, w):
        L.conv3d_c4_fwd_tc(x, wt, sig, b, y, n, dd, h, w, 3, 0.1)
    if w % 64 == 0:
        L.conv3d_c4_wgrad(x, y, gw, n, dd, h, w, 3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
