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
    for d in model.gl:
        if d['use']:
            li = d['li']
            d['rconv'].fwd(model.Bf['rin%d' % li][3], model.Bf['gpre%d' % li][3])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
