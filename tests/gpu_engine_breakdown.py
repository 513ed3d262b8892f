"""Per-geometry breakdown of the tensor-core engine over ONE eager cfg2 training step (B = 16): calls, summed CUDA-event time,
algorithmic TFLOP/s and the engine the autotuner chose (0 box / 1 halo / -1 wgrad).  Every call is timed as a CUDA-graph replay
(lib.profile_engine(replay=True)): eager event pairs around kernels of a few microseconds measure the host launch path.  Usage: python tests/gpu_engine_breakdown.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from video_prediction_b200 import lib as L  # noqa: E402
from video_prediction_b200.models import get_model_class  # noqa: E402

os.environ['VP_CONCURRENT_D'] = '0'
model = get_model_class('savp')(mode='train', hparams_dict=dict(bench.SAVP_HPARAMS), num_gpus=1)
batch = bench.synthetic_batch(bench.PER_GPU_BATCH, seed=0)
model.build_graph(batch)
for _ in range(2):
    model.stage_step()
    model._step_device(None)
    model.global_step += 1
torch.cuda.synchronize()
model.stage_step()
L.profile_engine(True, replay=True)
model._step_device(None)
prof = L.profile_engine(False, by_geometry=True)
rows = sorted(prof.items(), key=lambda kv: -kv[1]['ms'])
tot = sum(v['ms'] for v in prof.values())
print('engine total %.2f ms over %d calls' % (tot, sum(v['calls'] for v in prof.values())))
print('%-6s %5s %8s %7s %7s  geometry (in n,d,h,w,c -> out n,d,h,w,c | k | s | transposed | n_pad,kc | extra) engine' % ('kind', 'calls', 'ms', 'share', 'TF/s'))
for (kind, key), v in rows[:60]:
    k, eng = key
    geom = 'in %s out %s k%s s%s t%d np%d kc%d %s' % (k[0:5], k[5:10], k[11:14], k[14:17], k[20], k[21], k[22], k[24])
    print('%-6s %5d %8.3f %6.1f%% %7.1f  %s eng=%d' % (kind, v['calls'], v['ms'], 100 * v['ms'] / tot, v['flops'] / v['ms'] / 1e9, geom, eng))
