"""Generates tests/golden/oracle_small.npz from the oracle (the reference itself cannot be imported here:
it needs TensorFlow 1.x, see oracle/savp_oracle.py header).  The vectors pin the oracle against accidental
edits and travel to the GPU box, where the CUDA path is compared with them as well."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import savp_oracle as O  # noqa: E402

hp = O.make_hparams(context_frames=2, sequence_length=5, nz=4, ngf=8, nef=8, ndf=8, clip_length=3)
params, _ = O.init_params(hp, (32, 32, 3), seed=3)
inputs, noise = O.make_synthetic_inputs(hp, 2, (32, 32, 3), seed=3)
with torch.no_grad():
    out = O.generator(O.Vars(params), hp, inputs, noise, O.ground_truth_mask(hp, 2))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_small.npz'),
                    gen_images=out['gen_images'].numpy(), gen_images_enc=out['gen_images_enc'].numpy(),
                    zs_mu_enc=out['zs_mu_enc'].numpy(), images=inputs['images'].numpy())
print('written')
