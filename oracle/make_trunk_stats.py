"""Calibrate pooled-trunk statistics for the synthetic weight sets (run once in the build container).

Writes ``deephumor_b200/utils/trunk_stats.npz`` with, per weight seed, the mean and variance over 32
synthetic images of the 2048 pooled ResNet-50 features.  ``synth_weights.image_encoder`` centres the
head BatchNorm1d on them.  Uses the oracle trunk (not the reference); values only shape the *inputs*.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deephumor_b200.utils import synth, synth_weights  # noqa: E402
from oracle import model  # noqa: E402

out = {}
for seed in (0, 1):
    sd = {}
    synth_weights.resnet_trunk(sd, seed, 'encoder.resnet')
    with torch.no_grad():
        feats = torch.cat([model.resnet50_trunk(sd, 'encoder.resnet', synth.images(1234, i, 8)).mean(dim=(2, 3))
                           for i in range(0, 32, 8)])
    out[f'mean_{seed}'] = feats.mean(0).numpy().astype(np.float32)
    out[f'var_{seed}'] = feats.var(0).numpy().astype(np.float32)
    print(seed, feats.mean().item(), feats.std(0).mean().item())
np.savez(os.path.join(os.path.dirname(synth_weights.__file__), 'trunk_stats.npz'), **out)
