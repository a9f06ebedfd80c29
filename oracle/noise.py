"""Counter-based Exp(1) race noise: the shared noise model for sampling parity.

The reference draws every choice with ``torch.multinomial`` (models/beam.py:39-48), which is
``topk(p / q, k)`` with ``q ~ Exp(1)`` iid per element.  ATen's
Philox bookkeeping cannot be reproduced once generation is batched, so both sides use
``q = float32(-log(u))``, ``u = (m + 0.5) * 2^-24`` (float64), ``m`` = 24 hash bits keyed on
``(seed, image, step, call, row, column)``.  ``call``: 0 token draw, 1 beam pruning, 2 final pick.
Modes: ``deterministic`` (q == 1, classical top-k / beam), ``injected`` (this generator).
"""
import numpy as np
import torch

from deephumor_b200.utils import synth

CALL_TOKEN, CALL_PRUNE, CALL_FINAL = 0, 1, 2
NOISE_STREAM = 0x4E5A  # 'NZ'


def row_key(seed, image, step, call, row):
    return synth.key(seed, NOISE_STREAM, image, step * 4 + call, row)


def exp_noise(seed, image, step, call, row, ncols):
    """float32[ncols] of Exp(1) variates for one row."""
    m = synth.bits24(row_key(seed, image, step, call, row), ncols).astype(np.float64)
    u = (m + 0.5) * (1.0 / 16777216.0)
    return torch.from_numpy((-np.log(u)).astype(np.float32))


class Noise:
    """mode in {'deterministic', 'injected'}; q(...) returns None for deterministic."""

    def __init__(self, mode='deterministic', seed=0):
        assert mode in ('deterministic', 'injected')
        self.mode, self.seed = mode, seed

    def q(self, image, step, call, nrows, ncols):
        if self.mode == 'deterministic':
            return None
        return torch.stack([exp_noise(self.seed, image, step, call, r, ncols) for r in range(nrows)])
