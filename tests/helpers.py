"""Shared helpers for parity tests (tolerances per BASELINE.json north_star / SURVEY.md Appendix D.5)."""
import os

import torch

from deephumor_b200.utils import synth, synth_weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KINDS = synth_weights.KINDS
TOL_FP32 = 1e-4      # fp32 check mode: ||d|| / ||ref|| per tensor
TOL_BF16 = 1e-2      # bf16 mode
NEAR_TIE = 3e-5      # decisions whose oracle margin is below this are reported, not asserted


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load_fixture(tag, kind):
    return torch.load(os.path.join(GOLDEN, f'{tag}_{kind}.pt'), weights_only=False)


def fixture_inputs(fx):
    """Recreate the inputs / weights the fixture was generated with (hash-based, machine independent)."""
    kind, hp, V = fx['kind'], fx['hp'], fx['V']
    sd = synth_weights.make_state_dict(kind, hp, seed=fx['wseed'])
    imgs = synth.images(0, 0, fx['n_img'])
    labs = synth.labels(0, 0, fx['n_img'], V) if kind == 'lstm_labels' else None
    caps, lens = synth.captions(0, 0, fx['n_img'], V, width=fx['max_len'], min_len=4)
    return sd, imgs, labs, caps, lens


def compare_ids(ids, lens, g, what):
    """Token parity with the near-tie policy: rows whose margin < NEAR_TIE may differ (and are listed)."""
    bad, near = [], []
    for n in range(ids.shape[0]):
        same = bool((ids[n].cpu() == g['ids'][n]).all()) and int(lens[n]) == int(g['lengths'][n])
        if not same:
            (near if float(g['gaps'][n]) < NEAR_TIE else bad).append(n)
    assert not bad, f'{what}: token mismatch on images {bad} (near-tie-excused: {near})'
    return near
