"""Shared helpers for parity tests (tolerances per BASELINE.json north_star / SURVEY.md Appendix D.5)."""
import os

import torch

from deephumor_b200.utils import synth, synth_weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KINDS = synth_weights.KINDS
TOL_FP32 = 1e-4      # fp32 check mode: ||d|| / ||ref|| per tensor
TOL_BF16 = 1e-2      # bf16 mode
NEAR_TIE = 3e-5      # decisions whose oracle margin is below this are reported, not asserted
# Tensor-core mode: a decision (top-k filter membership, race order of a draw, beam pruning) can legitimately flip when its
# oracle margin, in logit units, is within the path's numeric error.  The north-star budget is 1e-2 relative on logits; the
# canonical classifiers produce logits with RMS ~2, i.e. up to ~2e-2 absolute per logit, and a margin is a difference of two
# of them, accumulated over the steps of a cumulative beam score.  Decisions with an oracle margin above this bound must agree.
# The bound is derived from the measured error of the mode: BOUND_FACTOR x the largest absolute error of the mode's
# teacher-forced logits against the reference fixture (a margin is a difference of two logits / cumulative scores).
BOUND_FACTOR = 4.0
BOUND_FP32 = 2e-3


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load_fixture(tag, kind):
    return torch.load(os.path.join(GOLDEN, f'{tag}_{kind}.pt'), weights_only=False)


def n_fwd(fx):
    """Images whose teacher-forced logits / spatial tokens the fixture stores (all of them in the small fixtures)."""
    return fx.get('n_fwd', fx['n_img'])


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


def oracle_traces(fx, g, sd, imgs, labs, caps, n=None):
    """Live CPU-oracle generation of the fixture's variant g with per-step beam states and absolute margins."""
    from oracle import model as omodel, noise as onoise
    kind = fx['kind']
    n = fx['n_img'] if n is None else n
    with torch.no_grad():
        enc = omodel.encode(kind, sd, imgs[:n], None if labs is None else labs[:n])
        traces = []
        prefix = caps[:1, :g['prefix_len']] if g['prefix_len'] else None
        ids, lens = omodel.generate_batch(kind, sd, fx['hp'], None, None if labs is None else labs[:n], max_len=fx['max_len'],
                                          encoded=enc, traces=traces, caption=prefix, beam_size=g['beam_size'],
                                          top_k=g['top_k'], temperature=g['temperature'],
                                          noise=onoise.Noise(g['mode'], g['noise_seed']))
    return ids, lens, traces


def compare_beam_states(trace_dev, traces, beam, val_tol):
    """Step-by-step comparison of the device beam state with the oracle's.  For every image the states (beam sequences,
    ended flags, cumulative scores within val_tol) are compared after every decode step until the first step at which they
    differ.  Returns one record per image: (first differing step or None, oracle margin of that step's decisions in logit
    units, steps matched, steps available).  Past a divergence the two decodes follow different beams, so nothing later is
    comparable; a divergence is legitimate only at a step whose oracle margin is within the mode's numeric error."""
    by_step = {st: (seq.cpu(), val.cpu(), ended.cpu()) for st, seq, val, ended, _ in trace_dev}
    out = []
    for n, tr in enumerate(traces):
        steps = sorted(tr.states)
        first, matched = None, 0
        for st in steps:
            oseq, oval, oend = tr.states[st]
            seq, val, ended = by_step[st]
            cols = oseq.shape[1]
            same = (torch.equal(seq[n * beam:(n + 1) * beam, :cols].long(), oseq)
                    and torch.equal(ended[n * beam:(n + 1) * beam].bool(), oend)
                    and torch.allclose(val[n * beam:(n + 1) * beam], oval, atol=val_tol, rtol=0))
            if not same:
                first = st
                break
            matched += 1
        out.append((first, tr.step_abs.get(first, float('inf')) if first is not None else None, matched, len(steps)))
    return out
