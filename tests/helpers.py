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
BOUND_BF16 = 0.10
BOUND_FP32 = 1e-4


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


def compare_beam_states(trace_dev, traces, beam, bound, val_tol):
    """Step-by-step parity of the device beam state with the oracle's: for every image, every step up to (not including)
    the first one whose oracle margin is below `bound` must show identical beam sequences / ended flags and scores within
    val_tol.  Returns (steps checked, steps available, images whose whole decode was checked)."""
    checked = total = whole = 0
    by_step = {st: (seq.cpu(), val.cpu(), ended.cpu()) for st, seq, val, ended, _ in trace_dev}
    for n, tr in enumerate(traces):
        steps = sorted(tr.states)
        total += len(steps)
        ok_all = True
        for st in steps:
            if tr.step_abs.get(st, float('inf')) < bound:
                ok_all = False
                break
            oseq, oval, oend = tr.states[st]
            seq, val, ended = by_step[st]
            cols = oseq.shape[1]
            mine = seq[n * beam:(n + 1) * beam, :cols].long()
            assert torch.equal(mine, oseq), f'image {n} step {st}: beam sequences differ\n{mine}\n{oseq}'
            assert torch.equal(ended[n * beam:(n + 1) * beam].bool(), oend), f'image {n} step {st}: ended flags differ'
            assert torch.allclose(val[n * beam:(n + 1) * beam], oval, atol=val_tol, rtol=0), \
                f'image {n} step {st}: beam scores differ {val[n * beam:(n + 1) * beam]} vs {oval}'
            checked += 1
        whole += ok_all
    return checked, total, whole
