"""Harness for running the UNMODIFIED reference from /root/reference in the build container.

Used only by ``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent, e.g. on the GPU box).  Nothing in the reference is modified:
 * ``torchvision.models.resnet50`` is wrapped to force ``weights=None`` (reference hard-codes
   ``pretrained=True`` -> download, models/encoders.py:34);
 * ``torch.multinomial`` (looked up at call time in models/beam.py:46) is patched inside a context
   manager with the shared noise model of ``oracle/noise.py`` (Appendix D.3).
"""
import contextlib
import os
import sys

import torch

from . import noise as _noise

REF_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'deephumor', 'models'))


def import_reference():
    """Returns the ``deephumor.models`` package of the reference, with the offline shim applied."""
    sys.dont_write_bytecode = True
    import torchvision
    if not getattr(torchvision.models.resnet50, '_dh_offline_shim', False):
        orig = torchvision.models.resnet50

        def resnet50(pretrained=False, **kw):
            return orig(weights=None, **kw)
        resnet50._dh_offline_shim = True
        torchvision.models.resnet50 = resnet50
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import deephumor.models as ref_models
    return ref_models


REF_CLASS = {'lstm': 'CaptioningLSTM', 'lstm_labels': 'CaptioningLSTMWithLabels',
             'xfmr_base': 'CaptioningTransformerBase', 'xfmr': 'CaptioningTransformer'}


def build_reference(kind, hp, sd):
    """Instantiate the reference class with ctor kwargs hp and load sd with strict=True (pins Appendix C)."""
    rm = import_reference()
    from deephumor.models import caption_models
    cls = getattr(caption_models, REF_CLASS[kind])
    model = cls(**hp)
    model.load_state_dict(sd, strict=True)
    return model.eval()


@contextlib.contextmanager
def patched_multinomial(mode, seed, image_index, p0, max_len, beam_size):
    """Replace torch.multinomial for ONE reference generate() call.

    The call schedule of a generate() is fixed (models/rnn_models.py:87-140, transformers.py:532-576):
    2-D input = token draw (first at step p0, then step+1 each time); 1-D input with k == beam = beam
    pruning at the current step; 1-D input with k == 1 (beam > 1) = final pick, keyed at step max_len+1.
    """
    orig = torch.multinomial
    state = {'step': None}

    def fake(p, k, *a, **kw):
        if torch.isnan(p).any() or (p.sum(-1) <= 0).any():       # what the real multinomial does (Q3)
            raise RuntimeError('invalid multinomial distribution (sum of probabilities <= 0)')
        if p.dim() == 2:
            state['step'] = p0 if state['step'] is None else state['step'] + 1
            call, step = _noise.CALL_TOKEN, state['step']
        elif k == 1 and beam_size > 1:
            call, step = _noise.CALL_FINAL, max_len + 1
        else:
            call, step = _noise.CALL_PRUNE, state['step']
        if mode == 'deterministic':
            score = p
        else:
            p2 = p if p.dim() == 2 else p.unsqueeze(0)
            q = torch.stack([_noise.exp_noise(seed, image_index, step, call, r, p2.shape[1])
                             for r in range(p2.shape[0])])
            score = p / (q if p.dim() == 2 else q[0])
        return torch.sort(score, dim=-1, descending=True, stable=True).indices[..., :k]

    torch.multinomial = fake
    try:
        yield
    finally:
        torch.multinomial = orig


def reference_generate_batch(model, kind, images, labels=None, first_index=0, mode='deterministic', seed=0,
                             caption=None, max_len=25, pad_index=0, **kw):
    N = images.shape[0]
    ids = torch.full((N, max_len), pad_index, dtype=torch.int64)
    lens = torch.zeros(N, dtype=torch.int64)
    p0 = 0 if caption is None else caption.shape[1]
    with torch.no_grad():
        for n in range(N):
            args = dict(image=images[n:n + 1], caption=caption, max_len=max_len, **kw)
            if kind == 'lstm_labels':
                args['label'] = labels[n:n + 1]
            with patched_multinomial(mode, seed, first_index + n, p0, max_len, kw.get('beam_size', 10)):
                s = model.generate(**args).reshape(-1)
            ids[n, :len(s)] = s
            lens[n] = len(s)
    return ids, lens
