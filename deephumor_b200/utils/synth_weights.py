"""Synthetic ``state_dict``s in the reference's exact key/shape/dtype layout.

Keys follow ``deephumor/models/caption_models.py:9-461`` (module attribute names),
``deephumor/models/encoders.py:22-44,91-94,124-127``, ``rnn_models.py:18-26``,
``transformers.py:66-80,329-338,413-430,602-607,675-692`` and torchvision's ``resnet50``
(children()[:-2] -> prefixes ``resnet.{0,1,4,5,6,7}``).  ``oracle/make_golden.py`` loads these
into the unmodified reference with ``strict=True``, which pins the layout.

Values come from ``deephumor_b200.utils.synth`` (hash-based, machine independent).  Scales are
chosen so that activations stay O(1) through the trunk (He-uniform convs, small last-BN gamma per
block).  The trunk tensors are keyed by their name *below* ``resnet`` so all four captioners share one
trunk per seed; the head ``BatchNorm1d`` running stats are centred on the calibrated pooled-feature
statistics in ``trunk_stats.npz`` (made by ``oracle/make_trunk_stats.py``) so that embeddings of
different images are not collinear.
"""
import os

import numpy as np
import math

import torch

from . import synth

_STATS = os.path.join(os.path.dirname(__file__), 'trunk_stats.npz')

RESNET_BLOCKS = (3, 4, 6, 3)
RESNET_PLANES = (64, 128, 256, 512)

KINDS = ('lstm', 'lstm_labels', 'xfmr_base', 'xfmr')


def _nm(p):
    """Generator name of a tensor: trunk tensors drop everything before 'resnet'."""
    i = p.find('resnet.')
    return p[i:] if i >= 0 else p


def _bn(sd, seed, p, c, gamma=1.0, gspread=0.2):
    n = _nm(p)
    sd[p + '.weight'] = synth.tensor(seed, n + '.weight', (c,), gspread, center=gamma)
    sd[p + '.bias'] = synth.tensor(seed, n + '.bias', (c,), 0.1)
    sd[p + '.running_mean'] = synth.tensor(seed, n + '.running_mean', (c,), 0.1)
    sd[p + '.running_var'] = synth.tensor(seed, n + '.running_var', (c,), 0.3, center=1.0)
    sd[p + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.int64)


def _conv(sd, seed, p, cout, cin, k):
    a = math.sqrt(6.0 / (cin * k * k))
    sd[p + '.weight'] = synth.tensor(seed, _nm(p) + '.weight', (cout, cin, k, k), a)


def _linear(sd, seed, p, nout, nin, bias=True):
    a = 1.0 / math.sqrt(nin)
    sd[p + '.weight'] = synth.tensor(seed, p + '.weight', (nout, nin), a)
    if bias:
        sd[p + '.bias'] = synth.tensor(seed, p + '.bias', (nout,), a)


def _ln(sd, seed, p, c):
    sd[p + '.weight'] = synth.tensor(seed, p + '.weight', (c,), 0.1, center=1.0)
    sd[p + '.bias'] = synth.tensor(seed, p + '.bias', (c,), 0.1)


def resnet_trunk(sd, seed, p):
    """torchvision ResNet-50 trunk: stem + layer1..4 (resnet.py:197-204, Bottleneck :108-163)."""
    _conv(sd, seed, p + '.0', 64, 3, 7)
    _bn(sd, seed, p + '.1', 64)
    inplanes = 64
    for li, (nblk, planes) in enumerate(zip(RESNET_BLOCKS, RESNET_PLANES)):
        for b in range(nblk):
            q = f'{p}.{4 + li}.{b}'
            _conv(sd, seed, q + '.conv1', planes, inplanes, 1)
            _bn(sd, seed, q + '.bn1', planes)
            _conv(sd, seed, q + '.conv2', planes, planes, 3)
            _bn(sd, seed, q + '.bn2', planes)
            _conv(sd, seed, q + '.conv3', planes * 4, planes, 1)
            _bn(sd, seed, q + '.bn3', planes * 4, gamma=0.35, gspread=0.1)
            if b == 0:
                _conv(sd, seed, q + '.downsample.0', planes * 4, inplanes, 1)
                _bn(sd, seed, q + '.downsample.1', planes * 4)
            inplanes = planes * 4


def image_encoder(sd, seed, p, emb_dim):
    resnet_trunk(sd, seed, p + '.resnet')
    _linear(sd, seed, p + '.linear', emb_dim, 2048)
    _bn(sd, seed, p + '.bn', emb_dim)
    # pooled trunk features are non-negative with a large common mean; centre BN1d on it.
    if os.path.exists(_STATS):
        st = np.load(_STATS)
        if f'mean_{seed}' in st:
            W = sd[p + '.linear.weight'].double().numpy()
            b = sd[p + '.linear.bias'].double().numpy()
            m, v = st[f'mean_{seed}'].astype(np.float64), st[f'var_{seed}'].astype(np.float64)
            sd[p + '.bn.running_mean'] = torch.from_numpy((W @ m + b).astype(np.float32))
            sd[p + '.bn.running_var'] = torch.from_numpy(((W * W) @ v).astype(np.float32))


def lstm_decoder(sd, seed, p, V, E, H, L, with_embedding=True):
    if with_embedding:
        sd[p + '.embedding.weight'] = synth.tensor(seed, p + '.embedding.weight', (V, E), synth.SQRT3 * 0.5)
    a = 1.0 / math.sqrt(H)
    for l in range(L):
        inp = E if l == 0 else H
        sd[f'{p}.lstm.weight_ih_l{l}'] = synth.tensor(seed, f'{p}.lstm.weight_ih_l{l}', (4 * H, inp), a)
        sd[f'{p}.lstm.weight_hh_l{l}'] = synth.tensor(seed, f'{p}.lstm.weight_hh_l{l}', (4 * H, H), a)
        sd[f'{p}.lstm.bias_ih_l{l}'] = synth.tensor(seed, f'{p}.lstm.bias_ih_l{l}', (4 * H,), a)
        sd[f'{p}.lstm.bias_hh_l{l}'] = synth.tensor(seed, f'{p}.lstm.bias_hh_l{l}', (4 * H,), a)
    _classifier(sd, seed, p + '.classifier', V, H, 12.0, 0.55)


def _classifier(sd, seed, p, V, H, wscale, eos_bias):
    # wider than torch's default so that top-1/top-2 logit gaps are comfortably above fp32 noise;
    # eos_bias lifts <eos> (id 3) into the sampled range so ended-beam logic (Q6, Q8) is exercised
    sd[p + '.weight'] = synth.tensor(seed, p + '.weight', (V, H), wscale / math.sqrt(H))
    b = synth.tensor(seed, p + '.bias', (V,), 1.0 / math.sqrt(H))
    b[3] += eos_bias
    sd[p + '.bias'] = b


def _mha(sd, seed, p, D, n_heads):
    sd[p + '.scale'] = torch.sqrt(torch.tensor(D // n_heads, dtype=torch.float32))
    for n in ('fc_q', 'fc_k', 'fc_v', 'fc_o'):
        _linear(sd, seed, f'{p}.{n}', D, D)


def xfmr_decoder(sd, seed, p, V, D, L, n_heads, pf, max_len, cross):
    sd[p + '.tok_embedding.weight'] = synth.tensor(seed, p + '.tok_embedding.weight', (V, D), synth.SQRT3 * 4.0)
    sd[p + '.pos_embedding.weight'] = synth.tensor(seed, p + '.pos_embedding.weight', (max_len, D), synth.SQRT3 * 0.3)
    for l in range(L):
        q = f'{p}.layers.{l}'
        _mha(sd, seed, q + '.self_attn', D, n_heads)
        _ln(sd, seed, q + '.self_attn_ln', D)
        if cross:
            _mha(sd, seed, q + '.enc_attn', D, n_heads)
            _ln(sd, seed, q + '.enc_attn_ln', D)
        _linear(sd, seed, q + '.pf.fc_1', pf, D)
        _linear(sd, seed, q + '.pf.fc_2', D, pf)
        _ln(sd, seed, q + '.pf_ln', D)
    sd[p + '.scale'] = torch.sqrt(torch.tensor(D, dtype=torch.float32))
    _classifier(sd, seed, p + '.classifier', V, D, 3.0, 3.3)


def make_state_dict(kind, hp, seed=0):
    """kind in KINDS; hp = the reference ctor kwargs (``_hp`` dict, caption_models.py:33-40,247-257)."""
    sd = {}
    V = hp['num_tokens']
    if kind == 'lstm':
        image_encoder(sd, seed, 'encoder', hp['emb_dim'])
        lstm_decoder(sd, seed, 'decoder', V, hp['emb_dim'], hp['hidden_size'], hp['num_layers'])
    elif kind == 'lstm_labels':
        E = hp['emb_dim']
        image_encoder(sd, seed, 'encoder.image_encoder', E)
        sd['encoder.label_encoder.embedding.weight'] = synth.tensor(
            seed, 'encoder.label_encoder.embedding.weight', (V, E), synth.SQRT3 * 0.5)
        _linear(sd, seed, 'encoder.linear', E, 2 * E)
        lstm_decoder(sd, seed, 'decoder', V, E, hp['hidden_size'], hp['num_layers'], with_embedding=False)
        # tied embedding: one tensor under two keys (caption_models.py:125)
        sd['decoder.embedding.weight'] = sd['encoder.label_encoder.embedding.weight']
    elif kind in ('xfmr_base', 'xfmr'):
        image_encoder(sd, seed, 'encoder', hp['hid_dim'])
        xfmr_decoder(sd, seed, 'decoder', V, hp['hid_dim'], hp['n_layers'], hp['n_heads'], hp['pf_dim'],
                     hp['max_len'], cross=(kind == 'xfmr'))
    else:
        raise ValueError(kind)
    return sd


def default_hp(kind, num_tokens, small=False):
    """Canonical (trained-checkpoint) dims from SURVEY.md section 2.4, or a small variant for fast tests."""
    if kind in ('lstm', 'lstm_labels'):
        if small:
            return dict(num_tokens=num_tokens, emb_dim=64, hidden_size=96, num_layers=2,
                        enc_dropout=0.3, dec_dropout=0.1)
        return dict(num_tokens=num_tokens, emb_dim=512, hidden_size=512, num_layers=3,
                    enc_dropout=0.3, dec_dropout=0.1)
    if small:
        return dict(num_tokens=num_tokens, hid_dim=64, n_layers=2, n_heads=4, pf_dim=128,
                    enc_dropout=0.3, dec_dropout=0.1, pad_index=0, max_len=64)
    return dict(num_tokens=num_tokens, hid_dim=512, n_layers=3, n_heads=8, pf_dim=2048,
                enc_dropout=0.3, dec_dropout=0.1, pad_index=0, max_len=128)
