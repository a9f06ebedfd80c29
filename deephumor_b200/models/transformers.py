"""Transformer decoders with the reference's module layout and signatures
(deephumor/models/transformers.py:43-165, 309-579, 582-825).

The layer classes are parameter containers whose names reproduce the reference ``state_dict``; compute goes
through ``runtime.xfmr.XfmrDecoderRT``.  ``TransformerEncoder`` / ``EncoderLayer`` of the reference are dead
code there (forward raises AttributeError) and are not provided.
"""
import torch
from torch import nn

from ..runtime.xfmr import XfmrDecoderRT
from ._base import RTModule, as_caption, finish, prefixed, resolve_noise


class MultiHeadAttentionLayer(nn.Module):
    def __init__(self, hid_dim=512, n_heads=8, dropout=0.):
        super().__init__()
        assert hid_dim % n_heads == 0, "hid_dim must be divisible by n_heads"
        self.hid_dim, self.n_heads, self.head_dim = hid_dim, n_heads, hid_dim // n_heads
        self.fc_q, self.fc_k = nn.Linear(hid_dim, hid_dim), nn.Linear(hid_dim, hid_dim)
        self.fc_v, self.fc_o = nn.Linear(hid_dim, hid_dim), nn.Linear(hid_dim, hid_dim)
        self.dropout = nn.Dropout(dropout)
        self.scale = nn.Parameter(torch.sqrt(torch.tensor(self.head_dim, dtype=torch.float32)), requires_grad=False)


class PositionwiseFeedforwardLayer(nn.Module):
    def __init__(self, hid_dim=512, pf_dim=2048, dropout=0.):
        super().__init__()
        self.fc_1, self.fc_2 = nn.Linear(hid_dim, pf_dim), nn.Linear(pf_dim, hid_dim)
        self.dropout = nn.Dropout(dropout)


class DecoderLayer(nn.Module):
    def __init__(self, hid_dim=512, n_heads=8, pf_dim=2048, dropout=0.):
        super().__init__()
        self.self_attn = MultiHeadAttentionLayer(hid_dim, n_heads, dropout)
        self.self_attn_ln = nn.LayerNorm(hid_dim)
        self.enc_attn = MultiHeadAttentionLayer(hid_dim, n_heads, dropout)
        self.enc_attn_ln = nn.LayerNorm(hid_dim)
        self.pf = PositionwiseFeedforwardLayer(hid_dim, pf_dim, dropout)
        self.pf_ln = nn.LayerNorm(hid_dim)
        self.dropout = nn.Dropout(dropout)


class SelfAttentionDecoderLayer(nn.Module):
    def __init__(self, hid_dim=512, n_heads=8, pf_dim=2048, dropout=0.):
        super().__init__()
        self.self_attn = MultiHeadAttentionLayer(hid_dim, n_heads, dropout)
        self.self_attn_ln = nn.LayerNorm(hid_dim)
        self.pf = PositionwiseFeedforwardLayer(hid_dim, pf_dim, dropout)
        self.pf_ln = nn.LayerNorm(hid_dim)
        self.dropout = nn.Dropout(dropout)


class _DecoderBase(RTModule):
    _cross = False
    _layer_cls = SelfAttentionDecoderLayer

    def __init__(self, num_tokens, hid_dim=512, n_layers=6, n_heads=8, pf_dim=2048, dropout=0., pad_index=None,
                 max_len=128):
        super().__init__()
        if pad_index is None:
            # the reference documents pad_index=None as "no masking"; its captioners always pass 0 (caption_models.py:211-212)
            raise NotImplementedError('pad_index=None (no key masking) is not supported: pass the pad token id (0)')
        self.pad_index = pad_index
        self.tok_embedding = nn.Embedding(num_tokens, hid_dim)
        self.pos_embedding = nn.Embedding(max_len, hid_dim)
        self.dropout = nn.Dropout(dropout)
        self.layers = nn.ModuleList([self._layer_cls(hid_dim, n_heads, pf_dim, dropout) for _ in range(n_layers)])
        self.scale = nn.Parameter(torch.sqrt(torch.tensor(hid_dim, dtype=torch.float32)), requires_grad=False)
        self.classifier = nn.Linear(hid_dim, num_tokens)
        self._hp = dict(hid_dim=hid_dim, n_layers=n_layers, n_heads=n_heads, pf_dim=pf_dim,
                        pad_index=0 if pad_index is None else pad_index)

    def _rt(self):
        return self._get_rt('dec', lambda dt, dev: XfmrDecoderRT(prefixed(self, 'm'), 'm', self._hp, self._cross, dt, dev))

    def _spatial(self, enc_out, dev):
        rt = self._rt()
        n = enc_out.shape[0]
        assert enc_out.shape[1] == 49, 'cross-attention runtime expects the 7x7 = 49 spatial tokens'
        return enc_out.to(dev).reshape(n * 49, -1).to(rt.dtype).contiguous()

    def token_logprob(self, x, enc_out, start_emb, targets):
        dev = self._device()
        sp = self._spatial(enc_out, dev) if self._cross else None
        return self._rt().token_logprob(start_emb.to(dev, torch.float32).contiguous(), sp, x.to(dev), targets)

    def _generate(self, start_emb, enc_out, caption, max_len, temperature, beam_size, top_k, eos_index, noise, seed,
                  image_base, unk_index):
        assert beam_size <= top_k, '`beam_size` should be less than `top_k`'
        if max_len + 1 > 160:
            raise ValueError(f'max_len={max_len}: the KV cache / slot tables hold at most 160 positions per beam (max_len <= 159)')
        dev = self._device()
        start = start_emb.to(dev, torch.float32).contiguous()
        sp = self._spatial(enc_out, dev) if self._cross else None
        mode, seed = resolve_noise(noise, seed)
        ids, lens, status = self._rt().generate(start, sp, as_caption(caption, dev), max_len, temperature, beam_size,
                                                top_k, eos_index, unk_index, mode, seed, image_base)
        return finish(ids, lens, status, start.shape[0])


class TransformerDecoder(_DecoderBase):
    """transformers.py:380-579 (cross-attention over encoder outputs)."""
    _cross = True
    _layer_cls = DecoderLayer

    def forward(self, x, enc_out, start_emb=None):
        dev = self._device()
        assert start_emb is not None, 'the captioners always pass start_emb (caption_models.py:404)'
        return self._rt().forward(start_emb.to(dev, torch.float32).contiguous(), self._spatial(enc_out, dev), x.to(dev))

    def generate(self, start_emb, enc_out, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50,
                 eos_index=3, *, noise=None, seed=None, image_base=0, unk_index=1):
        return self._generate(start_emb, enc_out, caption, max_len, temperature, beam_size, top_k, eos_index, noise,
                              seed, image_base, unk_index)


class SelfAttentionTransformerDecoder(_DecoderBase):
    """transformers.py:639-825 (image embedding prepended, no encoder attention)."""

    def forward(self, x, start_emb):
        dev = self._device()
        return self._rt().forward(start_emb.to(dev, torch.float32).contiguous(), None, x.to(dev))

    def generate(self, start_emb, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50, eos_index=3,
                 *, noise=None, seed=None, image_base=0, unk_index=1):
        return self._generate(start_emb, None, caption, max_len, temperature, beam_size, top_k, eos_index, noise, seed,
                              image_base, unk_index)
