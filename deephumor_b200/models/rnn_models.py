"""LSTM decoder with the reference's module layout and signatures (deephumor/models/rnn_models.py:8-143)."""
import torch
from torch import nn

from ..runtime.lstm import LSTMDecoderRT
from ._base import RTModule, as_caption, finish, prefixed, resolve_noise


class LSTMDecoder(RTModule):
    def __init__(self, num_tokens, emb_dim=256, hidden_size=512, num_layers=3, dropout=0.1, embedding=None):
        super().__init__()
        self.num_tokens = num_tokens
        self.embedding = embedding if embedding is not None else nn.Embedding(num_tokens, emb_dim)
        self.lstm = nn.LSTM(emb_dim, hidden_size, num_layers, batch_first=True,
                            dropout=(0 if num_layers == 1 else dropout))      # parameter container only
        self.classifier = nn.Linear(hidden_size, num_tokens)

    def _rt(self):
        return self._get_rt('dec', lambda dt, dev: LSTMDecoderRT(prefixed(self, 'm'), 'm', dt, dev))

    def forward(self, image_emb, captions, lengths=None):
        dev = self._device()
        return self._rt().forward(image_emb.to(dev, torch.float32).contiguous(), captions.to(dev), lengths)

    def token_logprob(self, image_emb, captions, lengths, targets):
        dev = self._device()
        return self._rt().token_logprob(image_emb.to(dev, torch.float32).contiguous(), captions.to(dev), lengths, targets)

    def generate(self, image_emb, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50, eos_index=3,
                 *, noise=None, seed=None, image_base=0, unk_index=1):
        """image_emb [N,1,E] or [N,E] (reference: [1,1,E]); returns the reference's 1-D ids for N == 1, else
        (ids [N,max_len], lengths [N])."""
        assert beam_size <= top_k, '`beam_size` should be less than `top_k`'
        dev = self._device()
        emb = image_emb.to(dev, torch.float32)
        emb = emb.reshape(emb.shape[0], emb.shape[-1]).contiguous()
        mode, seed = resolve_noise(noise, seed)
        ids, lens, status = self._rt().generate(emb, as_caption(caption, dev), max_len, temperature, beam_size, top_k,
                                                eos_index, unk_index, mode, seed, image_base)
        return finish(ids, lens, status, emb.shape[0])
