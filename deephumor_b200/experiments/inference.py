"""ids <-> text helpers, behaviour-compatible with deephumor/experiments/inference.py:11-89 (host-side only)."""
import re

import torch

from ..data import SPECIAL_TOKENS

PUNCT_PATTERN = re.compile(r"( )([!#$%&\()*+,\-.\/:;<=>?@\\^{|}~]+)")


def text_to_seq(text, vocab, tokenizer):
    """str -> int64 [1, n_tokens]; out-of-vocabulary tokens map to <unk>."""
    unk = vocab.stoi[SPECIAL_TOKENS['UNK']]
    ids = [vocab.stoi.get(tok, unk) for tok in tokenizer.tokenize(text.lower())]
    return torch.tensor(ids).unsqueeze(0)


def seq_to_text(seq, vocab, delimiter=' '):
    """1-D id tensor -> text, cut at the first <eos>."""
    eos = vocab.stoi[SPECIAL_TOKENS['EOS']]
    hits = torch.where(seq == eos)[0]
    if len(hits) > 0:
        seq = seq[:hits[0]]
    return delimiter.join(vocab.itos[int(i)] for i in seq.cpu().numpy())


def split_caption(text, num_blocks=None):
    """Split on <sep>, strip remaining <...> tokens and outer whitespace, re-attach punctuation."""
    def clean(block):
        block = re.sub(r'<\w+>', '', block)
        block = re.sub(r'^\s+', '', block)
        block = re.sub(r'\s+$', '', block)
        return PUNCT_PATTERN.sub('\\2', block)

    blocks = [clean(b) for b in text.split(SPECIAL_TOKENS['SEP'])]
    if num_blocks is None:
        return blocks
    return (blocks + [''] * max(0, num_blocks - len(blocks)))[:num_blocks]


# ---------------------------------------------------------------------------------------- batched forms
# The reference converts one caption at a time (inference.py:11-58); the batched generate() of this package returns
# (ids [N, max_len], lengths [N]) for thousands of images, so the callers either side of the hot path get batched forms
# with the same per-item semantics.

def texts_to_seqs(texts, vocab, tokenizer, pad_index=0):
    """List[str] -> (int64 [N, T] padded with pad_index, int64 [N] lengths); row n == text_to_seq(texts[n]) (a batch of
    `caption=` prefixes of equal length can be passed to generate() directly)."""
    unk = vocab.stoi[SPECIAL_TOKENS['UNK']]
    stoi = vocab.stoi
    rows = [[stoi.get(tok, unk) for tok in tokenizer.tokenize(t.lower())] for t in texts]
    lengths = torch.tensor([len(r) for r in rows], dtype=torch.int64)
    out = torch.full((len(rows), int(lengths.max()) if rows else 0), pad_index, dtype=torch.int64)
    for n, r in enumerate(rows):
        out[n, :len(r)] = torch.tensor(r, dtype=torch.int64)
    return out, lengths


def seqs_to_texts(ids, lengths, vocab, delimiter=' '):
    """(ids [N, max_len], lengths [N]) as returned by the batched generate() -> List[str]; item n ==
    seq_to_text(ids[n, :lengths[n]], vocab, delimiter).  One device->host copy and one vectorised <eos> search for the
    whole batch instead of N tensor slices."""
    ids = ids.detach().cpu().numpy()
    lengths = lengths.detach().cpu().numpy() if torch.is_tensor(lengths) else lengths
    eos = vocab.stoi[SPECIAL_TOKENS['EOS']]
    n, width = ids.shape
    import numpy as np
    pos = np.arange(width)[None, :]
    is_eos = (ids == eos) & (pos < np.asarray(lengths)[:, None])
    first = np.where(is_eos.any(axis=1), is_eos.argmax(axis=1), np.asarray(lengths))     # cut at the first <eos>
    itos = vocab.itos
    return [delimiter.join(itos[int(t)] for t in ids[i, :first[i]]) for i in range(n)]


def split_captions(texts, num_blocks=None):
    """List[str] -> List[List[str]]; item n == split_caption(texts[n], num_blocks)."""
    return [split_caption(t, num_blocks) for t in texts]
