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
