"""Regex tokenizers with the behaviour of deephumor/data/tokenizers.py:14-29 (host-side text I/O).

Both reference tokenizers are "find all matches of one pattern"; here that is a single base class parameterised by the
pattern, with a batch helper for the caller side of batched generation.
"""
import re


class Tokenizer:
    """Splits text into the non-overlapping matches of `pattern`, left to right."""
    pattern = None

    def __init__(self):
        if self.pattern is None:
            raise NotImplementedError('Tokenizer is abstract: use WordPunctTokenizer or CharTokenizer')
        self.token_pattern = re.compile(self.pattern)      # attribute name kept from the reference

    def tokenize(self, text):
        return self.token_pattern.findall(text)

    def tokenize_batch(self, texts):
        find = self.token_pattern.findall
        return [find(t) for t in texts]

    __call__ = tokenize


class WordPunctTokenizer(Tokenizer):
    # runs of word characters / apostrophes / angle brackets (so `<sep>` stays whole), else runs of other symbols
    pattern = r"[<\w'>]+|[^\w\s]+"


class CharTokenizer(Tokenizer):
    # single characters, except that `<word>` special tokens stay whole
    pattern = r"<\w+>|."
