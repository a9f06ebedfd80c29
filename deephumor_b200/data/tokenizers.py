"""Regex tokenizers, behaviour-compatible with deephumor/data/tokenizers.py:14-29."""
import re


class Tokenizer:
    def tokenize(self, text):
        raise NotImplementedError


class WordPunctTokenizer(Tokenizer):
    """Runs of word characters / apostrophes / angle brackets, or runs of other non-space symbols."""
    token_pattern = re.compile(r"[<\w'>]+|[^\w\s]+")

    def tokenize(self, text):
        return self.token_pattern.findall(text)


class CharTokenizer(Tokenizer):
    """Single characters, except that ``<word>`` special tokens stay whole."""
    token_pattern = re.compile(r"<\w+>|.")

    def tokenize(self, text):
        return self.token_pattern.findall(text)
