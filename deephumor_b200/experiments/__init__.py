from .inference import seq_to_text, split_caption, text_to_seq
from .metrics import perplexity

__all__ = ['text_to_seq', 'seq_to_text', 'split_caption', 'perplexity']
