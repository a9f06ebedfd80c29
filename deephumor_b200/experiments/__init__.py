from .inference import seq_to_text, seqs_to_texts, split_caption, split_captions, text_to_seq, texts_to_seqs
from .metrics import perplexity

__all__ = ['text_to_seq', 'seq_to_text', 'split_caption', 'perplexity', 'texts_to_seqs', 'seqs_to_texts', 'split_captions']
