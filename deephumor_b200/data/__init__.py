from .tokenizers import CharTokenizer, Tokenizer, WordPunctTokenizer
from .vocab import SPECIAL_TOKENS, Vocab, build_vocab, build_vocab_from_file

__all__ = ['SPECIAL_TOKENS', 'Vocab', 'build_vocab', 'build_vocab_from_file', 'Tokenizer', 'WordPunctTokenizer',
           'CharTokenizer']
