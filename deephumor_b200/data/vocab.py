"""Token vocabulary, behaviour-compatible with deephumor/data/vocab.py:5-90.

Ids 0-5 are the six special tokens in the order below; every other token follows in sorted order, so
``<pad>``=0, ``<unk>``=1, ``<bos>``=2, ``<eos>``=3, ``<sep>``=4, ``<emp>``=5 (the decode kernels' eos / unk / pad
defaults rely on this).
"""
import collections

SPECIAL_TOKENS = dict(PAD='<pad>', UNK='<unk>', BOS='<bos>', EOS='<eos>', SEP='<sep>', EMPTY='<emp>')


class Vocab:
    def __init__(self, tokens, special_tokens=tuple(SPECIAL_TOKENS.values())):
        specials = list(special_tokens)
        rest = sorted(t for t in tokens if t not in special_tokens)     # duplicates kept, like the reference
        self.tokens = specials + rest
        self.stoi, self.itos = {}, {}
        for i, t in enumerate(self.tokens):
            self.stoi[t] = i                                             # last duplicate wins (dict-comprehension order)
            self.itos[i] = t

    def __len__(self):
        return len(self.tokens)

    def __iter__(self):
        return iter(self.tokens)

    def save(self, filepath):
        with open(filepath, 'w') as fh:
            fh.writelines(t + '\n' for t in self.tokens)

    @staticmethod
    def load(filepath):
        with open(filepath) as fh:
            return Vocab([line.strip('\n') for line in fh])


def build_vocab(documents, tokenizer, min_df=7):
    """Tokens whose document frequency (lower-cased) is at least min_df."""
    df = collections.Counter()
    for doc in documents:
        df.update(set(tokenizer.tokenize(doc.lower())))
    return Vocab([t for t, n in df.items() if n >= min_df])


def build_vocab_from_file(captions_file, tokenizer, min_df=7):
    """captions_file: tab-separated ``label<TAB>?<TAB>caption`` lines; only the third field is used."""
    docs = []
    with open(captions_file) as fh:
        for line in fh:
            _, _, caption = line.strip().split('\t')
            docs.append(caption)
    return build_vocab(docs, tokenizer, min_df=min_df)
