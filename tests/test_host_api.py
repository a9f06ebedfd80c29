"""CPU tests: drop-in API surface, state_dict layout / known-answer parameter counts, host text helpers,
and the C-ABI library (loads, exports every symbol the header declares)."""
import ctypes
import os

import pytest
import torch

from deephumor_b200 import _lib, models
from deephumor_b200.data import CharTokenizer, SPECIAL_TOKENS, Vocab, WordPunctTokenizer, build_vocab
from deephumor_b200.experiments import seq_to_text, split_caption, text_to_seq
from deephumor_b200.utils import synth_weights as W
from oracle import refshim

CLS = {'lstm': models.CaptioningLSTM, 'lstm_labels': models.CaptioningLSTMWithLabels,
       'xfmr_base': models.CaptioningTransformerBase, 'xfmr': models.CaptioningTransformer}
# trainable-parameter counts printed by the reference notebook (deephumor_demo.ipynb cells 17-32)
NOTEBOOK_COUNTS = {('lstm', 36541): 44808381, ('lstm', 71): 7426631, ('lstm_labels', 36541): 45333181,
                   ('lstm_labels', 71): 7951431, ('xfmr_base', 36541): 48027325, ('xfmr_base', 71): 10645575,
                   ('xfmr', 36541): 51182269, ('xfmr', 71): 13800519}


@pytest.mark.parametrize('kind,V', sorted(NOTEBOOK_COUNTS))
def test_parameter_counts_match_reference_notebook(kind, V):
    m = CLS[kind](**W.default_hp(kind, V))
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == NOTEBOOK_COUNTS[(kind, V)]


@pytest.mark.parametrize('kind', W.KINDS)
def test_state_dict_layout_and_checkpoint_roundtrip(kind, tmp_path):
    hp = W.default_hp(kind, 300, small=True)
    sd = W.make_state_dict(kind, hp, seed=1)
    m = CLS[kind](**hp)
    m.load_state_dict(sd, strict=True)
    own = m.state_dict()
    assert set(own) == set(sd) and all(own[k].shape == sd[k].shape and own[k].dtype == sd[k].dtype for k in sd)
    assert m._hp == hp
    path = str(tmp_path / 'ckpt.pth')
    m.save(path)
    ck = torch.load(path, map_location='cpu')
    assert set(ck) == {'model', 'hp'} and ck['hp'] == hp
    m2 = CLS[kind].from_pretrained(path)
    assert all(torch.equal(m2.state_dict()[k], sd[k]) for k in sd)
    if kind == 'lstm_labels':      # tied embedding under two keys (Q22)
        assert m2.decoder.embedding is m2.encoder.label_encoder.embedding
    if refshim.available():        # the reference accepts our checkpoint and vice versa
        ref = refshim.build_reference(kind, hp, own)
        assert set(ref.state_dict()) == set(own)


def test_no_cpu_fallback():
    m = CLS['lstm'](**W.default_hp('lstm', 50, small=True)).eval()
    with pytest.raises(RuntimeError, match='CUDA'):
        m.generate(torch.zeros(1, 3, 224, 224), max_len=4, beam_size=1, top_k=2)
    with pytest.raises(AssertionError):
        m.decoder.generate(torch.zeros(1, 1, 64), beam_size=5, top_k=2)


def test_library_exports_every_declared_symbol():
    from deephumor_b200 import build
    build.build()
    protos = _lib.parse_header()
    assert len(protos) >= 20 and 'dh_select_tokens' in protos
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f'{name} declared in include/deephumor_b200.h but not exported'
    _lib.LIB.load()
    assert _lib.LIB.load().dh_version() == _lib.LIB._header_version()
    out = os.popen(f'cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null').read()
    assert 'sm_100a' in out


def test_vocab_tokenizers_and_text_helpers():
    wt, ct = WordPunctTokenizer(), CharTokenizer()
    assert wt.tokenize("don't <sep> stop!!") == ["don't", '<sep>', 'stop', '!!']
    assert ct.tokenize('ab<sep>c') == ['a', 'b', '<sep>', 'c']
    v = Vocab(['zebra', 'apple', '<eos>', 'mango'])
    assert v.tokens[:6] == list(SPECIAL_TOKENS.values()) and v.tokens[6:] == ['apple', 'mango', 'zebra']
    assert (v.stoi['<pad>'], v.stoi['<unk>'], v.stoi['<eos>']) == (0, 1, 3) and len(v) == 9
    vb = build_vocab(['a b', 'a c', 'A d'], wt, min_df=2)
    assert vb.tokens[6:] == ['a']
    seq = text_to_seq('Apple kiwi zebra', v, wt)
    assert seq.tolist() == [[6, 1, 8]]
    assert seq_to_text(torch.tensor([6, 8, 3, 7]), v) == 'apple zebra'
    assert split_caption('hello , world <sep> bye <emp>', 3) == ['hello, world', 'bye', '']


def test_vocab_save_load(tmp_path):
    v = Vocab(['b', 'a'])
    p = str(tmp_path / 'v.txt')
    v.save(p)
    assert Vocab.load(p).tokens == v.tokens

def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The host structs passed by pointer through the C ABI (dh_beam_state, dh_lstm_operands) are mirrored by hand in
    deephumor_b200/_lib.py: compile the header with gcc and compare sizes and field offsets."""
    import ctypes
    import shutil
    import subprocess
    from deephumor_b200 import _lib
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('gcc not available')
    structs = {'dh_beam_state': _lib.BeamState, 'dh_lstm_operands': _lib.LstmOperands}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{_lib.HEADER}"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run([gcc, '-o', str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f'{cname}.{fname}']) == getattr(cls, fname).offset, (cname, fname)


def test_host_batch_copy_schedule():
    """runtime/encoder.h2d_schedule: every image is copied exactly once, no pass exceeds the trunk chunk, the ramp is only
    taken when enough images are left, and the default schedule of the benchmark batch is 64 / 192 / 256."""
    from deephumor_b200.runtime.encoder import h2d_schedule
    assert h2d_schedule(512, 512, (64, 192, 256)) == [64, 192, 256]
    assert h2d_schedule(65, 512, (64, 192, 256)) == [65]
    assert h2d_schedule(4096, 512, (64, 192, 256))[:5] == [64, 192, 256, 512, 512]
    for n in (0, 1, 63, 64, 127, 128, 300, 511, 512, 513, 1000, 4096, 8191):
        for chunk in (100, 256, 512, 1024):
            sizes = h2d_schedule(n, chunk, (64, 192, 256))
            assert sum(sizes) == n and all(0 < s <= chunk for s in sizes)      # a pass never exceeds the staging buffer

