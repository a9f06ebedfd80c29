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


def test_entry_points_validate_arguments_before_touching_the_device():
    """The C entries check shapes / alignment first and return an argument error (no CUDA call, so this runs without a GPU):
    the chained convolution and the LayerNorm contraction added in round 2."""
    L = _lib.LIB
    P = 1 << 20                                                   # a 16-byte aligned, never dereferenced "device pointer"
    good = dict(y2=P, x2=P, src=0, w=P, bias=P, out=P, n=2, H=8, W=8, C1=64, C2=256, H2=8, W2=8, s2=1, Cout=256, wn=P, bn=P,
                z=P, N2=64, dt=1, stream=None)

    def chain(**kw):
        a = dict(good, **kw)
        L.call('dh_conv1x1_chain_tc', a['y2'], a['x2'], a['src'], a['w'], a['bias'], a['out'], a['n'], a['H'], a['W'], a['C1'],
               a['C2'], a['H2'], a['W2'], a['s2'], a['Cout'], a['wn'], a['bn'], a['z'], a['N2'], a['dt'], a['stream'])

    chain(n=0)                                                    # nothing to do: returns before any device work
    for bad in (dict(N2=96), dict(Cout=300), dict(C1=60), dict(C2=128), dict(src=1, C2=64, H2=15, W2=15, s2=3), dict(y2=P + 8),
                dict(dt=0), dict(z=None)):
        with pytest.raises(_lib.DeepHumorLibError, match='argument error'):
            chain(**bad)
    with pytest.raises(_lib.DeepHumorLibError, match='argument error'):           # LayerNorm rows are 512 wide
        L.call('dh_gemm_tc_ln', P, 512, P, 512, 1, P, P, 256, P, P, 1e-5, P, 256, 128, 256, 512, None)


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


def test_batched_text_helpers_equal_the_per_item_reference_forms():
    """batched ids -> text / text -> ids over (ids [N, max_len], lengths [N]) == the reference's
    one-string-at-a-time helpers (experiments/inference.py:11-58) item by item; when /root/reference is present the
    reference's own functions are the comparison."""
    from deephumor_b200.experiments import seqs_to_texts, split_captions, texts_to_seqs
    wt = WordPunctTokenizer()
    words = [f'w{i}' for i in range(200)] + ['hello', ',', 'world', '!', '<sep>']
    v = Vocab(words)
    g = torch.Generator().manual_seed(0)
    N, L = 300, 32
    ids = torch.randint(0, len(v), (N, L), generator=g)
    ids[::3, 7] = 3                                   # <eos> in the middle
    ids[::5, 0] = 3                                   # <eos> first: empty text
    lengths = torch.randint(1, L + 1, (N,), generator=g)
    texts = seqs_to_texts(ids, lengths, v)
    ref_fn = seq_to_text
    try:
        from oracle import refshim
        if refshim.available():
            refshim.import_reference()
            from deephumor.experiments.inference import seq_to_text as ref_fn      # the unmodified reference
    except Exception:
        pass
    for n in range(N):
        assert texts[n] == ref_fn(ids[n, :int(lengths[n])], v)
    strings = ['Hello , world !', 'w1 w2 <sep> w3 unknownword', '', 'w5']
    seqs, lens = texts_to_seqs(strings, v, wt)
    for n, t in enumerate(strings):
        one = text_to_seq(t, v, wt) if t else torch.zeros(1, 0, dtype=torch.int64)
        assert seqs[n, :int(lens[n])].tolist() == one[0].tolist() and int(lens[n]) == one.shape[1]
        assert bool((seqs[n, int(lens[n]):] == 0).all())
    assert split_captions(['a , b <sep> c', 'd'], 2) == [split_caption('a , b <sep> c', 2), split_caption('d', 2)]


def test_plan_cache_is_an_lru():
    from deephumor_b200.runtime.ops import PlanCache
    c = PlanCache(capacity=2)
    c.put('a', 1); c.put('b', 2)
    assert c.get('a') == 1                               # 'a' becomes the most recent
    c.put('c', 3)                                        # evicts 'b', not 'a'
    assert c.get('b') is None and c.get('a') == 1 and c.get('c') == 3 and len(c) == 2


def test_api_guards_raise_instead_of_silently_differing():
    """training mode and pad_index=None are not implemented by the runtime -- they must raise, not diverge."""
    from deephumor_b200.models import TransformerDecoder
    with pytest.raises(NotImplementedError):
        TransformerDecoder(num_tokens=50, hid_dim=64, n_layers=1, n_heads=4, pf_dim=64, pad_index=None)
    m = TransformerDecoder(num_tokens=50, hid_dim=64, n_layers=1, n_heads=4, pf_dim=64, pad_index=0)
    m.train()
    with pytest.raises(RuntimeError):
        m._device()


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
    structs = {'dh_beam_state': _lib.BeamState, 'dh_lstm_operands': _lib.LstmOperands, 'dh_vocab_sparse': _lib.VocabSparse,
               'dh_resnet50_weights': _lib.Resnet50Weights, 'dh_xfmr_layer': _lib.XfmrLayer, 'dh_xfmr_weights': _lib.XfmrWeights,
               'dh_xfmr_buffers': _lib.XfmrBuffers}
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

