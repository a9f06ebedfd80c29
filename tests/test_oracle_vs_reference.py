"""Live check of the oracle against the imported reference (only where /root/reference exists)."""
import pytest
import torch

from deephumor_b200.utils import synth, synth_weights
from oracle import model, noise, refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason='reference tree not present on this box')


@pytest.mark.parametrize('kind', synth_weights.KINDS)
def test_live_reference_agrees(kind):
    V = 500
    hp = synth_weights.default_hp(kind, V, small=True)
    sd = synth_weights.make_state_dict(kind, hp, seed=0)
    ref = refshim.build_reference(kind, hp, sd)       # strict=True load: pins the state_dict layout
    assert set(ref.state_dict().keys()) == set(sd.keys())
    imgs = synth.images(5, 100, 2)
    labs = synth.labels(5, 100, 2, V) if kind == 'lstm_labels' else None
    with torch.no_grad():
        r = ref.encoder(imgs, labs) if kind == 'lstm_labels' else ref.encoder(imgs)
        o = model.encode(kind, sd, imgs, labs)
        assert torch.allclose(o[0], r[0] if kind == 'xfmr' else r, atol=1e-5)
        for mode in ('deterministic', 'injected'):
            rid, rl = refshim.reference_generate_batch(ref, kind, imgs, labs, first_index=100, mode=mode, seed=3,
                                                       max_len=10, beam_size=4, top_k=12, temperature=1.2)
            oid, ol = model.generate_batch(kind, sd, hp, None, labs, first_index=100, max_len=10, encoded=o,
                                           beam_size=4, top_k=12, temperature=1.2, noise=noise.Noise(mode, 3))
            assert torch.equal(rid, oid) and torch.equal(rl, ol)
