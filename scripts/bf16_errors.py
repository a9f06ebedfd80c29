"""Prints tensor-core-mode errors of every model against the reference fixtures (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200 import models
from deephumor_b200.experiments import perplexity
from tests import helpers as H

CLS = {'lstm': models.CaptioningLSTM, 'lstm_labels': models.CaptioningLSTMWithLabels,
       'xfmr_base': models.CaptioningTransformerBase, 'xfmr': models.CaptioningTransformer}
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
for tag in ('small', 'canon'):
    for kind in H.KINDS:
        fx = H.load_fixture(tag, kind)
        sd, imgs, labs, caps, lens = H.fixture_inputs(fx)
        m = CLS[kind](**fx['hp']); m.load_state_dict(sd, strict=True); m = m.cuda().eval().set_precision(prec)
        with torch.no_grad():
            enc = m.encoder(imgs.cuda(), labs.cuda()) if kind == 'lstm_labels' else m.encoder(imgs.cuda())
            emb = enc[0] if kind == 'xfmr' else enc
            e_emb = H.rel_err(emb, fx['emb'])
            e_sp = H.rel_err(enc[1], fx['spatial']) if kind == 'xfmr' else float('nan')
            args = (imgs.cuda(), caps[:, :-1].cuda(), lens.cuda()) + ((labs.cuda(),) if kind == 'lstm_labels' else ())
            logits = m(*args)
            e_log = H.rel_err(logits[..., :fx['logits'].shape[-1]], fx['logits'])
            T = min(logits.shape[1], caps.shape[1])
            pp = float(perplexity(logits[:, :T], caps[:, :T].cuda(), lens.cuda()))
            toks = []
            for g in fx['gen']:
                prefix = caps[:1, :g['prefix_len']] if g['prefix_len'] else None
                kw = dict(caption=prefix, max_len=fx['max_len'], temperature=g['temperature'], beam_size=g['beam_size'],
                          top_k=g['top_k'], noise=g['mode'], seed=g['noise_seed'])
                ids, ln = m.generate(imgs.cuda(), labs.cuda(), **kw) if kind == 'lstm_labels' else m.generate(imgs.cuda(), **kw)
                same = sum(bool((ids[n].cpu() == g['ids'][n]).all()) for n in range(ids.shape[0]))
                toks.append(f'{same}/{ids.shape[0]}')
        print(f'{tag:6s}{kind:12s} emb {e_emb:.2e} spatial {e_sp:.2e} logits {e_log:.2e} pp {pp:.1f} vs {fx["perplexity"]:.1f} '
              f'rows-identical {" ".join(toks)}', flush=True)
