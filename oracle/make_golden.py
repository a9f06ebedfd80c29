"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.pt

Inputs and weights are re-creatable anywhere from ``deephumor_b200.utils.synth[_weights]`` (hash-based),
so fixtures hold only OUTPUTS of the reference: encoder embeddings, teacher-forced logits, perplexity,
and generated ids under the shared noise model (oracle/noise.py), plus the
per-image minimum decision margin measured by the oracle so near-ties can be excluded (Appendix D.5).
Loading the synthetic state_dict into the reference with strict=True pins the Appendix-C layout.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deephumor_b200.utils import synth, synth_weights  # noqa: E402
from oracle import model, noise, refshim  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
GEN_VARIANTS = [  # (mode, beam, top_k, temperature, prefix_len)
    ('deterministic', 5, 50, 1.0, 0), ('deterministic', 1, 2, 1.0, 0), ('deterministic', 3, 7, 0.8, 2),
    ('injected', 5, 50, 1.0, 0), ('injected', 1, 20, 1.3, 0), ('injected', 4, 9, 1.1, 3),
]


N_FWD = 8   # images whose teacher-forced logits / spatial tokens are stored (generation + embeddings cover all n_img)


def run(kind, V, small, n_img, max_len, wseed, variants, tag, hp_update=None):
    hp = synth_weights.default_hp(kind, V, small=small)
    hp.update(hp_update or {})
    sd = synth_weights.make_state_dict(kind, hp, seed=wseed)
    ref = refshim.build_reference(kind, hp, sd)
    imgs = synth.images(0, 0, n_img)
    labs = synth.labels(0, 0, n_img, V) if kind == 'lstm_labels' else None
    caps, lens = synth.captions(0, 0, n_img, V, width=max_len, min_len=4)
    nf = min(n_img, N_FWD)
    fx = {'kind': kind, 'hp': hp, 'wseed': wseed, 'n_img': n_img, 'n_fwd': nf, 'max_len': max_len, 'V': V,
          'state_dict_keys': len(sd), 'gen': []}
    t0 = time.time()
    with torch.no_grad():
        enc = ref.encoder(imgs, labs) if kind == 'lstm_labels' else ref.encoder(imgs)
        fx['emb'] = (enc[0] if kind == 'xfmr' else enc).clone()
        if kind == 'xfmr':
            fx['spatial'] = enc[1][:nf].clone()
        args = (imgs[:nf], caps[:nf, :-1], lens[:nf]) + ((labs[:nf],) if kind == 'lstm_labels' else ())
        logits = ref(*args)
        fx['logits_shape'] = tuple(logits.shape)
        fx['logits'] = logits[:, :, :256].clone() if not small else logits.clone()
        fx['logits_rowsum'] = logits.double().sum(-1).float()
        T = min(logits.shape[1], caps.shape[1])
        sys.path.insert(0, refshim.REF_ROOT)
        from deephumor.experiments.metrics import perplexity as ref_pp
        fx['perplexity'] = float(ref_pp(logits[:, :T].clone(), caps[:nf, :T], lens[:nf]))
        o_enc = model.encode(kind, sd, imgs, labs)
        for (mode, B, K, T_, plen) in variants:
            prefix = caps[:1, :plen].clone() if plen else None
            ids, ln = refshim.reference_generate_batch(ref, kind, imgs, labs, mode=mode, seed=7, caption=prefix,
                                                       max_len=max_len, beam_size=B, top_k=K, temperature=T_)
            gaps, traces = [], []
            oids, oln = model.generate_batch(kind, sd, hp, None, labs, max_len=max_len, encoded=o_enc, gaps=gaps,
                                             traces=traces,
                                             caption=prefix, beam_size=B, top_k=K, temperature=T_,
                                             noise=noise.Noise(mode, 7))
            agree = bool((ids == oids).all() and (ln == oln).all())
            n_eos = int((ids == 3).any(1).sum())
            print(f'  {tag} {kind} {mode} B={B} K={K} T={T_} prefix={plen}: oracle==reference {agree} '
                  f'min_gap {min(gaps):.2e} rows-with-eos {n_eos}/{n_img}')
            fx['gen'].append(dict(mode=mode, beam_size=B, top_k=K, temperature=T_, prefix_len=plen, noise_seed=7,
                                  ids=ids, lengths=ln, gaps=torch.tensor(gaps, dtype=torch.float64),
                                  abs_gaps=torch.tensor([t.abs_gap for t in traces], dtype=torch.float64),
                                  oracle_agrees=agree))
    print(f'{tag} {kind}: {time.time() - t0:.1f}s  pp={fx["perplexity"]:.4f}')
    torch.save(fx, os.path.join(OUT, f'{tag}_{kind}.pt'))


if __name__ == '__main__':
    assert refshim.available(), 'needs /root/reference'
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    which = sys.argv[1:] or ['small', 'canon', 'cfg1']
    for kind in synth_weights.KINDS:
        if 'small' in which:
            run(kind, 1000, True, 8, 14, 1, GEN_VARIANTS, 'small')
        if 'canon' in which:
            # BASELINE shapes (V = 36 541, 32 tokens): beam 5 / top-k 50 (configs 2, 5), top-k 50 sampling (config 4)
            run(kind, 36541, False, 32, 32, 0, [('deterministic', 5, 50, 1.0, 0), ('injected', 5, 50, 1.0, 0),
                                                 ('injected', 1, 50, 1.0, 0)], 'canon')
    if 'cfg1' in which:
        # BASELINE.json configs[0] exactly: CaptioningLSTM, 1-layer LSTMDecoder, emb 256, "greedy" = beam 1 / top-k 1
        run('lstm', 36541, False, 8, 32, 0, [('deterministic', 1, 1, 1.0, 0), ('injected', 1, 1, 1.0, 0)], 'cfg1',
            hp_update=dict(emb_dim=256, num_layers=1))
