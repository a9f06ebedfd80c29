"""GPU debug: which part of the tensor-core decode diverges from the reference fixture ids (fused stride 8 / stride 1 /
materialised logits / unfused LSTM), per image: first differing token position against the fixture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200 import models
from deephumor_b200.runtime import ops
from tests import helpers as H

CLS = {'lstm': models.CaptioningLSTM, 'lstm_labels': models.CaptioningLSTMWithLabels,
       'xfmr_base': models.CaptioningTransformerBase, 'xfmr': models.CaptioningTransformer}


def first_diff(a, b):
    d = (a != b).nonzero()
    return int(d[0]) if d.numel() else None


def run(kind, variant, n):
    fx = H.load_fixture('canon', kind)
    sd, imgs, labs, caps, lens = H.fixture_inputs(fx)
    g = fx['gen'][variant]
    kw = dict(max_len=fx['max_len'], temperature=g['temperature'], beam_size=g['beam_size'], top_k=g['top_k'],
              noise=g['mode'], seed=g['noise_seed'])
    res = {}
    for name, env in (('stride8', {}), ('stride1', {'DH_VOCAB_STRIDE': '1'}), ('materialised', {'fused': False}),
                      ('nofusedlstm', {'lstm': False}), ('fp32', {'precision': 'fp32'})):
        os.environ.pop('DH_VOCAB_STRIDE', None)
        if 'DH_VOCAB_STRIDE' in env:
            os.environ['DH_VOCAB_STRIDE'] = env['DH_VOCAB_STRIDE']
        ops.FUSED_VOCAB = env.get('fused', True)
        ops.FUSED_LSTM = env.get('lstm', True)
        m = CLS[kind](**fx['hp'])
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval().set_precision(env.get('precision', 'bf16'))
        with torch.no_grad():
            a = (imgs[:n].cuda(), labs[:n].cuda()) if kind == 'lstm_labels' else (imgs[:n].cuda(),)
            ids, ln = m.generate(*a, **kw)
        res[name] = ids.cpu()
        ops.FUSED_VOCAB = ops.FUSED_LSTM = True
    ref = g['ids'][:n]
    print(f'== {kind} variant {variant} (B={g["beam_size"]}, top_k={g["top_k"]}, {g["mode"]}): first differing position vs the fixture')
    for i in range(n):
        row = {k: first_diff(v[i], ref[i]) for k, v in res.items()}
        if any(v is not None for v in row.values()):
            print(f'  image {i:2d} gap {float(g["gaps"][i]):.1e}: ' + ', '.join(f'{k}={v}' for k, v in row.items()))


for kind in sys.argv[1].split(','):
    for variant in (1, 2):
        run(kind, variant, 32 if kind.startswith('lstm') else 16)
