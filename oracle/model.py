"""Functional torch-CPU fp32 restatement of the DeepHumor captioning path (TEST INFRASTRUCTURE).

All functions take a reference-layout ``state_dict`` (see ``oracle/weights.py``) and plain tensors.
``generate_*`` are strictly per image, like the reference; ``generate_batch`` loops.
Citations are ``/root/reference/deephumor/...`` file:line.
"""
import math

import torch
import torch.nn.functional as F

from .noise import CALL_FINAL, CALL_PRUNE, CALL_TOKEN, Noise
from deephumor_b200.utils.synth_weights import RESNET_BLOCKS

NEG_INF = float('-inf')


# ------------------------------------------------------------------------------ encoder
def _conv_bn(x, sd, conv, bn, stride, pad, relu):
    # torchvision resnet.py:143-163 (conv -> eval-mode BN (eps 1e-5) -> ReLU)
    y = F.conv2d(x, sd[conv + '.weight'], None, stride, pad)
    y = F.batch_norm(y, sd[bn + '.running_mean'], sd[bn + '.running_var'],
                     sd[bn + '.weight'], sd[bn + '.bias'], False, 0.0, 1e-5)
    return F.relu(y) if relu else y


def resnet50_trunk(sd, p, images, taps=None):
    """torchvision resnet50 children()[:-2] (models/encoders.py:34-38,56; resnet.py:266-279)."""
    x = _conv_bn(images, sd, p + '.0', p + '.1', 2, 3, True)
    if taps is not None:
        taps['stem'] = x
    x = F.max_pool2d(x, 3, 2, 1)
    if taps is not None:
        taps['pool'] = x
    for li, nblk in enumerate(RESNET_BLOCKS):
        for b in range(nblk):
            q = f'{p}.{4 + li}.{b}'
            stride = 2 if (b == 0 and li > 0) else 1          # stride on conv2 (resnet.py:109-110,135)
            y = _conv_bn(x, sd, q + '.conv1', q + '.bn1', 1, 0, True)
            y = _conv_bn(y, sd, q + '.conv2', q + '.bn2', stride, 1, True)
            y = _conv_bn(y, sd, q + '.conv3', q + '.bn3', 1, 0, False)
            idn = _conv_bn(x, sd, q + '.downsample.0', q + '.downsample.1', stride, 0, False) if b == 0 else x
            x = F.relu(y + idn)
            if taps is not None:
                taps[f'layer{li + 1}.{b}'] = x
    return x


def image_encoder(sd, p, images, spatial):
    """models/encoders.py:46-70.  Dropout is identity in eval mode (Q28)."""
    f = resnet50_trunk(sd, p + '.resnet', images)
    bs, dim = f.shape[:2]
    pooled = f.mean(dim=(2, 3))                                            # :60 AdaptiveAvgPool2d(1)
    emb = F.linear(pooled, sd[p + '.linear.weight'], sd[p + '.linear.bias'])
    emb = F.batch_norm(emb, sd[p + '.bn.running_mean'], sd[p + '.bn.running_var'],
                       sd[p + '.bn.weight'], sd[p + '.bn.bias'], False, 0.0, 1e-5)   # :61
    if not spatial:
        return emb
    tok = f.reshape(bs, dim, -1).transpose(2, 1)                           # :65-66 token = y*7+x
    sp = F.linear(tok, sd[p + '.linear.weight'], sd[p + '.linear.bias'])   # :67 shared Linear, no BN (Q20)
    return emb, sp


def label_encoder(sd, p, labels):
    """models/encoders.py:96-106: mean over the FULL label width, incl. EOS / pads (Q23)."""
    return sd[p + '.embedding.weight'][labels].mean(dim=1)


def image_label_encoder(sd, p, images, labels):
    """models/encoders.py:129-144."""
    ie = image_encoder(sd, p + '.image_encoder', images, False)
    le = label_encoder(sd, p + '.label_encoder', labels)
    return F.linear(torch.cat([ie, le], dim=1), sd[p + '.linear.weight'], sd[p + '.linear.bias'])


def encode(kind, sd, images, labels=None):
    """Returns (start_emb [N,E], enc_out [N,49,E] or None) for the four captioners (caption_models.py)."""
    if kind == 'lstm':
        return image_encoder(sd, 'encoder', images, False), None
    if kind == 'lstm_labels':
        return image_label_encoder(sd, 'encoder', images, labels), None
    if kind == 'xfmr_base':
        return image_encoder(sd, 'encoder', images, False), None
    if kind == 'xfmr':
        return image_encoder(sd, 'encoder', images, True)
    raise ValueError(kind)


# ------------------------------------------------------------------------------ LSTM
def lstm_cell(sd, p, l, x, h, c):
    """nn.LSTM cell: gate order i,f,g,o, both biases (rnn_models.py:23-24)."""
    g = (F.linear(x, sd[f'{p}.weight_ih_l{l}'], sd[f'{p}.bias_ih_l{l}'])
         + F.linear(h, sd[f'{p}.weight_hh_l{l}'], sd[f'{p}.bias_hh_l{l}']))
    i, f, gg, o = g.chunk(4, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def lstm_step(sd, p, L, x, h, c):
    """x [R,E]; h,c [L,R,H] -> out [R,H], new h,c."""
    hs, cs = [], []
    for l in range(L):
        h2, c2 = lstm_cell(sd, p, l, x, h[l], c[l])
        hs.append(h2)
        cs.append(c2)
        x = h2
    return x, torch.stack(hs), torch.stack(cs)


def lstm_forward(sd, hp, image_emb, captions, lengths=None):
    """LSTMDecoder.forward (rnn_models.py:28-46): packed LSTM == outputs zeroed at t >= length,
    trimmed to max(lengths) (Q24); classifier applied to the zero rows too."""
    L, H = hp['num_layers'], hp['hidden_size']
    emb = sd['decoder.embedding.weight'][captions]
    x = torch.cat([image_emb.unsqueeze(1), emb], dim=1)
    N, T = x.shape[:2]
    if lengths is None:
        lengths = torch.full((N,), T, dtype=torch.int64)
    h = torch.zeros(L, N, H)
    c = torch.zeros(L, N, H)
    outs = []
    tmax = int(lengths.max())
    for t in range(tmax):
        o, h, c = lstm_step(sd, 'decoder.lstm', L, x[:, t], h, c)
        outs.append(torch.where((t < lengths).unsqueeze(1), o, torch.zeros_like(o)))
    out = torch.stack(outs, dim=1)
    return F.linear(out, sd['decoder.classifier.weight'], sd['decoder.classifier.bias'])


# ------------------------------------------------------------------------------ transformer
def _mha(sd, p, xq, xkv, mask, n_heads):
    """MultiHeadAttentionLayer.forward (transformers.py:82-129); K/V viewed with the QUERY's seq_len (:94)."""
    bs, S, D = xq.shape
    hd = D // n_heads
    q = F.linear(xq, sd[p + '.fc_q.weight'], sd[p + '.fc_q.bias']).view(bs, S, n_heads, hd).permute(0, 2, 1, 3)
    k = F.linear(xkv, sd[p + '.fc_k.weight'], sd[p + '.fc_k.bias']).view(bs, S, n_heads, hd).permute(0, 2, 3, 1)
    v = F.linear(xkv, sd[p + '.fc_v.weight'], sd[p + '.fc_v.bias']).view(bs, S, n_heads, hd).permute(0, 2, 1, 3)
    e = (q @ k) / sd[p + '.scale']
    e = e.masked_fill(mask.unsqueeze(1), -1e8)                              # :111 (-1e8, not -inf; Q16)
    a = torch.softmax(e, dim=-1)
    x = (a @ v).permute(0, 2, 1, 3).reshape(bs, S, D)
    return F.linear(x, sd[p + '.fc_o.weight'], sd[p + '.fc_o.bias'])


def _ln(sd, p, x):
    return F.layer_norm(x, x.shape[-1:], sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def xfmr_hidden(sd, hp, cross, tokens, start_emb, enc_out=None):
    """Decoder stack up to (not incl.) the classifier.

    cross=False: SelfAttentionTransformerDecoder.forward (transformers.py:694-738).
    cross=True : TransformerDecoder.forward (:432-490) incl. padding of tokens and enc_out to
                 seq_len = max(T+1, 49) (Q15) and the any-zero-feature encoder mask (Q16).
    tokens [bs,T] int64, start_emb [bs,D] -> hidden [bs, S_out, D].
    """
    p = 'decoder'
    pad = hp['pad_index']
    bs, T = tokens.shape
    D = start_emb.shape[1]
    S = T + 1
    if cross:
        Se = enc_out.shape[1]
        S = max(S, Se)
        tokens = torch.cat([tokens, torch.full((bs, S - T - 1), pad, dtype=torch.int64)], dim=1)
        enc_out = torch.cat([enc_out, torch.zeros(bs, S - Se, D)], dim=1)
    x = torch.cat([start_emb.unsqueeze(1), sd[p + '.tok_embedding.weight'][tokens]], dim=1)
    x = x / sd[p + '.scale']                                                # division, image slot too (Q18)
    x = x + sd[p + '.pos_embedding.weight'][:S].unsqueeze(0)                # IndexError-equivalent if S > rows (Q19)
    if S > sd[p + '.pos_embedding.weight'].shape[0]:
        raise IndexError('index out of range in self')
    ids = torch.cat([torch.ones(bs, 1, dtype=torch.int64), tokens], dim=1)  # dummy id 1 at the image slot (Q17)
    key_pad = (ids == pad).unsqueeze(1).expand(bs, S, S)
    causal = torch.triu(torch.ones(S, S), 1).bool().unsqueeze(0)
    in_mask = key_pad | causal
    if cross:
        enc_live = (enc_out != 0.).all(dim=-1).long()                       # :480
        enc_mask = (enc_live == pad).unsqueeze(1).expand(bs, S, S)          # :481 via get_pad_mask
    for l in range(hp['n_layers']):
        q = f'{p}.layers.{l}'
        x = _ln(sd, q + '.self_attn_ln', x + _mha(sd, q + '.self_attn', x, x, in_mask, hp['n_heads']))
        if cross:
            x = _ln(sd, q + '.enc_attn_ln', x + _mha(sd, q + '.enc_attn', x, enc_out, enc_mask, hp['n_heads']))
        ff = F.linear(F.relu(F.linear(x, sd[q + '.pf.fc_1.weight'], sd[q + '.pf.fc_1.bias'])),
                      sd[q + '.pf.fc_2.weight'], sd[q + '.pf.fc_2.bias'])
        x = _ln(sd, q + '.pf_ln', x + ff)
    return x


def xfmr_forward(sd, hp, cross, tokens, start_emb, enc_out=None):
    h = xfmr_hidden(sd, hp, cross, tokens, start_emb, enc_out)
    return F.linear(h, sd['decoder.classifier.weight'], sd['decoder.classifier.bias'])


def forward(kind, sd, hp, images, captions, lengths=None, labels=None):
    """Captioning*.forward (caption_models.py:42-46,138-142,259-272,393-406)."""
    start, enc = encode(kind, sd, images, labels)
    if kind in ('lstm', 'lstm_labels'):
        return lstm_forward(sd, hp, start, captions, lengths)
    return xfmr_forward(sd, hp, kind == 'xfmr', captions, start, enc)


def perplexity(logits, targets, lengths, pad_index=0):
    """experiments/metrics.py:4-9 (divide by length BEFORE zeroing pads, Q27)."""
    lv = logits.log_softmax(-1).gather(-1, targets.unsqueeze(-1)).squeeze(-1)
    lv = lv / lengths.unsqueeze(1)
    lv = torch.where(targets == pad_index, torch.zeros_like(lv), lv)
    return (-lv.sum(dim=-1)).exp().mean()


# ------------------------------------------------------------------------------ selection algebra
class Trace:
    """Collects decision margins of a generation (near-tie policy, Appendix D.5) and, optionally, the beam state
    after every step.

    min_gap   smallest RELATIVE margin between consecutive race scores among the first k+1 (fp32 check-mode policy).
    abs_gap   smallest ABSOLUTE margin in logit units: |log s_i - log s_j| * T between consecutive race scores
              among the first k+1 (what a logit perturbation has to exceed to reorder a draw), and the gap between the
              k-th and (k+1)-th largest logit of a top-k filter whenever the boundary token could reach the draw.
    step_abs  {step: smallest abs_gap among that step's decisions}; states {step: (seq, val, ended)} when keep_states.
    """

    def __init__(self, keep_states=False):
        self.min_gap = float('inf')
        self.abs_gap = float('inf')
        self.steps = 0
        self.error = None
        self.step = None
        self.step_abs = {}
        self.states = {} if keep_states else None

    def at(self, step):
        self.step = step

    def _abs(self, g):
        self.abs_gap = min(self.abs_gap, g)
        if self.step is not None:
            self.step_abs[self.step] = min(self.step_abs.get(self.step, float('inf')), g)

    def see(self, score_sorted, k, temperature=1.0):
        s = score_sorted[..., :k + 1].double()
        if s.shape[-1] < 2:
            return
        hi, lo = s[..., :-1], s[..., 1:]
        denom = hi.abs().clamp_min(1e-30)
        gap = ((hi - lo) / denom)
        live = hi > 0
        gap = gap[live] if live.any() else gap
        if gap.numel():
            self.min_gap = min(self.min_gap, float(gap.min()))
        if live.any():
            lg = (hi[live].log() - lo[live].clamp_min(1e-300).log()) * temperature
            self._abs(float(lg.min()))

    FILTER_WINDOW = 64

    def see_filter(self, logits, top_k, unk, B, temperature, q):
        """Margin of the top-k filter (beam.py:32-37).  Membership of token j flips when a perturbation of the logits
        carries it across the filter boundary: a member at rank r < k leaves once the (k+1)-th largest logit overtakes it
        (distance x_j - x_(k+1)); an outsider enters once it overtakes the k-th largest (distance x_(k) - x_j).  With
        near-tied logits MANY tokens around rank k are that close, not only the two boundary ones, so every token within
        FILTER_WINDOW ranks of the boundary is examined.  A flip only changes the outcome if the token can be drawn, i.e.
        its (hypothetical) race score reaches the first B+1 of the row."""
        V = logits.shape[-1]
        if V <= top_k:
            return
        w = min(V, top_k + self.FILTER_WINDOW)
        top = torch.topk(logits, w, dim=-1)
        for r in range(logits.shape[0]):
            x = logits[r].double()
            kth, k1th = top.values[r, top_k - 1].double(), top.values[r, top_k].double()
            fl = x.clone()
            fl[x < kth] = NEG_INF
            fl[unk] = NEG_INF
            sc = torch.softmax(fl / temperature, -1)
            z = torch.exp((x - x.max()) / temperature)
            zs = float(z[fl > NEG_INF].sum())
            qq = torch.ones_like(x) if q is None else q[r].double()
            race = sc / qq
            cut = float(torch.topk(race, min(B + 1, race.numel())).values[-1])
            idx = top.indices[r, max(0, top_k - self.FILTER_WINDOW):]
            xs = x[idx]
            inside = xs >= kth
            dist = torch.where(inside, xs - k1th, kth - xs)
            hyp = z[idx] / zs / qq[idx]                             # race score if the token is (or were) inside the filter
            reach = (hyp >= cut) & (idx != unk)
            if bool(reach.any()):
                self._abs(float(dist[reach].min()))

    def state(self, step, seq, val, ended):
        if self.states is not None:
            self.states[step] = (seq.clone(), val.clone(), ended.clone())


def filter_top_k(logits, top_k, unk):
    """beam.py:32-37: strict '<' keeps ties; <unk> always masked but still counted toward the k-th value (Q3)."""
    kth = torch.topk(logits, top_k, dim=-1).values[:, -1:]
    drop = logits < kth
    drop[:, unk] = True
    return logits.masked_fill(drop, NEG_INF)


def draw(values, k, temperature, q, trace=None):
    """beam.py:39-48 with multinomial(p,k) == topk(p / q, k) (Q2); q None -> deterministic (q == 1).
    Raises like torch.multinomial when the row is all -inf (Q3)."""
    p = torch.softmax(values / temperature, dim=-1)
    if torch.isnan(p).any() or (p.sum(-1) <= 0).any():
        raise RuntimeError('invalid multinomial distribution (sum of probabilities <= 0)')
    score = p if q is None else p / q
    srt = torch.sort(score, dim=-1, descending=True, stable=True)
    if trace is not None:
        trace.see(srt.values, k, temperature)
    return srt.indices[..., :k]


def select_tokens(logits, B, T, top_k, unk, q, trace):
    """filter -> draw B per row -> log_softmax over the B picked raw logits (Q4).  logits [R,V]."""
    if trace is not None:
        trace.see_filter(logits, top_k, unk, B, T, q)
    fl = filter_top_k(logits, top_k, unk)
    ind = draw(fl, B, T, q, trace)
    val = torch.gather(fl, 1, ind).log_softmax(-1)
    return ind, val


def expand_candidates(ind, val, has_ended, eos):
    """beam.py:83-102 (Appendix A.1): ended row -> 1 candidate (token 0, dscore 0), live row -> B."""
    R, B = ind.shape
    parent, tok, dv, ended = [], [], [], []
    for r in range(R):
        if has_ended[r]:
            parent.append(r); tok.append(0); dv.append(0.0); ended.append(True)
        else:
            for j in range(B):
                t = int(ind[r, j])
                parent.append(r); tok.append(t); dv.append(float(val[r, j])); ended.append(t == eos)
    return (torch.tensor(parent), torch.tensor(tok), torch.tensor(dv, dtype=torch.float32),
            torch.tensor(ended))


# ------------------------------------------------------------------------------ generation (per image)
def generate_lstm(sd, hp, image_emb, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50,
                  eos_index=3, unk_index=1, noise=None, image_index=0, trace=None):
    """LSTMDecoder.generate (rnn_models.py:48-143; Appendix A.2).  image_emb [1,E]; caption [1,p] or None."""
    assert beam_size <= top_k, '`beam_size` should be less than `top_k`'
    noise = noise or Noise()
    B, L, H = beam_size, hp['num_layers'], hp['hidden_size']
    V = sd['decoder.classifier.weight'].shape[0]
    W, bW = sd['decoder.classifier.weight'], sd['decoder.classifier.bias']
    x = image_emb
    if caption is not None:
        x = torch.cat([image_emb, sd['decoder.embedding.weight'][caption[0]]], dim=0)
    h = torch.zeros(L, 1, H)
    c = torch.zeros(L, 1, H)
    for t in range(x.shape[0]):
        out, h, c = lstm_step(sd, 'decoder.lstm', L, x[t:t + 1], h, c)
    logits = F.linear(out, W, bW)
    h, c = h.repeat(1, B, 1), c.repeat(1, B, 1)                              # :84
    p0 = 0 if caption is None else caption.shape[1]
    if trace is not None:
        trace.at(p0)
    ind, val = select_tokens(logits, B, temperature, top_k, unk_index,
                             noise.q(image_index, p0, CALL_TOKEN, 1, V), trace)
    last, val = ind[0].clone(), val[0].clone()
    seq = last.unsqueeze(1)
    if caption is not None:
        seq = torch.cat([caption.repeat(B, 1), seq], dim=1)
    ended = last == eos_index                                                # :103 (Q9)
    if trace is not None:
        trace.state(p0, seq, val, ended)
    for i in range(seq.shape[1], max_len):                                   # :105 (Q11)
        if trace is not None:
            trace.at(i)
        out, h, c = lstm_step(sd, 'decoder.lstm', L, sd['decoder.embedding.weight'][last], h, c)
        logits = F.linear(out, W, bW)
        ind, nv = select_tokens(logits, B, temperature, top_k, unk_index,
                                noise.q(image_index, i, CALL_TOKEN, B, V), trace)
        parent, tok, dv, cend = expand_candidates(ind, nv, ended, eos_index)
        cval = val[parent] + dv
        f = draw(cval.unsqueeze(0), B, temperature, noise.q(image_index, i, CALL_PRUNE, 1, len(cval)), trace)[0]
        val = cval[f]
        seq = torch.cat([seq[parent], tok.unsqueeze(1)], dim=1)[f]
        last = seq[:, -1]
        ended = cend[f]
        if trace is not None:
            trace.steps = i
            trace.state(i, seq, val, ended)
        if bool(ended.all()):                                                # :131
            break
        sp = f // B                                                          # :135-137 misaligned parent (Q8)
        h, c = h[:, sp], c[:, sp]
    if trace is not None:
        trace.at(max_len + 1)
    pick = draw(val.unsqueeze(0), 1, temperature, noise.q(image_index, max_len + 1, CALL_FINAL, 1, B), trace)[0, 0]
    return seq[pick]                                                         # :140-143 (Q13)


def generate_xfmr(sd, hp, cross, start_emb, enc_out=None, caption=None, max_len=25, temperature=1.0,
                  beam_size=10, top_k=50, eos_index=3, unk_index=1, noise=None, image_index=0, trace=None,
                  faithful_cost=False):
    """(SelfAttention)TransformerDecoder.generate (transformers.py:492-579, 740-825; Appendix A.3).
    faithful_cost=True also runs the classifier over every position like the reference (:488,:736) -- used
    only when this port is timed as the CPU baseline; the selected logits row is the same."""
    assert beam_size <= top_k, '`beam_size` should be less than `top_k`'
    noise = noise or Noise()
    B = beam_size
    pad = hp['pad_index']
    W, bW = sd['decoder.classifier.weight'], sd['decoder.classifier.bias']
    V = W.shape[0]

    def logits_at(seq, start, enc, pos):
        hid = xfmr_hidden(sd, hp, cross, seq, start, enc)
        if faithful_cost:
            return F.linear(hid, W, bW)[:, pos]
        return F.linear(hid[:, pos], W, bW)

    seq = torch.full((1, max_len), pad, dtype=torch.int64)
    p0 = 0
    if caption is not None:
        p0 = caption.shape[1]
        seq[:, :p0] = caption
    logits = logits_at(seq, start_emb, enc_out, p0)
    if trace is not None:
        trace.at(p0)
    ind, val = select_tokens(logits, B, temperature, top_k, unk_index,
                             noise.q(image_index, p0, CALL_TOKEN, 1, V), trace)
    val = val[0].clone()
    seq = seq.repeat(B, 1)
    seq[:, p0] = ind[0]
    start_b = start_emb.repeat(B, 1)
    enc_b = enc_out.repeat(B, 1, 1) if cross else None
    ended = torch.zeros(B, dtype=torch.bool)                                 # NOT initialised from tokens (Q9)
    if trace is not None:
        trace.state(p0, seq, val, ended)
    i = p0 + 1
    for i in range(p0 + 1, max_len + 1):                                     # inclusive upper bound (Q10)
        if trace is not None:
            trace.at(i)
        logits = logits_at(seq, start_b, enc_b, i)
        ind, nv = select_tokens(logits, B, temperature, top_k, unk_index,
                                noise.q(image_index, i, CALL_TOKEN, B, V), trace)
        parent, tok, dv, cend = expand_candidates(ind, nv, ended, eos_index)
        cseq = seq[parent].clone()
        if i < max_len:
            cseq[:, i] = tok                                                 # no-op column at i == max_len (Q10)
        cval = val[parent] + dv
        f = draw(cval.unsqueeze(0), B, temperature, noise.q(image_index, i, CALL_PRUNE, 1, len(cval)), trace)[0]
        val, seq, ended = cval[f], cseq[f], cend[f]
        if trace is not None:
            trace.steps = i
            trace.state(i, seq, val, ended)
        if bool(ended.all()):
            break
    if trace is not None:
        trace.at(max_len + 1)
    pick = draw(val.unsqueeze(0), 1, temperature, noise.q(image_index, max_len + 1, CALL_FINAL, 1, B), trace)[0, 0]
    return seq[pick, :i]


def generate(kind, sd, hp, image, label=None, caption=None, image_index=0, encoded=None, **kw):
    """Captioning*.generate for ONE image [1,3,H,W] (caption_models.py:48-74,144-171,274-300,408-434).
    encoded = precomputed (start_emb [1,E], enc_out [1,49,E] | None) skips the encoder."""
    start, enc = encoded if encoded is not None else encode(kind, sd, image, label)
    if kind in ('lstm', 'lstm_labels'):
        kw.pop('faithful_cost', None)
        return generate_lstm(sd, hp, start, caption, image_index=image_index, **kw)
    return generate_xfmr(sd, hp, kind == 'xfmr', start, enc, caption, image_index=image_index, **kw)


def generate_batch(kind, sd, hp, images, labels=None, first_index=0, pad_index=0, max_len=25, encoded=None,
                   gaps=None, traces=None, **kw):
    """Batched oracle = python loop over images (Appendix D.4): ids [N,max_len] padded + lengths [N].
    encoded = (start [N,E], enc [N,49,E] | None) from ``encode``; gaps = list receiving per-image min relative margins;
    traces = list receiving each image's Trace (per-step beam states and absolute margins)."""
    N = images.shape[0] if images is not None else encoded[0].shape[0]
    ids = torch.full((N, max_len), pad_index, dtype=torch.int64)
    lens = torch.zeros(N, dtype=torch.int64)
    for n in range(N):
        lab = None if labels is None else labels[n:n + 1]
        enc_n = None
        if encoded is not None:
            enc_n = (encoded[0][n:n + 1], None if encoded[1] is None else encoded[1][n:n + 1])
        tr = Trace(keep_states=traces is not None) if (gaps is not None or traces is not None) else None
        s = generate(kind, sd, hp, None if images is None else images[n:n + 1], lab, image_index=first_index + n,
                     max_len=max_len, encoded=enc_n, trace=tr, **kw)
        if gaps is not None:
            gaps.append(tr.min_gap)
        if traces is not None:
            traces.append(tr)
        s = s.reshape(-1)
        ids[n, :len(s)] = s
        lens[n] = len(s)
    return ids, lens
