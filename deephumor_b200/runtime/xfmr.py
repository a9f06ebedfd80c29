"""Transformer decoder runtime: incremental (KV-cached) stochastic-beam generation and teacher-forced forward.

Reference: models/transformers.py:432-490 / 694-738 (forward), :492-579 / 740-825 (generate);
SURVEY.md Appendix A.3.  The reference re-runs the whole decoder every step; logits at position i depend only
on tokens < i (Q14), so each step here processes ONE new position per row against a KV cache.  Beam reorder
never copies the cache: a per-image slot table maps (beam, position) -> physical slot (dh_beam_step).
Cross-attention K/V over the 49 spatial tokens are projected once per image and shared by its beams.
"""
import torch

from . import ops


class XfmrDecoderRT:
    def __init__(self, sd, prefix, hp, cross, dtype, device):
        self.dtype, self.device, self.cross = dtype, device, cross
        self.D, self.n_heads, self.L = hp['hid_dim'], hp['n_heads'], hp['n_layers']
        self.pad = hp['pad_index']
        to = lambda t, dt=dtype: t.float().contiguous().to(device=device, dtype=dt)
        f32 = torch.float32
        self.tok, self.pos = to(sd[prefix + '.tok_embedding.weight']), to(sd[prefix + '.pos_embedding.weight'])
        self.V = self.tok.shape[0]
        self.scale = float(sd[prefix + '.scale'])
        self.layers = []
        for l in range(self.L):
            q = f'{prefix}.layers.{l}'
            lay = {}
            for att in (('self_attn', 'enc_attn') if cross else ('self_attn',)):
                for n in ('fc_q', 'fc_k', 'fc_v', 'fc_o'):
                    lay[f'{att}.{n}.w'] = to(sd[f'{q}.{att}.{n}.weight'])
                    lay[f'{att}.{n}.b'] = to(sd[f'{q}.{att}.{n}.bias'], f32)
                lay[f'{att}.scale'] = float(sd[f'{q}.{att}.scale'])
                lay[f'{att}_ln.g'] = to(sd[f'{q}.{att}_ln.weight'], f32)
                lay[f'{att}_ln.b'] = to(sd[f'{q}.{att}_ln.bias'], f32)
            for n in ('fc_1', 'fc_2'):
                lay[f'pf.{n}.w'], lay[f'pf.{n}.b'] = to(sd[f'{q}.pf.{n}.weight']), to(sd[f'{q}.pf.{n}.bias'], f32)
            lay['pf_ln.g'], lay['pf_ln.b'] = to(sd[f'{q}.pf_ln.weight'], f32), to(sd[f'{q}.pf_ln.bias'], f32)
            if dtype != torch.float32 and self.D % 128 == 0:
                # tensor-core mode: Q | K | V of the new position in ONE contraction (dh_gemm_tc_split3)
                lay['self_attn.qkv.w'] = torch.cat([lay[f'self_attn.fc_{n}.w'] for n in 'qkv']).contiguous()
                lay['self_attn.qkv.b'] = torch.cat([lay[f'self_attn.fc_{n}.b'] for n in 'qkv']).contiguous()
            self.layers.append(lay)
        self.pf = self.layers[0]['pf.fc_1.w'].shape[0]
        self.Wc, self.bc = to(sd[prefix + '.classifier.weight']), to(sd[prefix + '.classifier.bias'], f32)
        self.ldv = (self.V + 3) // 4 * 4
        self._plans = ops.PlanCache()

    # ------------------------------------------------------------------ shared layer pieces
    def _cross_kv(self, spatial, n_img):
        """K/V of the 49 spatial tokens per layer: [n_img*49, D] each + the any-zero-feature mask (Q16)."""
        kv = []
        for lay in self.layers:
            K = torch.empty(n_img * 49, self.D, dtype=self.dtype, device=self.device)
            V = torch.empty_like(K)
            ops.gemm(spatial, lay['enc_attn.fc_k.w'], K, bias=lay['enc_attn.fc_k.b'])
            ops.gemm(spatial, lay['enc_attn.fc_v.w'], V, bias=lay['enc_attn.fc_v.b'])
            kv.append((K, V))
        mask = torch.empty(n_img * 49, dtype=torch.uint8, device=self.device)
        ops.enc_mask(spatial, mask)
        return kv, mask

    def _post_attn(self, lay, att, x, attn, tmp, rows):
        """x <- LN(x + fc_o(attn))."""
        D, es = self.D, x.element_size()
        if ops.gemm_ln_supported(x, D):
            # fc_o + residual + LayerNorm in one launch, in place on x (dh_gemm_tc_ln)
            with ops.PROFILE.range('xfmr_proj_ln', 2.0 * rows * D * D):
                ops.gemm_ln(attn[:rows], lay[f'{att}.fc_o.w'], lay[f'{att}.fc_o.b'], x[:rows], lay[f'{att}_ln.g'],
                            lay[f'{att}_ln.b'], x[:rows])
            return
        with ops.PROFILE.range('xfmr_proj', 2.0 * rows * D * D):
            ops.gemm(attn[:rows], lay[f'{att}.fc_o.w'], tmp[:rows], bias=lay[f'{att}.fc_o.b'], residual=x[:rows])
        with ops.PROFILE.range('xfmr_layernorm', nbytes=2.0 * rows * D * es):
            ops.add_layernorm(tmp[:rows], None, lay[f'{att}_ln.g'], lay[f'{att}_ln.b'], x[:rows])

    def _ffn(self, lay, x, h1, tmp, rows):
        D, es = self.D, x.element_size()
        if ops.gemm_ln_supported(x, D):
            with ops.PROFILE.range('xfmr_ffn', 4.0 * rows * D * self.pf):
                ops.gemm(x[:rows], lay['pf.fc_1.w'], h1[:rows], bias=lay['pf.fc_1.b'], relu=True)
                ops.gemm_ln(h1[:rows], lay['pf.fc_2.w'], lay['pf.fc_2.b'], x[:rows], lay['pf_ln.g'], lay['pf_ln.b'], x[:rows])
            return
        with ops.PROFILE.range('xfmr_ffn', 4.0 * rows * D * self.pf):
            ops.gemm(x[:rows], lay['pf.fc_1.w'], h1[:rows], bias=lay['pf.fc_1.b'], relu=True)
            ops.gemm(h1[:rows], lay['pf.fc_2.w'], tmp[:rows], bias=lay['pf.fc_2.b'], residual=x[:rows])
        with ops.PROFILE.range('xfmr_layernorm', nbytes=2.0 * rows * D * es):
            ops.add_layernorm(tmp[:rows], None, lay['pf_ln.g'], lay['pf_ln.b'], x[:rows])

    # ------------------------------------------------------------------ generation
    def _decode(self, pl, p0, max_len, temperature, B, top_k, eos_index, unk_index, noise_mode):
        """Whole incremental decode on the static buffers of plan `pl` (no host sync: CUDA-graph capturable)."""
        N, D = pl['N'], self.D
        R, S = N * B, max_len + 1
        start_emb, spatial, caption = pl['start'], pl['spatial'], pl['caption']
        x, qb, attn, tmp, h1, logits = pl['x'], pl['qb'], pl['attn'], pl['tmp'], pl['h1'], pl['logits']
        Kc, Vc, beam, ind, val, dyn = pl['Kc'], pl['Vc'], pl['beam'], pl['ind'], pl['val'], pl['dynw'].dev
        beam.status.zero_()
        with ops.PROFILE.range('xfmr_cross_kv', 4.0 * N * 49 * D * D * self.L if self.cross else 0.0):
            xkv, emask = self._cross_kv(spatial, N) if self.cross else (None, None)

        vsel = pl['vsel']

        def select(rows, rpi, step_i, done, beam_step=False):
            """classifier + BeamSearchHelper selection (transformers.py:488/736 -> beam.py:32-53): fused two-pass vocab
            projection in tensor-core mode (logits never stored), materialised fp32 logits in check mode."""
            if vsel is not None:
                vsel.run(x[:rows], self.Wc, self.bc, B, temperature, unk_index, rpi, noise_mode, step_i, done, ind, val,
                         beam.status, dyn, beam_step=(beam, max_len, eos_index, False) if beam_step else None)
            else:
                with ops.PROFILE.range('vocab_gemm', 2.0 * rows * self.V * D):
                    ops.gemm(x[:rows], self.Wc, logits[:rows, :self.V], bias=self.bc)
                with ops.PROFILE.range('select_beam'):
                    ops.select_tokens(logits[:rows, :self.V], self.V, B, top_k, temperature, unk_index, rpi, noise_mode,
                                      0, 0, step_i, done, ind, val, beam.status, dyn)
                    if beam_step:
                        beam.step(ind, val, step_i, max_len, eos_index, False, temperature, noise_mode, 0, 0, dyn)

        use_path = (ops.PATH_ENTRIES and not ops.PROFILE.on and all('self_attn.qkv.w' in lay for lay in self.layers)
                    and self.L <= 8 and ops.FUSED_LN)
        if use_path:
            if getattr(self, '_ctx', None) is None:
                self._ctx = ops.xfmr_ctx(self)
            bufs = ops.xfmr_buffers(x, qb, attn, tmp, h1, Kc, Vc, xkv, emask, start_emb, B, S)

        def step(rows, rpi, pos, tokens, seq, src):
            """One new position `pos` for `rows` rows (rpi rows per image)."""
            if use_path:                                           # the whole stack behind one path-level C entry
                ops.xfmr_step(self._ctx, bufs, rows, rpi, pos, tokens, seq, src, self.L, self.cross)
                return
            ops.xfmr_embed(self.tok, self.pos, start_emb, rpi, tokens, None, pos, self.scale, x[:rows])
            slot_stride = (B if rpi == 1 else 1)                 # prefix phase writes slot 0 of each image
            for l, lay in enumerate(self.layers):
                kdst = Kc[l].view(R * S, D)[pos::S * slot_stride][:rows]
                vdst = Vc[l].view(R * S, D)[pos::S * slot_stride][:rows]
                es = x.element_size()
                with ops.PROFILE.range('xfmr_proj', 6.0 * rows * D * D):
                    if 'self_attn.qkv.w' in lay:
                        # one contraction; the K / V rows of this position land straight in their cache slots
                        ops.gemm_split3(x[:rows], lay['self_attn.qkv.w'], lay['self_attn.qkv.b'], (qb[:rows], kdst, vdst))
                    else:
                        ops.gemm(x[:rows], lay['self_attn.fc_q.w'], qb[:rows], bias=lay['self_attn.fc_q.b'])
                        ops.gemm(x[:rows], lay['self_attn.fc_k.w'], kdst, bias=lay['self_attn.fc_k.b'])
                        ops.gemm(x[:rows], lay['self_attn.fc_v.w'], vdst, bias=lay['self_attn.fc_v.b'])
                # self-attention over the cache: every row reads its pos + 1 cached K and V rows once (1 KB each at D = 512)
                with ops.PROFILE.range('xfmr_self_attn', nbytes=2.0 * rows * (pos + 1) * D * es + 2.0 * rows * D * es):
                    ops.attention(qb[:rows], Kc[l], Vc[l], attn[:rows], self.n_heads, rpi, B, S, lay['self_attn.scale'],
                                  src=src, slot_shared=(rpi == 1), n_keys=pos + 1, seq=seq, seq_per_image=False,
                                  pad=self.pad)
                self._post_attn(lay, 'self_attn', x, attn, tmp, rows)
                if self.cross:
                    with ops.PROFILE.range('xfmr_proj', 2.0 * rows * D * D):
                        ops.gemm(x[:rows], lay['enc_attn.fc_q.w'], qb[:rows], bias=lay['enc_attn.fc_q.b'])
                    # cross-attention: an image's 49 K and V rows are read once per step for all of its beam rows
                    with ops.PROFILE.range('xfmr_cross_attn', nbytes=2.0 * (rows // rpi) * 49 * D * es + 2.0 * rows * D * es):
                        ops.attention(qb[:rows], xkv[l][0], xkv[l][1], attn[:rows], self.n_heads, rpi, 1, 49,
                                      lay['enc_attn.scale'], slot_shared=True, n_keys=49, enc_mask=emask)
                    self._post_attn(lay, 'enc_attn', x, attn, tmp, rows)
                self._ffn(lay, x, h1, tmp, rows)

        # ---- prefix phase: positions 0..p0, one row per image (transformers.py:517-529)
        for t in range(p0 + 1):
            tok = None if t == 0 else caption[:, t - 1].contiguous()
            step(N, 1, t, tok, caption, None)
        select(N, 1, p0, None)
        beam.init(ind, val, caption, eos_index, False)
        ops.trace_beam(p0, beam)
        # ---- beam phase: i = p0+1 .. max_len inclusive (Q10); fixed trip count, frozen-at-break on the device
        for i in range(p0 + 1, max_len + 1):
            step(R, B, i, beam.last_tok, beam.seq, beam.src)
            select(R, B, i, beam.done, beam_step=True)
            ops.trace_beam(i, beam)
        beam.final(temperature, noise_mode, 0, 0, max_len + 1, max_len, self.pad, max_len, pl['ids'], pl['lens'], dyn)

    def generate(self, *args, **kw):
        """Optimistic first try with a sampled pass 1 of the fused vocab projection; if a candidate list overflowed
        (status bit 2) the whole generation is redone with the exhaustive pass 1, which cannot overflow."""
        out = self._generate(*args, robust=False, **kw)
        if out[3] and int(out[2].item()) & 2:
            out = self._generate(*args, robust=True, **kw)
        return out[:3]

    def _generate(self, start_emb, spatial, caption, max_len, temperature, beam_size, top_k, eos_index, unk_index,
                  noise_mode, seed, image_base, robust=False):
        """start_emb fp32 [N,D]; spatial [N*49,D] (cross) or None; caption int32 [N or 1,p] or None.
        The decode is captured once per configuration into a CUDA graph over static buffers and replayed."""
        N, B, D, dev, dt = start_emb.shape[0], beam_size, self.D, self.device, self.dtype
        if N == 0:                                   # empty batch: nothing to launch
            z = lambda *sh: torch.zeros(*sh, dtype=torch.int64, device=dev)
            return z(0, max_len), z(0), torch.zeros(1, dtype=torch.int32, device=dev), False
        R = N * B
        p0 = 0 if caption is None else caption.shape[1]
        S = max_len + 1                                          # cached positions 0..max_len
        need = max(S, 49) if self.cross else S                   # reference pads to max(T+1, 49) (Q15, Q19)
        if need > self.pos.shape[0]:
            raise IndexError('index out of range in self')       # what nn.Embedding raises in the reference (Q19)
        key = (N, B, p0, max_len, float(temperature), top_k, eos_index, unk_index, noise_mode, robust)
        pl = self._plans.get(key)
        if pl is None:
            rows_alloc = max(R, N)
            mk = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype, device=dev)
            fused = ops.FUSED_VOCAB and ops.VocabSelect.supported(self.Wc, self.V, top_k)
            pl = dict(N=N, vsel=ops.VocabSelect(rows_alloc, self.V, top_k, dev, stride=1 if robust else None) if fused else None, x=mk(rows_alloc, D), qb=mk(rows_alloc, D), attn=mk(rows_alloc, D), tmp=mk(rows_alloc, D),
                      h1=mk(rows_alloc, self.pf), logits=None if fused else mk(rows_alloc, self.ldv, dtype=torch.float32),
                      Kc=[torch.zeros(R, S, D, dtype=dt, device=dev) for _ in range(self.L)],
                      Vc=[torch.zeros(R, S, D, dtype=dt, device=dev) for _ in range(self.L)],
                      beam=ops.Beam(N, B, max_len, dev, kv_slots=S),
                      ind=mk(R, B, dtype=torch.int32), val=mk(R, B, dtype=torch.float32),
                      dynw=ops.DynWords(dev),
                      start=mk(N, D, dtype=torch.float32), spatial=mk(N * 49, D) if self.cross else None,
                      caption=None if caption is None else mk(N, p0, dtype=torch.int32),
                      ids=mk(N, max_len, dtype=torch.int64), lens=mk(N, dtype=torch.int64), graph=None)
            self._plans.put(key, pl)
        pl['start'].copy_(start_emb)
        if self.cross:
            pl['spatial'].copy_(spatial)
        if caption is not None:
            pl['caption'].copy_(caption.expand(N, p0))
        pl['dynw'].set(seed, image_base)
        args = (pl, p0, max_len, temperature, B, top_k, eos_index, unk_index, noise_mode)
        if ops.PROFILE.on or not ops.USE_GRAPHS or ops.TRACE is not None:
            self._decode(*args)
        else:
            if pl['graph'] is None:
                self._decode(*args)                     # eager warm-up (lazy one-time initialisation in the library)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._decode(*args)
                pl['graph'] = g
            pl['graph'].replay()
        sampled = pl['vsel'] is not None and pl['vsel'].stride > 1
        return pl['ids'].clone(), pl['lens'].clone(), pl['beam'].status.clone(), sampled

    # ------------------------------------------------------------------ teacher-forced forward
    def hidden(self, start_emb, spatial, captions):
        """captions int64 [N,T] -> hidden [N*S, D] with S = T+1 (Base) or max(T+1, 49) (cross; Q15)."""
        N, T = captions.shape
        D, dev, dt = self.D, self.device, self.dtype
        S = max(T + 1, 49) if self.cross else T + 1
        if S > self.pos.shape[0]:
            raise IndexError('index out of range in self')
        toks = torch.full((N, S), self.pad, dtype=torch.int32, device=dev)     # column s = token feeding position s
        toks[:, 1:T + 1] = captions.to(device=dev, dtype=torch.int32)
        positions = torch.arange(S, dtype=torch.int32, device=dev).repeat(N)
        keyseq = toks[:, 1:].contiguous()                                       # key t>=1 masked iff token t-1 == pad
        rows = N * S
        x = torch.empty(rows, D, dtype=dt, device=dev)
        qb, kb, vb, attn, tmp = (torch.empty_like(x) for _ in range(5))
        h1 = torch.empty(rows, self.pf, dtype=dt, device=dev)
        ops.xfmr_embed(self.tok, self.pos, start_emb, S, toks.view(-1), positions, 0, self.scale, x)
        xkv, emask = self._cross_kv(spatial, N) if self.cross else (None, None)
        for l, lay in enumerate(self.layers):
            ops.gemm(x, lay['self_attn.fc_q.w'], qb, bias=lay['self_attn.fc_q.b'])
            ops.gemm(x, lay['self_attn.fc_k.w'], kb, bias=lay['self_attn.fc_k.b'])
            ops.gemm(x, lay['self_attn.fc_v.w'], vb, bias=lay['self_attn.fc_v.b'])
            ops.attention(qb, kb, vb, attn, self.n_heads, S, 1, S, lay['self_attn.scale'], slot_shared=True,
                          causal_full=True, seq=keyseq, seq_per_image=True, pad=self.pad)
            self._post_attn(lay, 'self_attn', x, attn, tmp, rows)
            if self.cross:
                ops.gemm(x, lay['enc_attn.fc_q.w'], qb, bias=lay['enc_attn.fc_q.b'])
                ops.attention(qb, xkv[l][0], xkv[l][1], attn, self.n_heads, S, 1, 49, lay['enc_attn.scale'],
                              slot_shared=True, n_keys=49, enc_mask=emask)
                self._post_attn(lay, 'enc_attn', x, attn, tmp, rows)
            self._ffn(lay, x, h1, tmp, rows)
        return x, S

    def token_logprob(self, start_emb, spatial, captions, targets):
        """log_softmax(decoder(captions))[n, t, targets[n, t]] for t < T = targets.shape[1] -> [N, T] fp32
        (experiments/metrics.py:5 fused into the classifier contraction; logits [N,S,V] are never materialised)."""
        N, T = targets.shape
        x, S = self.hidden(start_emb, spatial, captions)
        assert T <= S
        tg = torch.full((N, S), self.pad, dtype=torch.int64, device=self.device)
        tg[:, :T] = targets.to(self.device)
        lp = torch.empty(N * S, dtype=torch.float32, device=self.device)
        if x.dtype == torch.float32:
            logits = torch.empty(N * S, self.V, dtype=torch.float32, device=self.device)
            ops.gemm(x, self.Wc, logits, bias=self.bc)
            ops.token_logprob(logits, tg.view(-1), lp)
        else:
            ops.vocab_logprob(x, self.Wc, self.bc, tg.view(-1), lp)
        return lp.view(N, S)[:, :T]

    def forward(self, start_emb, spatial, captions):
        N = captions.shape[0]
        x, S = self.hidden(start_emb, spatial, captions)
        logits = torch.empty(N, S, self.V, dtype=torch.float32, device=self.device)
        ops.gemm(x, self.Wc, logits.view(N * S, self.V), bias=self.bc)
        return logits
