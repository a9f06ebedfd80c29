"""LSTM decoder runtime: batched stochastic-beam generation and teacher-forced forward.

Reference: models/rnn_models.py:28-46 (forward), :48-143 (generate); SURVEY.md Appendix A.2.
Per layer the gate product is ONE contraction over K = in+H against [W_ih | W_hh] (bias = b_ih + b_hh), the
cell update reads c through the beam-parent index (the reference's h/c regather, incl. its f // B misalignment,
Q8, is folded into loads), and the loop has a fixed trip count with per-image frozen-at-break flags on the
device, so there is no host synchronisation inside it (Q12).
"""
import torch

from . import ops


class LSTMDecoderRT:
    def __init__(self, sd, prefix, dtype, device):
        self.dtype, self.device = dtype, device
        to = lambda t, dt=dtype: t.float().contiguous().to(device=device, dtype=dt)
        self.table = to(sd[prefix + '.embedding.weight'])
        self.V, self.E = self.table.shape
        self.L = 0
        self.Wcat, self.bias = [], []
        while f'{prefix}.lstm.weight_ih_l{self.L}' in sd:
            l = self.L
            self.Wcat.append(to(torch.cat([sd[f'{prefix}.lstm.weight_ih_l{l}'].float(),
                                           sd[f'{prefix}.lstm.weight_hh_l{l}'].float()], dim=1)))
            self.bias.append(to(sd[f'{prefix}.lstm.bias_ih_l{l}'].float() + sd[f'{prefix}.lstm.bias_hh_l{l}'].float(),
                                torch.float32))
            self.L += 1
        self.H = self.Wcat[0].shape[0] // 4
        # tensor-core mode: gate-packed copies for the fused cell epilogue (csrc/gemm_tc.cu, epi_mode 3)
        self.fused_cell = ops.FUSED_LSTM and dtype != torch.float32 and self.H % 64 == 0
        if self.fused_cell:
            self.Wpk = [ops.pack_lstm_gates(w, self.H) for w in self.Wcat]
            self.bpk = [ops.pack_lstm_gates(b, self.H) for b in self.bias]
        # ... and, for num_layers > 1, one stacked copy so that a whole time step is a single persistent launch
        self.in_dims = [self.E if l == 0 else self.H for l in range(self.L)]
        self.Kmax = max(self.in_dims) + self.H
        self.stacked = self.fused_cell and ops.LSTM_STACK and 1 < self.L <= 8 and self.Kmax % 8 == 0
        if self.stacked:
            self.Wpk_all = torch.zeros(self.L * 4 * self.H, self.Kmax, dtype=dtype, device=device)
            for l, w in enumerate(self.Wpk):
                self.Wpk_all[l * 4 * self.H:(l + 1) * 4 * self.H, :w.shape[1]] = w
            self.bpk_all = torch.cat(self.bpk).contiguous()
        self.Wc, self.bc = to(sd[prefix + '.classifier.weight']), to(sd[prefix + '.classifier.bias'], torch.float32)
        self.ldv = (self.V + 3) // 4 * 4
        self._plans = ops.PlanCache()

    def _alloc(self, rows, logits=True):
        d, dev, H, E, L = self.dtype, self.device, self.H, self.E, self.L
        # layer l's operand [x | h_prev] lives in columns [0, in_l + H) of one stacked buffer (dh_lstm_stack_tc)
        A_all = torch.zeros(L, rows, self.Kmax, dtype=d, device=dev)
        ws = dict(
            A_all=A_all, A=[A_all[l][:, :self.in_dims[l] + H] for l in range(L)],
            ready=torch.zeros(1, dtype=torch.int32, device=dev), ready_next=0,
            gates=torch.empty(rows, 4 * H, dtype=torch.float32, device=dev),
            c=[torch.zeros(L, rows, H, dtype=torch.float32, device=dev) for _ in range(2)],
            hs=torch.zeros(L, rows, H, dtype=d, device=dev),
            top=torch.empty(rows, H, dtype=d, device=dev),
            logits=torch.empty(rows, self.ldv, dtype=torch.float32, device=dev) if logits else None)
        return ws

    def _step(self, ws, rows, cur, parent, logits=True, top_out=None):
        """One LSTM time step over `rows` rows; A[l][:, in:] must already hold the (gathered) recurrent h.  The top layer's
        h goes to top_out ([rows, H], any row stride) or ws['top']."""
        L, H = self.L, self.H
        top = ws['top'][:rows] if top_out is None else top_out
        if self.stacked:
            # one zeroed counter region per launch (the decode zeroes the pool once, before its first step)
            per = (L - 1) * ((rows + 127) // 128)
            off = ws['ready_next']
            assert off + per <= ws['ready'].numel(), 'ready-counter pool exhausted'
            ws['ready_next'] = off + per
            with ops.PROFILE.range('lstm_layers', sum(2.0 * rows * 4 * H * (i + H) for i in self.in_dims)):
                ops.lstm_stack_tc(ws['A_all'], self.in_dims, self.Wpk_all, self.bpk_all, ws['c'][cur], parent,
                                  ws['c'][1 - cur], top, ws['hs'], ws['ready'][off:off + per], rows)
            if logits:
                with ops.PROFILE.range('vocab_gemm', 2.0 * rows * self.V * H):
                    ops.gemm(top, self.Wc, ws['logits'][:rows, :self.V], bias=self.bc)
            return
        for l in range(L):
            A = ws['A'][l][:rows]
            nxt = ws['A'][l + 1][:rows, :H] if l + 1 < L else top
            if self.fused_cell:
                with ops.PROFILE.range('lstm_layers', 2.0 * rows * 4 * H * A.shape[1]):
                    ops.lstm_layer_tc(A, self.Wpk[l], self.bpk[l], ws['c'][cur][l], parent, ws['c'][1 - cur][l][:rows],
                                      nxt, ws['hs'][l][:rows])
            else:
                gates = ws['gates'][:rows]
                ops.gemm(A, self.Wcat[l], gates, bias=self.bias[l])
                ops.lstm_cell(gates, ws['c'][cur][l], parent, ws['c'][1 - cur][l][:rows], nxt, ws['hs'][l][:rows])
        if logits:
            with ops.PROFILE.range('vocab_gemm', 2.0 * rows * self.V * H):
                ops.gemm(top, self.Wc, ws['logits'][:rows, :self.V], bias=self.bc)

    def _select(self, pl, rows, rpi, step, done, B, top_k, temperature, unk_index, noise_mode, beam_step=None,
                lstm_next=None):
        """classifier + BeamSearchHelper selection for `rows` rows (rnn_models.py:81,87-92,109-113): fused two-pass
        vocab projection (logits never stored) in tensor-core mode, materialised fp32 logits in check mode."""
        ws, beam = pl['ws'], pl['beam']
        if pl['vsel'] is not None:
            pl['vsel'].run(ws['top'][:rows], self.Wc, self.bc, B, temperature, unk_index, rpi, noise_mode, step, done,
                           pl['ind'], pl['val'], beam.status, pl['dynw'].dev,
                           beam_step=None if beam_step is None else (beam,) + beam_step, lstm_next=lstm_next)
        else:
            with ops.PROFILE.range('select_beam'):
                ops.select_tokens(ws['logits'][:rows, :self.V], self.V, B, top_k, temperature, unk_index, rpi, noise_mode,
                                  0, 0, step, done, pl['ind'], pl['val'], beam.status, pl['dynw'].dev)
                if beam_step is not None:
                    max_len, eos_index, lstm_sem = beam_step
                    beam.step(pl['ind'], pl['val'], step, max_len, eos_index, lstm_sem, temperature, noise_mode, 0, 0,
                              pl['dynw'].dev)

    def _recur(self, ws, rows, parent):
        """A[l][:, in:] <- hs[l][parent] for every layer (recurrent operand of the next step)."""
        for l in range(self.L):
            in_l = self.E if l == 0 else self.H
            ops.gather_rows(ws['hs'][l], parent, ws['A'][l][:rows, in_l:])

    def _decode(self, pl, p0, max_len, temperature, B, top_k, eos_index, unk_index, noise_mode, pad_index):
        """Launches the whole decode (prefix phase, first selection, beam loop, final pick) on static buffers of plan
        `pl`; fixed trip count and no host synchronisation, so it can be captured in a CUDA graph."""
        ws, beam, ind, val, dyn, N = pl['ws'], pl['beam'], pl['ind'], pl['val'], pl['dynw'].dev, pl['N']
        start_emb, caption = pl['start'], pl['caption']
        R = N * B
        fused = pl['vsel'] is not None
        beam.status.zero_()
        for c in ws['c']:
            c.zero_()
        if self.stacked:
            steps = p0 + 1 + max(0, max_len - p0 - 1)
            need = steps * (self.L - 1) * ((max(R, N) + 127) // 128)
            if ws['ready'].numel() < need:
                assert not torch.cuda.is_current_stream_capturing()
                ws['ready'] = torch.zeros(need, dtype=torch.int32, device=self.device)
            ws['ready'].zero_()
            ws['ready_next'] = 0
        # ---- prefix phase: 1 row per image (rnn_models.py:73-81)
        cur = 0
        ops.gather_rows(start_emb, None, ws['A'][0][:N, :self.E])
        for l in range(self.L):
            ws['A'][l][:N, (self.E if l == 0 else self.H):].zero_()
        for t in range(p0 + 1):
            if t > 0:
                ops.gather_rows(self.table, caption[:, t - 1].contiguous(), ws['A'][0][:N, :self.E])
                self._recur(ws, N, None)
            self._step(ws, N, cur, None, logits=(not fused) and t == p0)
            cur = 1 - cur
        self._select(pl, N, 1, p0, None, B, top_k, temperature, unk_index, noise_mode)
        beam.init(ind, val, caption, eos_index, True)
        ops.trace_beam(p0, beam)
        # ---- beam phase (rnn_models.py:105-137): fixed trip count, frozen-at-break on the device
        # the gathers of a step's operands (next token embedding, recurrent h through the beam parent) ride in the previous
        # step's select + beam launch when the selection is fused (dh_select_beam_step_lstm); the first step takes them
        # from a launch of their own
        two_byte = self.dtype != torch.float32 and self.L <= 8
        nxt_ops = None
        if fused and two_byte and ops.FUSED_PREPARE:
            nxt_ops = pl.get('lstm_next')
            if nxt_ops is None:
                nxt_ops = pl['lstm_next'] = ops.lstm_operands(self.table, [ws['hs'][l] for l in range(self.L)],
                                                              [ws['A'][l][:R] for l in range(self.L)],
                                                              [self.E if l == 0 else self.H for l in range(self.L)])
        for i in range(p0 + 1, max_len):
            if two_byte:
                if nxt_ops is None or i == p0 + 1:
                    ops.lstm_prepare(self.table, beam.last_tok, beam.parent_state, [ws['hs'][l] for l in range(self.L)],
                                     [ws['A'][l][:R] for l in range(self.L)],
                                     [self.E if l == 0 else self.H for l in range(self.L)], R)
            else:
                ops.gather_rows(self.table, beam.last_tok, ws['A'][0][:R, :self.E])
                self._recur(ws, R, beam.parent_state)
            self._step(ws, R, cur, beam.parent_state, logits=not fused)
            cur = 1 - cur
            self._select(pl, R, B, i, beam.done, B, top_k, temperature, unk_index, noise_mode,
                         beam_step=(max_len, eos_index, True), lstm_next=nxt_ops if i + 1 < max_len else None)
            ops.trace_beam(i, beam)
        beam.final(temperature, noise_mode, 0, 0, max_len + 1, max(p0 + 1, max_len), pad_index, max_len, pl['ids'],
                   pl['lens'], dyn)

    def generate(self, *args, **kw):
        """Optimistic first try with a sampled pass 1 of the fused vocab projection; if a candidate list overflowed
        (status bit 2) the whole generation is redone with the exhaustive pass 1, which cannot overflow."""
        out = self._generate(*args, robust=False, **kw)
        if out[3] and int(out[2].item()) & 2:
            out = self._generate(*args, robust=True, **kw)
        return out[:3]

    def _generate(self, start_emb, caption, max_len, temperature, beam_size, top_k, eos_index, unk_index, noise_mode,
                  seed, image_base, pad_index=0, robust=False):
        """start_emb fp32 [N,E]; caption int32 [N or 1, p] or None -> (ids int64 [N,max_len], lengths int64 [N], status).

        The decode loop is captured once per (batch, beam, lengths, sampling parameters) into a CUDA graph over
        static buffers and replayed; seed / image_base reach the kernels through a device word pair."""
        N, B, dev = start_emb.shape[0], beam_size, self.device
        if N == 0:                                   # empty batch: nothing to launch
            z = lambda *sh: torch.zeros(*sh, dtype=torch.int64, device=dev)
            return z(0, max_len), z(0), torch.zeros(1, dtype=torch.int32, device=dev), False
        p0 = 0 if caption is None else caption.shape[1]
        key = (N, B, p0, max_len, float(temperature), top_k, eos_index, unk_index, noise_mode, pad_index, robust)
        pl = self._plans.get(key)
        if pl is None:
            R = N * B
            fused = ops.FUSED_VOCAB and ops.VocabSelect.supported(self.Wc, self.V, top_k)
            pl = dict(N=N, ws=self._alloc(max(R, N), logits=not fused),
                      vsel=ops.VocabSelect(max(R, N), self.V, top_k, dev, stride=1 if robust else None) if fused else None, beam=ops.Beam(N, B, max(max_len, p0 + 1), dev),
                      ind=torch.empty(R, B, dtype=torch.int32, device=dev),
                      val=torch.empty(R, B, dtype=torch.float32, device=dev),
                      dynw=ops.DynWords(dev),
                      start=torch.empty(N, self.E, dtype=torch.float32, device=dev),
                      caption=None if caption is None else torch.empty(N, p0, dtype=torch.int32, device=dev),
                      ids=torch.empty(N, max_len, dtype=torch.int64, device=dev),
                      lens=torch.empty(N, dtype=torch.int64, device=dev), graph=None)
            self._plans.put(key, pl)
        pl['start'].copy_(start_emb)
        if caption is not None:
            pl['caption'].copy_(caption.expand(N, p0))
        pl['dynw'].set(seed, image_base)
        args = (pl, p0, max_len, temperature, B, top_k, eos_index, unk_index, noise_mode, pad_index)
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

    def hidden(self, image_emb, captions, lengths=None):
        """Teacher-forced top-layer outputs [N, max(lengths), H], zero at t >= length (rnn_models.py:28-44;
        packed-sequence semantics, Q24)."""
        N, T = captions.shape
        dev, H = self.device, self.H
        S = T + 1
        if lengths is None:
            lengths = torch.full((N,), S, dtype=torch.int64, device=dev)
        lengths = lengths.to(dev)
        tmax = int(lengths.max())
        ws = self._alloc(N, logits=False)
        tops = torch.zeros(N, tmax, H, dtype=self.dtype, device=dev)
        cap32 = captions.to(device=dev, dtype=torch.int32)
        if self.stacked:                                 # one zeroed ready-counter region per time step's launch
            ws['ready'] = torch.zeros(tmax * (self.L - 1) * ((N + 127) // 128), dtype=torch.int32, device=dev)
            ws['ready_next'] = 0
        cur = 0
        for t in range(tmax):
            if t == 0:
                ops.gather_rows(image_emb, None, ws['A'][0][:N, :self.E])
            else:
                ops.gather_rows(self.table, cap32[:, t - 1].contiguous(), ws['A'][0][:N, :self.E])
                self._recur(ws, N, None)
            # the same step as generate(): in tensor-core mode every layer of the step is one persistent launch with the cell
            # update in the gate contraction's epilogue (dh_lstm_stack_tc), the top layer writing straight into tops[:, t]
            self._step(ws, N, cur, None, logits=False, top_out=tops[:, t])
            cur = 1 - cur
        # pad_packed_sequence zero-fills outputs at t >= length before the classifier (rnn_models.py:41-44)
        keep = (torch.arange(tmax, device=dev).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(-1)
        return tops * keep.to(tops.dtype), tmax

    def forward(self, image_emb, captions, lengths=None):
        """Teacher-forced logits [N, max(lengths), V] fp32."""
        N = captions.shape[0]
        tops, tmax = self.hidden(image_emb, captions, lengths)
        logits = torch.empty(N, tmax, self.V, dtype=torch.float32, device=self.device)
        ops.gemm(tops.view(N * tmax, self.H), self.Wc, logits.view(N * tmax, self.V), bias=self.bc)
        return logits

    def token_logprob(self, image_emb, captions, lengths, targets):
        """log_softmax(classifier(lstm(...)))[n, t, targets[n, t]] for t < min(max(lengths), targets width) -> [N, T]
        (experiments/metrics.py:5 fused into the classifier contraction in tensor-core mode)."""
        N = captions.shape[0]
        tops, tmax = self.hidden(image_emb, captions, lengths)
        T = min(tmax, targets.shape[1])
        tg = torch.zeros(N, tmax, dtype=torch.int64, device=self.device)
        tg[:, :T] = targets[:, :T].to(self.device)
        lp = torch.empty(N * tmax, dtype=torch.float32, device=self.device)
        x = tops.view(N * tmax, self.H)
        if x.dtype == torch.float32:
            logits = torch.empty(N * tmax, self.V, dtype=torch.float32, device=self.device)
            ops.gemm(x, self.Wc, logits, bias=self.bc)
            ops.token_logprob(logits, tg.view(-1), lp)
        else:
            ops.vocab_logprob(x, self.Wc, self.bc, tg.view(-1), lp)
        return lp.view(N, tmax)[:, :T]
