"""Stress the cross-CTA layer hand-off of dh_lstm_stack_tc: many consecutive steps at the benchmark shape (2 560 rows, 3 layers),
each compared bit for bit with the layer-by-layer launches (a publish / acquire ordering bug shows up as a rare mismatch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
DEV = 'cuda'
rows, H_, E_, L_ = 2560, 512, 512, 3
steps = int(os.environ.get('STEPS', 400))
g = torch.Generator().manual_seed(1)
in_dims = [E_] + [H_] * (L_ - 1)
Kmax = max(in_dims) + H_
Ws = [(torch.randn(4 * H_, i + H_, generator=g) * 0.1).to(torch.bfloat16).to(DEV) for i in in_dims]
bs = [torch.randn(4 * H_, generator=g).to(DEV) for _ in in_dims]
Wp_all = torch.zeros(L_ * 4 * H_, Kmax, dtype=torch.bfloat16, device=DEV)
for l, w in enumerate(Ws):
    Wp_all[l * 4 * H_:(l + 1) * 4 * H_, :w.shape[1]] = ops.pack_lstm_gates(w, H_)
b_all = torch.cat([ops.pack_lstm_gates(b, H_) for b in bs]).contiguous()
Wpk = [ops.pack_lstm_gates(w, H_) for w in Ws]
bpk = [ops.pack_lstm_gates(b, H_) for b in bs]
A_all = torch.zeros(L_, rows, Kmax, dtype=torch.bfloat16, device=DEV)
A_ref = [torch.zeros(rows, i + H_, dtype=torch.bfloat16, device=DEV) for i in in_dims]
c = [torch.zeros(L_, rows, H_, device=DEV) for _ in range(2)]
c_ref = [torch.zeros(L_, rows, H_, device=DEV) for _ in range(2)]
hs, hs_ref = (torch.zeros(L_, rows, H_, dtype=torch.bfloat16, device=DEV) for _ in range(2))
top, top_ref = (torch.zeros(rows, H_, dtype=torch.bfloat16, device=DEV) for _ in range(2))
per = (L_ - 1) * ((rows + 127) // 128)
ready = torch.zeros(steps * per, dtype=torch.int32, device=DEV)
gd = torch.Generator(device=DEV).manual_seed(2)
cur, bad = 0, 0
for t in range(steps):
    x = (torch.randn(rows, E_, device=DEV, generator=gd) * 0.5).to(torch.bfloat16)
    parent = torch.randint(0, rows, (rows,), device=DEV, generator=gd).to(torch.int32) if t else None
    A_all[0, :, :E_] = x
    A_ref[0][:, :E_] = x
    for l in range(L_):
        src, src_ref = hs[l], hs_ref[l]
        if parent is not None:
            src, src_ref = src[parent.long()], src_ref[parent.long()]
        A_all[l, :, in_dims[l]:in_dims[l] + H_] = src
        A_ref[l][:, in_dims[l]:] = src_ref
    ops.lstm_stack_tc(A_all, in_dims, Wp_all, b_all, c[cur], parent, c[1 - cur], top, hs, ready[t * per:(t + 1) * per], rows, rotate=0)
    for l in range(L_):
        nxt = A_ref[l + 1][:, :H_] if l + 1 < L_ else top_ref
        ops.lstm_layer_tc(A_ref[l], Wpk[l], bpk[l], c_ref[cur][l], parent, c_ref[1 - cur][l], nxt, hs_ref[l])
    cur = 1 - cur
    if not (torch.equal(top, top_ref) and torch.equal(hs, hs_ref) and torch.equal(c[cur], c_ref[cur])):
        bad += 1
torch.cuda.synchronize()
print(f'{steps} stacked steps against layer-by-layer launches: {bad} mismatching steps')
sys.exit(1 if bad else 0)
