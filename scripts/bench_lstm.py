"""Graph-timed device time of one fused LSTM layer step (gate contraction + cell epilogue) and of the step's gathers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
dev = 'cuda'
rows, H, E = 2560, 512, 512
A = (torch.randn(rows, E + H, device=dev) * 0.3).to(torch.bfloat16)
W = ops.pack_lstm_gates((torch.randn(4 * H, E + H, device=dev) * 0.05).to(torch.bfloat16), H)
b = ops.pack_lstm_gates(torch.randn(4 * H, device=dev), H)
c0 = torch.randn(rows, H, device=dev); c1 = torch.empty_like(c0)
parent = torch.randint(0, rows, (rows,), device=dev, dtype=torch.int32)
h0 = torch.empty(rows, E + H, dtype=torch.bfloat16, device=dev); h1 = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return min(ts)
us = timeit(lambda: ops.lstm_layer_tc(A, W, b, c0, parent, c1, h0[:, :H], h1))
print(f'lstm_layer_tc rows={rows} H={H} K={E+H}: {us:.1f} us  {2.0*rows*4*H*(E+H)/us/1e6:.0f} TF/s')
gates = torch.empty(rows, 4 * H, device=dev)
us2 = timeit(lambda: ops.gemm(A, W, gates, bias=b))
print(f'plain gate GEMM (fp32 out): {us2:.1f} us')
os.environ['X'] = '1'
# ---- all layers of a time step in one persistent launch (dh_lstm_stack_tc) against L single-layer launches
L = 3
A_all = (torch.randn(L, rows, E + H, device=dev) * 0.3).to(torch.bfloat16)
W_all = torch.cat([W] * L).contiguous()
b_all = torch.cat([b] * L).contiguous()
cc = [torch.randn(L, rows, H, device=dev) for _ in range(2)]
hs = torch.empty(L, rows, H, dtype=torch.bfloat16, device=dev)
top = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
per = (L - 1) * ((rows + 127) // 128)
pool = torch.zeros(64 * per, dtype=torch.int32, device=dev)
state = {'i': 0}
def stack(rot, layers=L):
    i = state['i'] % 64
    state['i'] += 1
    if i == 0:
        pool.zero_()
    ops.lstm_stack_tc(A_all[:layers], [E] + [H] * (layers - 1), W_all[:layers * 4 * H], b_all[:layers * 4 * H], cc[0][:layers],
                      parent, cc[1][:layers], top, hs[:layers], pool[i * per:(i + 1) * per], rows, rotate=rot)
def three():
    for l in range(L):
        ops.lstm_layer_tc(A_all[l], W, b, cc[0][l], parent, cc[1][l], A_all[(l + 1) % L][:, :H], hs[l])
print(f'{L} x lstm_layer_tc: {timeit(three):.1f} us')
for rot in (0, 1):
    state['i'] = 0
    print(f'lstm_stack_tc L={L} rotate={rot}: {timeit(lambda: stack(rot)):.1f} us')
state['i'] = 0
print(f'lstm_stack_tc L=1: {timeit(lambda: stack(0, 1)):.1f} us')
state['i'] = 0
print(f'lstm_stack_tc L=2 rotate=1: {timeit(lambda: stack(1, 2)):.1f} us')
