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
