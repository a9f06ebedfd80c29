"""Wave time vs fixed launch cost of the fused LSTM layer kernel: rows chosen so that the pair tiles fill 0.5 / 1 / 2 / 4 waves."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
dev = 'cuda'
H, E = 512, 512
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
W = ops.pack_lstm_gates((torch.randn(4 * H, E + H, device=dev) * 0.05).to(torch.bfloat16), H)
b = ops.pack_lstm_gates(torch.randn(4 * H, device=dev), H)
for rows in (256, 1184, 2368, 2560, 4736, 9472, 18944):
    A = (torch.randn(rows, E + H, device=dev) * 0.3).to(torch.bfloat16)
    c0 = torch.randn(rows, H, device=dev); c1 = torch.empty_like(c0)
    parent = torch.randint(0, rows, (rows,), device=dev, dtype=torch.int32)
    h0 = torch.empty(rows, E + H, dtype=torch.bfloat16, device=dev); h1 = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
    out16 = torch.empty(rows, 4 * H, dtype=torch.bfloat16, device=dev)
    t_l = timeit(lambda: ops.lstm_layer_tc(A, W, b, c0, parent, c1, h0[:, :H], h1))
    t_np = timeit(lambda: ops.lstm_layer_tc(A, W, b, c0, None, c1, h0[:, :H], h1))
    t_g = timeit(lambda: ops.gemm(A, W, out16, bias=b))
    print(f'rows {rows:6d} pair-tiles {((rows + 255) // 256) * 8:4d}: lstm_layer {t_l:6.1f} us (no parent {t_np:6.1f})  plain bf16-out GEMM {t_g:6.1f} us')
