"""One fused LSTM layer launch at 4 full waves (rows 9472) for `ncu --set full -k regex:gemm_tc_kernel`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
dev = 'cuda'
H, E, rows = 512, 512, int(os.environ.get('ROWS', 9472))
W = ops.pack_lstm_gates((torch.randn(4 * H, E + H, device=dev) * 0.05).to(torch.bfloat16), H)
b = ops.pack_lstm_gates(torch.randn(4 * H, device=dev), H)
A = (torch.randn(rows, E + H, device=dev) * 0.3).to(torch.bfloat16)
c0 = torch.randn(rows, H, device=dev); c1 = torch.empty_like(c0)
parent = torch.randint(0, rows, (rows,), device=dev, dtype=torch.int32)
h0 = torch.empty(rows, E + H, dtype=torch.bfloat16, device=dev); h1 = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
for _ in range(4):
    ops.lstm_layer_tc(A, W, b, c0, parent, c1, h0[:, :H], h1)
torch.cuda.synchronize()
