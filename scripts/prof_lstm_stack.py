"""Four launches of the 3-layer LSTM stack kernel at the config-2 shape (2560 rows, E = H = 512) for
`ncu --set full -k regex:gemm_tc_kernel -s 2 -c 1`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
dev = 'cuda'
rows, H, E, L = 2560, 512, 512, 3
A_all = (torch.randn(L, rows, E + H, device=dev) * 0.3).to(torch.bfloat16)
W_all = torch.cat([ops.pack_lstm_gates((torch.randn(4 * H, E + H, device=dev) * 0.05).to(torch.bfloat16), H) for _ in range(L)]).contiguous()
b_all = torch.cat([ops.pack_lstm_gates(torch.randn(4 * H, device=dev), H) for _ in range(L)]).contiguous()
cc = [torch.randn(L, rows, H, device=dev) for _ in range(2)]
hs = torch.empty(L, rows, H, dtype=torch.bfloat16, device=dev)
top = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
parent = torch.randint(0, rows, (rows,), device=dev, dtype=torch.int32)
per = (L - 1) * ((rows + 127) // 128)
pool = torch.zeros(4 * per, dtype=torch.int32, device=dev)
for i in range(4):
    ops.lstm_stack_tc(A_all, [E] + [H] * (L - 1), W_all, b_all, cc[0], parent, cc[1], top, hs, pool[i * per:(i + 1) * per], rows)
torch.cuda.synchronize()
