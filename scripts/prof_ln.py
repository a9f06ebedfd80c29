"""A few dh_gemm_tc_ln launches at the cfg5 decode shape (40 960 x 512 x K) for `ncu -k regex:gemm_tc_kernel -s 2 -c 1`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
M, K = int(os.environ.get('ROWS', 40960)), int(os.environ.get('K', 512))
dev, dt = 'cuda', torch.bfloat16
A = torch.randn(M, K, device=dev).to(dt)
W = (torch.randn(512, K, device=dev) * 0.05).to(dt)
b, g, be = torch.randn(512, device=dev), torch.rand(512, device=dev) + 0.5, torch.randn(512, device=dev)
x = torch.randn(M, 512, device=dev).to(dt)
out = torch.empty_like(x)
for _ in range(4):
    ops.gemm_ln(A, W, b, x, g, be, out)
torch.cuda.synchronize()
