#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/kernel_times.py cfg2 1024 > gpurun_out/ap_kernels_cfg2.txt 2>&1; head -24 gpurun_out/ap_kernels_cfg2.txt | cut -c1-170
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 2 -c 1 -f -o gpurun_out/ap_chain python - > gpurun_out/ap_ncu.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from deephumor_b200.runtime import ops
dt = torch.float16
n = 256
y2 = torch.randn(n, 56, 56, 64, device='cuda').to(dt)
x2 = torch.randn(n, 56, 56, 256, device='cuda').to(dt)
w = (torch.randn(256, 64, device='cuda') * 0.05).to(dt)
b = torch.randn(256, device='cuda')
wn = (torch.randn(64, 256, device='cuda') * 0.05).to(dt)
bn = torch.randn(64, device='cuda')
out = torch.empty(n, 56, 56, 256, dtype=dt, device='cuda')
z = torch.empty(n, 56, 56, 64, dtype=dt, device='cuda')
for _ in range(4):
    ops.conv1x1_chain(y2, x2, False, w, b, out, wn, bn, z)
torch.cuda.synchronize()
PY
tail -2 gpurun_out/ap_ncu.log
