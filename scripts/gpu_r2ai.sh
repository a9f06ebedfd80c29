#!/bin/bash
# l1.ds / l1.c3 (N = 256, K = 64 convs): what bounds them at ~3 us per tile? + read-only / write-only HBM rates
mkdir -p gpurun_out
python - > gpurun_out/ai_bw.txt 2>&1 <<'PY'
import torch
x = torch.empty(1 << 30, dtype=torch.float16, device='cuda')
y = torch.empty(1 << 30, dtype=torch.float16, device='cuda')
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_()); print(f'write-only 2 GiB: {ms:.3f} ms {2**31/ms/1e6:.0f} GB/s')
ms = t(lambda: x.fill_(1.0)); print(f'fill 2 GiB: {ms:.3f} ms {2**31/ms/1e6:.0f} GB/s')
ms = t(lambda: torch.sum(x.view(torch.int32))); print(f'read-only 2 GiB: {ms:.3f} ms {2**31/ms/1e6:.0f} GB/s')
ms = t(lambda: y.copy_(x)); print(f'copy 2+2 GiB: {ms:.3f} ms {2**32/ms/1e6:.0f} GB/s')
PY
cat gpurun_out/ai_bw.txt
for L in l1.ds l1.c3; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 3 -c 1 -f -o gpurun_out/ai_$L python scripts/bench_conv.py 256 $L > gpurun_out/ai_ncu_$L.log 2>&1
  tail -2 gpurun_out/ai_ncu_$L.log
done
