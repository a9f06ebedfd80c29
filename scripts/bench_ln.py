"""dh_gemm_tc_ln (fc_o / fc_2 + residual + LayerNorm in one launch) against dh_gemm_tc + dh_add_layernorm at the decode shapes
of BASELINE configs[4]: rows = 40 960, N = 512, K = 512 / 2048.  L2 is flushed between launches (inputs 42 MB + 42 MB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops

M = int(os.environ.get('ROWS', 40960))
dev, dt = 'cuda', torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for K in (512, 2048):
    A = torch.randn(M, K, device=dev).to(dt)
    W = (torch.randn(512, K, device=dev) * 0.05).to(dt)
    b, g, be = torch.randn(512, device=dev), torch.rand(512, device=dev) + 0.5, torch.randn(512, device=dev)
    x = torch.randn(M, 512, device=dev).to(dt)
    tmp, out = torch.empty_like(x), torch.empty_like(x)
    for _ in range(3):
        ops.gemm_ln(A, W, b, x, g, be, out)
        ops.gemm(A, W, tmp, bias=b, residual=x)
        ops.add_layernorm(tmp, None, g, be, out)
    t_f = timed(lambda: ops.gemm_ln(A, W, b, x, g, be, out))
    t_g = timed(lambda: ops.gemm(A, W, tmp, bias=b, residual=x))
    t_l = timed(lambda: ops.add_layernorm(tmp, None, g, be, out))
    fl = 2.0 * M * 512 * K
    print(f'K={K}: fused {t_f:.1f} us ({fl / t_f / 1e6:.0f} TF/s) | gemm {t_g:.1f} + layernorm {t_l:.1f} = {t_g + t_l:.1f} us')
