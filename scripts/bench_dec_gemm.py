"""Decoder-shaped contractions (bf16 in / bf16 out, the slab epilogue): CUDA-event time per launch, L2 flushed by size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops

dev = 'cuda'


def timeit(fn, reps=10, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return min(ts)


for (M, N, K, relu) in [(40960, 1536, 512, False), (40960, 512, 512, False), (40960, 2048, 512, True), (40960, 1024, 512, False),
                        (401408, 1024, 512, False), (2560, 2048, 512, True), (10240, 1536, 512, False)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    us = timeit(lambda: ops.gemm(A, W, out, bias=b, relu=relu))
    print(f'gemm M={M} N={N} K={K} relu={int(relu)}: {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s', flush=True)
