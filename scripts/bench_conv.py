"""Micro-benchmark of the tensor-core contraction kernel on every distinct ResNet-50 conv shape (n images) and the
decoder GEMM shapes: CUDA-event time, TFLOP/s and GB/s of algorithmic traffic (in + weights + out (+ residual))."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = sys.argv[2] if len(sys.argv) > 2 else ''
tile_n = int(os.environ.get('TILE_N', '0'))
dt = torch.float16
dev = 'cuda'
# (name, H, Cin, Cout, k, stride, pad, residual, count in trunk)
CONVS = [('l1.c1', 56, 64, 64, 1, 1, 0, 0, 1), ('l1.c2', 56, 64, 64, 3, 1, 1, 0, 3), ('l1.c3', 56, 64, 256, 1, 1, 0, 1, 3),
         ('l1.ds', 56, 64, 256, 1, 1, 0, 0, 1), ('l1.c1b', 56, 256, 64, 1, 1, 0, 0, 2),
         ('l2.c1', 56, 256, 128, 1, 1, 0, 0, 1), ('l2.c2s', 56, 128, 128, 3, 2, 1, 0, 1), ('l2.c3', 28, 128, 512, 1, 1, 0, 1, 4),
         ('l2.ds', 56, 256, 512, 1, 2, 0, 0, 1), ('l2.c1b', 28, 512, 128, 1, 1, 0, 0, 3), ('l2.c2', 28, 128, 128, 3, 1, 1, 0, 3),
         ('l3.c1', 28, 512, 256, 1, 1, 0, 0, 1), ('l3.c2s', 28, 256, 256, 3, 2, 1, 0, 1), ('l3.c3', 14, 256, 1024, 1, 1, 0, 1, 6),
         ('l3.ds', 28, 512, 1024, 1, 2, 0, 0, 1), ('l3.c1b', 14, 1024, 256, 1, 1, 0, 0, 5), ('l3.c2', 14, 256, 256, 3, 1, 1, 0, 5),
         ('l4.c1', 14, 1024, 512, 1, 1, 0, 0, 1), ('l4.c2s', 14, 512, 512, 3, 2, 1, 0, 1), ('l4.c3', 7, 512, 2048, 1, 1, 0, 1, 3),
         ('l4.ds', 14, 1024, 2048, 1, 2, 0, 0, 1), ('l4.c1b', 7, 2048, 512, 1, 1, 0, 0, 2), ('l4.c2', 7, 512, 512, 3, 1, 1, 0, 2)]


def timeit(fn, iters=5, reps=10):
    """Device time per call: `reps` back-to-back launches captured in one CUDA graph (no host launch gaps; the
    activations of one call exceed L2 at n >= 128, so consecutive calls do not find their inputs cached)."""
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


tot = 0.0
tot_ideal = 0.0
for name, H, Cin, Cout, k, s, p, res, cnt in CONVS:
    if only and only not in name:
        continue
    Ho = (H + 2 * p - k) // s + 1
    x = torch.randn(n, H, H, Cin, device=dev).to(dt)
    w = (torch.randn(Cout, k, k, Cin, device=dev) * 0.05).to(dt)
    b = torch.randn(Cout, device=dev)
    y = torch.empty(n, Ho, Ho, Cout, dtype=dt, device=dev)
    r = torch.randn(n, Ho, Ho, Cout, device=dev).to(dt) if res else None
    us = timeit(lambda: ops.conv2d(x, w, b, y, s, p, True, residual=r, tile_n=tile_n))
    flops = 2.0 * n * Ho * Ho * Cout * k * k * Cin
    byts = 2.0 * (x.numel() / (s * s if k == 1 else 1) + w.numel() + y.numel() * (2 if res else 1))
    ideal = max(flops / 1.4e15, byts / 6.5e12) * 1e6
    tot += us * cnt
    tot_ideal += ideal * cnt
    print(f'{name:7s} M={n*Ho*Ho:7d} N={Cout:4d} K={k*k*Cin:4d} res={res} x{cnt}: {us:8.1f} us  {flops/us/1e6:7.1f} TF/s  {byts/us/1e3:7.1f} GB/s  ideal {ideal:6.1f} us  ({us/ideal:4.1f}x)', flush=True)
print(f'trunk convs (excl. stem) for {n} images: {tot/1e3:.2f} ms, ideal {tot_ideal/1e3:.2f} ms')
if not only:
    for (M, N, K, odt) in [(2560, 36541, 512, torch.float32), (2560, 2048, 1024, torch.float32), (40960, 36541, 512, torch.float32),
                           (8192, 8192, 8192, torch.bfloat16), (125440, 512, 2048, torch.bfloat16)]:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        W = torch.randn(N, K, device=dev).to(torch.bfloat16)
        ldc = (N + 3) // 4 * 4
        out = torch.empty(M, ldc, dtype=odt, device=dev)
        us = timeit(lambda: ops.gemm(A, W, out[:, :N]))
        flops = 2.0 * M * N * K
        print(f'gemm M={M} N={N} K={K} out={odt}: {us:8.1f} us {flops/us/1e6:7.1f} TF/s  out-write {out.numel()*out.element_size()/us/1e3:7.1f} GB/s')
