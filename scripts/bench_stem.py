"""Device time of the fused stem kernel (conv7x7/2 + BN + ReLU + maxpool) for n images, graph-timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = 'cuda'
img = torch.randn(n, 3, 224, 224, device=dev)
w = (torch.randn(64, 192, device=dev) * 0.1).half()
b = torch.randn(64, device=dev)
out = torch.empty(n, 56, 56, 64, dtype=torch.float16, device=dev)
fn = lambda: ops.stem_pool(img, w, b, out)
fn(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(5):
        fn()
g.replay(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / 5)
us = min(ts)
byts = img.numel() * 4 + out.numel() * 2
print(f'stem_pool n={n}: {us:.1f} us  {byts/us/1e3:.0f} GB/s of algorithmic traffic (ideal {byts/6.5e12*1e6:.0f} us), '
      f'{2*n*12544*64*147/us/1e6:.0f} TF/s of useful conv FLOPs')
