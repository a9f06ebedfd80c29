"""Layer1 identity-bottleneck tail at 512 images: dh_bottleneck_tail_tc (conv2 -> conv3 + identity, one launch) against the
two launches (dh_conv3x3_halo_tc, dh_conv2d_tc with the residual).  Inputs 205 + 822 MB: larger than L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
n, H, dt, dev = int(os.environ.get('N', 512)), 56, torch.float16, 'cuda'
g = torch.Generator(device=dev).manual_seed(0)
y1 = torch.randn(n, H, H, 64, device=dev, generator=g).to(dt)
x = torch.randn(n, H, H, 256, device=dev, generator=g).to(dt)
w2 = (torch.randn(64, 3, 3, 64, device=dev, generator=g) * 0.05).to(dt)
w3 = (torch.randn(256, 1, 1, 64, device=dev, generator=g) * 0.1).to(dt)
b2, b3 = torch.randn(64, device=dev, generator=g) * 0.1, torch.randn(256, device=dev, generator=g) * 0.1
y2, out, out2 = torch.empty_like(y1), torch.empty_like(x), torch.empty_like(x)
def timed(fn, reps=8):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
fused = lambda: ops.bottleneck_tail(y1, w2, b2, w3, b3, x, out)
c2 = lambda: ops.conv2d(y1, w2, b2, y2, 1, 1, True)
c3 = lambda: ops.conv2d(y2, w3, b3, out2, 1, 0, True, residual=x)
for f in (fused, c2, c3):
    f()
torch.cuda.synchronize()
tf, t2, t3 = timed(fused), timed(c2), timed(c3)
gb = (y1.numel() + 2 * x.numel()) * 2 / 1e9
print(f'N={n}: fused {tf:.1f} us ({gb / tf * 1e6:.0f} GB/s of compulsory traffic) | conv2 {t2:.1f} + conv3 {t3:.1f} = {t2 + t3:.1f} us; max |diff| {float((out.float() - out2.float()).abs().max()):.4f}')
