"""Trunk (dh_resnet50_forward) time for N images, per layer1 L2-chunk setting (DH_TRUNK_L2_CHUNK; one subprocess per value,
the library reads it once).  Inputs larger than L2 (N x 602 KB)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if 'CHILD' not in os.environ:
    for n in (512, 1024):
        for ch in os.environ.get('CHUNKS', '0,16,24,32,48,64').split(','):
            subprocess.run([sys.executable, __file__], env=dict(os.environ, CHILD='1', DH_TRUNK_L2_CHUNK=ch, N=str(n)))
    sys.exit(0)
import torch
import bench
from deephumor_b200.runtime import ops
n = int(os.environ['N'])
m, hp, sd = bench.build_model('lstm', 'bf16')
enc = m.encoder._rt() if hasattr(m.encoder, '_rt') else None
images = torch.empty(n, 3, 224, 224, device='cuda')
ops.synth_images(images, 0, 0)
pooled = torch.empty(n, 2048, device='cuda')
def run():
    return enc.trunk(images, pooled)
for _ in range(3):
    feat, _ = run()
torch.cuda.synchronize()
ref = feat.clone()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f'N={n} layer1 L2 chunk {os.environ["DH_TRUNK_L2_CHUNK"]:>3s}: trunk {ts[len(ts) // 2]:.3f} ms (min {ts[0]:.3f})  checksum {float(feat.float().sum()):.6e} pooled {float(pooled.sum()):.6e}')
