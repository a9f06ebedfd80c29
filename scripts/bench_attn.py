"""Graph-timed cross-attention (49 cached spatial keys per image) and incremental self-attention at the config-5 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
dev = 'cuda'
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return min(ts)
D, H = 512, 8
for n_img, rpi in ((2048, 5), (4096, 1), (8192, 5)):
    rows = n_img * rpi
    L = 3    # rotate over 3 layers' K/V like the decoder does (615 MB at 2048 images: nothing stays in L2)
    q = torch.randn(rows, D, device=dev).to(torch.bfloat16)
    Ks = [torch.randn(n_img, 49, D, device=dev).to(torch.bfloat16) for _ in range(L)]
    Vs = [torch.randn(n_img, 49, D, device=dev).to(torch.bfloat16) for _ in range(L)]
    em = (torch.rand(n_img, 49, device=dev) < 0.1).to(torch.uint8)
    out = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    def run():
        for l in range(L):
            ops.attention(q, Ks[l], Vs[l], out, H, rpi, 1, 49, 8.0, slot_shared=True, n_keys=49, enc_mask=em)
    us = timeit(run) / L
    mb = 2 * n_img * 49 * D * 2 / 1e6
    print(f'cross-attention n_img={n_img} rpi={rpi}: {us:.1f} us  ({mb:.0f} MB of K/V -> {mb / us * 1e3:.0f} GB/s)')
# ---- incremental self-attention over the KV cache: rows = images x beams, keys through the slot table
for n_img, B, nk in ((2048, 5, 17), (2048, 5, 33), (8192, 5, 17)):
    rows, S = n_img * B, 33
    q = torch.randn(rows, D, device=dev).to(torch.bfloat16)
    Kc = torch.randn(n_img * B, S, D, device=dev).to(torch.bfloat16)
    Vc = torch.randn(n_img * B, S, D, device=dev).to(torch.bfloat16)
    # genealogy-like slot table: old positions point to few distinct slots, recent ones to the beam's own
    src = torch.zeros(n_img, B, S, dtype=torch.int32, device=dev)
    for t in range(S):
        src[:, :, t] = torch.arange(B, device=dev).view(1, B) if t >= nk - 3 else torch.randint(0, 2, (n_img, 1), device=dev).int()
    seq = torch.randint(1, 100, (rows, S), dtype=torch.int32, device=dev)
    out = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    us = timeit(lambda: ops.attention(q, Kc, Vc, out, H, B, B, S, 8.0, src=src, n_keys=nk, seq=seq, seq_per_image=False, pad=0))
    mb = 2 * rows * nk * D * 2 / 1e6
    print(f'self-attention rows={rows} keys={nk}: {us:.1f} us  ({mb:.0f} MB of K/V row reads -> {mb / us * 1e3:.0f} GB/s)')
