"""Vocab projection passes at ROWS rows (V = 36 541, K = 512), L2 flushed between launches: pass 1 (sampled, stride 8),
pass 2 (sparse materialisation) with A resident (default) and with A streamed (DH_TC_NO_ARES=1 in a second process)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200._lib import LIB, ptr, stream
from deephumor_b200.runtime import ops
dev = 'cuda'
M, N, K = int(os.environ.get('ROWS', 40960)), 36541, 512
g = torch.Generator(device=dev).manual_seed(1)
W = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
A = (torch.randn(M, K, device=dev, generator=g) * 0.5).to(torch.bfloat16)
b = torch.randn(N, device=dev, generator=g) * 0.1
vs = ops.VocabSelect(M, N, 50, dev)
args = (ptr(A), K, ptr(W), K, 1, ptr(b), M, N, K)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=12):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
p1 = lambda: LIB.call('dh_vocab_groupmax', *args, vs.stride, 0, ptr(vs.gmax), vs.n_groups_full, stream())
th = lambda: LIB.call('dh_vocab_threshold', ptr(vs.gmax), vs.n_groups_full, M, vs.groups(0), vs.rank, ptr(vs.thresh), ptr(vs.count), stream())
def p2():
    LIB.call('dh_vocab_candidates', *args, ptr(vs.thresh), ptr(vs.count), ptr(vs.sp_logits), vs.sp_ld, ptr(vs.hitmap), vs.hit_ld, stream())
p1(); th(); p2(); torch.cuda.synchronize()
t1, t2 = timed(p1), timed(p2)
fl = 2.0 * M * N * K
tag = 'A streamed' if os.environ.get('DH_TC_NO_ARES') else 'A resident'
print(f'rows {M} [{tag}]: pass 1 (1/{vs.stride}) {t1:.1f} us | pass 2 {t2:.1f} us = {fl / t2 / 1e6:.0f} TFLOP/s; groups/row {float(vs.count.float().mean()) / 12:.0f}')
if not os.environ.get('DH_TC_NO_ARES'):
    subprocess.run([sys.executable, __file__], env=dict(os.environ, DH_TC_NO_ARES='1'))
