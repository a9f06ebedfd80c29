"""Three rounds of pass 1 (sampled) -> threshold -> pass 2 of the fused vocab projection at ROWS rows; under
`ncu -k regex:gemm_tc_kernel -s 5 -c 1` the captured launch is the third pass 2 (roofline kernel of bench.py)."""
import os, sys
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
for it in range(3):
    LIB.call('dh_vocab_groupmax', *args, vs.stride, it % vs.stride, ptr(vs.gmax), vs.n_groups_full, stream())
    LIB.call('dh_vocab_threshold', ptr(vs.gmax), vs.n_groups_full, M, vs.groups(it % vs.stride), vs.rank, ptr(vs.thresh), ptr(vs.count), stream())
    LIB.call('dh_vocab_candidates', *args, ptr(vs.thresh), ptr(vs.count), ptr(vs.sp_logits), vs.sp_ld, ptr(vs.hitmap), vs.hit_ld, stream())
torch.cuda.synchronize()
print('rows', M, 'stored groups per row', float(vs.count.float().mean()))
