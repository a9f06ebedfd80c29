"""One decode step's selection chain (pass 1 -> threshold -> pass 2 -> select) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full`.  Shapes = BASELINE configs[1]: rows 2560, V 36541, K 512, top-k 50, beam 5."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
dev = 'cuda'
M, N, K = 2560, 36541, 512
A = (torch.randn(M, K, device=dev) * 0.3).to(torch.bfloat16)
W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(N, device=dev) * 0.1
ind = torch.empty(M, 5, dtype=torch.int32, device=dev); val = torch.empty(M, 5, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
vs = ops.VocabSelect(M, N, 50, dev)
run = lambda: vs.run(A, W, b, 5, 1.0, 1, 5, 1, 3, None, ind, val, status, None, seed=1)
for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('candidates per row: mean %.1f max %d' % (float(vs.count.float().mean()), int(vs.count.max())))
