import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import ops
dev='cuda'
M,N,K=2560,36541,512
W=(torch.randn(N,K,device=dev)*0.5).to(torch.bfloat16)
W=(torch.randn(N,K,device=dev)*0.5).to(torch.bfloat16)
A=torch.randn(M,K,device=dev).to(torch.bfloat16)
if os.environ.get('CORR'):   # strongly correlated rows (what random-init LSTM states look like)
    A=(torch.randn(1,K,device=dev)+0.2*torch.randn(M,K,device=dev)).to(torch.bfloat16)
b=torch.randn(N,device=dev)
ldc=(N+3)//4*4
out=torch.empty(M,ldc,device=dev)
ind=torch.empty(M,5,dtype=torch.int32,device=dev); val=torch.empty(M,5,device=dev); status=torch.zeros(1,dtype=torch.int32,device=dev)
def t(fn,it=10,flush=True):
    fn(); torch.cuda.synchronize()
    fl=torch.empty(256*1024*1024,dtype=torch.uint8,device=dev)
    ts=[]
    for _ in range(it):
        if flush: fl.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return min(ts), sum(ts)/len(ts)
print('nobias', t(lambda: ops.gemm(A,W,out[:,:N])))
print('bias  ', t(lambda: ops.gemm(A,W,out[:,:N],bias=b)))
print('bias noflush', t(lambda: ops.gemm(A,W,out[:,:N],bias=b),flush=False))
for tn in (64,128,256):
    print('bias tile',tn, t(lambda: ops.gemm(A,W,out[:,:N],bias=b,tile_n=tn)))
print('select', t(lambda: ops.select_tokens(out[:,:N],N,5,50,1.0,1,5,1,1,0,3,None,ind,val,status)))
def both():
    ops.gemm(A,W,out[:,:N],bias=b); ops.select_tokens(out[:,:N],N,5,50,1.0,1,5,1,1,0,3,None,ind,val,status)
print('gemm+select', t(both))
# bf16 logits variant
out16=torch.empty(M,ldc+(-ldc)%8,dtype=torch.bfloat16,device=dev)
print('bf16 out', t(lambda: ops.gemm(A,W,out16[:,:N],bias=b)))
# fused two-pass path (logits never stored)
from deephumor_b200._lib import LIB, ptr, stream
vs = ops.VocabSelect(M, N, 50, dev, stride=int(os.environ.get('STRIDE', '0')) or None)
print('pass-1 tile stride', vs.stride, 'rank', vs.rank, 'groups', vs.groups(0))
Ab = A
args = (ptr(Ab), K, ptr(W), K, 1, ptr(b), M, N, K)
print('pass1 groupmax', t(lambda: LIB.call('dh_vocab_groupmax', *args, vs.stride, 0, ptr(vs.gmax), vs.n_groups_full, stream())))
print('threshold     ', t(lambda: LIB.call('dh_vocab_threshold', ptr(vs.gmax), vs.n_groups_full, M, vs.groups(0), vs.rank, ptr(vs.thresh), ptr(vs.count), stream())))
lists = (ptr(vs.thresh), ptr(vs.count), ptr(vs.sp_logits), vs.sp_ld, ptr(vs.hitmap), vs.hit_ld)
def p2():
    vs.count.zero_()
    LIB.call('dh_vocab_candidates', *args, *lists, stream())
print('pass2 candidates (+zero)', t(p2))
import ctypes
print('select_candidates', t(lambda: LIB.call('dh_select_candidates', ctypes.byref(vs.c), M, 5, 50, 1.0, 1, 5, 1, 1, 0, 3, None, ptr(ind), ptr(val), ptr(status), None, stream())))
print('fused total', t(lambda: vs.run(A, W, b, 5, 1.0, 1, 5, 1, 3, None, ind, val, status, None, seed=1)))
print('stored groups per row: mean %.1f max %d' % (float(vs.count.float().mean()), int(vs.count.max())), ' status', int(status))
if vs.stride > 1:
    lo, hi = 50, vs.GROUP_CAP
    def fix():
        LIB.call('dh_vocab_groupmax_fix', *args, ptr(vs.gmax), vs.n_groups_full, ptr(vs.count), lo, hi, ptr(vs.flag), stream())
        LIB.call('dh_vocab_threshold_fix', ptr(vs.gmax), vs.n_groups_full, M, vs.n_groups_full, 50, ptr(vs.thresh), ptr(vs.count), lo, hi, ptr(vs.redo), ptr(vs.flag), stream())
        LIB.call('dh_vocab_candidates_fix', *args, *lists, ptr(vs.redo), ptr(vs.flag), stream())
    print('fix-up x3 (nothing to fix)', t(fix))
    print('rows below top_k:', int((vs.count < 50).sum()), 'min count', int(vs.count.min()), 'redo', int(vs.redo.sum()))
