"""Per-block relative error of the bf16 tensor-core trunk against the fp32 check-mode trunk (both on the GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deephumor_b200.runtime import encoder as E, ops
from deephumor_b200.utils import synth, synth_weights

kind = 'lstm'
hp = synth_weights.default_hp(kind, 1000, small=True)
sd = synth_weights.make_state_dict(kind, hp, seed=1)
imgs = synth.images(0, 0, 4).cuda()
rts = {dt: E.EncoderRT(sd, 'encoder', dt, torch.device('cuda')) for dt in (torch.float32, torch.bfloat16)}


def trunk_taps(rt, images):
    taps = {}
    n, _, H, W = images.shape
    if rt.dtype == torch.bfloat16:
        Ho = 112
        A = rt._buf('stemA', (n * Ho * Ho, 192)); ops.im2col_stem(images, A, 7, 7, 2, 3)
        x = rt._buf('stem', (n, Ho, Ho, 64)); ops.gemm(A, rt.stem_w, x.view(n * Ho * Ho, 64), bias=rt.stem.bias, relu=True)
    else:
        x = rt._buf('in', (n, H, W, 4)); ops.nchw_to_nhwc4(images, x, halo=0); x = rt._conv('stem', x, rt.stem, True)
    taps['stem'] = x.float().clone()
    y = rt._buf('pool', (n, 56, 56, 64)); ops.maxpool3x3s2(x, y); x = y
    taps['pool'] = x.float().clone()
    for i, blk in enumerate(rt.blocks):
        y1 = rt._conv(f'b{i}c1', x, blk['c1'], True)
        y2 = rt._conv(f'b{i}c2', y1, blk['c2'], True)
        idn = rt._conv(f'b{i}ds', x, blk['down'], False) if blk['down'] is not None else x
        x = rt._conv(f'b{i}c3', y2, blk['c3'], True, residual=idn)
        taps[f'b{i}.c1'] = y1.float().clone(); taps[f'b{i}.c2'] = y2.float().clone(); taps[f'b{i}'] = x.float().clone()
    return taps

with torch.no_grad():
    t32 = trunk_taps(rts[torch.float32], imgs)
    t16 = trunk_taps(rts[torch.bfloat16], imgs)
    for k in t32:
        a, b = t16[k].double(), t32[k].double()
        print(f'{k:8s} rel {float((a-b).norm()/b.norm()):.4e}  mean {float(b.mean()):+.3f} std {float(b.std()):.3f} absmax {float(b.abs().max()):.1f}')
    s32, _ = rts[torch.float32].forward(imgs); s16, _ = rts[torch.bfloat16].forward(imgs)
    print('emb rel', float((s16.double()-s32.double()).norm()/s32.double().norm()))
    p32 = t32['b15'].mean(dim=(1, 2)); p16 = t16['b15'].mean(dim=(1, 2))
    print('pooled rel', float((p16-p32).norm()/p32.norm()), 'centered rel', float((p16-p32).norm()/(p32-p32.mean(0,keepdim=True)).norm()))
