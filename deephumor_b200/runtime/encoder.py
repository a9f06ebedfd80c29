"""Encoder runtime: packed ResNet-50 trunk + embedding heads on the C-ABI kernels.

Reference: models/encoders.py:22-70 (ImageEncoder), :96-106 (LabelEncoder), :129-144 (ImageLabelEncoder);
torchvision resnet.py:143-163,197-204,266-279.  Packing (one-time, torch ops on the device): eval-mode BN folded
into conv weights and a bias, OIHW -> [Cout][kh][kw][Cin] for NHWC implicit GEMM,
BatchNorm1d folded into the global head, cast to the storage type of the precision mode.
"""
import os

import torch

from . import ops

RESNET_BLOCKS = (3, 4, 6, 3)
EPS = 1e-5


def _fold_bn(sd, conv, bn):
    w = sd[conv + '.weight'].float()
    g, b = sd[bn + '.weight'].float(), sd[bn + '.bias'].float()
    m, v = sd[bn + '.running_mean'].float(), sd[bn + '.running_var'].float()
    s = g / torch.sqrt(v + EPS)
    return w * s.view(-1, 1, 1, 1), b - m * s


class PackedConv:
    __slots__ = ('w', 'bias', 'stride', 'pad', 'cout', 'kh')

    def __init__(self, sd, conv, bn, stride, pad, dtype, device, cin_pad=None):
        w, bias = _fold_bn(sd, conv, bn)
        w = w.permute(0, 2, 3, 1)                                # [Cout, kh, kw, Cin]
        if cin_pad is not None and cin_pad > w.shape[3]:
            w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[3]))
        self.w = w.contiguous().to(device=device, dtype=dtype)
        self.bias = bias.contiguous().to(device=device, dtype=torch.float32)
        self.stride, self.pad, self.cout, self.kh = stride, pad, w.shape[0], w.shape[1]


import os as _os
TRUNK_CHUNK = int(_os.environ.get('DH_TRUNK_CHUNK', '512'))     # images per trunk pass (host batches: also the copy granule)
# device-resident batches: larger passes run layer3/4 with fewer partial waves (config 4: 88.2 -> 86.5 ms per 4096 images);
# host batches keep 512 (with 1024 the copy / trunk overlap gets coarser and end to end loses 5 %)
TRUNK_CHUNK_DEVICE = int(_os.environ.get('DH_TRUNK_CHUNK_DEVICE', '1024'))
H2D_RAMP = tuple(int(v) for v in _os.environ.get('DH_H2D_RAMP', '64,192,256').split(','))


def h2d_schedule(n_images, chunk, ramp=None):
    """Sizes of the host->device copies / trunk passes of a pinned host batch: a ramp (64, 192, 256 by default), then full
    trunk chunks (512 images run the layer3/4 convolutions ~6 % faster per image than 256: fewer partial waves;
    profiles/r01_bench_conv_n512.txt).  Copies are sequential and faster than the trunk, so the GPU only starves when a
    chunk grows faster than the trunk of the previous one takes: a ramp step below 256 is taken while at least twice its
    size is left, the 256 step only when more than a full chunk is left."""
    sizes, left = [], n_images
    for want in (H2D_RAMP if ramp is None else ramp):
        if want <= chunk and ((left >= 2 * want) if want < 256 else (left > chunk)):
            sizes.append(want)
            left -= want
    while left > 0:
        sizes.append(min(chunk, left))
        left -= sizes[-1]
    return sizes


class EncoderRT:
    """prefix = 'encoder' (ImageEncoder) or 'encoder.image_encoder' (inside ImageLabelEncoder)."""

    def __init__(self, sd, prefix, dtype, device, spatial=False, label_prefix=None, fuse_prefix=None, chunk=None):
        chunk = chunk or TRUNK_CHUNK
        self.dtype, self.device, self.spatial, self.chunk = dtype, device, spatial, chunk
        # Tensor-core mode stores the trunk (weights + activations) in fp16, not bf16: same tcgen05 kind::f16 rate,
        # 3 more significand bits.  bf16 storage leaves ~0.7 % error on the 7x7 feature map, which the mean-centring
        # BatchNorm1d of the global head amplifies beyond the 1e-2 budget (numerics); activations of a
        # BN-folded ResNet-50 stay far below the fp16 range.  The decoders stay bf16.
        self.tdtype = tdtype = torch.float16 if dtype == torch.bfloat16 else dtype
        p = prefix + '.resnet'
        mk = lambda conv, bn, s, pad, cin_pad=None: PackedConv(sd, conv, bn, s, pad, tdtype, device, cin_pad)
        self.stem = mk(p + '.0', p + '.1', 2, 3, cin_pad=4)
        if tdtype != torch.float32:
            # tensor-core mode: stem as an explicit (r,s,c) gather + contraction, K = 147 zero-padded to 192
            w, _ = _fold_bn(sd, p + '.0', p + '.1')
            wp = torch.zeros(w.shape[0], 192)
            wp[:, :147] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 147)
            self.stem_w = wp.contiguous().to(device=device, dtype=tdtype)
            # fused stem + maxpool kernel: K slot r*22 + s*3 + c (csrc/stem_tc.cu)
            wf = torch.zeros(w.shape[0], 7, 22)
            wf[:, :, :21] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 7, 21)
            wq = torch.zeros(w.shape[0], 192)
            wq[:, :154] = wf.reshape(w.shape[0], 154)
            self.stem_wq = wq.contiguous().to(device=device, dtype=tdtype)
        self.blocks = []
        for li, nblk in enumerate(RESNET_BLOCKS):
            for b in range(nblk):
                q = f'{p}.{4 + li}.{b}'
                stride = 2 if (b == 0 and li > 0) else 1
                self.blocks.append(dict(
                    c1=mk(q + '.conv1', q + '.bn1', 1, 0), c2=mk(q + '.conv2', q + '.bn2', stride, 1),
                    c3=mk(q + '.conv3', q + '.bn3', 1, 0),
                    down=mk(q + '.downsample.0', q + '.downsample.1', stride, 0) if b == 0 else None))
        if tdtype != torch.float32:
            # first block of every stage: conv3 and the downsample branch as ONE contraction over [y2 | x] (dh_conv1x1_dual_tc)
            for blk in self.blocks:
                if blk['down'] is not None:
                    c3, dn = blk['c3'], blk['down']
                    blk['dual_w'] = torch.cat([c3.w.reshape(c3.cout, -1), dn.w.reshape(dn.cout, -1)], dim=1).contiguous()
                    blk['dual_b'] = (c3.bias + dn.bias).contiguous()
        W = sd[prefix + '.linear.weight'].float()
        b = sd[prefix + '.linear.bias'].float()
        g, be = sd[prefix + '.bn.weight'].float(), sd[prefix + '.bn.bias'].float()
        m, v = sd[prefix + '.bn.running_mean'].float(), sd[prefix + '.bn.running_var'].float()
        s = g / torch.sqrt(v + EPS)
        to = lambda t, dt=dtype: t.contiguous().to(device=device, dtype=dt)
        # Global head (Linear + eval BN1d folded, label mean, fusion Linear): always true fp32 (FFMA kernel).  It is
        # 2 MFLOP/image, and BN1d subtracts the common mean of the pooled features, which would amplify bf16
        # rounding of the operands several-fold (numerics).
        f32 = torch.float32
        self.Wg, self.bg = to(W * s.view(-1, 1), f32), to((b - m) * s + be, f32)            # Linear + BN1d (eval)
        self.Wsp, self.bsp = to(W, tdtype), to(b, torch.float32)                                   # spatial: no BN (Q20)
        self.E = W.shape[0]
        self.label_table = self.Wl = self.bl = None
        if label_prefix is not None:
            self.label_table = to(sd[label_prefix + '.embedding.weight'].float(), f32)
            self.Wl, self.bl = to(sd[fuse_prefix + '.linear.weight'].float(), f32), to(sd[fuse_prefix + '.linear.bias'].float(), f32)
        self._ws = {}

    # ---------------------------------------------------------------- trunk
    def _buf(self, name, shape, dtype=None):
        """Workspace `name` viewed as `shape`: ONE allocation per name, grown to the largest size requested so far, so that
        varying batch / chunk sizes (copy ramp, tail chunks, a server's changing N) do not accumulate a set of buffers each."""
        dtype = dtype or self.tdtype
        numel = 1
        for d in shape:
            numel *= int(d)
        t = self._ws.get(name)
        if t is None or t.dtype != dtype or t.numel() < numel:
            t = torch.empty(max(numel, 1), dtype=dtype, device=self.device)
            self._ws[name] = t
        return t[:numel].view(*shape)

    def _conv(self, name, x, pc, relu, residual=None):
        n, H, W, _ = x.shape
        Ho = (H + 2 * pc.pad - pc.kh) // pc.stride + 1
        Wo = (W + 2 * pc.pad - pc.kh) // pc.stride + 1          # square kernels; H and W may differ
        y = self._buf(name, (n, Ho, Wo, pc.cout))
        ops.conv2d(x, pc.w, pc.bias, y, pc.stride, pc.pad, relu, residual)
        return y

    # ImageNet statistics of the reference's preprocessing (deephumor_demo.ipynb cell 11: ToTensor + Normalize)
    pixel_mean = (0.485, 0.456, 0.406)
    pixel_std = (0.229, 0.224, 0.225)

    def normalize(self, images_u8):
        """uint8 [n,3,H,W] -> fp32 (x / 255 - mean) / std, the operation order of ToTensor + Normalize (check mode and
        sizes the fused stem does not cover)."""
        m = torch.tensor(self.pixel_mean, dtype=torch.float32, device=images_u8.device).view(1, 3, 1, 1)
        sd = torch.tensor(self.pixel_std, dtype=torch.float32, device=images_u8.device).view(1, 3, 1, 1)
        return ((images_u8.float() / 255.0) - m) / sd

    def trunk(self, images, pooled=None):
        """images [n,3,224,224] fp32 NCHW (device) -> features [n,7,7,2048] NHWC in the trunk storage dtype.
        pooled (fp32 [n,2048], optional) receives the global average pool; returns (features, pooled_written)."""
        n, _, H, W = images.shape
        if (ops.PATH_ENTRIES and self.tdtype != torch.float32 and H == 224 and W == 224 and ops.FUSED_STEM and ops.DUAL_CONV
                and ops.HALO_CONV and ops.FUSED_POOL and images.dtype in (torch.uint8, torch.float32)):
            # the whole trunk (+ pooled epilogue) behind one path-level C entry: dh_resnet50_forward
            if getattr(self, '_ctx', None) is None:
                self._ctx = ops.resnet50_ctx(self)
            feat = self._buf('feat', (n, 7, 7, 2048))
            ops.resnet50_forward(self._ctx, images, feat, pooled, lambda nb: self._buf('trunk_ws', (nb,), torch.uint8))
            return feat, pooled is not None
        if images.dtype == torch.uint8:
            if self.tdtype != torch.float32 and H == 224 and W == 224 and ops.FUSED_STEM:
                x = self._buf('pool', (n, 56, 56, 64))
                ops.stem_pool_u8(images, self.pixel_mean, self.pixel_std, self.stem_wq, self.stem.bias, x)
                return self._layers(x, pooled)
            images = self.normalize(images)
        if self.tdtype != torch.float32 and H == 224 and W == 224 and ops.FUSED_STEM:
            x = self._buf('pool', (n, 56, 56, 64))
            ops.stem_pool(images, self.stem_wq, self.stem.bias, x)
            return self._layers(x, pooled)
        if self.tdtype != torch.float32:
            Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
            A = self._buf('stemA', (n * Ho * Wo, 192))
            ops.im2col_stem(images, A, 7, 7, 2, 3)
            x = self._buf('stem', (n, Ho, Wo, 64))
            ops.gemm(A, self.stem_w, x.view(n * Ho * Wo, 64), bias=self.stem.bias, relu=True)
        else:
            x = self._buf('in', (n, H, W, 4))
            ops.nchw_to_nhwc4(images, x, halo=0)
            x = self._conv('stem', x, self.stem, True)
        y = self._buf('pool', (n, (x.shape[1] - 1) // 2 + 1, (x.shape[2] - 1) // 2 + 1, 64))
        ops.maxpool3x3s2(x, y)
        return self._layers(y, pooled)

    def _layers(self, x, pooled=None):
        last = len(self.blocks) - 1
        for i, blk in enumerate(self.blocks):
            y1 = self._conv(f'b{i}c1', x, blk['c1'], True)
            y2 = self._conv(f'b{i}c2', y1, blk['c2'], True)
            if 'dual_w' in blk and ops.DUAL_CONV:
                n, Ho, Wo, _ = y2.shape
                out = self._buf(f'b{i}c3', (n, Ho, Wo, blk['c3'].cout))
                ops.conv1x1_dual(y2, x, blk['dual_w'], blk['dual_b'], out, blk['down'].stride, True)
                x = out
                continue
            idn = self._conv(f'b{i}ds', x, blk['down'], False) if blk['down'] is not None else x
            n, Ho, Wo, C2 = y2.shape
            if (i == last and pooled is not None and ops.FUSED_POOL and self.tdtype != torch.float32 and Ho * Wo <= 128
                    and blk['down'] is None):
                # conv3 + bn3 + identity + ReLU of the last bottleneck with AdaptiveAvgPool2d((1,1)) in its epilogue
                # (encoders.py:39,60): M tiles of whole images, the map is never re-read for the pooled vector
                c3 = blk['c3']
                out = self._buf(f'b{i}c3', (n, Ho, Wo, c3.cout))
                ops.gemm_pool(y2.view(n * Ho * Wo, C2), c3.w.view(c3.cout, C2), out.view(n * Ho * Wo, c3.cout), pooled,
                              Ho * Wo, bias=c3.bias, residual=idn.view(n * Ho * Wo, c3.cout), relu=True)
                return out, True
            x = self._conv(f'b{i}c3', y2, blk['c3'], True, residual=idn)
        return x, False

    def _host_chunks(self, images):
        """Pinned HOST images -> device chunks, copied on a side stream into two staging buffers so the H2D transfer
        of chunk k+1 overlaps the trunk of chunk k (yields (first index, device chunk, event to record when consumed)).
        The schedule starts small (64, then 192, then 256 images) so the trunk starts after 0.7 ms of copying, then uses
        the full trunk chunk, which runs the convolutions at their best rate."""
        N = images.shape[0]
        main = torch.cuda.current_stream()
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._copied = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
        cs = self._copy_stream
        sizes = h2d_schedule(N, self.chunk)
        bufs = [self._buf(f'h2d{b}.{images.dtype}', (self.chunk,) + tuple(images.shape[1:]), images.dtype) for b in range(2)]
        cs.wait_stream(main)                           # earlier readers of the staging buffers are done
        i0 = 0
        for k, n in enumerate(sizes):
            b = k & 1
            with torch.cuda.stream(cs):
                if k >= 2:
                    cs.wait_event(self._consumed[b])
                bufs[b][:n].copy_(images[i0:i0 + n], non_blocking=True)
                self._copied[b].record(cs)
            main.wait_event(self._copied[b])
            yield i0, bufs[b][:n], self._consumed[b]
            i0 += n

    # ---------------------------------------------------------------- heads
    def forward(self, images, labels=None):
        """-> (start_emb fp32 [N,E], spatial [N*49,E] storage dtype or None).  Processes `chunk` images at a time."""
        N = images.shape[0]
        E = self.E
        start = torch.empty(N, E, dtype=torch.float32, device=self.device)
        sp = torch.empty(N * 49, E, dtype=self.dtype, device=self.device) if self.spatial else None
        if N == 0:
            return start, sp
        pooled = self._buf('pooled', (N, 2048), torch.float32)
        dchunk = max(self.chunk, TRUNK_CHUNK_DEVICE) if N >= 2 * TRUNK_CHUNK_DEVICE else self.chunk
        chunks = self._host_chunks(images) if not images.is_cuda else \
            ((i0, images[i0:i0 + dchunk], None) for i0 in range(0, N, dchunk))
        for i0, img, consumed in chunks:
            n = img.shape[0]
            with ops.PROFILE.range('encoder_trunk', 8.174e9 * n):
                feat, has_pool = self.trunk(img.contiguous(), pooled[i0:i0 + n])
            if consumed is not None:
                consumed.record()                      # the staging buffer may be overwritten by the next-but-one copy
            hw = feat.shape[1] * feat.shape[2]
            if not has_pool:
                ops.avgpool(feat.view(n, hw, 2048), pooled[i0:i0 + n])
            if self.spatial and hw != 49:
                raise ValueError(f'spatial features need 224 x 224 images (7 x 7 = 49 tokens for the cross-attention decoder); '
                                 f'got a {feat.shape[1]} x {feat.shape[2]} feature map')
            if self.spatial:
                ops.gemm(feat.view(n * hw, 2048), self.Wsp, sp[i0 * 49:(i0 + n) * 49], bias=self.bsp)
        # global heads once for the whole batch (fp32 FFMA, 64x64 tiles)
        with ops.PROFILE.range('encoder_heads'):
            if self.label_table is None:
                ops.gemm(pooled, self.Wg, start, bias=self.bg)
            else:
                cat = self._buf('cat', (N, 2 * E), torch.float32)
                ops.gemm(pooled, self.Wg, cat[:, :E], bias=self.bg)
                ops.embed_mean(self.label_table, labels.contiguous(), cat[:, E:])
                ops.gemm(cat, self.Wl, start, bias=self.bl)
        return start, sp
