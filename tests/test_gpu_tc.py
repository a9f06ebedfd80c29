"""Tensor-core (tcgen05 / TMEM / TMA) kernels against plain PyTorch fp32 references of the same op on the same
bf16-rounded operands.  Tolerance: fp32 accumulation order only (1e-5 relative for fp32 output; bf16 output adds
one rounding, 2^-8 relative per element -> 4e-3 on the norm)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from deephumor_b200.runtime import ops
from tests import helpers as H

DEV = 'cuda'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def bf(x):
    return x.to(torch.bfloat16)


GEMM_SHAPES = [(128, 128, 64), (128, 256, 128), (1, 64, 64), (5, 36541, 512), (130, 129, 72), (300, 512, 1024),
               (64, 2048, 512), (2560, 1000, 160), (777, 200, 2048), (4096, 384, 512)]


@pytest.mark.parametrize('tile_n', [0, 64, 128, 256])
@pytest.mark.parametrize('M,N,K', GEMM_SHAPES)
def test_gemm_bf16_fp32_out(M, N, K, tile_n):
    A, W, b, r = bf(rnd(M, K, seed=1)), bf(rnd(N, K, seed=2)), rnd(N, seed=3), rnd(M, N, seed=4)
    ldc = (N + 3) // 4 * 4
    out = torch.full((M, ldc), 7.0, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), out[:, :N], bias=b.to(DEV), residual=r.to(DEV), relu=True, tile_n=tile_n)
    ref = F.relu(A.double() @ W.double().T + b.double() + r.double()).float()
    assert H.rel_err(out[:, :N], ref) < 1e-5
    if ldc > N:
        assert float((out[:, N:] - 7.0).abs().max()) == 0.0           # padding columns untouched


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (333, 96, 160), (2560, 2048, 1024)])
def test_gemm_bf16_bf16_out_strided(M, N, K):
    """bf16 output and residual, A a column slice of a wider buffer (the LSTM [x|h] operand), no bias / relu."""
    wide = bf(rnd(M, K + 64, seed=1)).to(DEV)
    A = wide[:, 64:]
    W, r = bf(rnd(N, K, seed=2)), bf(rnd(M, N, seed=4))
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W.to(DEV), out, residual=r.to(DEV))
    ref = (A.cpu().double() @ W.double().T + r.double()).float()
    assert H.rel_err(out.float(), ref) < 4e-3
    out32 = torch.empty(M, N, device=DEV)
    ops.gemm(A, W.to(DEV), out32, residual=r.to(DEV))
    assert H.rel_err(out32, ref) < 1e-5


CONV_SHAPES = [  # n, H, Cin, Cout, k, stride, pad
    (2, 56, 64, 64, 1, 1, 0), (2, 56, 64, 64, 3, 1, 1), (3, 56, 256, 128, 1, 1, 0), (2, 56, 128, 128, 3, 2, 1),
    (2, 56, 256, 512, 1, 2, 0), (3, 14, 256, 256, 3, 1, 1), (5, 7, 512, 512, 3, 1, 1), (5, 7, 512, 2048, 1, 1, 0),
    (2, 14, 1024, 2048, 1, 2, 0), (1, 9, 64, 64, 3, 2, 1), (7, 14, 512, 512, 3, 2, 1),
]


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('n,H_,Cin,Cout,k,s,p', CONV_SHAPES)
def test_conv2d_tc_im2col_tma(n, H_, Cin, Cout, k, s, p, dt):
    x, w, b = rnd(n, Cin, H_, H_, seed=1).to(dt), rnd(Cout, Cin, k, k, seed=2, scale=0.05).to(dt), rnd(Cout, seed=3)
    ref = F.conv2d(x.double(), w.double(), b.double(), s, p)
    res = rnd(*ref.shape, seed=4).to(dt)
    ref = F.relu(ref + res.double()).float().permute(0, 2, 3, 1).contiguous()
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wd = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    rd = res.permute(0, 2, 3, 1).contiguous().to(DEV)
    y = torch.empty(ref.shape, dtype=dt, device=DEV)
    ops.conv2d(xd, wd, b.to(DEV), y, s, p, True, residual=rd)
    assert H.rel_err(y.float(), ref) < (4e-3 if dt == torch.bfloat16 else 5e-4)
    # the explicit-gather route through the same contraction kernel must agree with the TMA route bit for bit
    Ho = ref.shape[1]
    A = torch.empty(n * Ho * Ho, k * k * Cin, dtype=dt, device=DEV)
    ops.im2col_nhwc(xd, A, k, k, s, p)
    y2 = torch.empty(n * Ho * Ho, Cout, dtype=dt, device=DEV)
    ops.gemm(A, wd.view(Cout, -1), y2, bias=b.to(DEV), residual=rd.view(-1, Cout), relu=True)
    assert torch.equal(y2.view_as(y), y)


@pytest.mark.parametrize('n,Ho,C1,C2,Cout,s2', [(3, 56, 64, 64, 256, 1), (5, 28, 128, 256, 512, 2), (9, 14, 256, 512, 1024, 2),
                                              (40, 7, 512, 1024, 2048, 2), (2, 5, 64, 128, 64, 2)])
def test_conv1x1_dual_conv3_plus_downsample(n, Ho, C1, C2, Cout, s2):
    """relu(conv3(y2) + downsample(x)) of a stage's first bottleneck as one contraction over [y2 | x] (dh_conv1x1_dual_tc:
    second im2col source with its own stride) against a float64 reference (torchvision resnet.py:154-161)."""
    dt = torch.float16
    H2 = Ho * s2 - (s2 - 1) if s2 > 1 and Ho == 5 else Ho * s2          # an odd input size too: (H2 - 1) // s2 + 1 == Ho
    y2 = rnd(n, C1, Ho, Ho, seed=1).to(dt)
    x = rnd(n, C2, H2, H2, seed=2).to(dt)
    w3, wd = rnd(Cout, C1, 1, 1, seed=3, scale=0.05).to(dt), rnd(Cout, C2, 1, 1, seed=4, scale=0.05).to(dt)
    b3, bd = rnd(Cout, seed=5), rnd(Cout, seed=6)
    ref = F.relu(F.conv2d(y2.double(), w3.double(), b3.double()) + F.conv2d(x.double(), wd.double(), bd.double(), s2))
    ref = ref.float().permute(0, 2, 3, 1).contiguous()
    w_cat = torch.cat([w3.view(Cout, C1), wd.view(Cout, C2)], 1).contiguous().to(DEV)
    out = torch.empty(n, Ho, Ho, Cout, dtype=dt, device=DEV)
    ops.conv1x1_dual(y2.permute(0, 2, 3, 1).contiguous().to(DEV), x.permute(0, 2, 3, 1).contiguous().to(DEV), w_cat,
                     (b3 + bd).to(DEV), out, s2, True)
    assert H.rel_err(out.float(), ref) < 5e-4


@pytest.mark.parametrize('dt', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('n,H_,W_,C1,source,Cout,N2,s2', [(3, 56, 56, 64, True, 256, 64, 1), (2, 56, 56, 64, False, 256, 64, 1),
                                                          (5, 56, 56, 64, False, 256, 128, 1), (1, 8, 8, 64, False, 256, 64, 1),
                                                          (1, 9, 7, 128, True, 256, 256, 1), (40, 56, 56, 64, False, 256, 64, 1),
                                                          (37, 28, 20, 64, True, 256, 128, 1), (9, 28, 28, 128, True, 512, 128, 2),
                                                          (33, 28, 28, 128, False, 512, 128, 1), (6, 28, 28, 128, False, 512, 256, 1),
                                                          (2, 5, 5, 64, True, 768, 64, 2)])
def test_conv1x1_chain_conv3_then_next_conv1(n, H_, W_, C1, source, Cout, N2, s2, dt):
    """dh_conv1x1_chain_tc (a bottleneck's conv3 + identity / downsample + ReLU and the next bottleneck's conv1 + ReLU in one
    launch, the second contraction reading the first one's tiles back from L2; torchvision resnet.py:154-161, :146-148) is
    BIT-IDENTICAL to the separate launches: one unit per CTA, many units per CTA, clipped last tiles, Cout of 1 - 3 column
    tiles (the layer1 / layer2 shapes) and a strided downsample source."""
    C2 = 64 * s2 if source else Cout
    H2, W2 = (H_ * s2 - (s2 - 1), W_ * s2 - (s2 - 1)) if source and s2 > 1 else (H_, W_)
    y2 = rnd(n, H_, W_, C1, seed=1).to(dt).to(DEV)
    x2 = rnd(n, H2, W2, C2, seed=2).to(dt).to(DEV)
    w = rnd(Cout, C1 + (C2 if source else 0), seed=3, scale=0.05).to(dt).to(DEV)
    b, wn, bn_ = rnd(Cout, seed=4).to(DEV), rnd(N2, Cout, seed=5, scale=0.05).to(dt).to(DEV), rnd(N2, seed=6).to(DEV)
    out = torch.zeros(n, H_, W_, Cout, dtype=dt, device=DEV)
    z = torch.zeros(n, H_, W_, N2, dtype=dt, device=DEV)
    ops.conv1x1_chain(y2, x2, source, w, b, out, wn, bn_, z, stride2=s2)
    out_ref = torch.zeros_like(out)
    z_ref = torch.zeros_like(z)
    if source:
        ops.conv1x1_dual(y2, x2, w, b, out_ref, s2, True)
    else:
        ops.conv2d(y2, w.view(Cout, 1, 1, C1), b, out_ref, 1, 0, True, residual=x2)
    ops.conv2d(out_ref, wn.view(N2, 1, 1, Cout), bn_, z_ref, 1, 0, True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_ref)
    assert torch.equal(z, z_ref)
    xs = x2[:, ::s2, ::s2] if source else x2
    ref = F.relu(y2.double().reshape(-1, C1) @ w.double()[:, :C1].T + (xs.double().reshape(-1, C2) @ w.double()[:, C1:].T if source
                                                                        else xs.double().reshape(-1, Cout)) + b.double())
    assert H.rel_err(out.float().view(-1, Cout), ref.float()) < (5e-3 if dt == torch.bfloat16 else 6e-4)


@pytest.mark.parametrize('dt', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('n,H_,Cin,Cout', [(2, 56, 64, 64), (3, 28, 128, 128), (1, 30, 64, 128), (5, 56, 128, 64), (2, 33, 192, 64)])
def test_conv3x3_halo_kernel(n, H_, Cin, Cout, dt):
    """Halo-tile 3x3 convolution (input read once, taps as shifted UMMA descriptor views) against torch and against
    the im2col-TMA kernel; H not a multiple of the 8 x 16 tile exercises the clipped stores and the zero-filled halo."""
    x, w, b = rnd(n, Cin, H_, H_, seed=1).to(dt), rnd(Cout, Cin, 3, 3, seed=2, scale=0.05).to(dt), rnd(Cout, seed=3)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), 1, 1)).float().permute(0, 2, 3, 1).contiguous()
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wd = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    y = torch.full(ref.shape, 7.0, dtype=dt, device=DEV)
    from deephumor_b200._lib import LIB, ptr, stream
    bd = b.to(DEV)
    LIB.call('dh_conv3x3_halo_tc', ptr(xd), ptr(wd), ptr(bd), ptr(y), n, H_, H_, Cin, Cout, 1, ops.code(xd), stream())
    y2 = torch.empty_like(y)
    ops.conv2d(xd, wd, b.to(DEV), y2, 1, 1, True, tile_n=64)           # forces the im2col-TMA kernel
    torch.cuda.synchronize()
    tol = 4e-3 if dt == torch.bfloat16 else 5e-4
    assert H.rel_err(y.float(), ref) < tol and H.rel_err(y2.float(), ref) < tol
    assert H.rel_err(y.float(), y2.float()) < tol


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
def test_stem_im2col_gemm(dt):
    n, H_ = 2, 64
    img = rnd(n, 3, H_, H_, seed=1)
    w, b = rnd(64, 3, 7, 7, seed=2, scale=0.1), rnd(64, seed=3)
    Ho = (H_ + 6 - 7) // 2 + 1
    A = torch.empty(n * Ho * Ho, 192, dtype=dt, device=DEV)
    ops.im2col_stem(img.to(DEV), A, 7, 7, 2, 3)
    wp = torch.zeros(64, 192)
    wp[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147)
    y = torch.empty(n * Ho * Ho, 64, dtype=dt, device=DEV)
    ops.gemm(A, wp.to(dt).to(DEV), y, bias=b.to(DEV), relu=True)
    ref = F.relu(F.conv2d(img.to(dt).double(), w.to(dt).double(), b.double(), 2, 3)).float().permute(0, 2, 3, 1).reshape(-1, 64)
    assert H.rel_err(y.float(), ref) < (4e-3 if dt == torch.bfloat16 else 5e-4)


@pytest.mark.parametrize('stride', [1, 4, 8])
@pytest.mark.parametrize('mode', ['deterministic', 'injected'])
@pytest.mark.parametrize('rows,V,K_,B,top_k,T', [(300, 36541, 512, 5, 50, 1.0), (7, 1000, 64, 3, 10, 0.8), (130, 4099, 128, 1, 1, 1.3),
                                                 (64, 2048, 256, 4, 64, 1.0), (5, 71, 64, 2, 2, 1.0), (5, 100, 64, 1, 1, 1.0),
                                                 (40, 50001, 64, 3, 20, 1.0), (300, 4099, 1024, 5, 50, 1.0)])
def test_fused_vocab_select_matches_materialised_logits(mode, rows, V, K_, B, top_k, T, stride):
    """Two-pass vocab projection (group maxima -> threshold -> candidate compaction, logits never stored) picks
    exactly what dh_select_tokens and the CPU oracle pick on the materialised logits of the same tcgen05 product."""
    from oracle import model as omodel, noise as onoise
    g = torch.Generator().manual_seed(V + rows)
    A = (torch.randn(rows, K_, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(V, K_, generator=g) * 0.2).to(torch.bfloat16).to(DEV)
    bias = torch.randn(V, generator=g).to(DEV)
    if top_k > 1 or rows == 5:
        bias[1] = 50.0                                    # <unk> is always in the top-k: masked but counted (Q3)
    rpi = 1 if rows % B else B
    ldv = (V + 3) // 4 * 4
    logits = torch.empty(rows, ldv, device=DEV)
    ops.gemm(A, W, logits[:, :V], bias=bias)
    mk = lambda: (torch.empty(rows, B, dtype=torch.int32, device=DEV), torch.empty(rows, B, device=DEV),
                  torch.zeros(1, dtype=torch.int32, device=DEV))
    ind0, val0, st0 = mk()
    ops.select_tokens(logits[:, :V], V, B, top_k, T, 1, rpi, ops.NOISE[mode], 11, 5, 3, None, ind0, val0, st0)
    assert ops.VocabSelect.supported(A, V, top_k)
    vs = ops.VocabSelect(rows, V, top_k, DEV, stride=stride)
    ind1, val1, st1 = mk()
    vs.run(A, W, bias, B, T, 1, rpi, ops.NOISE[mode], 3, None, ind1, val1, st1, None, seed=11, image_base=5)
    torch.cuda.synchronize()
    assert int(st0.item()) == int(st1.item())
    if top_k == 1 and rows == 5:
        assert int(st1.item()) & 1                        # every row filtered (arg-max is <unk>): EMPTY_ROW like Q3
        return
    assert torch.equal(ind0, ind1) and torch.equal(val0, val1)
    assert int(vs.candidates(rows).min()) >= min(top_k, V)            # the row's top_k logits are all in its stored groups
    lc = logits[:, :V].cpu()
    for r in range(0, rows, max(1, rows // 7)):
        q = None if mode == 'deterministic' else torch.stack([onoise.exp_noise(11, 5 + r // rpi, 3, 0, r % rpi, V)])
        oi, ov = omodel.select_tokens(lc[r:r + 1], B, T, top_k, 1, q, None)
        assert ind1[r].cpu().tolist() == oi[0].tolist()
        assert torch.allclose(val1[r].cpu(), ov[0], atol=1e-5)


@pytest.mark.parametrize('stride', [4, 8])
def test_sampled_threshold_miss_is_repaired_in_stream(stride):
    """A row whose largest logits ALL sit in the tiles the sampled pass 1 visits gets a threshold above its top_k-th largest
    logit (fewer than top_k candidates); the fix-up launches must detect it from the candidate counts and redo that row
    exhaustively, in stream order, so the picks equal the materialised-logits path; untouched rows keep their lists."""
    rows, V, K_, B, top_k, step = 200, 36541, 512, 5, 50, 5
    g = torch.Generator().manual_seed(3)
    A = (torch.randn(rows, K_, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(V, K_, generator=g) * 0.2).to(torch.bfloat16).to(DEV)
    bias = torch.randn(V, generator=g)
    vs = ops.VocabSelect(rows, V, top_k, DEV, stride=stride)
    assert vs.stride == stride and vs.rank < top_k
    off = step % stride
    cols = [(off + stride * t) * 256 + 32 * gidx + 5 for t in range(7) for gidx in range(8)][:49]     # distinct sampled groups
    bias[cols] += 40.0                                      # the 49 largest logits of EVERY row; the 50th is an ordinary one
    bias = bias.to(DEV)
    logits = torch.empty(rows, (V + 3) // 4 * 4, device=DEV)
    ops.gemm(A, W, logits[:, :V], bias=bias)
    mk = lambda: (torch.empty(rows, B, dtype=torch.int32, device=DEV), torch.empty(rows, B, device=DEV),
                  torch.zeros(1, dtype=torch.int32, device=DEV))
    ind0, val0, st0 = mk()
    ops.select_tokens(logits[:, :V], V, B, top_k, 1.0, 1, B, ops.NOISE['injected'], 11, 5, step, None, ind0, val0, st0)
    ind1, val1, st1 = mk()
    vs.run(A, W, bias, B, 1.0, 1, B, ops.NOISE['injected'], step, None, ind1, val1, st1, None, seed=11, image_base=5)
    torch.cuda.synchronize()
    assert int(vs.redo.sum()) == rows and int(vs.flag) == 1                     # every row missed and was repaired
    assert int(vs.count.min()) >= top_k and int(vs.count.max()) <= vs.GROUP_CAP
    assert int(st1.item()) == 0 and torch.equal(ind0, ind1) and torch.equal(val0, val1)
    # next step: ordinary rows -> nothing to repair, the flag is cleared again
    bias2 = torch.randn(V, generator=g).to(DEV)
    vs.run(A, W, bias2, B, 1.0, 1, B, ops.NOISE['injected'], step + 1, None, ind1, val1, st1, None, seed=11, image_base=5)
    torch.cuda.synchronize()
    assert int(vs.redo.sum()) == 0 and int(vs.flag) == 0 and int(vs.count.min()) >= top_k


def test_top_k_above_64_takes_the_materialised_path():
    """the fused selection ranks at most 64 group maxima exactly; larger top_k must not be routed to it."""
    A = torch.zeros(4, 64, dtype=torch.bfloat16, device=DEV)
    assert ops.VocabSelect.supported(A, 36541, 64) and not ops.VocabSelect.supported(A, 36541, 100)
    assert not ops.VocabSelect.supported(A, 1000, 50)       # fewer than top_k 32-column groups


def _pack_stem(w, dt):
    wf = torch.zeros(64, 7, 22)
    wf[:, :, :21] = w.permute(0, 2, 3, 1).reshape(64, 7, 21)
    wq = torch.zeros(64, 192)
    wq[:, :154] = wf.reshape(64, 154)
    return wq.to(dt).to(DEV)


@pytest.mark.parametrize('n', [1, 3, 160])
@pytest.mark.parametrize('dt', [torch.float16, torch.bfloat16])
def test_fused_stem_pool(dt, n):
    """conv 7x7/2 + bias + ReLU + maxpool 3x3/2 in one tcgen05 kernel (im2col built in smem, pooling on the
    accumulators) against torch on the same half-rounded image / weights; edges exercise both paddings."""
    img = rnd(n, 3, 224, 224, seed=n)
    img[0, :, :4, :] = 3.0                                 # make the top/left borders matter
    img[0, :, :, :4] = -2.0
    w, b = rnd(64, 3, 7, 7, seed=2, scale=0.1), rnd(64, seed=3)
    out = torch.empty(n, 56, 56, 64, dtype=dt, device=DEV)
    ops.stem_pool(img.to(DEV), _pack_stem(w, dt), b.to(DEV), out)
    torch.cuda.synchronize()
    k = min(n, 4)
    sel = list(range(k - 1)) + [n - 1]
    x = img[sel].to(dt).to(DEV).double()
    ref = F.max_pool2d(F.relu(F.conv2d(x, w.to(dt).to(DEV).double(), b.to(DEV).double(), 2, 3)), 3, 2, 1)
    ref = ref.float().permute(0, 2, 3, 1)
    got = out[sel].float()
    assert H.rel_err(got, ref) < (4e-3 if dt == torch.bfloat16 else 5e-4)
    assert float((got - ref).abs().max()) < (0.05 if dt == torch.bfloat16 else 0.01)


@pytest.mark.parametrize('rows,H_,in_,use_parent', [(2560, 512, 512, True), (37, 64, 32, False), (300, 128, 256, True)])
def test_lstm_layer_fused_cell_epilogue(rows, H_, in_, use_parent):
    """Gate contraction with the cell update in the tcgen05 epilogue (MUFU.TANH activations, rel. error 2^-11) against
    dh_gemm_tc + dh_lstm_cell (precise expf / tanhf) and nn.LSTMCell on the same bf16-rounded operands: c within 1e-3,
    h within the bf16 rounding of the output."""
    g = torch.Generator().manual_seed(rows)
    K_ = in_ + H_
    A = (torch.randn(rows, K_, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(4 * H_, K_, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    b = torch.randn(4 * H_, generator=g).to(DEV)
    c_prev = torch.randn(rows, H_, generator=g).to(DEV)
    parent = torch.randint(0, rows, (rows,), generator=g).to(torch.int32).to(DEV) if use_parent else None
    gates = torch.empty(rows, 4 * H_, device=DEV)
    ops.gemm(A, W, gates, bias=b)
    c0, h0a, h0b = torch.empty(rows, H_, device=DEV), torch.empty(rows, H_ + 8, dtype=torch.bfloat16, device=DEV), \
        torch.empty(rows, H_, dtype=torch.bfloat16, device=DEV)
    ops.lstm_cell(gates, c_prev, parent, c0, h0a[:, :H_], h0b)
    c1, h1a, h1b = torch.empty_like(c0), torch.empty_like(h0a), torch.empty_like(h0b)
    ops.lstm_layer_tc(A, ops.pack_lstm_gates(W, H_), ops.pack_lstm_gates(b, H_), c_prev, parent, c1, h1a[:, :H_], h1b)
    torch.cuda.synchronize()
    assert torch.equal(h1a[:, :H_], h1b)
    assert H.rel_err(c1, c0) < 1e-3 and H.rel_err(h1b.float(), h0b.float()) < 4e-3
    cell = torch.nn.LSTMCell(in_, H_).double()
    with torch.no_grad():
        cell.weight_ih.copy_(W[:, :in_].double().cpu()); cell.weight_hh.copy_(W[:, in_:].double().cpu())
        cell.bias_ih.copy_(b.double().cpu()); cell.bias_hh.zero_()
        cp = c_prev.cpu().double() if parent is None else c_prev.cpu().double()[parent.cpu().long()]
        hr, cr = cell(A[:, :in_].double().cpu(), (A[:, in_:].double().cpu(), cp))
    assert H.rel_err(c1, cr) < 1e-3 and H.rel_err(h1b.float(), hr) < 5e-3


@pytest.mark.parametrize('rows,H_,E_,L_,rotate', [(2560, 512, 512, 3, 1), (2560, 512, 512, 3, 0), (512, 512, 512, 3, 1),
                                                 (300, 128, 64, 2, 1), (37, 64, 32, 4, 1), (1100, 256, 320, 2, 0)])
def test_lstm_stack_one_launch_equals_layer_by_layer(rows, H_, E_, L_, rotate):
    """dh_lstm_stack_tc (all layers of a time step in one persistent launch, cross-CTA ready counters) against L
    dh_lstm_layer_tc launches on the same operands: bit-identical with the natural K order, within fp32 summation-order
    noise when the upper layers start on the recurrent half.  Three consecutive steps exercise the counter regions."""
    g = torch.Generator().manual_seed(rows + L_)
    ra = rows + 5                                              # allocation taller than the live rows
    in_dims = [E_] + [H_] * (L_ - 1)
    Kmax = max(in_dims) + H_
    Ws = [(torch.randn(4 * H_, i + H_, generator=g) * 0.1).to(torch.bfloat16).to(DEV) for i in in_dims]
    bs = [torch.randn(4 * H_, generator=g).to(DEV) for _ in in_dims]
    Wp_all = torch.zeros(L_ * 4 * H_, Kmax, dtype=torch.bfloat16, device=DEV)
    for l, w in enumerate(Ws):
        Wp_all[l * 4 * H_:(l + 1) * 4 * H_, :w.shape[1]] = ops.pack_lstm_gates(w, H_)
    b_all = torch.cat([ops.pack_lstm_gates(b, H_) for b in bs]).contiguous()
    A_all = torch.zeros(L_, ra, Kmax, dtype=torch.bfloat16, device=DEV)
    A_ref = [torch.zeros(ra, i + H_, dtype=torch.bfloat16, device=DEV) for i in in_dims]
    c = [torch.zeros(L_, ra, H_, device=DEV) for _ in range(2)]
    c_ref = [torch.zeros(L_, ra, H_, device=DEV) for _ in range(2)]
    hs, hs_ref = (torch.zeros(L_, ra, H_, dtype=torch.bfloat16, device=DEV) for _ in range(2))
    top, top_ref = (torch.zeros(ra, H_, dtype=torch.bfloat16, device=DEV) for _ in range(2))
    per = (L_ - 1) * ((rows + 127) // 128)
    ready = torch.zeros(3 * per, dtype=torch.int32, device=DEV)
    cur = 0
    for t in range(3):
        x = (torch.randn(rows, E_, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        parent = torch.randint(0, rows, (rows,), generator=g).to(torch.int32).to(DEV) if t else None
        A_all[0, :rows, :E_] = x
        A_ref[0][:rows, :E_] = x
        for l in range(L_):                                    # recurrent halves: h of the previous step through parent
            src, src_ref = hs[l, :rows], hs_ref[l, :rows]
            if parent is not None:
                src, src_ref = src[parent.long()], src_ref[parent.long()]
            A_all[l, :rows, in_dims[l]:in_dims[l] + H_] = src
            A_ref[l][:rows, in_dims[l]:] = src_ref
        ops.lstm_stack_tc(A_all, in_dims, Wp_all, b_all, c[cur], parent, c[1 - cur], top[:rows], hs,
                          ready[t * per:(t + 1) * per], rows, rotate=rotate)
        for l in range(L_):
            nxt = A_ref[l + 1][:rows, :H_] if l + 1 < L_ else top_ref[:rows]
            ops.lstm_layer_tc(A_ref[l][:rows], ops.pack_lstm_gates(Ws[l], H_), ops.pack_lstm_gates(bs[l], H_), c_ref[cur][l],
                              parent, c_ref[1 - cur][l][:rows], nxt, hs_ref[l][:rows])
        cur = 1 - cur
        torch.cuda.synchronize()
        if rotate:
            assert H.rel_err(c[cur][:, :rows], c_ref[cur][:, :rows]) < 2e-3
            assert H.rel_err(top[:rows].float(), top_ref[:rows].float()) < 8e-3
            assert H.rel_err(hs[:, :rows].float(), hs_ref[:, :rows].float()) < 8e-3
        else:
            assert torch.equal(c[cur][:, :rows], c_ref[cur][:, :rows])
            assert torch.equal(top[:rows], top_ref[:rows]) and torch.equal(hs[:, :rows], hs_ref[:, :rows])
    assert int(ready.min()) == (4 * H_) // 256 and int(ready.max()) == (4 * H_) // 256
    assert ops.tc_error_flag() == 0


def test_fused_stem_from_uint8_pixels_is_bit_identical_to_the_float_path():
    """uint8 input: ToTensor + Normalize fused into the stem's band loader give exactly the features of the float
    path fed with torchvision-style preprocessed images (same fp32 operation order)."""
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (5, 3, 224, 224), generator=g, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    m = torch.tensor(mean).view(1, 3, 1, 1)
    s = torch.tensor(std).view(1, 3, 1, 1)
    flt = ((u8.float() / 255.0) - m) / s                  # ToTensor then Normalize
    w, b = rnd(64, 3, 7, 7, seed=2, scale=0.1), rnd(64, seed=3)
    wq = _pack_stem(w, torch.float16)
    o_f = torch.empty(5, 56, 56, 64, dtype=torch.float16, device=DEV)
    o_u = torch.empty_like(o_f)
    ops.stem_pool(flt.to(DEV), wq, b.to(DEV), o_f)
    ops.stem_pool_u8(u8.to(DEV), mean, std, wq, b.to(DEV), o_u)
    torch.cuda.synchronize()
    assert torch.equal(o_f, o_u)


@pytest.mark.parametrize('rows,V,K_', [(300, 4099, 1024), (70, 36541, 512), (5, 71, 64), (129, 1000, 256)])
def test_vocab_logprob_fused_epilogue(rows, V, K_):
    """log_softmax(A W^T + b)[m, target[m]] from the contraction's (max, sum exp) epilogue (K = 1024 runs on CTA pairs)
    against torch on the materialised logits of the same product."""
    g = torch.Generator().manual_seed(rows + V)
    A = (torch.randn(rows, K_, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(V, K_, generator=g) * 0.2).to(torch.bfloat16).to(DEV)
    bias = torch.randn(V, generator=g).to(DEV)
    tg = torch.randint(0, V, (rows,), generator=g).to(DEV)
    tg[0], tg[-1] = 0, V - 1
    out = torch.empty(rows, device=DEV)
    ops.vocab_logprob(A, W, bias, tg, out)
    logits = torch.empty(rows, (V + 3) // 4 * 4, device=DEV)
    ops.gemm(A, W, logits[:, :V], bias=bias)
    ref = torch.log_softmax(logits[:, :V].double(), dim=-1).gather(1, tg.view(-1, 1)).view(-1)
    assert float((out.double() - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('dt', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('n,hw,N,K', [(5, 49, 2048, 512), (1, 49, 2048, 512), (4, 49, 256, 64), (7, 49, 320, 192), (3, 64, 512, 128),
                                      (3, 100, 256, 64), (6, 16, 128, 64), (300, 49, 2048, 512), (2, 128, 256, 64)])
def test_gemm_pool_equals_gemm_then_avgpool(n, hw, N, K, dt):
    """dh_gemm_tc_pool (conv3 of the last bottleneck with the global average pool in its epilogue, encoders.py:60) against
    the two-launch form on the same operands: the stored map and the pooled means are identical bit for bit, and both
    match a float64 restatement."""
    M = n * hw
    A, W, b, r = rnd(M, K, seed=1).to(dt).to(DEV), rnd(N, K, seed=2, scale=0.1).to(dt).to(DEV), rnd(N, seed=3).to(DEV), rnd(M, N, seed=4).to(dt).to(DEV)
    out = torch.zeros(M, N, dtype=dt, device=DEV)
    pool = torch.full((n, N), 7.0, device=DEV)
    ops.gemm_pool(A, W, out, pool, hw, bias=b, residual=r, relu=True)
    out2 = torch.zeros(M, N, dtype=dt, device=DEV)
    pool2 = torch.zeros(n, N, device=DEV)
    ops.gemm(A, W, out2, bias=b, residual=r, relu=True)
    ops.avgpool(out2.view(n, hw, N), pool2)
    assert torch.equal(out, out2)
    assert torch.equal(pool, pool2)
    ref = F.relu(A.double() @ W.double().T + b.double() + r.double())
    assert H.rel_err(out.float(), ref.float()) < 4e-3
    assert H.rel_err(pool, ref.view(n, hw, N).mean(1).float()) < 2e-3


def test_encoder_fused_pool_equals_separate_avgpool():
    """EncoderRT with the pooled epilogue on and off: identical embeddings (the trunk's last launch is the only difference)."""
    from deephumor_b200 import models
    from deephumor_b200.utils import synth, synth_weights
    hp = synth_weights.default_hp('xfmr', 1000, small=True)
    sd = synth_weights.make_state_dict('xfmr', hp, seed=1)
    m = models.CaptioningTransformer(**hp)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.set_precision('bf16')
    imgs = synth.images(0, 0, 5).cuda()
    outs = []
    for flag in (True, False):
        ops.FUSED_POOL = flag
        try:
            with torch.no_grad():
                outs.append([t.clone() for t in m.encoder(imgs)])
        finally:
            ops.FUSED_POOL = True
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16])
@pytest.mark.parametrize('M,K', [(128, 512), (1, 512), (130, 64), (2560, 2048), (40960, 512), (777, 1024), (300, 128)])
def test_gemm_layernorm_epilogue(M, K, dt):
    """dh_gemm_tc_ln (fc_o / fc_2 + residual + nn.LayerNorm of a decoder sublayer, transformers.py:355-356,374-375) against a
    float64 restatement on the same rounded operands, out of place and in place on the residual; and against the two-launch
    form (dh_gemm_tc + dh_add_layernorm), which rounds the pre-norm sum to 2 bytes first."""
    N = 512
    A, W = rnd(M, K, seed=1).to(dt).to(DEV), rnd(N, K, seed=2, scale=0.1).to(dt).to(DEV)
    b, r = rnd(N, seed=3).to(DEV), rnd(M, N, seed=4).to(dt).to(DEV)
    g, be = (1.0 + 0.2 * rnd(N, seed=5)).to(DEV), rnd(N, seed=6, scale=0.3).to(DEV)
    ref = F.layer_norm(A.double() @ W.double().T + b.double() + r.double(), (N,), g.double(), be.double(), 1e-5).float()
    out = torch.zeros(M, N, dtype=dt, device=DEV)
    ops.gemm_ln(A, W, b, r, g, be, out)
    tol = 4e-3 if dt == torch.bfloat16 else 6e-4                  # one rounding of the output
    assert H.rel_err(out.float(), ref) < tol
    x = r.clone()
    ops.gemm_ln(A, W, b, x, g, be, x)                             # in place on the residual (how the decoder runs it)
    assert torch.equal(x, out)
    tmp = torch.zeros(M, N, dtype=dt, device=DEV)
    two = torch.zeros(M, N, dtype=dt, device=DEV)
    ops.gemm(A, W, tmp, bias=b, residual=r)
    ops.add_layernorm(tmp, None, g, be, two)
    assert H.rel_err(two.float(), ref) >= H.rel_err(out.float(), ref) * 0.5      # the fused form is no less accurate
    assert H.rel_err(out.float(), two.float()) < 3 * tol
    out2 = torch.zeros(M, N, dtype=dt, device=DEV)
    ops.gemm_ln(A, W, None, None, g, be, out2)                    # no bias / residual
    ref2 = F.layer_norm(A.double() @ W.double().T, (N,), g.double(), be.double(), 1e-5).float()
    assert H.rel_err(out2.float(), ref2) < tol


@pytest.mark.parametrize('flag', ['DH_TC_LN_UNSPLIT', 'DH_TC_LN_NO_PAIR'])
def test_gemm_layernorm_other_forms(flag):
    """The LayerNorm contraction has three forms: the row split over a cluster of two CTA pairs (default), over two plain
    CTAs (DH_TC_LN_NO_PAIR) and the whole row on one CTA / pair (DH_TC_LN_UNSPLIT).  The library reads the switches once, so the
    other two run the same checks in a child process."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_tc.py'), '-m', 'gpu', '-q', '-x',
                        '-k', 'test_gemm_layernorm_epilogue'], env=dict(os.environ, **{flag: '1'}), cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize('dt', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('n,H_,W_', [(2, 56, 56), (1, 32, 8), (3, 40, 24), (1, 56, 60), (5, 64, 16)])
def test_bottleneck_tail_fused_conv2_conv3(n, H_, W_, dt):
    """dh_bottleneck_tail_tc (conv2 3x3 + bn2 + relu -> conv3 1x1 + bn3 + identity + relu of a layer1 bottleneck, torchvision
    resnet.py:150-161, in one launch) against a float64 restatement on the same rounded operands (conv2's output rounded to
    the storage type, as the two-launch form stores it) and against the two launches themselves."""
    y1 = rnd(n, 64, H_, W_, seed=1).to(dt)
    w2, b2 = rnd(64, 64, 3, 3, seed=2, scale=0.06).to(dt), rnd(64, seed=3, scale=0.2)
    w3, b3 = rnd(256, 64, 1, 1, seed=4, scale=0.15).to(dt), rnd(256, seed=5, scale=0.2)
    x = rnd(n, 256, H_, W_, seed=6).to(dt)
    y2 = F.relu(F.conv2d(y1.double(), w2.double(), b2.double(), 1, 1)).to(dt)
    ref = F.relu(F.conv2d(y2.double(), w3.double(), b3.double()) + x.double()).float().permute(0, 2, 3, 1).contiguous()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV)
    y1d, xd = nhwc(y1), nhwc(x)
    w2d, w3d = w2.permute(0, 2, 3, 1).contiguous().to(DEV), w3.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.zeros(n, H_, W_, 256, dtype=dt, device=DEV)
    ops.bottleneck_tail(y1d, w2d, b2.to(DEV), w3d, b3.to(DEV), xd, out)
    tol = 4e-3 if dt == torch.bfloat16 else 6e-4
    assert H.rel_err(out.float(), ref) < tol
    y2d = torch.zeros(n, H_, W_, 64, dtype=dt, device=DEV)
    two = torch.zeros_like(out)
    ops.conv2d(y1d, w2d, b2.to(DEV), y2d, 1, 1, True)
    ops.conv2d(y2d, w3d, b3.to(DEV), two, 1, 0, True, residual=xd)
    assert H.rel_err(out.float(), two.float()) < tol
    assert float((out.float() - two.float()).abs().max()) <= 2 * float(ref.abs().max()) * (2 ** -8 if dt == torch.bfloat16 else 2 ** -11)
