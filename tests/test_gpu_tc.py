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
