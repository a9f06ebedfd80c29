"""Op-level parity of the CUDA kernels (called through the C ABI) against torch-CPU fp32 restatements."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from deephumor_b200.runtime import ops
from deephumor_b200.utils import synth
from oracle import model as omodel, noise as onoise
from tests import helpers as H

DEV = 'cuda'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def test_synth_images_bit_identical():
    out = torch.empty(3, 3, 32, 32, device=DEV)
    ops.synth_images(out, 11, 5)
    assert torch.equal(out.cpu(), synth.images(11, 5, 3, size=32))


@pytest.mark.parametrize('M,N,K', [(1, 7, 4), (5, 36541, 64), (130, 129, 68), (300, 512, 1024), (64, 2048, 512)])
def test_gemm_f32(M, N, K):
    A, W, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), out, bias=b.to(DEV), residual=r.to(DEV), relu=True)
    ref = F.relu(A.double() @ W.double().T + b.double() + r.double()).float()
    assert H.rel_err(out, ref) < 1e-6


@pytest.mark.parametrize('n,H_,Cin,Cout,k,s,p', [(2, 20, 4, 64, 7, 2, 3), (3, 14, 64, 64, 3, 1, 1), (2, 14, 128, 128, 3, 2, 1),
                                                (2, 9, 256, 512, 1, 2, 0), (5, 7, 512, 2048, 1, 1, 0)])
def test_conv2d_f32(n, H_, Cin, Cout, k, s, p):
    x, w, b = rnd(n, Cin, H_, H_, seed=1), rnd(Cout, Cin, k, k, seed=2, scale=0.1), rnd(Cout, seed=3)
    ref = F.conv2d(x.double(), w.double(), b.double(), s, p)
    res = rnd(*ref.shape, seed=4)
    ref = F.relu(ref + res.double()).float().permute(0, 2, 3, 1).contiguous()
    y = torch.empty(ref.shape, device=DEV)
    ops.conv2d(x.permute(0, 2, 3, 1).contiguous().to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV), y, s, p,
               True, residual=res.permute(0, 2, 3, 1).contiguous().to(DEV))
    assert H.rel_err(y, ref) < 1e-6


def test_layout_pool_kernels():
    img = rnd(2, 3, 12, 12, seed=1)
    out = torch.empty(2, 12, 12, 4, device=DEV)
    ops.nchw_to_nhwc4(img.to(DEV), out)
    assert torch.equal(out.cpu()[..., :3], img.permute(0, 2, 3, 1)) and float(out[..., 3].abs().max()) == 0
    x = rnd(2, 8, 11, 11, seed=2)
    y = torch.empty(2, 6, 6, 8, device=DEV)
    ops.maxpool3x3s2(x.permute(0, 2, 3, 1).contiguous().to(DEV), y)
    assert torch.equal(y.cpu(), F.max_pool2d(x, 3, 2, 1).permute(0, 2, 3, 1))
    p = torch.empty(2, 8, device=DEV)
    ops.avgpool(x.permute(0, 2, 3, 1).reshape(2, 121, 8).contiguous().to(DEV), p)
    assert H.rel_err(p, x.mean(dim=(2, 3))) < 1e-6


@pytest.mark.parametrize('rows,D_,dt,use_y', [(1000, 512, torch.bfloat16, True), (37, 256, torch.float16, True),
                                             (4097, 1024, torch.bfloat16, False)])
def test_add_layernorm_vectorised_two_byte(rows, D_, dt, use_y):
    """Single-read 16-byte add + LayerNorm kernel for 2-byte activations (transformers.py:360,368,375,627,634) on strided
    rows, against torch's layer_norm on the same rounded inputs."""
    x = rnd(rows, D_ + 8, seed=1).to(dt).to(DEV)[:, :D_]
    y = rnd(rows, D_, seed=2).to(dt).to(DEV) if use_y else None
    g, b = rnd(D_, seed=3).to(DEV), rnd(D_, seed=4).to(DEV)
    out = torch.empty(rows, D_ + 8, dtype=dt, device=DEV)[:, :D_]
    ops.add_layernorm(x, y, g, b, out)
    ref = F.layer_norm(x.float() + (y.float() if use_y else 0.), (D_,), g, b, 1e-5)
    assert H.rel_err(out.float(), ref) < (4e-3 if dt == torch.bfloat16 else 6e-4)


def test_avgpool_vectorised_matches_scalar_order():
    """avgpool over 49 positions of fp16 NHWC features (encoders.py:60): 16-byte loads, per channel the ascending-row
    fp32 summation of the scalar kernel."""
    x = rnd(5, 49, 2048, seed=11).to(torch.float16).to(DEV)
    out = torch.empty(5, 2048, device=DEV)
    ops.avgpool(x, out)
    ref = torch.zeros(5, 2048, device=DEV)
    for j in range(49):
        ref += x[:, j].float()
    assert torch.equal(out, ref / torch.full_like(ref, 49.0))     # tensor / tensor: IEEE division (scalar division multiplies by 1/49)


def test_layernorm_lstm_cell_embed():
    x, y, g, b = rnd(37, 96, seed=1), rnd(37, 96, seed=2), rnd(96, seed=3), rnd(96, seed=4)
    out = torch.empty(37, 96, device=DEV)
    ops.add_layernorm(x.to(DEV), y.to(DEV), g.to(DEV), b.to(DEV), out)
    assert H.rel_err(out, F.layer_norm(x + y, (96,), g, b, 1e-5)) < 1e-5
    R, Hh = 21, 40
    gates, c = rnd(R, 4 * Hh, seed=5, scale=3), rnd(30, Hh, seed=6)
    parent = torch.randint(0, 30, (R,), generator=torch.Generator().manual_seed(0), dtype=torch.int32)
    c2, h2 = torch.empty(R, Hh, device=DEV), torch.empty(R, Hh, device=DEV)
    ops.lstm_cell(gates.to(DEV), c.to(DEV), parent.to(DEV), c2, h2, None)
    i, f, gg, o = gates.chunk(4, -1)
    cr = torch.sigmoid(f) * c[parent.long()] + torch.sigmoid(i) * torch.tanh(gg)
    assert H.rel_err(c2, cr) < 1e-5 and H.rel_err(h2, torch.sigmoid(o) * torch.tanh(cr)) < 1e-5
    table, ids = rnd(50, 16, seed=7), torch.randint(0, 50, (6, 4), generator=torch.Generator().manual_seed(1))
    e = torch.empty(6, 16, device=DEV)
    ops.embed_mean(table.to(DEV), ids.to(DEV), e)
    assert H.rel_err(e, table[ids].mean(1)) < 1e-6


@pytest.mark.parametrize('mode', ['deterministic', 'injected'])
@pytest.mark.parametrize('V,B,K,T', [(1000, 5, 50, 1.0), (36541, 5, 50, 1.0), (71, 3, 7, 0.8), (500, 1, 1, 1.3), (300, 4, 4, 1.0)])
def test_select_tokens_matches_oracle(mode, V, B, K, T):
    rows, rpi = 6, 3
    logits = rnd(rows, V, seed=V, scale=3.0)
    logits[1, 1] = 10.0                                   # <unk> is the arg-max of row 1 (Q3: masked but counted)
    nz = onoise.Noise(mode, 9)
    ind = torch.empty(rows, B, dtype=torch.int32, device=DEV)
    val = torch.empty(rows, B, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.select_tokens(logits.to(DEV), V, B, K, T, 1, rpi, ops.NOISE[mode], 9, 40, 7, None, ind, val, status)
    for r in range(rows):
        if K == 1 and int(logits[r].argmax()) == 1:
            continue                                      # reference raises; covered by the status test below
        q = None if mode == 'deterministic' else torch.stack([onoise.exp_noise(9, 40 + r // rpi, 7, 0, r % rpi, V)])
        oi, ov = omodel.select_tokens(logits[r:r + 1], B, T, K, 1, q, None)
        assert ind[r].cpu().tolist() == oi[0].tolist(), (r, ind[r].cpu().tolist(), oi[0].tolist())
        assert torch.allclose(val[r].cpu(), ov[0], atol=1e-5)
    if K == 1:
        assert int(status.item()) & 1                    # row 1 is entirely filtered -> EMPTY_ROW flag


@pytest.mark.parametrize('rpi,nk,causal,use_enc,use_pad', [(5, 49, False, True, False), (1, 49, False, True, False),
                                                        (32, 0, True, False, True), (49, 49, False, True, False),
                                                        (3, 7, False, False, False), (17, 0, True, False, True)])
def test_attention_shared_keys_tensor_core_path(rpi, nk, causal, use_enc, use_pad):
    """dh_attention with keys shared by an image's rows (bf16, head_dim 64): the mma.sync path against a float64 torch
    restatement of MultiHeadAttentionLayer.forward's core (transformers.py:100-121) with the -1e8 masks."""
    n_img, n_heads, D_ = 6, 8, 512
    S = rpi if causal else nk
    rows = n_img * rpi
    g = torch.Generator().manual_seed(rpi * 100 + S)
    q = (torch.randn(rows, D_, generator=g)).to(torch.bfloat16)
    K = (torch.randn(n_img, S, D_, generator=g)).to(torch.bfloat16)
    Vv = (torch.randn(n_img, S, D_, generator=g)).to(torch.bfloat16)
    enc = (torch.rand(n_img, S, generator=g) < 0.2).to(torch.uint8) if use_enc else None
    seq = torch.randint(0, 4, (n_img, S), generator=g).to(torch.int32) if use_pad else None     # 0 == pad
    out = torch.empty(rows, D_, dtype=torch.bfloat16, device=DEV)
    scale = 8.0
    ops.attention(q.to(DEV), K.to(DEV), Vv.to(DEV), out, n_heads, rpi, 1, S, scale, slot_shared=True,
                  n_keys=0 if causal else nk, causal_full=causal, seq=None if seq is None else seq.to(DEV),
                  seq_per_image=True, pad=0, enc_mask=None if enc is None else enc.to(DEV))
    qd = q.double().view(n_img, rpi, n_heads, 64).permute(0, 2, 1, 3)
    kd = K.double().view(n_img, S, n_heads, 64).permute(0, 2, 1, 3)
    vd = Vv.double().view(n_img, S, n_heads, 64).permute(0, 2, 1, 3)
    e = qd @ kd.transpose(-1, -2) / scale                                  # [img, head, rpi, S]
    if enc is not None:
        e = e.masked_fill(enc.bool().view(n_img, 1, 1, S), -1e8)
    if seq is not None:
        padk = torch.zeros(n_img, S, dtype=torch.bool)
        padk[:, 1:] = seq[:, :S - 1] == 0                                  # key t >= 1 masked iff token t-1 is pad
        e = e.masked_fill(padk.view(n_img, 1, 1, S), -1e8)
    if causal:
        e = e.masked_fill(torch.triu(torch.ones(rpi, S, dtype=torch.bool), 1), float('-inf'))
    ref = (torch.softmax(e, -1) @ vd).permute(0, 2, 1, 3).reshape(rows, D_)
    assert H.rel_err(out.float(), ref.float()) < 1e-2


@pytest.mark.parametrize('n_img,rpi,nk,n_heads,S_alloc,slots', [(700, 5, 49, 8, 49, 1), (1000, 1, 49, 8, 49, 1),
                                                               (333, 5, 20, 4, 24, 2), (150, 33, 49, 8, 49, 1)])
def test_attention_streaming_cross_kv_ring(n_img, rpi, nk, n_heads, S_alloc, slots):
    """Persistent bulk-copy cross-attention (attn_stream_kernel): more images than CTAs, so every CTA wraps its stage
    ring several times; strided K/V allocations (S_alloc > n_keys, slots > 1: slot 0 is read); float64 torch reference of
    transformers.py:100-121 with the any-zero encoder mask (-1e8)."""
    D_ = n_heads * 64
    rows = n_img * rpi
    g = torch.Generator().manual_seed(n_img + rpi)
    q = torch.randn(rows, D_, generator=g).to(torch.bfloat16)
    K = torch.randn(n_img * slots, S_alloc, D_, generator=g).to(torch.bfloat16)
    Vv = torch.randn(n_img * slots, S_alloc, D_, generator=g).to(torch.bfloat16)
    enc = (torch.rand(n_img, S_alloc, generator=g) < 0.2).to(torch.uint8)
    out = torch.empty(rows, D_, dtype=torch.bfloat16, device=DEV)
    scale = 8.0
    qd_, Kd_, Vd_, ed_ = q.to(DEV), K.to(DEV), Vv.to(DEV), enc.to(DEV)
    ops.attention(qd_, Kd_, Vd_, out, n_heads, rpi, slots, S_alloc, scale, slot_shared=True, n_keys=nk, enc_mask=ed_)
    out2 = torch.empty_like(out)
    ops.attention(qd_, Kd_, Vd_, out2, n_heads, rpi, slots, S_alloc, scale, slot_shared=True, n_keys=nk, enc_mask=ed_)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)                       # the stage ring hands every image over completely: run-to-run stable
    Ks = K.view(n_img, slots, S_alloc, D_)[:, 0, :nk]
    Vs = Vv.view(n_img, slots, S_alloc, D_)[:, 0, :nk]
    qd = q.double().view(n_img, rpi, n_heads, 64).permute(0, 2, 1, 3)
    kd = Ks.double().reshape(n_img, nk, n_heads, 64).permute(0, 2, 1, 3)
    vd = Vs.double().reshape(n_img, nk, n_heads, 64).permute(0, 2, 1, 3)
    e = (qd @ kd.transpose(-1, -2) / scale).masked_fill(enc[:, :nk].bool().view(n_img, 1, 1, nk), -1e8)
    ref = (torch.softmax(e, -1) @ vd).permute(0, 2, 1, 3).reshape(rows, D_)
    assert H.rel_err(out.float(), ref.float()) < 1e-2


@pytest.mark.parametrize('n_img,B,nk,S_alloc,use_src', [(37, 5, 18, 33, True), (64, 1, 7, 33, False), (9, 10, 33, 40, True),
                                                       (3, 5, 1, 33, True), (5, 4, 130, 160, True)])
def test_attention_incremental_self_row_kernel(n_img, B, nk, S_alloc, use_src):
    """Incremental bf16 self-attention over the KV cache (attn_row_kernel: one warp per row, all 8 heads, keys through
    the per-(beam, position) slot table) against a float64 torch restatement of transformers.py:100-121 with the pad-key
    mask (-1e8) of transformers.py:473-477."""
    D_, n_heads = 512, 8
    rows = n_img * B
    g = torch.Generator().manual_seed(n_img * 7 + nk)
    q = torch.randn(rows, D_, generator=g).to(torch.bfloat16)
    K = torch.randn(n_img * B, S_alloc, D_, generator=g).to(torch.bfloat16)
    Vv = torch.randn(n_img * B, S_alloc, D_, generator=g).to(torch.bfloat16)
    src = torch.randint(0, B, (n_img, B, S_alloc), generator=g).to(torch.int32) if use_src else None
    seq = torch.randint(0, 3, (rows, S_alloc), generator=g).to(torch.int32)                 # 0 == pad
    out = torch.empty(rows, D_, dtype=torch.bfloat16, device=DEV)
    scale = 8.0
    ops.attention(q.to(DEV), K.to(DEV), Vv.to(DEV), out, n_heads, B, B, S_alloc, scale,
                  src=None if src is None else src.to(DEV), slot_shared=(B == 1), n_keys=nk, seq=seq.to(DEV),
                  seq_per_image=False, pad=0)
    torch.cuda.synchronize()
    Kp, Vp = K.view(n_img, B, S_alloc, D_), Vv.view(n_img, B, S_alloc, D_)
    t = torch.arange(nk)
    ref = torch.empty(rows, D_, dtype=torch.float64)
    for r in range(rows):
        i, b = divmod(r, B)
        slot = src[i, b, :nk].long() if src is not None else torch.full((nk,), 0 if B == 1 else b)
        kr = Kp[i, slot, t].double().view(nk, n_heads, 64).permute(1, 0, 2)              # [head, key, 64]
        vr = Vp[i, slot, t].double().view(nk, n_heads, 64).permute(1, 0, 2)
        e = (q[r].double().view(n_heads, 1, 64) @ kr.transpose(-1, -2)).squeeze(1) / scale  # [head, key]
        padk = torch.zeros(nk, dtype=torch.bool)
        padk[1:] = seq[r, :nk - 1] == 0
        e = e.masked_fill(padk.view(1, nk), -1e8)
        ref[r] = (torch.softmax(e, -1).unsqueeze(1) @ vr).reshape(D_)
    assert H.rel_err(out.float(), ref.float()) < 1e-2



@pytest.mark.parametrize('dt', [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize('rows,D,rps,pos', [(10, 512, 5, 0), (10, 512, 5, 7), (37, 64, 1, 3), (2560, 512, 5, 31), (6, 40, 2, 0)])
def test_xfmr_embed_vectorised_equals_formula(rows, D, rps, pos, dt):
    """(start | tok_embedding[token]) / scale + pos_embedding[pos] (transformers.py:455-470): the 16-byte vectorised kernel of
    the 2-byte tables and the scalar kernel give exactly the rounded fp32 expression."""
    g = torch.Generator().manual_seed(rows + D + pos)
    V = 97
    tok = torch.randn(V, D, generator=g).to(dt).cuda()
    pe = torch.randn(40, D, generator=g).to(dt).cuda()
    start = torch.randn((rows + rps - 1) // rps, D, generator=g).cuda()
    tokens = torch.randint(0, V, (rows,), generator=g, dtype=torch.int32).cuda()
    out = torch.empty(rows, D, dtype=dt, device='cuda')
    scale = 22.627416997969522
    ops.xfmr_embed(tok, pe, start, rps, tokens, None, pos, scale, out)
    e = start.repeat_interleave(rps, 0)[:rows] if pos == 0 else tok[tokens.long()].float()
    ref = (e / torch.tensor(scale, dtype=torch.float32).cuda() + pe[pos].float()).to(dt)
    assert torch.equal(out, ref)
