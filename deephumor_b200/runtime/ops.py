"""Typed wrappers over the C ABI (include/deephumor_b200.h): torch tensors in, kernel launches out.

torch is used for device memory and streams only; every call below launches hand-written CUDA from
libdeephumor_sm100.so on torch's current stream.
"""
import os

import torch

from .._lib import (LIB, BeamState, Ctx, LstmOperands, Resnet50Weights, VocabSparse, XfmrBuffers, XfmrLayer, XfmrWeights,
                    ptr, stream)

F32, BF16, F16 = 0, 1, 2
NOISE = {'deterministic': 0, 'injected': 1}


class _Profile:
    """Optional CUDA-event ranges on torch's current stream (bench.py: per-stage time and the roofline kernel)."""

    def __init__(self):
        self.on, self.ev = False, []

    def start(self):
        self.on, self.ev = True, []

    class _Range:
        def __init__(self, prof, tag, flops, nbytes):
            self.prof, self.tag, self.flops, self.nbytes = prof, tag, flops, nbytes

        def __enter__(self):
            if self.prof.on:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *a):
            if self.prof.on:
                self.e1.record()
                self.prof.ev.append((self.tag, self.e0, self.e1, self.flops, self.nbytes))

    def range(self, tag, flops=0.0, nbytes=0.0):
        """flops: algorithmic FLOPs of a tensor-bound stage; nbytes: algorithmic HBM bytes of a bandwidth-bound one."""
        return self._Range(self, tag, flops, nbytes)

    def stop(self):
        """-> {tag: (count, total_ms, total_flops, total_bytes)}"""
        self.on = False
        torch.cuda.synchronize()
        out = {}
        for tag, e0, e1, fl, nb in self.ev:
            n, ms, f, b = out.get(tag, (0, 0.0, 0.0, 0.0))
            out[tag] = (n + 1, ms + e0.elapsed_time(e1), f + fl, b + nb)
        self.ev = []
        return out


PROFILE = _Profile()
TRACE = None    # tests: a list that receives (step, seq, val, ended, done) clones of the beam state after every decode step


def trace_beam(step, beam):
    if TRACE is not None:
        TRACE.append((step, beam.seq.clone(), beam.val.clone(), beam.ended.clone(), beam.done.clone()))
USE_GRAPHS = os.environ.get('DH_NO_GRAPH', '') == ''   # capture decode loops into CUDA graphs
HALO_CONV = os.environ.get('DH_NO_HALO_CONV', '') == ''       # 3x3 stride-1 convs of layer1 via the halo-tile kernel
HALO_COUT = 64
FUSED_STEM = os.environ.get('DH_NO_FUSED_STEM', '') == ''     # conv1 + ReLU + maxpool in one tcgen05 kernel
FUSED_LSTM = os.environ.get('DH_NO_FUSED_LSTM', '') == ''     # LSTM cell update in the gate GEMM's epilogue
LSTM_STACK = os.environ.get('DH_NO_LSTM_STACK', '') == ''     # all LSTM layers of a step in one persistent launch
LSTM_ROTATE = os.environ.get('DH_LSTM_ROTATE', '') != ''      # ... upper layers start on the recurrent half of K (measured: no gain)
FUSED_PREPARE = os.environ.get('DH_NO_FUSED_PREPARE', '') == ''   # next LSTM step's gathers in the select + beam launch
DUAL_CONV = os.environ.get('DH_NO_DUAL_CONV', '') == ''       # conv3 + downsample of a stage's first block as one contraction
PATH_ENTRIES = os.environ.get('DH_NO_PATH_ENTRIES', '') == ''   # trunk / decoder step through the path-level C entries
FUSED_LN = os.environ.get('DH_NO_FUSED_LN', '') == ''           # LayerNorm + residual in the fc_o / fc_2 contraction's epilogue
FUSED_POOL = os.environ.get('DH_NO_FUSED_POOL', '') == ''       # global average pool in the last conv3's epilogue
FUSED_VOCAB = os.environ.get('DH_NO_FUSED_VOCAB', '') == ''   # two-pass vocab projection, logits never stored


class DynWords:
    """{seed, image_base} device words read by a captured decode graph.  The host values go through a small ring of pinned
    buffers, each guarded by an event, so a second generate() issued before the first copy has executed cannot overwrite
    the words the first one is about to read."""

    def __init__(self, device, slots=4):
        self.dev = torch.zeros(2, dtype=torch.int64, device=device)
        self.host = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.next = 0

    def set(self, seed, image_base):
        i = self.next
        self.next = (i + 1) % len(self.host)
        if self.events[i] is not None:
            self.events[i].synchronize()
        self.host[i][0], self.host[i][1] = seed, image_base
        self.dev.copy_(self.host[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


class PlanCache:
    """Small LRU of decode plans (static buffers + captured CUDA graph per configuration): alternating between a few
    shapes must not re-capture a graph on every call, and a long-running server must not keep every shape it ever saw."""

    def __init__(self, capacity=4):
        from collections import OrderedDict
        self.capacity, self.items = capacity, OrderedDict()

    def get(self, key):
        pl = self.items.get(key)
        if pl is not None:
            self.items.move_to_end(key)
        return pl

    def put(self, key, plan):
        self.items[key] = plan
        while len(self.items) > self.capacity:
            self.items.popitem(last=False)

    def clear(self):
        self.items.clear()

    def __len__(self):
        return len(self.items)


def code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f'unsupported dtype {t.dtype}')


def torch_dtype(c):
    return {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}[c]


def _rows(t):
    assert t.dim() == 2 and t.stride(1) == 1, 'expect a row-major 2-D view'
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def synth_images(out, seed, first_index):
    n, _, size, _ = out.shape
    assert out.is_contiguous() and out.dtype == torch.float32
    LIB.call('dh_synth_images', ptr(out), seed, first_index, n, size, stream())


def nchw_to_nhwc4(images, out, halo=0):
    n, c, H, W = images.shape
    assert c == 3 and images.is_contiguous() and images.dtype == torch.float32 and out.is_contiguous()
    LIB.call('dh_nchw_to_nhwc4', ptr(images), ptr(out), n, H, W, halo, code(out), stream())


def conv2d(x, w, bias, y, stride, pad, relu, residual=None, tile_n=0):
    """x [n,H,W,Cin], w [Cout,kh,kw,Cin], y [n,Ho,Wo,Cout] (all contiguous NHWC)."""
    n, H, W, Cin = x.shape
    Cout, kh, kw, _ = w.shape
    assert x.is_contiguous() and w.is_contiguous() and y.is_contiguous() and w.shape[3] == Cin
    assert tuple(y.shape) == (n, (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1, Cout), 'conv2d output shape'
    if x.dtype == torch.float32:
        LIB.call('dh_conv2d_f32', ptr(x), ptr(w), ptr(bias), ptr(residual), ptr(y), n, H, W, Cin, Cout, kh, kw,
                 stride, pad, int(relu), stream())
    elif (HALO_CONV and kh == 3 and kw == 3 and stride == 1 and pad == 1 and residual is None and H >= 28
          and Cout == HALO_COUT and Cin % 64 == 0 and tile_n == 0):
        # layer1 conv2: halo tile in shared memory, nine taps as shifted descriptor views (input read once, weights
        # resident).  Measured: 159 -> 122 us per 256 images at 64 -> 64 channels; at 128 -> 128 (layer2) the streamed
        # weights dominate the operand traffic and the im2col kernel is as fast, so only Cout == 64 is routed here.
        assert x.dtype == w.dtype == y.dtype
        LIB.call('dh_conv3x3_halo_tc', ptr(x), ptr(w), ptr(bias), ptr(y), n, H, W, Cin, Cout, int(relu), code(x), stream())
    else:
        assert x.dtype == w.dtype == y.dtype and (residual is None or residual.dtype == x.dtype)
        LIB.call('dh_conv2d_tc', ptr(x), ptr(w), ptr(bias), ptr(residual), ptr(y), n, H, W, Cin, Cout, kh, kw,
                 stride, pad, int(relu), code(x), tile_n, stream())


def conv1x1_dual(x1, x2, w_cat, bias, y, stride2, relu, tile_n=0):
    """y = act(x1 * W1 + x2[::stride2, ::stride2] * W2 + bias) with w_cat = [W1 | W2] ([Cout, C1 + C2]): conv3 and the
    downsample branch of a stage's first bottleneck in one contraction (dh_conv1x1_dual_tc)."""
    n, Ho, Wo, C1 = x1.shape
    _, H2, W2, C2 = x2.shape
    Cout = w_cat.shape[0]
    assert x1.is_contiguous() and x2.is_contiguous() and y.is_contiguous() and w_cat.is_contiguous()
    assert w_cat.shape[1] == C1 + C2 and x1.dtype == x2.dtype == w_cat.dtype == y.dtype and y.shape == (n, Ho, Wo, Cout)
    LIB.call('dh_conv1x1_dual_tc', ptr(x1), ptr(x2), ptr(w_cat), ptr(bias), ptr(y), n, Ho, Wo, C1, H2, W2, C2, stride2, Cout,
             int(relu), code(x1), tile_n, stream())


def conv1x1_chain(y2, x2, x2_is_source, w, bias, out, w_next, bias_next, z, stride2=1):
    """out = relu(y2 * W[:, :C1] + (x2[::stride2, ::stride2] * W[:, C1:] if x2_is_source else x2) + bias) and
    z = relu(out * w_next + bias_next) in one launch (dh_conv1x1_chain_tc: a bottleneck's conv3 and the next bottleneck's
    conv1, the latter fed from L2)."""
    n, H, W, C1 = y2.shape
    _, H2, W2, C2 = x2.shape
    Cout, N2 = w.shape[0], w_next.shape[0]
    assert all(t.is_contiguous() for t in (y2, x2, w, out, w_next, z))
    assert out.shape == (n, H, W, Cout) and z.shape == (n, H, W, N2) and w_next.shape[1] == Cout
    assert w.shape[1] == (C1 + C2 if x2_is_source else C1)
    assert y2.dtype == x2.dtype == w.dtype == out.dtype == w_next.dtype == z.dtype
    LIB.call('dh_conv1x1_chain_tc', ptr(y2), ptr(x2), int(x2_is_source), ptr(w), ptr(bias), ptr(out), n, H, W, C1, C2, H2, W2,
             stride2, Cout, ptr(w_next), ptr(bias_next), ptr(z), N2, code(y2), stream())


def bottleneck_tail(y1, w2, b2, w3, b3, x, y):
    """y = relu(conv3(relu(conv2(y1) + b2)) + b3 + x): conv2 3x3 (64 -> 64) and conv3 1x1 (64 -> 256) of a layer1 identity
    bottleneck in one launch (dh_bottleneck_tail_tc)."""
    n, H, W, C = y1.shape
    assert C == 64 and w2.shape == (64, 3, 3, 64) and w3.shape[0] == 256 and w3.numel() == 256 * 64
    assert x.shape == (n, H, W, 256) and y.shape == (n, H, W, 256)
    assert all(t.is_contiguous() and t.dtype == y1.dtype for t in (y1, w2, w3, x, y))
    LIB.call('dh_bottleneck_tail_tc', ptr(y1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), ptr(x), ptr(y), n, H, W, code(y1), stream())


def im2col_stem(images, A, kh, kw, stride, pad):
    """images [n,3,H,W] fp32 NCHW -> A [n*Ho*Wo, Kp] bf16, k = (r*kw+s)*3 + c, zero padded."""
    n, c, H, W = images.shape
    assert c == 3 and images.is_contiguous() and images.dtype == torch.float32 and A.is_contiguous()
    assert A.shape[0] == n * ((H + 2 * pad - kh) // stride + 1) * ((W + 2 * pad - kw) // stride + 1), 'im2col rows'
    LIB.call('dh_im2col_stem', ptr(images), ptr(A), n, H, W, kh, kw, stride, pad, A.shape[1], code(A), stream())


def im2col_nhwc(x, A, kh, kw, stride, pad):
    n, H, W, C = x.shape
    assert x.is_contiguous() and A.is_contiguous() and x.dtype in (torch.bfloat16, torch.float16)
    LIB.call('dh_im2col_nhwc', ptr(x), ptr(A), n, H, W, C, kh, kw, stride, pad, stream())


def stem_pool(images, w_packed, bias, out):
    """images [n,3,224,224] fp32 NCHW -> out [n,56,56,64] NHWC (conv 7x7/2 + BN + ReLU + maxpool 3x3/2, fused)."""
    n, c, H, W = images.shape
    assert c == 3 and images.is_contiguous() and images.dtype == torch.float32 and out.is_contiguous()
    assert w_packed.shape == (64, 192) and w_packed.dtype == out.dtype and w_packed.is_contiguous()
    LIB.call('dh_stem_pool_tc', ptr(images), ptr(w_packed), ptr(bias), ptr(out), n, H, W, code(out), stream())


def stem_pool_u8(images, mean, std, w_packed, bias, out):
    """uint8 images [n,3,224,224] (0..255) -> out [n,56,56,64]: ToTensor + Normalize(mean, std) fused into the stem."""
    import ctypes
    n, c, H, W = images.shape
    assert c == 3 and images.is_contiguous() and images.dtype == torch.uint8 and out.is_contiguous()
    assert w_packed.shape == (64, 192) and w_packed.dtype == out.dtype and w_packed.is_contiguous()
    f3 = ctypes.c_float * 3
    LIB.call('dh_stem_pool_tc_u8', ptr(images), f3(*[float(v) for v in mean]), f3(*[float(v) for v in std]),
             ptr(w_packed), ptr(bias), ptr(out), n, H, W, code(out), stream())


def maxpool3x3s2(x, y):
    n, H, W, C = x.shape
    assert tuple(y.shape) == (n, (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1, C) and x.is_contiguous() and y.is_contiguous()
    LIB.call('dh_maxpool3x3s2', ptr(x), ptr(y), n, H, W, C, code(x), stream())


def avgpool(x, out):
    """x [n,HW,C] contiguous -> out [n,C] (row-major view)."""
    n, HW, C = x.shape
    assert x.is_contiguous()
    LIB.call('dh_avgpool', ptr(x), ptr(out), _rows(out), n, HW, C, code(x), code(out), stream())


def embed_mean(table, ids, out):
    n, L = ids.shape
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    LIB.call('dh_embed_mean', ptr(table), _rows(table), ptr(ids), L, ptr(out), _rows(out), n, table.shape[1],
             code(table), code(out), stream())


def gemm(A, W, out, bias=None, residual=None, relu=False, tile_n=0):
    """out[M,N] = act(A[M,K] @ W[N,K]^T + bias + residual).  fp32 operands -> FFMA check kernel (fp32 out);
    bf16 / f16 operands -> tcgen05 kernel (out fp32, bf16 or f16)."""
    M, K = A.shape
    N = W.shape[0]
    assert W.shape[1] == K and out.shape[0] == M and out.shape[1] == N
    if M == 0:
        return
    if A.dtype == torch.float32:
        assert W.dtype == torch.float32 and out.dtype == torch.float32
        assert residual is None or residual.dtype == torch.float32
        LIB.call('dh_gemm_f32', ptr(A), _rows(A), ptr(W), _rows(W), ptr(bias), ptr(residual),
                 0 if residual is None else _rows(residual), ptr(out), _rows(out), M, N, K, int(relu), stream())
    else:
        assert W.dtype == A.dtype
        LIB.call('dh_gemm_tc', ptr(A), _rows(A), ptr(W), _rows(W), code(A), ptr(bias), ptr(residual),
                 0 if residual is None else _rows(residual), 0 if residual is None else code(residual),
                 ptr(out), _rows(out), code(out), M, N, K, int(relu), tile_n, stream())


def gemm_pool(A, W, out, pool, pool_hw, bias=None, residual=None, relu=False):
    """out = act(A @ W^T + bias + residual) and pool[i] = mean of out's rows [i * pool_hw, (i + 1) * pool_hw) (fp32), one
    launch (dh_gemm_tc_pool: conv3 of the last bottleneck with the global average pool in its epilogue)."""
    M, K = A.shape
    N = W.shape[0]
    assert W.shape == (N, K) and out.shape == (M, N) and A.dtype == W.dtype == out.dtype and M % pool_hw == 0
    assert pool.dtype == torch.float32 and pool.shape == (M // pool_hw, N) and pool.stride(1) == 1
    assert residual is None or (residual.dtype == A.dtype and residual.shape == (M, N))
    if M == 0:
        return
    LIB.call('dh_gemm_tc_pool', ptr(A), _rows(A), ptr(W), _rows(W), code(A), ptr(bias), ptr(residual),
             0 if residual is None else _rows(residual), ptr(out), _rows(out), M, N, K, int(relu), pool_hw, ptr(pool),
             _rows(pool), stream())


def gemm_ln_supported(x, N):
    return FUSED_LN and x.dtype in (torch.bfloat16, torch.float16) and N == 512


def gemm_ln(A, W, bias, residual, gamma, beta, out, eps=1e-5):
    """out = LayerNorm(A @ W^T + bias + residual) * gamma + beta for N == 512 in one launch (dh_gemm_tc_ln); out may be the
    residual buffer itself."""
    M, K = A.shape
    N = W.shape[0]
    assert W.shape == (N, K) and out.shape == (M, N) and A.dtype == W.dtype == out.dtype and N == 512
    assert residual is None or (residual.dtype == A.dtype and residual.shape == (M, N))
    assert gamma.dtype == beta.dtype == torch.float32 and gamma.numel() == beta.numel() == N
    if M == 0:
        return
    LIB.call('dh_gemm_tc_ln', ptr(A), _rows(A), ptr(W), _rows(W), code(A), ptr(bias), ptr(residual),
             0 if residual is None else _rows(residual), ptr(gamma), ptr(beta), float(eps), ptr(out), _rows(out), M, N, K,
             stream())


def gemm_split3(A, W, bias, outs):
    """[outs[0] | outs[1] | outs[2]] = A @ W^T + bias with W [3n, K]: one tcgen05 contraction, three row-major destinations
    (dh_gemm_tc_split3; the fused Q | K | V projection of a transformer decode step)."""
    M, K = A.shape
    n = W.shape[0] // 3
    assert W.shape == (3 * n, K) and W.dtype == A.dtype and len(outs) == 3
    assert all(o.shape == (M, n) and o.dtype == outs[0].dtype and o.stride(1) == 1 for o in outs)
    if M == 0:
        return
    LIB.call('dh_gemm_tc_split3', ptr(A), _rows(A), ptr(W), _rows(W), code(A), ptr(bias), ptr(outs[0]), _rows(outs[0]),
             ptr(outs[1]), _rows(outs[1]), ptr(outs[2]), _rows(outs[2]), code(outs[0]), n, M, K, stream())


def gather_rows(src, idx, dst, width=None):
    """dst[r, :width] = src[idx[r] (or r), :width] with dtype conversion."""
    rows = dst.shape[0]
    width = dst.shape[1] if width is None else width
    assert idx is None or idx.dtype == torch.int32
    LIB.call('dh_gather_rows', ptr(src), _rows(src), src.shape[0], ptr(idx), ptr(dst), _rows(dst), rows, width,
             code(src), code(dst), stream())


def lstm_cell(gates, c_prev, parent, c_out, h_out0, h_out1):
    rows, H4 = gates.shape
    H = H4 // 4
    assert gates.dtype == torch.float32 and c_out.is_contiguous() and (c_prev is None or c_prev.is_contiguous())
    h = h_out0 if h_out0 is not None else h_out1
    LIB.call('dh_lstm_cell', ptr(gates), _rows(gates), ptr(c_prev), ptr(parent), ptr(c_out),
             ptr(h_out0), 0 if h_out0 is None else _rows(h_out0), ptr(h_out1), 0 if h_out1 is None else _rows(h_out1),
             rows, H, code(h), stream())


def lstm_layer_tc(A, Wp, bias_p, c_prev, parent, c_out, h_out0, h_out1):
    """One LSTM layer step: gate contraction on tcgen05 with the cell update in the epilogue (Wp / bias_p gate-packed
    per 64 hidden units, see pack_lstm_gates)."""
    rows, K = A.shape
    H = Wp.shape[0] // 4
    assert Wp.shape[1] == K and Wp.dtype == A.dtype and c_out.is_contiguous() and (c_prev is None or c_prev.is_contiguous())
    assert all(h is None or h.dtype == A.dtype for h in (h_out0, h_out1))
    LIB.call('dh_lstm_layer_tc', ptr(A), _rows(A), ptr(Wp), _rows(Wp), code(A), ptr(bias_p), ptr(c_prev), ptr(parent),
             ptr(c_out), ptr(h_out0), 0 if h_out0 is None else _rows(h_out0), ptr(h_out1),
             0 if h_out1 is None else _rows(h_out1), rows, H, K, stream())


def tc_error_flag():
    """Watchdog code left by a tcgen05 kernel before it trapped (0 = none; dh_tc_error_flag)."""
    import ctypes
    out = ctypes.c_int(0)
    LIB.call('dh_tc_error_flag', ctypes.byref(out))
    return out.value


def lstm_stack_tc(A_all, in_dims, Wp_all, bias_all, c_prev, parent, c_out, h_top, hs, ready, rows, rotate=None):
    """Every layer of one LSTM time step in one persistent launch (dh_lstm_stack_tc): A_all [L, rows_alloc, Kmax] holds
    layer l's [x | h_prev] in columns [0, in_l + H), Wp_all [L * 4H, Kmax] / bias_all [L * 4H] are gate-packed per layer
    (zero past in_l + H), c_prev / c_out [L, rows_alloc, H] fp32, hs [L, rows_alloc, H]; `ready` is a zeroed int32 region
    of (L - 1) * ceil(rows / 128) counters that the launch consumes."""
    import ctypes
    L, ra, Kmax = A_all.shape
    H = Wp_all.shape[0] // (4 * L)
    assert A_all.is_contiguous() and Wp_all.is_contiguous() and Wp_all.dtype == A_all.dtype and len(in_dims) == L
    assert c_out.is_contiguous() and c_out.shape[0] == L and (c_prev is None or c_prev.shape == c_out.shape)
    assert hs is None or (hs.is_contiguous() and hs.shape[0] == L and hs.dtype == A_all.dtype)
    assert ready.dtype == torch.int32 and ready.numel() >= (L - 1) * ((rows + 127) // 128)
    LIB.call('dh_lstm_stack_tc', ptr(A_all), Kmax, ra, (ctypes.c_int * L)(*in_dims), ptr(Wp_all), Wp_all.shape[1],
             code(A_all), ptr(bias_all), ptr(c_prev), ptr(parent), ptr(c_out), c_out.stride(0), ptr(h_top), _rows(h_top),
             ptr(hs), 0 if hs is None else hs.stride(0), ptr(ready), rows, H, L,
             int(LSTM_ROTATE if rotate is None else rotate), stream())


def lstm_operands(table, hs, A, in_off):
    """struct dh_lstm_operands for dh_select_beam_step_lstm: the gathers of lstm_prepare, described once per plan."""
    L, H, E = len(hs), hs[0].shape[-1], table.shape[1]
    assert L <= 8 and all(h.is_contiguous() and h.dtype == table.dtype and h.element_size() == 2 for h in hs)
    o = LstmOperands()
    o.table, o.ldt, o.n_tok_rows, o.E, o.H, o.L = ptr(table), _rows(table), table.shape[0], E, H, L
    for l in range(L):
        o.hs[l], o.A[l], o.lda[l], o.in_off[l] = ptr(hs[l]), ptr(A[l]), _rows(A[l]), in_off[l]
    return o


def lstm_prepare(table, tok, parent, hs, A, in_off, rows):
    """One launch for every operand of an LSTM step: A[0][:, :E] = table[tok], A[l][:, in_off[l]:+H] = hs[l][parent]."""
    import ctypes
    L, H, E = len(hs), hs[0].shape[-1], table.shape[1]
    assert all(h.is_contiguous() and h.dtype == table.dtype and h.element_size() == 2 for h in hs)
    arr_p = ctypes.c_void_p * L
    LIB.call('dh_lstm_prepare', ptr(table), _rows(table), table.shape[0], ptr(tok), E, ptr(parent),
             arr_p(*[ptr(h) for h in hs]), arr_p(*[ptr(a) for a in A]), (ctypes.c_longlong * L)(*[_rows(a) for a in A]),
             (ctypes.c_int * L)(*in_off), L, H, rows, stream())


def pack_lstm_gates(t, H):
    """[4H, ...] in nn.LSTM gate order (i, f, g, o) -> rows grouped per 64 hidden units as [i | f | g | o] blocks."""
    rest = t.shape[1:]
    return t.view(4, H // 64, 64, *rest).transpose(0, 1).reshape(4 * H, *rest).contiguous()


def add_layernorm(x, y, gamma, beta, out):
    rows, D = x.shape
    LIB.call('dh_add_layernorm', ptr(x), _rows(x), ptr(y), 0 if y is None else _rows(y), ptr(gamma), ptr(beta),
             ptr(out), _rows(out), rows, D, code(x), stream())


def xfmr_embed(tok_table, pos_table, start, rows_per_start, tokens, positions, pos_const, scale, out):
    rows, D = out.shape
    assert start.dtype == torch.float32 and tok_table.dtype == pos_table.dtype == out.dtype
    LIB.call('dh_xfmr_embed', ptr(tok_table), ptr(pos_table), _rows(tok_table), ptr(start), _rows(start),
             rows_per_start, ptr(tokens), ptr(positions), pos_const, float(scale), ptr(out), _rows(out), rows, D,
             code(out), stream())


def cast(src, dst):
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    LIB.call('dh_cast', ptr(src), ptr(dst), src.numel(), code(src), code(dst), stream())


def attention(q, K, V, out, n_heads, rows_per_image, slots, S_alloc, scale, src=None, slot_shared=False, n_keys=0,
              causal_full=False, seq=None, seq_per_image=False, pad=0, enc_mask=None):
    rows, D = q.shape
    LIB.call('dh_attention', ptr(q), _rows(q), ptr(K), ptr(V), ptr(out), _rows(out), rows, D, n_heads, rows_per_image,
             slots, S_alloc, ptr(src), int(slot_shared), n_keys, int(causal_full), ptr(seq),
             0 if seq is None else _rows(seq), int(seq_per_image), pad, ptr(enc_mask), float(scale), code(q), stream())


def enc_mask(spatial, mask):
    rows, D = spatial.shape
    LIB.call('dh_enc_mask', ptr(spatial), ptr(mask), rows, D, code(spatial), stream())


def select_tokens(logits, V, beam, top_k, temperature, unk, rows_per_image, noise_mode, seed, image_base, step, done,
                  ind, val, status, dyn=None):
    rows = logits.shape[0]
    assert logits.dtype == torch.float32
    LIB.call('dh_select_tokens', ptr(logits), _rows(logits), rows, V, beam, top_k, float(temperature), unk,
             rows_per_image, noise_mode, seed, image_base, step, ptr(done), ptr(ind), ptr(val), ptr(status), ptr(dyn),
             stream())


def sampled_rank(top_k, frac, eps=1e-7):
    """Rank j for a SAMPLED pass 1 that sees a fraction `frac` of the vocabulary: the j-th largest sampled value lies above
    the top_k-th largest logit only if at least j of the row's top_k - 1 largest logits fell into the sample, which -- for a
    sample unrelated to the logits -- has probability P(Bin(top_k - 1, frac) >= j).  Returns the smallest j that pushes it
    below eps (the rare miss is repaired on the device by the *_fix launches, so eps only bounds their cost)."""
    import math
    n = top_k - 1
    for j in range(1, top_k + 1):
        tail = sum(math.comb(n, i) * frac ** i * (1.0 - frac) ** (n - i) for i in range(j, n + 1))
        if tail < eps:
            return j
    return top_k


class VocabSelect:
    """Vocab projection fused with token selection for `rows` rows, dense logits never stored (dh_vocab_* in the header).

    Pass 2 stores only the 32-column groups that hold a logit >= the row's threshold (sparse materialisation: a hit map plus
    their 128-byte lines in a dense-pitch buffer) and the selection kernels rank exactly what they find there, so they are
    exact for any per-row threshold that leaves at least top_k candidates.  Default: a SAMPLED pass 1 over every stride-th
    256-column tile (1/stride of a full contraction) whose rank is chosen so that the threshold is low enough except with
    probability < 1e-7 per row, one full contraction for pass 2, and three fix-up launches that return at once unless a row
    stored too few / too many groups -- then they redo that step's threshold exhaustively, inside the same CUDA graph.
    stride 1 = the exhaustive two-pass form (two full contractions, ~top_k + 1 candidates per row)."""

    GROUP_CAP = 400          # stored groups per row the selection kernel accepts (its list holds 480); more -> fix-up

    def __init__(self, rows, V, top_k, device, stride=None):
        self.rows, self.V, self.top_k = rows, V, top_k
        bn = self.bn = 64 if V <= 64 else 128 if V <= 128 else 256
        self.n_blocks = n_blocks = (V + bn - 1) // bn
        if stride is None:
            stride = int(os.environ.get('DH_VOCAB_STRIDE', '8'))
        # a sampled pass needs enough sampled groups for the rank statistics (rank <= 64, a few times fewer than the groups)
        while stride > 1 and (n_blocks // stride) * (bn // 32) < max(64, 2 * top_k):
            stride //= 2
        self.stride = max(1, stride)
        self.n_groups_full = n_blocks * (bn // 32)
        if self.stride > 1:
            frac = ((n_blocks + self.stride - 1) // self.stride) / n_blocks      # largest sampled fraction over the offsets
            self.rank = min(sampled_rank(top_k, frac), 64)
        else:
            self.rank = top_k
        f32, i32 = dict(dtype=torch.float32, device=device), dict(dtype=torch.int32, device=device)
        self.gmax = torch.empty(rows, self.n_groups_full, **f32)
        self.thresh = torch.empty(rows, **f32)
        self.count = torch.zeros(rows, **i32)
        self.sp_ld = n_blocks * bn
        self.sp_logits = torch.empty(rows, self.sp_ld, **f32)               # only the stored groups are ever written / read
        self.hit_ld = 2 * n_blocks
        self.hitmap = torch.zeros(rows, self.hit_ld, dtype=torch.uint8, device=device)
        self.redo = torch.zeros(rows, dtype=torch.uint8, device=device) if self.stride > 1 else None
        self.flag = torch.zeros(1, **i32) if self.stride > 1 else None
        self.c = VocabSparse(ptr(self.thresh), ptr(self.hitmap), self.hit_ld, ptr(self.sp_logits), self.sp_ld, V)

    def groups(self, offset):
        return (self.n_blocks - offset + self.stride - 1) // self.stride * (self.bn // 32)

    @staticmethod
    def supported(A, V, top_k):
        # the threshold kernel ranks at most 64 values exactly and the warp-level selection keeps 160 survivors: larger
        # top_k takes the materialised-logits path (dh_select_tokens), which has no such limits
        return A.dtype in (torch.bfloat16, torch.float16) and top_k <= 64 and top_k <= (V + 31) // 32

    def candidates(self, rows):
        """Number of logits >= thresh per row (host-side diagnostic; walks the hit map)."""
        hm = self.hitmap[:rows].cpu()
        th = self.thresh[:rows].cpu()
        lg = self.sp_logits[:rows].cpu().view(rows, -1, 32)
        cpb = self.bn // 64
        out = torch.zeros(rows, dtype=torch.int64)
        for r in range(rows):
            nz = torch.nonzero(hm[r]).flatten()
            for bi in nz.tolist():
                byte = int(hm[r, bi])
                for bit in range(cpb):
                    if byte >> bit & 1:
                        out[r] += int((lg[r, bi * cpb + bit] >= th[r]).sum())
        return out

    def run(self, A, W, bias, beam, temperature, unk, rows_per_image, noise_mode, step, done, ind, val, status, dyn,
            seed=0, image_base=0, beam_step=None, lstm_next=None):
        """beam_step = (Beam, max_len, eos, lstm_semantics): also run that image's beam step in the same launch;
        lstm_next (ops.lstm_operands): ... and gather the next LSTM step's operands for the image's rows."""
        import ctypes
        rows, K = A.shape
        assert rows <= self.rows and W.shape == (self.V, K) and W.dtype == A.dtype
        args = (ptr(A), _rows(A), ptr(W), _rows(W), code(A), ptr(bias), rows, self.V, K)
        offset = step % self.stride                       # the sampled tiles rotate from step to step
        ng = self.groups(offset)
        lists = (ptr(self.thresh), ptr(self.count), ptr(self.sp_logits), self.sp_ld, ptr(self.hitmap), self.hit_ld)
        with PROFILE.range('vocab_pass1', 2.0 * rows * min(self.V, ng * 32) * K):
            LIB.call('dh_vocab_groupmax', *args, self.stride, offset, ptr(self.gmax), self.n_groups_full, stream())
        with PROFILE.range('vocab_threshold', nbytes=4.0 * rows * ng):
            LIB.call('dh_vocab_threshold', ptr(self.gmax), self.n_groups_full, rows, ng, self.rank,
                     ptr(self.thresh), ptr(self.count), stream())
        with PROFILE.range('vocab_gemm', 2.0 * rows * self.V * K):          # one launch = one full [rows,V,K] product
            LIB.call('dh_vocab_candidates', *args, *lists, stream())
        if self.stride > 1:
            # stored groups per row: at least top_k of them guarantee top_k candidates; the selection kernel lists <= 480
            lo, hi = min(self.top_k, self.n_groups_full), self.GROUP_CAP
            with PROFILE.range('vocab_fixup'):
                LIB.call('dh_vocab_groupmax_fix', *args, ptr(self.gmax), self.n_groups_full, ptr(self.count), lo, hi,
                         ptr(self.flag), stream())
                LIB.call('dh_vocab_threshold_fix', ptr(self.gmax), self.n_groups_full, rows, self.n_groups_full, self.top_k,
                         ptr(self.thresh), ptr(self.count), lo, hi, ptr(self.redo), ptr(self.flag), stream())
                LIB.call('dh_vocab_candidates_fix', *args, *lists, ptr(self.redo), ptr(self.flag), stream())
        with PROFILE.range('select_beam'):
            if beam_step is None:
                LIB.call('dh_select_candidates', ctypes.byref(self.c), rows, beam,
                         self.top_k, float(temperature), unk, rows_per_image, noise_mode, seed, image_base, step,
                         ptr(done), ptr(ind), ptr(val), ptr(status), ptr(dyn), stream())
            else:
                bm, max_len, eos, lstm_sem = beam_step
                assert rows == bm.n_img * bm.beam and rows_per_image == bm.beam == beam
                common = (ctypes.byref(self.c), ctypes.byref(bm.c), ptr(ind), ptr(val),
                          ptr(status), bm.n_img, beam, self.top_k, float(temperature), unk, step, max_len, eos,
                          int(lstm_sem), noise_mode, seed, image_base, ptr(dyn))
                if lstm_next is None:
                    LIB.call('dh_select_beam_step', *common, stream())
                else:
                    LIB.call('dh_select_beam_step_lstm', *common, ctypes.byref(lstm_next), stream())


class Beam:
    """Device-side beam state for n_img images x beam rows (struct dh_beam_state)."""

    def __init__(self, n_img, beam, seq_ld, device, kv_slots=0):
        i32 = dict(dtype=torch.int32, device=device)
        self.n_img, self.beam = n_img, beam
        self.seq = torch.zeros(n_img * beam, seq_ld, **i32)
        self.val = torch.zeros(n_img * beam, dtype=torch.float32, device=device)
        self.ended = torch.zeros(n_img * beam, dtype=torch.uint8, device=device)
        self.done = torch.zeros(n_img, dtype=torch.uint8, device=device)
        self.final_len = torch.zeros(n_img, **i32)
        self.last_tok = torch.zeros(n_img * beam, **i32)
        self.parent_state = torch.zeros(n_img * beam, **i32)
        self.src = torch.zeros(n_img * beam, kv_slots, **i32) if kv_slots else None
        self.status = torch.zeros(1, **i32)
        self.c = BeamState(ptr(self.seq), seq_ld, ptr(self.val), ptr(self.ended), ptr(self.done), ptr(self.final_len),
                           ptr(self.last_tok), ptr(self.parent_state), ptr(self.src), kv_slots)

    def init(self, ind0, val0, prefix, eos, lstm_semantics):
        import ctypes
        plen = 0 if prefix is None else prefix.shape[1]
        LIB.call('dh_beam_init', ctypes.byref(self.c), ptr(ind0), ptr(val0), ptr(prefix),
                 0 if prefix is None else prefix.stride(0), 1 if prefix is None else prefix.shape[0], plen,
                 self.n_img, self.beam, eos, int(lstm_semantics), stream())

    def step(self, new_ind, new_val, step, max_len, eos, lstm_semantics, temperature, noise_mode, seed, image_base,
             dyn=None):
        import ctypes
        LIB.call('dh_beam_step', ctypes.byref(self.c), ptr(new_ind), ptr(new_val), self.n_img, self.beam, step, max_len,
                 eos, int(lstm_semantics), float(temperature), noise_mode, seed, image_base, ptr(dyn), stream())

    def final(self, temperature, noise_mode, seed, image_base, final_step, len_if_running, pad, max_len, out_ids, out_len,
              dyn=None):
        import ctypes
        LIB.call('dh_beam_final', ctypes.byref(self.c), self.n_img, self.beam, float(temperature), noise_mode, seed,
                 image_base, final_step, len_if_running, pad, max_len, ptr(out_ids), ptr(out_len), ptr(dyn), stream())


def vocab_logprob(A, W, bias, targets, out):
    """out[m] = log_softmax(A W^T + bias)[m, targets[m]] on tensor cores, logits never stored (dh_vocab_logprob)."""
    rows, K = A.shape
    V = W.shape[0]
    assert W.shape[1] == K and W.dtype == A.dtype and targets.dtype == torch.int64 and targets.is_contiguous()
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == rows
    bn = 64 if V <= 64 else 128 if V <= 128 else 256
    n_groups = (V + bn - 1) // bn * (bn // 32)
    gmax = torch.empty(rows, n_groups, dtype=torch.float32, device=A.device)
    gsum = torch.empty_like(gmax)
    tl = torch.empty(rows, dtype=torch.float32, device=A.device)
    LIB.call('dh_vocab_logprob', ptr(A), _rows(A), ptr(W), _rows(W), code(A), ptr(bias), rows, V, K, ptr(targets),
             ptr(gmax), ptr(gsum), n_groups, ptr(tl), ptr(out), stream())


def token_logprob(logits, targets, out):
    rows, V = logits.shape
    assert logits.dtype == torch.float32 and targets.dtype == torch.int64
    LIB.call('dh_token_logprob', ptr(logits), _rows(logits), rows, V, ptr(targets), ptr(out), stream())


def resize_images(images, out_size=224):
    """torchvision `Resize((out_size, out_size))` of PIL RGB images on the device (dh_resize_bilinear_u8, bit-exact with
    Pillow): images = list of uint8 HWC arrays / tensors of different sizes -> uint8 [n,3,out_size,out_size] NCHW on the
    current device (the input format of the fused uint8 stem).  One pinned staging copy, three launches."""
    import ctypes
    import numpy as np
    n = len(images)
    dev = torch.device('cuda', torch.cuda.current_device())
    out = torch.empty(n, 3, out_size, out_size, dtype=torch.uint8, device=dev)
    if n == 0:
        return out
    def as_tensor(im):
        if torch.is_tensor(im):
            return im.contiguous()
        a = np.ascontiguousarray(im)
        return torch.from_numpy(a if a.flags.writeable else a.copy())      # np.asarray(PIL image) is read-only
    ts = [as_tensor(im) for im in images]
    for t in ts:
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError('resize_images expects uint8 images of shape [H, W, 3] (RGB, as PIL.Image.convert("RGB") gives)')
    hs, ws = [int(t.shape[0]) for t in ts], [int(t.shape[1]) for t in ts]
    offs, total = [], 0
    for h, w in zip(hs, ws):
        offs.append(total)
        total += (h * w * 3 + 255) // 256 * 256
    if all(t.is_cuda for t in ts):
        packed = torch.empty(total, dtype=torch.uint8, device=dev)
        for t, o in zip(ts, offs):
            packed[o:o + t.numel()].copy_(t.reshape(-1))
    else:
        host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        for t, o in zip(ts, offs):
            host[o:o + t.numel()].copy_(t.reshape(-1))
        packed = host.to(dev, non_blocking=True)
    ia = (ctypes.c_int * n)
    nbytes = ctypes.c_longlong(0)
    LIB.call('dh_resize_workspace_bytes', n, ia(*hs), ia(*ws), out_size, ctypes.byref(nbytes))
    ws_buf = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    LIB.call('dh_resize_bilinear_u8', ptr(packed), (ctypes.c_longlong * n)(*offs), ia(*hs), ia(*ws), n, out_size, ptr(out),
             ptr(ws_buf), nbytes.value, stream())
    return out


# ------------------------------------------------------------------------------------------------ path-level entries
DH_STAGE_RESNET50 = 1


def resnet50_ctx(enc):
    """dh_ctx holding the packed trunk of an EncoderRT (dh_ctx_set_resnet50)."""
    import ctypes
    ctx = Ctx()
    w = Resnet50Weights()
    w.stem_w, w.stem_b = ptr(enc.stem_wq), ptr(enc.stem.bias)
    stage = 0
    for i, blk in enumerate(enc.blocks):
        for j, name in enumerate(('c1', 'c2', 'c3')):
            w.conv_w[i][j], w.conv_b[i][j] = ptr(blk[name].w), ptr(blk[name].bias)
        if 'dual_w' in blk:
            w.dual_w[stage], w.dual_b[stage] = ptr(blk['dual_w']), ptr(blk['dual_b'])
            stage += 1
    for c in range(3):
        w.mean[c], w.std[c] = enc.pixel_mean[c], enc.pixel_std[c]
    w.dtype = code(enc.stem_wq)
    LIB.call('dh_ctx_set_resnet50', ctx.handle, ctypes.byref(w), launches=0)
    return ctx


def resnet50_forward(ctx, images, feat, pooled, workspace_of):
    """images fp32 / uint8 [n,3,224,224] -> feat [n,7,7,2048] (+ pooled fp32 [n,2048]) through dh_resnet50_forward;
    workspace_of(nbytes) returns a uint8 device buffer of at least that size."""
    import ctypes
    n, _, H, W = images.shape
    assert images.is_contiguous() and feat.is_contiguous() and (pooled is None or (pooled.is_contiguous() and pooled.dtype == torch.float32))
    nbytes = ctypes.c_longlong(0)
    LIB.call('dh_workspace_bytes', ctx.handle, DH_STAGE_RESNET50, n, H, W, ctypes.byref(nbytes), launches=0)
    ws = workspace_of(nbytes.value)
    LIB.call('dh_resnet50_forward', ctx.handle, ptr(images), int(images.dtype == torch.uint8), n, H, W, ptr(feat), ptr(pooled),
             ptr(ws), nbytes.value, stream(), launches=49)


def xfmr_ctx(rt):
    """dh_ctx holding the packed decoder stack of an XfmrDecoderRT (dh_ctx_set_xfmr)."""
    import ctypes
    ctx = Ctx()
    w = XfmrWeights()
    w.n_layers, w.D, w.n_heads, w.pf, w.cross, w.dtype, w.pad = rt.L, rt.D, rt.n_heads, rt.pf, int(rt.cross), code(rt.tok), rt.pad
    w.scale = rt.scale
    w.tok, w.pos, w.ld_tok = ptr(rt.tok), ptr(rt.pos), _rows(rt.tok)
    for l, lay in enumerate(rt.layers):
        y = w.layer[l]
        y.qkv_w, y.qkv_b = ptr(lay['self_attn.qkv.w']), ptr(lay['self_attn.qkv.b'])
        y.so_w, y.so_b = ptr(lay['self_attn.fc_o.w']), ptr(lay['self_attn.fc_o.b'])
        y.sln_g, y.sln_b, y.s_scale = ptr(lay['self_attn_ln.g']), ptr(lay['self_attn_ln.b']), lay['self_attn.scale']
        if rt.cross:
            y.cq_w, y.cq_b = ptr(lay['enc_attn.fc_q.w']), ptr(lay['enc_attn.fc_q.b'])
            y.co_w, y.co_b = ptr(lay['enc_attn.fc_o.w']), ptr(lay['enc_attn.fc_o.b'])
            y.cln_g, y.cln_b, y.c_scale = ptr(lay['enc_attn_ln.g']), ptr(lay['enc_attn_ln.b']), lay['enc_attn.scale']
        y.f1_w, y.f1_b, y.f2_w, y.f2_b = ptr(lay['pf.fc_1.w']), ptr(lay['pf.fc_1.b']), ptr(lay['pf.fc_2.w']), ptr(lay['pf.fc_2.b'])
        y.fln_g, y.fln_b = ptr(lay['pf_ln.g']), ptr(lay['pf_ln.b'])
    LIB.call('dh_ctx_set_xfmr', ctx.handle, ctypes.byref(w), launches=0)
    return ctx


def xfmr_buffers(x, qb, attn, tmp, h1, Kc, Vc, xkv, enc_mask, start, slots, S):
    b = XfmrBuffers()
    b.x, b.qb, b.attn, b.tmp, b.h1 = ptr(x), ptr(qb), ptr(attn), ptr(tmp), ptr(h1)
    for l in range(len(Kc)):
        b.Kc[l], b.Vc[l] = ptr(Kc[l]), ptr(Vc[l])
        if xkv is not None:
            b.xK[l], b.xV[l] = ptr(xkv[l][0]), ptr(xkv[l][1])
    b.enc_mask = ptr(enc_mask)
    b.start, b.ld_start = ptr(start), _rows(start)
    b.slots, b.S = slots, S
    return b


def xfmr_step(ctx, bufs, rows, rpi, pos, tokens, seq, src, n_layers, cross):
    """One decoder-stack step through dh_xfmr_step (embed + per layer 5 / 8 launches)."""
    import ctypes
    bufs.seq, bufs.seq_ld = ptr(seq), (0 if seq is None else _rows(seq))
    bufs.src = ptr(src)
    LIB.call('dh_xfmr_step', ctx.handle, ctypes.byref(bufs), rows, rpi, pos, ptr(tokens), stream(),
             launches=1 + n_layers * (8 if cross else 5))
