// Path-level entries of the C ABI: a whole stage of the caption path behind ONE call, so that a consumer that is not the
// Python runtime (deephumor_b200/runtime) can run the path without re-implementing its launch order (SURVEY.md 8(b):
// dh_ctx / dh_workspace_bytes / dh_resnet50_forward / dh_xfmr_step).  A dh_ctx is a host-side table of DEVICE pointers to
// the packed weights plus the geometry derived from them; it owns no device memory and launches nothing when created.
//
//   dh_resnet50_forward   torchvision ResNet-50 trunk as ImageEncoder uses it (models/encoders.py:34-39,56: `children()[:-2]`,
//                         torchvision resnet.py:143-163 Bottleneck, :197-204 / :266-279 layer stack) + AdaptiveAvgPool2d
//                         (encoders.py:39,60): 1 stem launch + 16 x 3 convolutions, the last one with the pooled epilogue.
//   dh_xfmr_step          one new position through the whole decoder stack with the KV cache (models/transformers.py:343-377
//                         per layer, :455-486 around it): embed -> per layer [Q|K|V, self-attention, fc_o + LN, (Q, cross-
//                         attention, fc_o + LN), fc_1, fc_2 + LN].
#include <stdlib.h>

#include <new>

#include "common.cuh"

namespace {

constexpr int kBlocks[4] = {3, 4, 6, 3};


}  // namespace

struct dh_ctx {
  bool has_resnet = false;
  dh_resnet50_weights rn{};
  bool has_xfmr = false;
  dh_xfmr_weights xf{};
};

extern "C" int dh_ctx_create(dh_ctx** out) {
  DH_ARG(out);
  *out = new (std::nothrow) dh_ctx();
  if (!*out) return dh_fail(DH_ERR_DEVICE, "host allocation failed", __FILE__, __LINE__);
  return DH_OK;
}

extern "C" int dh_ctx_destroy(dh_ctx* ctx) {
  delete ctx;
  return DH_OK;
}

extern "C" int dh_ctx_set_resnet50(dh_ctx* ctx, const dh_resnet50_weights* w) {
  DH_ARG(ctx && w && (w->dtype == DH_F16 || w->dtype == DH_BF16) && w->stem_w && w->stem_b);
  for (int b = 0; b < 16; ++b) {
    for (int c = 0; c < 3; ++c) DH_ARG(w->conv_w[b][c] && w->conv_b[b][c] && ((uintptr_t)w->conv_w[b][c] % 16) == 0);
  }
  for (int s = 0; s < 4; ++s) {
    DH_ARG(w->dual_w[s] && w->dual_b[s] && ((uintptr_t)w->dual_w[s] % 16) == 0);
  }
  ctx->rn = *w;
  ctx->has_resnet = true;
  return DH_OK;
}

namespace {

// activation workspace of the trunk for n images of H x W: two ping-pong block buffers (largest: layer1 output,
// (H/4)(W/4) x 256), the conv1 output (largest: layer2 block 0, (H/4)(W/4) x 128) and the conv2 output ((H/4)(W/4) x 64)
struct TrunkLayout { long long x_elems, y1_elems, y2_elems, total_bytes; };

TrunkLayout trunk_layout(int n, int H, int W) {
  const long long hp = (((H + 6 - 7) / 2 + 1) - 1) / 2 + 1, wp = (((W + 6 - 7) / 2 + 1) - 1) / 2 + 1;   // after stem + maxpool
  TrunkLayout L;
  L.x_elems = (long long)n * hp * wp * 256;
  L.y1_elems = (long long)n * hp * wp * 128;
  L.y2_elems = (long long)n * hp * wp * 64;
  auto al = [](long long v) { return (v + 255) / 256 * 256; };
  L.total_bytes = 2 * al(L.x_elems * 2) + al(L.y1_elems * 2) + al(L.y2_elems * 2);
  return L;
}

}  // namespace

extern "C" int dh_workspace_bytes(const dh_ctx* ctx, int stage, int n, int H, int W, long long* bytes_out) {
  DH_ARG(ctx && bytes_out && n >= 0);
  if (stage == DH_STAGE_RESNET50) {
    DH_ARG(H >= 32 && W >= 32);
    *bytes_out = trunk_layout(n, H, W).total_bytes;
    return DH_OK;
  }
  return dh_fail(DH_ERR_ARG, "unknown stage", __FILE__, __LINE__);
}

extern "C" int dh_resnet50_forward(const dh_ctx* ctx, const void* images_nchw, int images_u8, int n, int H, int W, void* feat,
                                   float* pooled, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  DH_ARG(ctx && ctx->has_resnet && images_nchw && feat && n >= 0);
  if (H != 224 || W != 224)
    return dh_fail(DH_ERR_UNSUPPORTED, "dh_resnet50_forward: the fused stem covers 224 x 224 inputs (use the per-op entries otherwise)",
                   __FILE__, __LINE__);
  if (n == 0) return DH_OK;
  const dh_resnet50_weights& w = ctx->rn;
  const TrunkLayout L = trunk_layout(n, H, W);
  DH_ARG(workspace && ((uintptr_t)workspace % 256) == 0 && workspace_bytes >= L.total_bytes);
  auto al = [](long long v) { return (v + 255) / 256 * 256; };
  unsigned char* ws = (unsigned char*)workspace;
  void* xb[2] = {ws, ws + al(L.x_elems * 2)};
  void* y1 = ws + 2 * al(L.x_elems * 2);
  void* y2 = ws + 2 * al(L.x_elems * 2) + al(L.y1_elems * 2);
  const int dt = w.dtype;
  int rc;
  // stem: conv 7x7/2 + BN + ReLU + maxpool 3x3/2 in one launch (ToTensor + Normalize folded in for uint8 pixels)
  if (images_u8)
    rc = dh_stem_pool_tc_u8((const unsigned char*)images_nchw, w.mean, w.std, w.stem_w, w.stem_b, xb[0], n, H, W, dt, stream);
  else
    rc = dh_stem_pool_tc((const float*)images_nchw, w.stem_w, w.stem_b, xb[0], n, H, W, dt, stream);
  if (rc) return rc;
  // One bottleneck (torchvision resnet.py:143-163) for `cnt` images: x -> out, through the conv1 / conv2 scratch buffers
  // `c1_done`: conv1 of this block was already computed into y1 by the previous block's chained launch; `next_blk` >= 0: chain
  // conv1 of block next_blk (of `next_mid` channels) behind this block's conv3 (dh_conv1x1_chain_tc; layer1's blocks only)
  static const bool chain_ok = getenv("DH_NO_CHAIN") == nullptr;
  auto bottleneck = [&](int s, int b, int blk, const void* x, void* out, int cnt, int hw, int cin, float* pool, bool c1_done,
                        int next_blk, int next_mid) -> int {
    const int mid = 64 << s, cout = 4 * mid;
    const int stride = (b == 0 && s > 0) ? 2 : 1;
    const int ho = (hw + 2 - 3) / stride + 1;
    // conv1 1x1 + bn1 + relu (resnet.py:146-148)
    int r = c1_done ? DH_OK
                    : dh_conv2d_tc(x, w.conv_w[blk][0], w.conv_b[blk][0], nullptr, y1, cnt, hw, hw, cin, mid, 1, 1, 1, 0, 1, dt, 0, stream);
    if (r) return r;
    // layer1 identity blocks: conv2 -> conv3 + identity in one launch, conv2's output stays on chip (dh_bottleneck_tail_tc).
    // Opt-in: correct (tests/test_gpu_tc.py) but not faster yet -- 590 us against 171 + 295 us for the two launches at 512
    // images (profiles/r02_bench_bottleneck_tail.txt): with the identity added in the epilogue instead of on the tensor core,
    // eight epilogue warps still spend 25 k cycles per 256-pixel tile against 12.7 k of HBM time.
    static const bool fused_tail = getenv("DH_FUSED_TAIL") != nullptr;
    if (fused_tail && b > 0 && mid == 64 && stride == 1 && cin == 256 && !pool)
      return dh_bottleneck_tail_tc(y1, w.conv_w[blk][1], w.conv_b[blk][1], w.conv_w[blk][2], w.conv_b[blk][2], x, out, cnt, hw, hw, dt,
                                   stream);
    // conv2 3x3 (stride on conv2: ResNet v1.5, resnet.py:109-110) + bn2 + relu (:150-152)
    if (stride == 1 && hw >= 28 && mid == 64)
      r = dh_conv3x3_halo_tc(y1, w.conv_w[blk][1], w.conv_b[blk][1], y2, cnt, hw, hw, mid, mid, 1, dt, stream);
    else
      r = dh_conv2d_tc(y1, w.conv_w[blk][1], w.conv_b[blk][1], nullptr, y2, cnt, hw, hw, mid, mid, 3, 3, stride, 1, 1, dt, 0, stream);
    if (r) return r;
    // conv3 1x1 + bn3 (+ downsample branch of a stage's first block, :157-158) + identity + relu (:154-161)
    if (next_blk >= 0)   // ... and the next block's conv1 on the tile just stored, fed from L2 (y1 is free again: conv2 has run)
      return dh_conv1x1_chain_tc(y2, x, b == 0, b == 0 ? w.dual_w[s] : w.conv_w[blk][2], b == 0 ? w.dual_b[s] : w.conv_b[blk][2], out,
                                 cnt, ho, ho, mid, cin, hw, hw, stride, cout, w.conv_w[next_blk][0], w.conv_b[next_blk][0], y1, next_mid,
                                 dt, stream);
    if (b == 0)
      return dh_conv1x1_dual_tc(y2, x, w.dual_w[s], w.dual_b[s], out, cnt, ho, ho, mid, hw, hw, cin, stride, cout, 1, dt, 0, stream);
    if (pool)
      return dh_gemm_tc_pool(y2, mid, w.conv_w[blk][2], mid, dt, w.conv_b[blk][2], x, cout, out, cout, cnt * ho * ho, cout, mid, 1,
                             ho * ho, pool, cout, stream);
    return dh_conv2d_tc(y2, w.conv_w[blk][2], w.conv_b[blk][2], x, out, cnt, ho, ho, mid, cout, 1, 1, 1, 0, 1, dt, 0, stream);
  };
  int cur = 0, hw = 56, cin = 64, blk = 0;
  bool c1_done = false;
  static const int chain_layers = getenv("DH_CHAIN_LAYERS") ? atoi(getenv("DH_CHAIN_LAYERS")) : 1;
  static const bool fused_tail_on = getenv("DH_FUSED_TAIL") != nullptr;
  // layer1 works on 56 x 56 maps of 64 - 256 channels (1.6 MB per image and tensor) and every one of its ten convolutions
  // runs at the HBM roofline when the batch goes through one layer at a time.  Taking `l2_chunk` images through the WHOLE
  // stage before the next ones keeps the block-to-block activations inside the 126 MB L2 (write-back): only the stage's
  // input and output cross HBM.  The two block-output scratch maps of a chunk live in the unused three quarters of xb[cur]
  // (the stem output has 64 of the 256 channels the buffer is sized for).
  static const int l2_chunk = getenv("DH_TRUNK_L2_CHUNK") ? atoi(getenv("DH_TRUNK_L2_CHUNK")) : 0;
  if (l2_chunk > 0 && n > l2_chunk && 2ll * l2_chunk * 256 <= (long long)n * 192) {
    const long long px = 56ll * 56;
    unsigned char* in0 = (unsigned char*)xb[0];
    unsigned char* scratch = in0 + al((long long)n * px * 64 * 2);
    for (int i0 = 0; i0 < n; i0 += l2_chunk) {
      const int cnt = n - i0 < l2_chunk ? n - i0 : l2_chunk;
      void* sa = scratch;
      void* sb = scratch + al((long long)l2_chunk * px * 256 * 2);
      const void* x = in0 + (long long)i0 * px * 64 * 2;
      void* outp = (unsigned char*)xb[1] + (long long)i0 * px * 256 * 2;
      rc = bottleneck(0, 0, 0, x, sa, cnt, 56, 64, nullptr, false, -1, 0);
      if (!rc) rc = bottleneck(0, 1, 1, sa, sb, cnt, 56, 256, nullptr, false, -1, 0);
      if (!rc) rc = bottleneck(0, 2, 2, sb, outp, cnt, 56, 256, nullptr, false, -1, 0);
      if (rc) return rc;
    }
    cur = 1; cin = 256; blk = 3;
  }
  for (int s = (blk ? 1 : 0); s < 4; ++s) {
    const int mid = 64 << s, cout = 4 * mid;
    for (int b = 0; b < kBlocks[s]; ++b, ++blk) {
      const int stride = (b == 0 && s > 0) ? 2 : 1;
      const int ho = (hw + 2 - 3) / stride + 1;
      const bool last = (s == 3 && b == kBlocks[3] - 1);
      void* out = last ? feat : xb[cur ^ 1];
      // layer1 (Cout == 256): conv3 of a block carries conv1 of the next block -- also the first block of layer2, whose conv1
      // is 1x1 / stride 1 on the same grid (ResNet v1.5 strides conv2).  DH_CHAIN_LAYERS=2 chains layer2 (Cout == 512) as well:
      // correct and tested, but measured equal to the separate launches (1021 against 1017 us per 512 images for the stage's
      // four boundaries) -- with K = 128 + 256 residual columns and K2 = 512 a unit streams 640 KB of operands from L2 for
      // 320 KB of HBM traffic, so the chained launch sits on the L2 -> SM operand rate instead of on HBM.
      const bool chain = chain_ok && s < chain_layers && !fused_tail_on && blk + 1 < 16;
      const int next_mid = (b + 1 < kBlocks[s]) ? mid : 2 * mid;
      rc = bottleneck(s, b, blk, xb[cur], out, n, hw, cin, (last && pooled && ho * ho <= 128) ? pooled : nullptr, c1_done,
                      chain ? blk + 1 : -1, next_mid);
      if (rc) return rc;
      c1_done = chain;
      cur ^= 1;
      hw = ho;
      cin = cout;
    }
  }
  return DH_OK;
}

extern "C" int dh_ctx_set_xfmr(dh_ctx* ctx, const dh_xfmr_weights* w) {
  DH_ARG(ctx && w && w->n_layers >= 1 && w->n_layers <= DH_XFMR_MAX_LAYERS && w->D > 0 && w->D % 128 == 0 && w->n_heads > 0);
  DH_ARG(w->D % w->n_heads == 0 && w->pf > 0 && w->pf % 8 == 0 && (w->dtype == DH_BF16 || w->dtype == DH_F16));
  DH_ARG(w->tok && w->pos && w->ld_tok >= w->D);
  for (int l = 0; l < w->n_layers; ++l) {
    const dh_xfmr_layer& y = w->layer[l];
    DH_ARG(y.qkv_w && y.qkv_b && y.so_w && y.so_b && y.sln_g && y.sln_b && y.f1_w && y.f1_b && y.f2_w && y.f2_b && y.fln_g && y.fln_b);
    if (w->cross) DH_ARG(y.cq_w && y.cq_b && y.co_w && y.co_b && y.cln_g && y.cln_b);
  }
  ctx->xf = *w;
  ctx->has_xfmr = true;
  return DH_OK;
}

// One new position `pos` for `rows` rows (rows_per_image rows per image) through the whole decoder stack:
//   x = (start | tok_embedding[token]) / scale + pos_embedding[pos]                         transformers.py:455-470 / :706-722
//   per layer: [Q | K | V] = x W_qkv^T (K / V rows land in cache slot `pos`), incremental self-attention over the cache through
//   the beam slot table, x = LN(x + fc_o(attn))                                             :97-127, :349-356 / :618-627
//   [cross: Q = x W_q^T, attention over the image's 49 cached K / V rows, x = LN(x + fc_o)] :358-366
//   x = LN(x + fc_2(relu(fc_1(x))))                                                          :153-165, :368-375
// With D == 512 every "x = LN(x + ...)" is one dh_gemm_tc_ln launch; otherwise dh_gemm_tc + dh_add_layernorm through `tmp`.
extern "C" int dh_xfmr_step(const dh_ctx* ctx, const dh_xfmr_buffers* b, int rows, int rows_per_image, int pos,
                            const int* tokens, cudaStream_t stream) {
  DH_ARG(ctx && ctx->has_xfmr && b && rows >= 0 && rows_per_image >= 1 && pos >= 0 && pos < b->S);
  DH_ARG(b->x && b->qb && b->attn && b->h1 && b->start && b->slots >= 1 && b->S >= 1);
  if (rows == 0) return DH_OK;
  const dh_xfmr_weights& w = ctx->xf;
  const int D = w.D, dt = w.dtype;
  const bool fused_ln = D == 512;
  DH_ARG(fused_ln || b->tmp);
  int rc = dh_xfmr_embed(w.tok, w.pos, w.ld_tok, b->start, b->ld_start, rows_per_image, tokens, nullptr, pos, w.scale, b->x, D, rows,
                         D, dt, stream);
  if (rc) return rc;
  // prefix phase (one row per image) writes slot 0 of each image's `slots` cache slots; beam phase writes every slot
  const long long slot_stride = rows_per_image == 1 ? b->slots : 1;
  const long long ld_cache = (long long)b->S * D * slot_stride;
  auto post = [&](const void* W, const float* bias, const float* g, const float* be) -> int {
    if (fused_ln) return dh_gemm_tc_ln(b->attn, D, W, D, dt, bias, b->x, D, g, be, 1e-5f, b->x, D, rows, D, D, stream);
    int r = dh_gemm_tc(b->attn, D, W, D, dt, bias, b->x, D, dt, b->tmp, D, dt, rows, D, D, 0, 0, stream);
    if (r) return r;
    return dh_add_layernorm(b->tmp, D, nullptr, 0, g, be, b->x, D, rows, D, dt, stream);
  };
  for (int l = 0; l < w.n_layers; ++l) {
    const dh_xfmr_layer& y = w.layer[l];
    DH_ARG(b->Kc[l] && b->Vc[l]);
    uint16_t* kdst = reinterpret_cast<uint16_t*>(b->Kc[l]) + (long long)pos * D;
    uint16_t* vdst = reinterpret_cast<uint16_t*>(b->Vc[l]) + (long long)pos * D;
    rc = dh_gemm_tc_split3(b->x, D, y.qkv_w, D, dt, y.qkv_b, b->qb, D, kdst, ld_cache, vdst, ld_cache, dt, D, rows, D, stream);
    if (rc) return rc;
    rc = dh_attention(b->qb, D, b->Kc[l], b->Vc[l], b->attn, D, rows, D, w.n_heads, rows_per_image, b->slots, b->S, b->src,
                      rows_per_image == 1, pos + 1, 0, b->seq, b->seq_ld, 0, w.pad, nullptr, y.s_scale, dt, stream);
    if (rc) return rc;
    rc = post(y.so_w, y.so_b, y.sln_g, y.sln_b);
    if (rc) return rc;
    if (w.cross) {
      DH_ARG(b->xK[l] && b->xV[l] && b->enc_mask);
      rc = dh_gemm_tc(b->x, D, y.cq_w, D, dt, y.cq_b, nullptr, 0, 0, b->qb, D, dt, rows, D, D, 0, 0, stream);
      if (rc) return rc;
      rc = dh_attention(b->qb, D, b->xK[l], b->xV[l], b->attn, D, rows, D, w.n_heads, rows_per_image, 1, 49, nullptr, 1, 49, 0,
                        nullptr, 0, 0, 0, b->enc_mask, y.c_scale, dt, stream);
      if (rc) return rc;
      rc = post(y.co_w, y.co_b, y.cln_g, y.cln_b);
      if (rc) return rc;
    }
    rc = dh_gemm_tc(b->x, D, y.f1_w, D, dt, y.f1_b, nullptr, 0, 0, b->h1, w.pf, dt, rows, w.pf, D, 1, 0, stream);
    if (rc) return rc;
    if (fused_ln) {
      rc = dh_gemm_tc_ln(b->h1, w.pf, y.f2_w, w.pf, dt, y.f2_b, b->x, D, y.fln_g, y.fln_b, 1e-5f, b->x, D, rows, D, w.pf, stream);
    } else {
      rc = dh_gemm_tc(b->h1, w.pf, y.f2_w, w.pf, dt, y.f2_b, b->x, D, dt, b->tmp, D, dt, rows, D, w.pf, 0, 0, stream);
      if (!rc) rc = dh_add_layernorm(b->tmp, D, nullptr, 0, y.fln_g, y.fln_b, b->x, D, rows, D, dt, stream);
    }
    if (rc) return rc;
  }
  return DH_OK;
}
