/* deephumor_b200.h -- C ABI of libdeephumor_sm100.so (hand-written sm_100a CUDA for the DeepHumor
 * caption-generation path).
 *
 * The reference (ilya16/deephumor) has no FFI layer: its boundary for this path is the Python model-class
 * API.  The classes in deephumor_b200/models mirror that API and bind the entry
 * points below through ctypes (deephumor_b200/_lib.py); INTEGRATION.md shows the stub a maintainer of the
 * reference would add.  Each entry cites the reference call site(s) it replaces, as
 * /root/reference/deephumor/<file>:<line> (torchvision lines are site-packages/torchvision/models/resnet.py).
 *
 * Conventions: every pointer is a DEVICE pointer unless it is a struct passed by pointer (host); sizes are
 * explicit; every launch goes to the given cudaStream_t; nothing allocates or frees caller memory; return
 * value 0 = ok, negative = DH_ERR_*, positive = cudaError_t; dh_last_error() gives text for the calling
 * thread.  dtype arguments are DH_F32 (fp32 check mode) or DH_BF16 (tensor-core mode).  Matrices are
 * row-major with explicit leading dimensions in elements.
 */
#ifndef DEEPHUMOR_B200_H_
#define DEEPHUMOR_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define DH_VERSION 203

#define DH_OK 0
#define DH_ERR_ARG (-1)
#define DH_ERR_DEVICE (-2)
#define DH_ERR_UNSUPPORTED (-3)

#define DH_F32 0
#define DH_BF16 1
#define DH_F16 2

/* noise model of the stochastic decoder (models/beam.py:39-48 uses torch.multinomial == topk(p / Exp(1))) */
#define DH_NOISE_DETERMINISTIC 0 /* q == 1: classical top-k / beam with the reference's scoring */
#define DH_NOISE_INJECTED 1      /* counter-based Exp(1) noise keyed on (seed, image, step, call, row, col) */
#define DH_CALL_TOKEN 0ull
#define DH_CALL_PRUNE 1ull
#define DH_CALL_FINAL 2ull

/* bits OR-ed into the int status word by the decode kernels */
#define DH_STATUS_EMPTY_ROW 1     /* a logits row was entirely filtered: the reference raises RuntimeError (beam.py:32-46) */
#define DH_STATUS_TOO_MANY_TIES 2 /* more than 4096 values tie at the top-k threshold */

const char* dh_last_error(void);
int dh_version(void);
int dh_check_device(int device);

/* ------------------------------------------------------------------------------------------ synthetic inputs */
/* images [count,3,size,size] fp32 keyed by GLOBAL image index (bit-identical to deephumor_b200/utils/synth.py). */
int dh_synth_images(float* out_nchw, unsigned long long seed, long long first_index, int count, int size,
                    cudaStream_t stream);

/* ------------------------------------------------------------------------------------------ encoder
 * models/encoders.py:46-70 (ImageEncoder.forward), :96-106 (LabelEncoder), :129-144 (ImageLabelEncoder);
 * torchvision resnet.py:143-163,266-279 (Bottleneck / trunk).  BN is folded into the conv weights by the
 * host packer; activations are NHWC. */
int dh_nchw_to_nhwc4(const float* images_nchw, void* out_nhwc4, int n, int H, int W, int halo, int dtype,
                     cudaStream_t stream);
/* fp32 check mode: y = act(conv(x, w) + bias + residual); w is [Cout][kh][kw][Cin], Cin % 4 == 0. */
int dh_conv2d_f32(const float* x, const float* w, const float* bias, const float* residual, float* y, int n, int H,
                  int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int relu, cudaStream_t stream);
int dh_maxpool3x3s2(const void* x, void* y, int n, int H, int W, int C, int dtype, cudaStream_t stream);
int dh_avgpool(const void* x, void* out, long long ldo, int n, int HW, int C, int dtype, int out_dtype,
               cudaStream_t stream);
int dh_embed_mean(const void* table, long long ldt, const long long* ids, int L, void* out, long long ldo, int n, int E,
                  int dtype, int out_dtype, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------ contractions
 * C[M,N] = act(A[M,K] * W[N,K]^T + bias[N] + residual[M,N]).  Replaces every nn.Linear on the path
 * (encoders.py:61,67,142; rnn_models.py:81,109; transformers.py:97,127,162-163,488,736) and, with [x|h]
 * concatenated along K, the nn.LSTM gate products (rnn_models.py:80,108). */
int dh_gemm_f32(const float* A, long long lda, const float* W, long long ldw, const float* bias, const float* residual,
                long long ldr, float* C, long long ldc, int M, int N, int K, int relu, cudaStream_t stream);

/* Tensor-core mode (tcgen05 / TMEM / TMA, csrc/gemm_tc.cu): same contract, A and W both ab_dtype (DH_BF16 or
 * DH_F16), fp32 accumulate, residual and output DH_F32 / DH_BF16 / DH_F16.  tile_n = 0 lets the library pick the
 * N tile (64 / 128 / 256). */
int dh_gemm_tc(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
               const void* residual, long long ldr, int res_dtype, void* C, long long ldc, int out_dtype, int M, int N,
               int K, int relu, int tile_n, cudaStream_t stream);
/* dh_gemm_tc plus a fused global average pool: pool[i, :] (fp32) = mean of rows [i * pool_hw, (i + 1) * pool_hw) of the
 * stored C.  The last bottleneck's conv3 + bn3 + identity + ReLU (torchvision resnet.py:154-161) with the
 * AdaptiveAvgPool2d((1,1)) of encoders.py:39,60 in its epilogue -- the 7x7x2048 map is written once (for the spatial
 * tokens) and never re-read for the pooled vector.  M % pool_hw == 0, pool_hw <= 128, C / residual of ab_dtype. */
int dh_gemm_tc_pool(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                    const void* residual, long long ldr, void* C, long long ldc, int M, int N, int K, int relu,
                    int pool_hw, float* pool, long long ld_pool, cudaStream_t stream);
/* out = LayerNorm(A W^T + bias + residual) * gamma + beta, rows of N == 512 columns, in one launch: the post-LN tail of every
 * decoder sublayer (transformers.py:355-356,365-366,374-375: fc_o / fc_2 -> dropout (eval) -> residual add -> nn.LayerNorm).
 * A CTA (pair) keeps the whole 128-row x 512-column block in tensor memory, so the pre-norm sums never reach HBM.
 * residual / out of ab_dtype; out may alias residual; bias / gamma / beta fp32, 16-byte aligned. */
int dh_gemm_tc_ln(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                  const void* residual, long long ldr, const float* gamma, const float* beta, float eps, void* out,
                  long long ldc, int M, int N, int K, cudaStream_t stream);
/* One contraction, three destinations: block j (of split_n columns) of A[M,K] W[3*split_n,K]^T + bias goes to Cj with its own
 * leading dimension.  The transformer decode step projects Q, K and V of the new position from the same activation
 * (transformers.py:97-99): W = [fc_q | fc_k | fc_v] stacked along N, C0 = the query buffer, C1 / C2 = this position's rows of
 * the K / V caches, so the K / V rows land in the cache without a copy.  split_n % 128 == 0; Cj of out_dtype (DH_BF16 /
 * DH_F16), 16 B aligned rows. */
int dh_gemm_tc_split3(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias, void* C0,
                      long long ldc0, void* C1, long long ldc1, void* C2, long long ldc2, int out_dtype, int split_n, int M,
                      int K, cudaStream_t stream);
/* Implicit-GEMM convolution (torchvision resnet.py:143-163 conv+BN+ReLU(+identity)): x NHWC, Cin % 64 == 0,
 * w [Cout][kh][kw][Cin] with BN folded, residual / y NHWC, all of dtype (DH_BF16 or DH_F16).  The A operand is
 * fetched by im2col-mode TMA. */
int dh_conv2d_tc(const void* x, const void* w, const float* bias, const void* residual, void* y, int n, int H, int W,
                 int Cin, int Cout, int kh, int kw, int stride, int pad, int relu, int dtype, int tile_n,
                 cudaStream_t stream);
/* relu(conv3(y2) + downsample(x)) of a ResNet stage's first bottleneck (torchvision resnet.py:154-161) as ONE contraction
 * over K = C1 + C2: x1 = y2 [n,Ho,Wo,C1] (1x1, stride 1), x2 = the block input [n,H2,W2,C2] (1x1, stride2, so that
 * (H2 - 1) / stride2 + 1 == Ho), w_cat [Cout][C1 + C2] = [W3 | Wd] (BN folded), bias = b3 + bd.  The downsample output
 * is neither written nor re-read.  C1, C2 % 64 == 0. */
int dh_conv1x1_dual_tc(const void* x1, const void* x2, const void* w_cat, const float* bias, void* y, int n, int Ho, int Wo,
                       int C1, int H2, int W2, int C2, int stride2, int Cout, int relu, int dtype, int tile_n,
                       cudaStream_t stream);
/* conv3 of a bottleneck AND conv1 of the next one in ONE launch (torchvision resnet.py:154-161, then :146-148 of the following
 * block; the layer1 / layer2 shapes, Cout % 256 == 0):
 *   out [n,H,W,Cout] = relu(y2 [n,H,W,C1] * W[:, :C1]^T + (x2_is_source ? x2 [n,H2,W2,C2] (1x1, stride2) * W[:, C1:]^T
 *                                                                        : x2 [n,H,W,Cout]) + bias)
 *   z   [n,H,W,N2]   = relu(out * w_next [N2][Cout]^T + bias_next),  N2 = 64, 128 or 256
 * x2 is the block input: the downsample branch's source (first block of a stage, W = [W3 | Wd], bias = b3 + bd,
 * (H2 - 1) / stride2 + 1 == H) or the identity (H2 / W2 / stride2 ignored).  A CTA computes the 128 pixels of z two tiles after
 * it has stored the same 128 pixels of out, reading them back through TMA while they are still in L2, so the next block's
 * conv1 has no HBM read.  Bit-identical to dh_conv2d_tc / dh_conv1x1_dual_tc followed by dh_conv2d_tc. */
int dh_conv1x1_chain_tc(const void* y2, const void* x2, int x2_is_source, const void* w, const float* bias, void* out, int n,
                        int H, int W, int C1, int C2, int H2, int W2, int stride2, int Cout, const void* w_next,
                        const float* bias_next, void* z, int N2, int dtype, cudaStream_t stream);
/* 3x3 / stride 1 / pad 1 convolution (+ bias + ReLU) that loads each input pixel ONCE per tile (torchvision resnet.py:146-148,
 * conv2 of the layer1 / layer2 bottlenecks): an 8 x 16 output rectangle per tile, its 10 x 18 halo fetched by one tiled TMA
 * load per 64-channel chunk, the nine taps contracted as shifted UMMA-descriptor views of the same shared memory
 * (csrc/conv3x3_tc.cu).  x / w / y as in dh_conv2d_tc; Cin % 64 == 0, Cout % 64 == 0. */
int dh_conv3x3_halo_tc(const void* x, const void* w, const float* bias, void* y, int n, int H, int W, int Cin, int Cout,
                       int relu, int dtype, cudaStream_t stream);
/* Tail of a layer1-shaped identity bottleneck in ONE launch (torchvision resnet.py:150-161): y = relu(bn3(conv3(relu(bn2(conv2(
 * y1))))) + x), conv2 3x3 / 1 / 1 (64 -> 64), conv3 1x1 (64 -> 256); y1 [n,H,W,64], x / y [n,H,W,256] NHWC, w2 [64][3][3][64],
 * w3 [256][64] (BN folded).  conv2's output tile stays in shared memory as conv3's operand (never written to HBM). */
int dh_bottleneck_tail_tc(const void* y1, const void* w2, const float* bias2, const void* w3, const float* bias3, const void* x,
                          void* y, int n, int H, int W, int dtype, cudaStream_t stream);
/* Explicit gathers: the C_in = 3 stem straight from the NCHW fp32 image into A[n*Ho*Wo, k_padded] (bf16 / f16) with
 * k = (r*kw + s)*3 + c (encoders.py:56 -> resnet.py:197 conv1), and a generic NHWC gather A[m, (r*kw+s)*C + c]. */
int dh_im2col_stem(const float* images_nchw, void* A, int n, int H, int W, int kh, int kw, int stride, int pad,
                   int k_padded, int out_dtype, cudaStream_t stream);
int dh_im2col_nhwc(const void* x, void* A, int n, int H, int W, int C, int kh, int kw, int stride, int pad,
                   cudaStream_t stream);
/* Fused stem (encoders.py:56 -> torchvision resnet.py:197-200,268-271): conv 7x7/2 pad 3 (3 -> 64, BN folded) + ReLU +
 * maxpool 3x3/2 pad 1 from the NCHW fp32 image to NHWC [n,56,56,64] of dtype (DH_F16 / DH_BF16); the im2col operand is
 * built in shared memory and contracted on tcgen05, pooling happens on the accumulators.  H = W = 224 only.
 * w_packed [64][192] of dtype: element [o][r*22 + s*3 + c] = folded weight [o][c][r][s], zero elsewhere. */
int dh_stem_pool_tc(const float* images_nchw, const void* w_packed, const float* bias, void* out, int n, int H, int W,
                    int dtype, cudaStream_t stream);
/* The same from raw uint8 pixels [n,3,224,224] (NCHW, 0..255): (x / 255 - mean[c]) / std[c] -- torchvision's ToTensor +
 * Normalize (deephumor_demo.ipynb cell 11) in the exact fp32 operation order -- is applied while the input band is staged,
 * so the float image never exists in HBM.  mean3_host / std3_host are HOST arrays of 3 floats. */
int dh_stem_pool_tc_u8(const unsigned char* images_nchw_u8, const float* mean3_host, const float* std3_host,
                       const void* w_packed, const float* bias, void* out, int n, int H, int W, int dtype,
                       cudaStream_t stream);
/* torchvision `Resize((out_size, out_size))` on PIL RGB images (deephumor_demo.ipynb cell 11; data/datasets.py:48-53,94-98),
 * i.e. Pillow's two-pass antialiased BILINEAR resampling (22-bit fixed-point taps, uint8 intermediate), bit-exact.
 * n images of different sizes, HWC uint8, packed in one device buffer at byte offsets_host[i] (heights / widths / offsets
 * are HOST arrays); out uint8 [n,3,out_size,out_size] NCHW = the input of dh_stem_pool_tc_u8.  The caller provides the
 * workspace (256-byte aligned) of dh_resize_workspace_bytes; in / out ratios up to 79 per axis, out_size <= 256. */
int dh_resize_workspace_bytes(int n, const int* heights_host, const int* widths_host, int out_size, long long* bytes_out);
int dh_resize_bilinear_u8(const unsigned char* packed_hwc, const long long* offsets_host, const int* heights_host,
                          const int* widths_host, int n, int out_size, unsigned char* out_nchw, void* workspace,
                          long long workspace_bytes, cudaStream_t stream);
/* watchdog code left by gemm_tc_kernel before it traps (0 = none). */
int dh_tc_error_flag(int* out_host);

/* ------------------------------------------------------------------------------------------ row-wise kernels */
int dh_gather_rows(const void* src, long long lds, long long n_src_rows, const int* idx, void* dst, long long ldd,
                   int rows, int width, int src_dtype, int dst_dtype, cudaStream_t stream);
/* nn.LSTM cell update, gate order i,f,g,o (rnn_models.py:23-24,80,108); c_prev rows read through parent[]
 * which folds the beam reorder of rnn_models.py:135-137 into the load. */
int dh_lstm_cell(const float* gates, long long ldg, const float* c_prev, const int* parent, float* c_out, void* h_out0,
                 long long ldh0, void* h_out1, long long ldh1, int rows, int H, int dtype, cudaStream_t stream);
/* All operands of one LSTM time step in one launch (2-byte elements): A[0][r, 0:E] = table[tok[r]] (rnn_models.py:107) and
 * A[l][r, in_off[l] : +H] = hs[l][parent[r]] for l < L (the h regather of rnn_models.py:135-137).  hs / A / lda / in_off
 * are HOST arrays of L entries (device pointers inside). */
int dh_lstm_prepare(const void* table, long long ldt, long long n_tok_rows, const int* tok, int E, const int* parent,
                    const void* const* hs, void* const* A, const long long* lda, const int* in_off, int L, int H, int rows,
                    cudaStream_t stream);
/* Tensor-core nn.LSTM layer step with the cell update fused into the contraction's epilogue (rnn_models.py:80,108):
 * gates = A[rows,K] Wp[4H,K]^T + bias_p with A = [x | h_prev], Wp = [W_ih | W_hh] whose rows are re-ordered per 64
 * hidden units as (i, f, g, o) blocks (bias_p = b_ih + b_hh likewise); c_prev is read through parent[] (nullable:
 * identity; c_prev nullable: zeros), c_out fp32 [rows,H], h (ab_dtype) goes to up to two row-major destinations.
 * The gate pre-activations never reach HBM.  H % 64 == 0. */
int dh_lstm_layer_tc(const void* A, long long lda, const void* Wp, long long ldw, int ab_dtype, const float* bias_p,
                     const float* c_prev, const int* parent, float* c_out, void* h_out0, long long ldh0, void* h_out1,
                     long long ldh1, int rows, int H, int K, cudaStream_t stream);
/* Every layer of one nn.LSTM time step in ONE persistent launch (rnn_models.py:23-24,80,108 with num_layers > 1).  Layer l
 * computes what dh_lstm_layer_tc computes, on stacked operands: A [layers, a_layer_rows, lda] with layer l's [x | h_prev] in
 * columns [0, in_l + H) (in_dims_host: HOST array; in_l == H for l > 0; columns past in_l + H must be zero in Wp), Wp
 * [layers * 4H, ldw] and bias_p [layers * 4H] gate-packed per layer, c_prev / c_out [layers, *, H] fp32 with
 * c_layer_stride elements between layers (c_prev rows read through parent[], shared by the layers), hs [layers, *, H]
 * (nullable) receives every layer's h.  Layer l < layers-1 writes its h into the x half of A[l+1]; the top layer writes
 * h_top [rows, ld_top].  Tiles are queued layer-major on the same CTAs; a tile of layer l > 0 starts on the recurrent half
 * of its K loop (rotate_k != 0) and takes the x half once `ready` says the layer below has stored those 128 rows.
 * ready: (layers - 1) * ceil(rows / 128) ints, ZERO on entry, left non-zero (one region per launch, or memset between). */
int dh_lstm_stack_tc(void* A, long long lda, long long a_layer_rows, const int* in_dims_host, const void* Wp,
                     long long ldw, int ab_dtype, const float* bias_p, const float* c_prev, const int* parent, float* c_out,
                     long long c_layer_stride, void* h_top, long long ld_top, void* hs, long long hs_layer_stride,
                     int* ready, int rows, int H, int layers, int rotate_k, cudaStream_t stream);
/* out = LayerNorm(x + y), eps 1e-5 (transformers.py:360,368,375,627,634). y may be null. */
int dh_add_layernorm(const void* x, long long ldx, const void* y, long long ldy, const float* gamma, const float* beta,
                     void* out, long long ldo, int rows, int D, int dtype, cudaStream_t stream);
/* (start_emb | tok_embedding[token]) / scale + pos_embedding[pos]  (transformers.py:455-470, 706-722). */
int dh_xfmr_embed(const void* tok_table, const void* pos_table, long long ldt, const float* start, long long lds,
                  int rows_per_start, const int* tokens, const int* positions, int pos_const, float scale, void* out,
                  long long ldo, int rows, int D, int dtype, cudaStream_t stream);
int dh_cast(const void* src, void* dst, long long n, int src_dtype, int dst_dtype, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------ attention
 * MultiHeadAttentionLayer.forward core (transformers.py:100-121) on cached K/V: one call covers incremental
 * self-attention with beam slot indirection, cross-attention over the 49 spatial tokens, and teacher-forced
 * causal attention (causal_full).  See csrc/attention.cu for the addressing rules. */
int dh_attention(const void* q, long long ldq, const void* K, const void* V, void* out, long long ldo, int rows, int D,
                 int n_heads, int rows_per_image, int slots, int S_alloc, const int* src, int slot_shared, int n_keys,
                 int causal_full, const int* seq, long long seq_ld, int seq_per_image, int pad,
                 const unsigned char* enc_mask, float scale, int dtype, cudaStream_t stream);
/* enc_mask[row] = any(spatial[row,:] == 0)  (transformers.py:480-481). */
int dh_enc_mask(const void* spatial, unsigned char* mask, int rows, int D, int dtype, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------ selection / beam
 * BeamSearchHelper (models/beam.py:32-108) and the generate() loops (rnn_models.py:84-143,
 * transformers.py:531-579, 778-825), batched over images with per-image "frozen at break" semantics. */
typedef struct dh_beam_state {
  int* seq;               /* [n_img, beam, seq_ld] token ids */
  long long seq_ld;
  float* val;             /* [n_img, beam] cumulative scores */
  unsigned char* ended;   /* [n_img, beam] */
  unsigned char* done;    /* [n_img] image has hit the reference's `break` */
  int* final_len;         /* [n_img] output length recorded at the break */
  int* last_tok;          /* [n_img*beam] next input token per row */
  int* parent_state;      /* [n_img*beam] row to gather recurrent state / KV from */
  int* src;               /* [n_img, beam, S_alloc] KV-cache slot table, or NULL (LSTM) */
  int S_alloc;
} dh_beam_state;

/* dyn (device, nullable): {seed, image_base} read at run time instead of the by-value arguments, so a captured CUDA
 * graph of the whole decode loop can be replayed with a new noise seed / shard offset. */
int dh_select_tokens(const float* logits, long long ld, int rows, int V, int beam, int top_k, float temperature, int unk,
                     int rows_per_image, int noise_mode, unsigned long long seed, long long image_base, int step,
                     const unsigned char* done, int* ind, float* val, int* status, const long long* dyn,
                     cudaStream_t stream);
/* Vocab projection fused with the selection, dense logits never stored (rnn_models.py:81,109 / transformers.py:488,736 ->
 * beam.py:32-53).  logits[M,N] = A[M,K] W[N,K]^T + bias (tcgen05, A / W ab_dtype):
 *   dh_vocab_groupmax    gmax[m, g] = max of the g-th 32-column group of logits[m, :] among the N tiles o, o + s, o + 2s, ...
 *                        (s = tile_stride, o = tile_offset; tile = 256 columns; groups past N hold -inf)   (pass 1)
 *   dh_vocab_threshold   thresh[m] = top_k-th largest of gmax[m, :]; cand_count[m] = 0
 *   dh_vocab_candidates  SPARSE MATERIALISATION (pass 2): every 32-column group of row m whose maximum is >= thresh[m] is
 *                        stored as is (one 128-byte line) at sp_logits[m, 32 g ..]; hitmap[m, 2 t + h] holds one bit per
 *                        stored group of N tile t, column half h (written for every (m, t, h), no atomics);
 *                        cand_count[m] += number of stored groups.  The epilogue does the same work wherever the candidates
 *                        sit, and a few per cent of the logits reach memory.
 *   dh_select_candidates gathers the stored groups, keeps the logits >= thresh[m], and runs dh_select_tokens on them.
 * The selection is exact for ANY threshold that leaves at least top_k candidates (and at most 480 stored groups / 512
 * candidates per row), because dh_select_candidates recomputes the exact top_k-th largest value.  Two ways to get one:
 *   exhaustive  tile_stride 1, rank top_k: the top_k-th largest group maximum is a lower bound of the top_k-th largest logit
 *               and at most top_k groups reach it.  Costs a second full contraction.
 *   sampled     tile_stride s > 1 and a rank j < top_k chosen so that the j-th largest SAMPLED group maximum lies below the
 *               top_k-th largest logit except with negligible probability (the host picks j from the binomial tail).  Pass 1
 *               costs 1/s.  The three *_fix entries then repair, inside the same stream / CUDA graph, the rare step in which
 *               a row stored fewer than count_min or more than count_max groups: each returns immediately unless such a row
 *               exists.  dh_vocab_groupmax_fix = exhaustive pass 1 (and *any_flag = 0); dh_vocab_threshold_fix = exact
 *               threshold, cand_count = 0, redo = 1 for the failing rows (redo = 0 for the others), *any_flag = 1;
 *               dh_vocab_candidates_fix = pass 2 for the rows with redo != 0, only if *any_flag != 0. */
typedef struct dh_vocab_sparse {
  const float* thresh;            /* [rows] */
  const unsigned char* hitmap;    /* [rows, hit_ld], hit_ld >= 2 * ceil(n_cols / 256) */
  long long hit_ld;
  const float* logits;            /* [rows, ld], ld >= ceil(n_cols / 256) * 256, ld % 4 == 0, 16-byte aligned */
  long long ld;
  int n_cols;                     /* N (vocabulary size) */
} dh_vocab_sparse;
int dh_vocab_groupmax(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias, int M,
                      int N, int K, int tile_stride, int tile_offset, float* gmax, long long ld_gmax, cudaStream_t stream);
int dh_vocab_threshold(const float* gmax, long long ld_gmax, int rows, int n_groups, int top_k, float* thresh,
                       int* cand_count, cudaStream_t stream);
int dh_vocab_candidates(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias, int M,
                        int N, int K, const float* thresh, int* cand_count, float* sp_logits, long long sp_ld,
                        unsigned char* hitmap, long long hit_ld, cudaStream_t stream);
int dh_vocab_groupmax_fix(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias, int M,
                          int N, int K, float* gmax, long long ld_gmax, const int* cand_count, int count_min, int count_max,
                          int* any_flag, cudaStream_t stream);
int dh_vocab_threshold_fix(const float* gmax, long long ld_gmax, int rows, int n_groups, int top_k, float* thresh,
                           int* cand_count, int count_min, int count_max, unsigned char* redo, int* any_flag,
                           cudaStream_t stream);
int dh_vocab_candidates_fix(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias, int M,
                            int N, int K, const float* thresh, int* cand_count, float* sp_logits, long long sp_ld,
                            unsigned char* hitmap, long long hit_ld, const unsigned char* redo, int* any_flag,
                            cudaStream_t stream);
int dh_select_candidates(const dh_vocab_sparse* cand, int rows, int beam, int top_k, float temperature, int unk,
                         int rows_per_image, int noise_mode, unsigned long long seed, long long image_base, int step,
                         const unsigned char* done, int* ind, float* val, int* status, const long long* dyn, cudaStream_t stream);
/* dh_select_candidates for the beam rows of every image followed by dh_beam_step of that image, in one launch
 * (one CTA per image, one warp per row).  ind / val [n_img*beam, beam] receive the picks as in dh_select_tokens. */
int dh_select_beam_step(const dh_vocab_sparse* cand, const dh_beam_state* st, int* ind, float* val, int* status, int n_img,
                        int beam, int top_k, float temperature, int unk, int step, int max_len, int eos, int lstm_semantics,
                        int noise_mode, unsigned long long seed, long long image_base, const long long* dyn, cudaStream_t stream);
/* The same launch, followed -- per image, once its beam step has chosen tokens and parents -- by the operand gathers of the
 * NEXT LSTM time step (what dh_lstm_prepare does in a launch of its own; rnn_models.py:107,135-137): A[0][r, 0:E] =
 * table[last_tok[r]] and A[l][r, in_off[l] : +H] = hs[l][parent_state[r]] for the image's beam rows.  Images that are done
 * keep their operands. */
typedef struct dh_lstm_operands {
  const void* table; long long ldt; long long n_tok_rows;   /* embedding table [n_tok_rows, ldt] (2-byte elements) */
  int E, H, L;                                               /* E, H multiples of 8; L <= 8 */
  const void* hs[8];                                         /* [*, H] contiguous per layer */
  void* A[8]; long long lda[8]; int in_off[8];               /* operand buffers, leading dimensions, column of the h half */
} dh_lstm_operands;
int dh_select_beam_step_lstm(const dh_vocab_sparse* cand, const dh_beam_state* st, int* ind, float* val, int* status, int n_img,
                             int beam, int top_k, float temperature, int unk, int step, int max_len, int eos, int lstm_semantics,
                             int noise_mode, unsigned long long seed, long long image_base, const long long* dyn,
                             const dh_lstm_operands* next, cudaStream_t stream);
int dh_beam_init(const dh_beam_state* st, const int* ind0, const float* val0, const int* prefix, long long prefix_ld,
                 int prefix_rows, int prefix_len, int n_img, int beam, int eos, int lstm_semantics, cudaStream_t stream);
int dh_beam_step(const dh_beam_state* st, const int* new_ind, const float* new_val, int n_img, int beam, int step,
                 int max_len, int eos, int lstm_semantics, float temperature, int noise_mode, unsigned long long seed,
                 long long image_base, const long long* dyn, cudaStream_t stream);
int dh_beam_final(const dh_beam_state* st, int n_img, int beam, float temperature, int noise_mode,
                  unsigned long long seed, long long image_base, int final_step, int len_if_running, int pad, int max_len,
                  long long* out_ids, long long* out_len, const long long* dyn, cudaStream_t stream);
/* out[m] = log_softmax(A[M,K] W[N,K]^T + bias)[m, targets[m]] with the logits kept on chip (experiments/metrics.py:5 on
 * the classifier of rnn_models.py:44 / transformers.py:488,736): tcgen05 contraction whose epilogue keeps (max, sum exp)
 * per 32-column group (gmax / gsum [M, ld_g], ld_g >= ceil(N/256)*8 for N > 128) and the target's logit (tlogit [M]),
 * then a per-row fold.  A / W ab_dtype (DH_BF16 / DH_F16). */
int dh_vocab_logprob(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias, int M,
                     int N, int K, const long long* targets, float* gmax, float* gsum, long long ld_g, float* tlogit,
                     float* out, cudaStream_t stream);
/* log_softmax(logits)[target] per row (experiments/metrics.py:5). */
int dh_token_logprob(const float* logits, long long ld, int rows, int V, const long long* targets, float* out,
                     cudaStream_t stream);

/* ------------------------------------------------------------------------------------------ path-level entries
 * A whole stage of the caption path behind ONE call (csrc/path.cu), for consumers that are not the Python runtime.  A
 * dh_ctx is a HOST table of device pointers to the packed weights (packing as documented per entry above; the Python
 * packer is deephumor_b200/runtime/encoder.py / xfmr.py); it owns no device memory.  Workspaces are the caller's. */
typedef struct dh_ctx dh_ctx;

#define DH_STAGE_RESNET50 1

typedef struct dh_resnet50_weights {
  const void* stem_w;          /* [64][192] fused-stem packing, K slot r*22 + s*3 + c (dh_stem_pool_tc) */
  const float* stem_b;
  const void* conv_w[16][3];   /* bottleneck b (3 + 4 + 6 + 3 in stage order): conv1 / conv2 / conv3, BN folded, */
  const float* conv_b[16][3];  /*   [Cout][kh][kw][Cin] in `dtype`; biases fp32 */
  const void* dual_w[4];       /* first bottleneck of each stage: [W_conv3 | W_downsample] along K (dh_conv1x1_dual_tc) */
  const float* dual_b[4];      /*   b_conv3 + b_downsample */
  float mean[3], std[3];       /* Normalize() statistics for uint8 input */
  int dtype;                   /* DH_F16 or DH_BF16: storage type of weights and activations */
} dh_resnet50_weights;

#define DH_XFMR_MAX_LAYERS 8
typedef struct dh_xfmr_layer {
  const void* qkv_w; const float* qkv_b;                    /* [3D, D] = [fc_q | fc_k | fc_v] of self_attn, [3D] */
  const void* so_w; const float* so_b; const float* sln_g; const float* sln_b; float s_scale;
  const void* cq_w; const float* cq_b;                      /* enc_attn (cross == 1): fc_q; K / V are projected per image */
  const void* co_w; const float* co_b; const float* cln_g; const float* cln_b; float c_scale;
  const void* f1_w; const float* f1_b; const void* f2_w; const float* f2_b; const float* fln_g; const float* fln_b;
} dh_xfmr_layer;
typedef struct dh_xfmr_weights {
  int n_layers, D, n_heads, pf, cross, dtype, pad;
  float scale;                                              /* decoder.scale = sqrt(hid_dim) */
  const void* tok; const void* pos; long long ld_tok;       /* embedding tables, `dtype` */
  dh_xfmr_layer layer[DH_XFMR_MAX_LAYERS];
} dh_xfmr_weights;
typedef struct dh_xfmr_buffers {
  void* x; void* qb; void* attn; void* tmp; void* h1;       /* [rows, D] x 4 (tmp only when D != 512), [rows, pf] */
  void* Kc[DH_XFMR_MAX_LAYERS]; void* Vc[DH_XFMR_MAX_LAYERS];   /* KV cache per layer [rows_total, S, D] */
  const void* xK[DH_XFMR_MAX_LAYERS]; const void* xV[DH_XFMR_MAX_LAYERS];   /* cross K / V per layer [n_img * 49, D] */
  const unsigned char* enc_mask;                            /* [n_img * 49] (dh_enc_mask) */
  const float* start; long long ld_start;                   /* image embedding [n_img, D] fp32 (position 0) */
  const int* seq; long long seq_ld;                         /* beam sequences (pad-key mask) */
  const int* src;                                           /* KV slot table [rows, S] (null in the prefix phase) */
  int slots, S;                                             /* cache slots per image (= beam), cached positions */
} dh_xfmr_buffers;

int dh_ctx_create(dh_ctx** out);
int dh_ctx_destroy(dh_ctx* ctx);
int dh_ctx_set_resnet50(dh_ctx* ctx, const dh_resnet50_weights* w);
int dh_ctx_set_xfmr(dh_ctx* ctx, const dh_xfmr_weights* w);
/* bytes of the caller-provided, 256-byte aligned workspace of a stage for n images of H x W */
int dh_workspace_bytes(const dh_ctx* ctx, int stage, int n, int H, int W, long long* bytes_out);
/* ResNet-50 trunk of ImageEncoder (encoders.py:34-39,56; torchvision resnet.py:143-163,197-204) + the global average pool
 * (encoders.py:39,60): images fp32 (or uint8 when images_u8) [n,3,224,224] NCHW -> feat [n,7,7,2048] NHWC (weights' dtype)
 * and pooled [n,2048] fp32 (nullable).  50 launches.  DH_ERR_UNSUPPORTED for other image sizes (use the per-op entries). */
int dh_resnet50_forward(const dh_ctx* ctx, const void* images_nchw, int images_u8, int n, int H, int W, void* feat,
                        float* pooled, void* workspace, long long workspace_bytes, cudaStream_t stream);
/* One new position for `rows` rows through the whole decoder stack against the KV cache (transformers.py:343-377 per layer,
 * :455-486 around it; SelfAttentionDecoderLayer :612-636 when cross == 0): embed, then per layer fused Q|K|V (K / V rows land
 * in cache slot `pos`), self-attention through the beam slot table, fc_o + residual + LayerNorm, [cross-attention over the
 * image's 49 tokens], FFN + residual + LayerNorm.  The hidden state of the new position is left in buffers->x. */
int dh_xfmr_step(const dh_ctx* ctx, const dh_xfmr_buffers* buffers, int rows, int rows_per_image, int pos, const int* tokens,
                 cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPHUMOR_B200_H_ */
