// Tensor-core contraction kernel for sm_100a: C[M,N] = act(A[M,K] * W[N,K]^T + bias[N] + residual[M,N]).
//
// One persistent, warp-specialised kernel serves every nn.Linear on the caption path (vocab projection,
// LSTM gate products, attention / FFN projections, embedding heads) and -- with the A operand fetched by
// im2col-mode TMA straight from the NHWC activation tensor -- every convolution of the ResNet-50 trunk
// (implicit GEMM: M = n*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin; BN folded, bias/ReLU/residual fused).
//
//   warp 0 (1 lane)  : TMA producer.  A tile 128 x 64 bf16 and W tile BN x 64 bf16 per stage, 128B-swizzled,
//                      completion on an mbarrier (cp.async.bulk.tensor, tiled or im2col mode).
//   warp 1 (1 lane)  : MMA issuer.  tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16, fp32 accumulators
//                      in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps tile i+1.
//   warp 2           : TMEM allocator / deallocator.
//   warps 4..11      : epilogue, two groups of four warps (a warp reads its own 32-lane quadrant of tensor memory, so a
//                      group covers the tile's 128 rows).  tcgen05.ld -> bias / ReLU / rounding -> 128B-swizzled slab ->
//                      TMA store, the slab rounds of a tile alternating between the groups (or, fallback for unaligned
//                      outputs: padded smem transpose -> coalesced 16 B global stores, one group).
//
// Tiles are scheduled statically (tile = blockIdx.x + i * gridDim.x, n fastest so CTAs running together
// share the A rows through L2); grid = min(tiles, #SMs).  M / N / K tails rely on TMA zero fill.
//
// Variants selected at compile time (template <BN, PAIR, EPI, ARES, G2>), each its own kernel in profiles:
//   PAIR   two CTAs of a cluster run one tcgen05.mma.cta_group::2 of M = 256, each staging half of the W tile
//   EPI 0  bias / ReLU / residual (R x I chunks on the tensor core) -> swizzled slab -> TMA store; optional fused global
//          average pool over image-aligned M tiles (dh_gemm_tc_pool); three destinations (dh_gemm_tc_split3)
//   EPI 1  maxima of 32-column groups (sampled pass 1 of the vocab projection)        EPI 2  sparse materialisation (pass 2)
//   EPI 3  LSTM cell update, all layers of a time step chained in one launch          EPI 4  log-softmax pieces (perplexity)
//   EPI 5  LayerNorm over the full 512-wide row: both N halves of a row block on one CTA (pair), packed-fp32 epilogue
//   EPI 6  the same with the row SPLIT over a cluster: two CTAs (or two CTA pairs, a 4-CTA cluster) hold the two 256-column
//          halves of a row block, exchange per-row (mean, M2) through st.async + transaction barriers, and keep the
//          accumulator double-buffered (the default form of dh_gemm_tc_ln)
//   ARES   the A row block stays resident in shared memory, the CTA walks a contiguous run of N tiles (pass 2)
//   G2     plain-store epilogue (EPI 0) form: 0 one group / two slabs (long K loops), 1 both groups / one slab each and the
//          full ring (single-CTA 256-wide residual launches), 2 both groups / two slabs each (K <= 1024), 3 chain mode:
//          after tile P a CTA also contracts the rows it has just stored with a second weight matrix (tile Q, operand read
//          back from L2): a bottleneck's conv3 + the next bottleneck's conv1 in one launch (dh_conv1x1_chain_tc)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 384;            // warps 0-2 producer / MMA / TMEM, 4-7 epilogue, 8-11 second epilogue group
constexpr int kStageLd = 36;             // floats per staged row (32 + 4: conflict-free for 16 B accesses)
constexpr int kSmemBudget = 196608;      // bytes of A/B ring
constexpr int kMaxLayers = 8;            // LSTM layers one dh_lstm_stack_tc launch can chain
constexpr int kChainLag = 2;             // chain mode: tile Q_j follows tile P_(j + kChainLag)

struct TcParams {
  int M, N, K;
  int m_blocks, n_blocks, k_chunks;
  int conv;                               // 0: A via tiled TMA {K, M}; 1: A via im2col TMA {C, W, H, N}
  int HoWo, Wo, c_chunks, kw, stride, pad;
  int k1_chunks, stride2;                 // > 0: K chunks >= k1_chunks come from a SECOND 1x1 source (map_r) of stride2
  const float* bias;
  const void* res; long long ldr; int res_dtype;
  void* out; long long ldc; int out_dtype;
  int relu;
  int ab_dtype;                           // DH_BF16 or DH_F16 operands
  int tma_store;                          // 1: epilogue stages 128 B-wide row slabs in smem and stores them by TMA
  int res_chunks;                         // > 0: residual added on the tensor core as BN/64 extra K chunks (R x I)
  int* error;                             // device flag set before a watchdog trap
  // Vocab-projection epilogues that keep the logits out of HBM (dh_vocab_groupmax / dh_vocab_candidates):
  int n_stride;                           // N blocks visited: 0, n_stride, 2 n_stride, ... (1 except for sampled group maxima)
  int epi_mode;                           // 0: store C; 1: maxima of 32-column groups; 2: compact logits >= thresh[row];
                                          // 3: LSTM cell; 4: per-group (max, sum exp) + target logit (log-softmax);
                                          // 5: LayerNorm over the full row (both N halves on one CTA)
  float* gsum; const long long* targets; float* tlogit;     // mode 4: [M, ld_gmax], [M] (int64), [M]
  float* gmax; long long ld_gmax;         // [M, ld_gmax] group maxima (mode 1)
  const float* thresh;                    // [M] lower bound of the row's top_k-th largest logit (mode 2)
  // mode 2 (sparse materialisation): every 32-column group whose maximum reaches thresh[row] is stored as is (128 B) into
  // the dense-pitch buffer sp_logits [M, sp_ld]; hitmap [M, hit_ld] gets one byte per (row, N tile, column half) with one
  // bit per stored group -- written exactly once, no atomics, no data-dependent control flow beyond the predicated stores;
  // cand_count[row] accumulates the number of stored groups (fire-and-forget reduction).
  float* sp_logits; long long sp_ld; unsigned char* hitmap; long long hit_ld; int* cand_count;
  // LSTM cell epilogue (epi_mode 3, dh_lstm_layer_tc): the N axis is packed per 64 hidden units as [i | f | g | o]
  const float* c_prev; const int* parent; float* c_out;            // [*, H] fp32, parent[M] (nullable), [M, H] fp32
  void* h0; long long ldh0; void* h1; long long ldh1; int H;       // bf16 h to up to two destinations
  // Multi-layer LSTM step in ONE launch (dh_lstm_stack_tc): tiles are ordered layer-major and walked in that order by every
  // CTA; the tiles of layer l > 0 covering rows [m0, m0 + 128) load their x half only after ready[(l-1) * mb128 + m0 / 128]
  // has reached n_blocks, i.e. every N tile of the layer below has stored h for those rows.  layers <= 1: plain single layer.
  int layers, tiles_per_layer, mb128;
  int a_layer_rows, w_layer_rows;                                  // row offset of layer l in the stacked A / Wp tensor maps
  int kch_l[kMaxLayers], rot_l[kMaxLayers];                        // K chunks of layer l; chunk the K loop starts from
  const float* bias_l[kMaxLayers]; const float* cprev_l[kMaxLayers]; float* cout_l[kMaxLayers];
  void* h0_l[kMaxLayers]; long long ldh0_l[kMaxLayers]; void* h1_l[kMaxLayers]; long long ldh1_l[kMaxLayers];
  int* ready;
  // Three row-major destinations for one contraction (dh_gemm_tc_split3: fused Q | K | V projection): N tile n0 goes to
  // destination n0 / split_n through map_c / map_r / map_i, at column n0 % split_n.  0: single destination.
  int split_n;
  // Sampled vocab selection (dh_vocab_*_fix): n_offset = first N block visited by a strided pass 1; cond_mode 1 = run only
  // if some row's candidate count lies outside [cond_min, cond_max] (the sampled threshold missed), 2 = run only if
  // *cond_flag != 0; redo (nullable) = rows whose candidates this launch may emit.
  int n_offset, cond_mode, cond_min, cond_max;
  const int* cond_count; int* cond_flag; const unsigned char* redo;
  // Image-aligned M tiles + fused global average pool (dh_gemm_tc_pool: the last bottleneck's conv3, encoders.py:60-61):
  // a tile covers bm_rows = floor(128 / pool_hw) * pool_hw rows, i.e. whole images, so the slab epilogue can reduce the
  // pool_hw rows of each image per column and store their mean (fp32) without a second pass over the feature map.
  int bm_rows;                            // rows of A / C per CTA tile (128 unless pooling)
  int pool_hw; float* pool_out; long long ld_pool;
  // LayerNorm epilogue (epi_mode 5, dh_gemm_tc_ln): N == 2 * BN, a CTA computes BOTH N halves of its row block into the two
  // accumulator buffers, its two epilogue groups exchange row sums, and out = LN(A W^T + bias + residual) * gamma + beta
  const float* ln_gamma; const float* ln_beta; float ln_eps;
  // Chained 1x1 convolution (dh_conv1x1_chain_tc, template flag G2 == 3): after tile P (this launch's C = act(...), N == BN) a
  // CTA also computes tile Q = act2(C[m0 : m0 + 128, :] W2^T + bias2) for the SAME rows, reading C back through TMA while it is
  // still in L2 -- the next bottleneck's conv1 without its HBM read.  W2 [N2, N] and the Q destination come in ChainMaps.
  int N2, k2_chunks, relu2;
  const float* bias2;
};

struct alignas(64) ChainMaps { CUtensorMap b2, c2; };
thread_local const ChainMaps* tl_chain = nullptr;         // set by dh_conv1x1_chain_tc around its dispatch

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      if (error) atomicExch(error, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h,
                                                int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the pair shares one UMMA of M = 256; both CTAs' loads signal the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h,
                                                     int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Default semantics (release at CTA scope), as CUTLASS' ClusterBarrier::arrive: an explicit .release.cluster compiles to
// MEMBAR.ALL.GPU and stalls the epilogue warp until all of its earlier global stores have drained (ncu: 25 % of the LSTM
// epilogue); the TMEM hand-off itself is ordered by tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Row statistics exchanged between the two CTAs of a cluster (LayerNorm split over the cluster, EPI 6): the partial lands in
// the peer's shared memory and counts as transaction bytes on the peer's barrier; the reader acquires at cluster scope.
// (st.async: the store itself completes 8 transaction bytes on the peer's barrier -- an explicit
// mbarrier.arrive.release.cluster after a plain st.shared::cluster compiles to MEMBAR.ALL.GPU: 35 % of this epilogue's samples)
__device__ __forceinline__ void st_async_f32x2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr),
               "f"(a), "f"(b), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int* error, int code) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      if (error) atomicExch(error, code);
      __threadfence_system();
      __trap();
    }
  }
}
// arrives on `bar` in BOTH CTAs of the pair (cluster ranks lead, lead + 1: mask 3 << lead)
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint32_t lead = 0) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)(3u << lead))
               : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld32_nw(uint32_t taddr, uint32_t* v) {   // no wait: pair with tc_wait_ld()
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// Packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100a): two IEEE fp32 operations per issued instruction
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void un2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ int dh_cdiv_dev(int a, int b) { return (a + b - 1) / b; }

// Cross-CTA hand-off of LSTM layer outputs inside one launch: the writer's epilogue stores h with ordinary (generic-proxy)
// stores, fences, and bumps a counter with release semantics; the reader's TMA producer spins with acquire loads and then
// orders its async-proxy (TMA) reads after them.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void wait_ready(const int* flag, int target, int* error) {
  if (ld_acquire_gpu(flag) >= target) { fence_proxy_async_all(); return; }
  const long long t0 = clock64();
  while (ld_acquire_gpu(flag) < target) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000ll) {
      if (error) atomicExch(error, 5);
      __threadfence_system();
      __trap();
    }
  }
  fence_proxy_async_all();
}

// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row atoms of 1024 B):
// start address >> 4 | SBO = 1024 B (bits 32..45) | descriptor version 1 (bits 46..47) | SWIZZLE_128B (bits 61..63).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16 (format 1) or f16 (format 0), both K-major, M = 128, N = bn.
__device__ __forceinline__ uint32_t umma_idesc(int bn, int ab_dtype, int m = BM) {
  const uint32_t fmt = ab_dtype == DH_BF16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// PAIR: two CTAs of a cluster (one TPC) run ONE tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own 128 A rows and
// HALF of the W tile (BN/2 rows) and owns the accumulator of its 128 rows -- a third less operand traffic per FLOP from L2,
// which is what bounds these contractions (profiles/: ~12 TB/s chip-wide TMA ceiling).
// ARES ("A resident"): the CTA keeps the whole K extent (<= 8 chunks = 128 KB) of its 128 A rows in shared memory and walks a
// CONTIGUOUS run of N tiles of that row block, so only the W tiles stream through the ring.  A contraction with K = 512 and
// 256 x 256 pair tiles moves 512 KB from L2 per 67 MFLOP when A and W both stream -- 17 TB/s at the tensor peak, above what L2
// delivers (~11.5 TB/s measured: profiles/r02_ncu_vocab_pass2_40960.csv shows 74 % tensor-active) -- and half of that with A
// resident.
template <int BN, bool PAIR = false, int EPI = 0, bool ARES = false, int G2 = 0>
struct Cfg {
  static constexpr int kMaxResChunks = 8;
  static constexpr int kAresBytes = ARES ? kMaxResChunks * BM * BK * 2 : 0;
  static constexpr int kStageBytes = ((ARES ? 0 : BM) + (PAIR ? BN / 2 : BN)) * BK * 2;
  // LayerNorm mode gives up ring stages for its parameter block (5 x 32 KB stages as a pair, 3 x 48 KB alone)
  // plain stores with both epilogue groups and two slabs each (EPI 0, G2 == 2): four slabs, paid for with ring depth
  static constexpr int kStages = EPI == 6 ? (PAIR ? 5 : 3) : EPI == 5 ? (PAIR ? 5 : 3) : ARES ? (PAIR ? 5 : 3) : (kSmemBudget - (EPI == 0 && G2 == 2 ? 2 * BM * 128 : 0)) / kStageBytes;
  static constexpr int kRingBytes = kAresBytes + kStages * kStageBytes;     // resident A block + ring
  static constexpr int kTmemCols = 2 * BN;
  // 128-row x 128 B slabs (TMA store; two per epilogue group in plain-store mode, one per group in LayerNorm mode) / per-warp
  // transpose scratch; the selection epilogues (1, 2, 4) stage nothing
  static constexpr int kStagingBytes = (EPI == 1 || EPI == 2 || EPI == 4) ? 0 : (EPI == 0 && G2 == 2 ? 4 : 2) * BM * 128;
  // bias slice of the current tile (BN <= 256 floats); double-buffered in the selection epilogues
  // (plain stores: one copy per epilogue group)
  static constexpr int kBiasBytes = (EPI == 1 || EPI == 2 || EPI == 4) ? 2048 : (EPI == 0 && G2 && BN == 256) ? 2048 : 1024;
  // LayerNorm mode: bias | gamma | beta of the whole 2 BN-wide row (fp32) + double-buffered per-row (mean, M2) partials of
  // the two column halves
  // (split form, EPI 6: four partials per row -- two column quarters of each of the cluster's two CTAs -- and two barriers)
  static constexpr int kLnBytes = EPI == 5 ? 3 * 2 * BN * 4 + 2 * 2 * BM * 2 * 4 : EPI == 6 ? 3 * 2 * BN * 4 + 2 * 4 * BM * 2 * 4 + 64 : 0;
  static constexpr int kLnOff = kRingBytes + kStagingBytes + kBiasBytes + 256;
  // 1 KB of slack to align the ring to the 1024-byte swizzle atom -- except where that would exceed the 227 KB a CTA may own
  // (<256, pair, plain stores>): there the kernel requires the dynamic shared memory window itself to be 1024-byte aligned
  // (it is: the window starts right after the driver's 1 KB reservation) and traps with error code 6 otherwise
  static constexpr int kAlignSlack = (kLnOff + kLnBytes + 1024 <= 232448) ? 1024 : 0;
  static constexpr int kSmemBytes = kAlignSlack + kLnOff + kLnBytes;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// EPI = TcParams::epi_mode as a compile-time constant: every epilogue is its own kernel (named in profiles, no dead code)
template <int BN, bool PAIR, int EPI, bool ARES = false, int G2 = 0>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
               const __grid_constant__ CUtensorMap map_i, const TcParams p, const __grid_constant__ ChainMaps cm) {
  using C = Cfg<BN, PAIR, EPI, ARES, G2>;
  constexpr int CG = PAIR ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (C::kAlignSlack == 0 && base != smem_u32(smem_raw)) {
    if (p.error) atomicExch(p.error, 6);
    __threadfence_system();
    __trap();
  }
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ares = base;                       // resident A block (ARES): chunk kc at ares + kc * 16 KB
  const uint32_t ring = base + C::kAresBytes;
  float* staging = reinterpret_cast<float*>(gen_base + C::kRingBytes);
  float* bias_s = reinterpret_cast<float*>(gen_base + C::kRingBytes + C::kStagingBytes);
  const uint32_t bars = base + C::kRingBytes + C::kStagingBytes + C::kBiasBytes;
  // barrier slots (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base word
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::kStages + 2 + s); };
  const uint32_t afull_bar = bars + 8u * (2 * C::kStages + 4), afree_bar = bars + 8u * (2 * C::kStages + 5);   // ARES only
  const uint32_t stored_bar0 = bars + 8u * (2 * C::kStages + 6);                    // chain mode: four "tile P is in L2" barriers
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(gen_base + C::kRingBytes + C::kStagingBytes +
                                                    C::kBiasBytes + 8 * (2 * C::kStages + 10));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.cond_mode) {
    // fix-up launches of the sampled vocab selection: nothing to do unless a row's sampled threshold missed.  Every CTA
    // (both CTAs of a pair) evaluates the same predicate on the same data, before any barrier / TMEM set-up.
    int bad = 0;
    if (p.cond_mode == 1) {
      for (int i = threadIdx.x; i < p.M; i += kThreads) {
        const int c = __ldg(p.cond_count + i);
        bad |= (c < p.cond_min) | (c > p.cond_max);
      }
      if (blockIdx.x == 0 && threadIdx.x == 0 && p.cond_flag) *p.cond_flag = 0;      // set again by dh_vocab_threshold_fix
    } else {
      bad = *reinterpret_cast<const volatile int*>(p.cond_flag) != 0;
    }
    if (!__syncthreads_or(bad)) return;
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_c);
    if (p.res_chunks) { tma_prefetch_desc(&map_r); tma_prefetch_desc(&map_i); }
    if (p.k1_chunks) tma_prefetch_desc(&map_r);
    if (p.split_n) { tma_prefetch_desc(&map_r); tma_prefetch_desc(&map_i); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      // the leader's barrier collects both CTAs' epilogue warps: both groups except in LayerNorm mode (one group drains each
      // buffer) and in the direct-store fallback of the plain epilogue (one group)
      mbar_init(tempty_bar(s), ((EPI == 5 || (EPI == 0 && (!p.tma_store || !G2))) ? 4 : 8) * CG);
    }
    if (ARES) { mbar_init(afull_bar, 1); mbar_init(afree_bar, 1); }
    if (G2 == 3) { for (int b = 0; b < 4; ++b) mbar_init(stored_bar0 + 8u * b, 2); }   // one arrive per epilogue group
    if (EPI == 6) {                       // row-statistics barriers: 256 local arrivals + 256 x 8 bytes stored by the peer
      const uint32_t sb = base + C::kLnOff + 3 * 2 * BN * 4 + 2 * 4 * BM * 2 * 4;
      mbar_init(sb, 256); mbar_init(sb + 8, 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)),
                   "r"((uint32_t)C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)),
                   "r"((uint32_t)C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR || EPI == 6) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_word);
  // tile = (M block of BM * CG rows, N block); a CTA pair walks the tiles together, CTA `rank` owning rows +rank * BM
  // (a LayerNorm cluster holds TWO pairs, ranks {0, 1} and {2, 3}: `lead` is the rank of this CTA's pair leader)
  const uint32_t cta_rank = PAIR ? (cluster_ctarank() & 1u) : 0u;
  const uint32_t lead = (PAIR && EPI == 6) ? (cluster_ctarank() & 2u) : 0u;
  const int tiles = p.m_blocks * p.n_blocks * ((EPI == 3 && p.layers > 1) ? p.layers : 1);
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // virtual tile sequence of this CTA (pair): tile_of(0), tile_of(1), ... until >= tiles.  LayerNorm mode walks ROW BLOCKS with
  // the stride and visits both N halves of a block back to back (accumulator buffer = half)
  // ARES: CTA (pair) c owns the contiguous run [c * per, (c + 1) * per) of the m-major tile order, i.e. consecutive N tiles of
  // (mostly) one row block
  const int ares_per = ARES ? dh_cdiv_dev(tiles, tstride) : 0;
  const int ares_end = ARES ? min(tiles, (tile0 + 1) * ares_per) : 0;
  auto tile_of = [&](int it) {
    if (ARES) { const int t = tile0 * ares_per + it; return t < ares_end ? t : tiles; }
    return EPI == 5 ? 2 * (tile0 + (it >> 1) * tstride) + (it & 1) : tile0 + it * tstride;
  };
  auto tile_m0 = [&](int tile) { return (tile / p.n_blocks) * (p.bm_rows * CG) + (int)cta_rank * p.bm_rows; };
  // Chain mode (G2 == 3, single CTA, N == BN): this CTA's own tiles t_j = tile0 + j * tstride, j < n_own, are walked as the
  // virtual sequence P0 P1 P2 Q0 P3 Q1 ... P(n-1) Q(n-3) Q(n-2) Q(n-1) by all three roles -- Q_j (the second contraction over
  // the rows of tile j) runs kChainLag tiles behind P_j, so P_j's TMA stores have landed in L2, and the "stored" hand-off
  // (store completion -> barrier -> producer -> operand loads -> MMAs) has two tiles' worth of epilogue work to hide behind.
  // (a unit = one 128-row block: its P part is n_blocks tiles of BN columns, its Q part one tile of N2 columns)
  const int n_own = (G2 == 3 && tile0 < p.m_blocks) ? (p.m_blocks - tile0 + tstride - 1) / tstride : 0;
  const int chain_lead = n_own < kChainLag + 1 ? n_own : kChainLag + 1;      // P tiles before the first Q
  auto chain_vt = [&](int v, bool& isq) {
    if (v < chain_lead) { isq = false; return v; }
    const int w = v - chain_lead, pairs = n_own - chain_lead;               // then Q_k, P_(lead + k) alternate
    if (w < 2 * pairs) { isq = !(w & 1); return (w & 1) ? chain_lead + (w >> 1) : (w >> 1); }
    isq = true;
    return pairs + (w - 2 * pairs);                                          // and the last Q tiles drain
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================================================================== TMA producer
      int stage = 0;
      uint32_t phase = 0;
      int ares_mb = -1, ares_loads = 0;
      if constexpr (G2 == 3) {
        for (int v = 0; v < 2 * n_own; ++v) {
          bool isq;
          const int j = chain_vt(v, isq);
          const int m0 = (tile0 + j * tstride) * p.bm_rows;
          if (isq) {
            // tile P_j is complete in L2 (both epilogue groups waited for their bulk stores): read it back as Q's A operand
            mbar_wait(stored_bar0 + 8u * (uint32_t)(j & 3), (uint32_t)((j >> 2) & 1), p.error, 1);
            fence_proxy_async_all();
            for (int kc = 0; kc < p.k2_chunks; ++kc) {
              mbar_wait(empty_bar(stage), phase ^ 1u, p.error, 1);
              const uint32_t sa = ring + stage * C::kStageBytes, sb = sa + BM * BK * 2;
              mbar_expect_tx(full_bar(stage), (uint32_t)(p.bm_rows + p.N2) * (BK * 2));
              tma_load_2d(sa, &map_c, full_bar(stage), kc * BK, m0);
              tma_load_2d(sb, &cm.b2, full_bar(stage), kc * BK, 0);
              if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
            }
            continue;
          }
          int img = 0, ph = 0, qw = 0;
          if (p.conv) {
            img = m0 / p.HoWo;
            const int rem = m0 - img * p.HoWo;
            ph = rem / p.Wo;
            qw = rem - ph * p.Wo;
          }
          for (int sub = 0; sub < p.n_blocks; ++sub) {
            const int n0 = sub * BN;
            for (int kc = 0; kc < p.k_chunks; ++kc) {
              mbar_wait(empty_bar(stage), phase ^ 1u, p.error, 1);
              const uint32_t sa = ring + stage * C::kStageBytes, sb = sa + BM * BK * 2;
              mbar_expect_tx(full_bar(stage), (uint32_t)(p.bm_rows + BN) * (BK * 2));
              if (p.conv && p.k1_chunks && kc >= p.k1_chunks)
                tma_load_im2col(sa, &map_r, full_bar(stage), (kc - p.k1_chunks) * BK, qw * p.stride2, ph * p.stride2, img, 0, 0);
              else if (p.conv)
                tma_load_im2col(sa, &map_a, full_bar(stage), kc * BK, qw, ph, img, 0, 0);          // 1x1, stride 1, no padding
              else
                tma_load_2d(sa, &map_a, full_bar(stage), kc * BK, m0);
              tma_load_2d(sb, &map_b, full_bar(stage), kc * BK, n0);
              if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
            }
            for (int r = 0; r < p.res_chunks; ++r) {
              mbar_wait(empty_bar(stage), phase ^ 1u, p.error, 1);
              const uint32_t sa = ring + stage * C::kStageBytes, sb = sa + BM * BK * 2;
              mbar_expect_tx(full_bar(stage), p.bm_rows * BK * 2 + 64 * BK * 2);
              tma_load_2d(sa, &map_r, full_bar(stage), n0 + r * BK, m0);
              tma_load_2d(sb, &map_i, full_bar(stage), 0, 0);
              if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      } else
      for (int it = 0, tile; (tile = tile_of(it)) < tiles; ++it) {
        int layer = 0, rt = tile;                                // EPI 3: layer-major tile order of a stacked LSTM step
        if (EPI == 3 && p.layers > 1) { layer = tile / p.tiles_per_layer; rt = tile - layer * p.tiles_per_layer; }
        const int m0 = tile_m0(rt), n0 = ((rt % p.n_blocks) * p.n_stride + p.n_offset) * BN;
        // an M block past the end (odd block count, second CTA of the last pair) re-loads the last valid block: its
        // accumulator is never stored, and every TMA coordinate stays inside the tensor
        const int m0l = min(m0, (dh_cdiv_dev(p.M, p.bm_rows) - 1) * p.bm_rows);
        const uint32_t stage_tx = (uint32_t)(p.bm_rows + BN / CG) * (BK * 2);   // bytes one CTA's loads of a stage deliver
        int img = 0, ph = 0, qw = 0;
        if (p.conv) {
          img = m0l / p.HoWo;
          const int rem = m0l - img * p.HoWo;
          ph = rem / p.Wo;
          qw = rem - ph * p.Wo;
        }
        const int a_row = m0l + ((EPI == 3 && p.layers > 1) ? layer * p.a_layer_rows : 0);
        const int nb0 = n0 + (int)cta_rank * (BN / CG) +         // this CTA's slice of the W tile
                        ((EPI == 3 && p.layers > 1) ? layer * p.w_layer_rows : 0);
        const int kch = (EPI == 3 && p.layers > 1) ? p.kch_l[layer] : p.k_chunks;
        const int rot = (EPI == 3 && p.layers > 1) ? p.rot_l[layer] : 0;
        if (ARES) {
          // resident A: (re)load the row block's K extent when the run crosses into a new row block -- after the MMAs that
          // read the previous one have retired -- then stream only W tiles through the ring
          const int mb = rt / p.n_blocks;
          if (mb != ares_mb) {
            if (ares_mb >= 0) mbar_wait(afree_bar, (uint32_t)((ares_loads - 1) & 1), p.error, 1);
            if (cta_rank == 0) mbar_expect_tx(afull_bar, (uint32_t)(kch * p.bm_rows * BK * 2) * CG);
            const uint32_t ab = PAIR ? mapa_u32(afull_bar, lead) : afull_bar;
            for (int kc = 0; kc < kch; ++kc) {
              if (PAIR) tma_load_2d_pair(ares + kc * (BM * BK * 2), &map_a, ab, kc * BK, a_row);
              else tma_load_2d(ares + kc * (BM * BK * 2), &map_a, ab, kc * BK, a_row);
            }
            ares_mb = mb;
            ++ares_loads;
          }
          for (int kc = 0; kc < kch; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.error, 1);
            const uint32_t sb = ring + stage * C::kStageBytes;
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), (uint32_t)C::kStageBytes * CG);
            if (PAIR) tma_load_2d_pair(sb, &map_b, mapa_u32(full_bar(stage), lead), kc * BK, nb0);
            else tma_load_2d(sb, &map_b, full_bar(stage), kc * BK, nb0);
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
          continue;
        }
        for (int j = 0; j < kch; ++j) {
          // a stacked layer starts its K loop at the recurrent half (chunk rot), which is ready at launch, and reaches
          // the x half -- written by the layer below during this launch -- last
          int kc = j + rot;
          if (kc >= kch) kc -= kch;
          if (EPI == 3 && layer > 0 && kc == 0) wait_ready(p.ready + (layer - 1) * p.mb128 + (m0l >> 7), p.n_blocks, p.error);
          mbar_wait(empty_bar(stage), phase ^ 1u, p.error, 1);
          const uint32_t sa = ring + stage * C::kStageBytes, sb = sa + BM * BK * 2;
          if (cta_rank == 0) mbar_expect_tx(full_bar(stage), stage_tx * CG);
          if (PAIR) {
            const uint32_t fb = mapa_u32(full_bar(stage), lead);
            if (p.conv && p.k1_chunks && kc >= p.k1_chunks) {
              // second source of a dual 1x1 convolution (the bottleneck's downsample branch, torchvision resnet.py:157-158)
              tma_load_im2col_pair(sa, &map_r, fb, (kc - p.k1_chunks) * BK, qw * p.stride2, ph * p.stride2, img, 0, 0);
            } else if (p.conv) {
              const int tap = kc / p.c_chunks, c0 = (kc - tap * p.c_chunks) * BK;
              const int r = tap / p.kw, s = tap - r * p.kw;
              tma_load_im2col_pair(sa, &map_a, fb, c0, qw * p.stride - p.pad, ph * p.stride - p.pad, img, (uint16_t)s,
                                   (uint16_t)r);
            } else {
              tma_load_2d_pair(sa, &map_a, fb, kc * BK, a_row);
            }
            tma_load_2d_pair(sb, &map_b, fb, kc * BK, nb0);
          } else {
            if (p.conv && p.k1_chunks && kc >= p.k1_chunks) {
              tma_load_im2col(sa, &map_r, full_bar(stage), (kc - p.k1_chunks) * BK, qw * p.stride2, ph * p.stride2, img, 0, 0);
            } else if (p.conv) {
              const int tap = kc / p.c_chunks, c0 = (kc - tap * p.c_chunks) * BK;
              const int r = tap / p.kw, s = tap - r * p.kw;
              tma_load_im2col(sa, &map_a, full_bar(stage), c0, qw * p.stride - p.pad, ph * p.stride - p.pad, img,
                              (uint16_t)s, (uint16_t)r);
            } else {
              tma_load_2d(sa, &map_a, full_bar(stage), kc * BK, a_row);
            }
            tma_load_2d(sb, &map_b, full_bar(stage), kc * BK, nb0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
        // residual as extra K chunks: D += R[m0:m0+128, n0+64j : +64] x I[:, 64j : +64]^T  (exact in fp32 accumulate)
        for (int j = 0; j < p.res_chunks; ++j) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.error, 1);
          const uint32_t sa = ring + stage * C::kStageBytes, sb = sa + BM * BK * 2;
          if (PAIR) {
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), stage_tx * CG);
            const uint32_t fb = mapa_u32(full_bar(stage), lead);
            tma_load_2d_pair(sa, &map_r, fb, n0 + j * BK, m0l);
            tma_load_2d_pair(sb, &map_i, fb, j * BK, (int)cta_rank * (BN / CG));
          } else {
            // single CTA: the chunk's B operand is just the 64 x 64 identity (8 KB, not BN x 64 of mostly zeros); its MMAs
            // write accumulator columns 64 j .. 64 j + 63 with N = 64
            mbar_expect_tx(full_bar(stage), p.bm_rows * BK * 2 + 64 * BK * 2);
            tma_load_2d(sa, &map_r, full_bar(stage), n0 + j * BK, m0l);
            tma_load_2d(sb, &map_i, full_bar(stage), 0, 0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      // ===================================================================== MMA issuer (the pair's leader CTA only)
      const uint32_t idesc = umma_idesc(BN, p.ab_dtype, BM * CG);
      const uint32_t idesc64 = umma_idesc(64, p.ab_dtype, BM);
      int stage = 0;
      uint32_t phase = 0;
      int ares_mb = -1, ares_cnt = 0;
      if constexpr (G2 == 3) {
        const uint32_t idesc2 = umma_idesc(p.N2, p.ab_dtype, BM);
        int vc = 0;                                              // accumulator buffers alternate per (sub-)tile
        for (int v = 0; v < 2 * n_own; ++v) {
          bool isq;
          chain_vt(v, isq);
          for (int sub = 0; sub < (isq ? 1 : p.n_blocks); ++sub, ++vc) {
            const int as = vc & 1;
            mbar_wait(tempty_bar(as), ((vc >> 1) & 1) ^ 1u, p.error, 2);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
            const int chunks = isq ? p.k2_chunks : p.k_chunks + p.res_chunks;
            for (int kc = 0; kc < chunks; ++kc) {
              mbar_wait(full_bar(stage), phase, p.error, 3);
              tc_fence_after();
              const uint32_t sa = ring + stage * C::kStageBytes, sb = sa + BM * BK * 2;
              const uint64_t da = umma_desc(sa), db = umma_desc(sb);
              if (isq) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) tc_mma(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (kc | k) ? 1u : 0u);
              } else if (kc >= p.k_chunks) {
                const uint32_t td = tmem_d + (uint32_t)((kc - p.k_chunks) * 64);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) tc_mma(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc64, 1u);
              } else {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) tc_mma(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | k) ? 1u : 0u);
              }
              tc_commit(empty_bar(stage));
              if (kc == chunks - 1) tc_commit(tfull_bar(as));
              if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      } else
      for (int it = 0, tile; (tile = tile_of(it)) < tiles; ++it) {
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1u, p.error, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        const int chunks = ((EPI == 3 && p.layers > 1) ? p.kch_l[tile / p.tiles_per_layer] : p.k_chunks) + p.res_chunks;
        if (ARES && tile / p.n_blocks != ares_mb) {              // a new row block: its resident A must have landed
          mbar_wait(afull_bar, (uint32_t)(ares_cnt & 1), p.error, 3);
          tc_fence_after();
          ares_mb = tile / p.n_blocks;
          ++ares_cnt;
        }
        for (int kc = 0; kc < chunks; ++kc) {
          mbar_wait(full_bar(stage), phase, p.error, 3);
          tc_fence_after();
          const uint32_t sa = ARES ? ares + kc * (BM * BK * 2) : ring + stage * C::kStageBytes;
          const uint32_t sb = ARES ? ring + stage * C::kStageBytes : sa + BM * BK * 2;
          const uint64_t da = umma_desc(sa), db = umma_desc(sb);
          if (!PAIR && kc >= p.k_chunks) {
            // residual chunk j: D[:, 64 j .. 64 j + 63] += R[:, n0 + 64 j ..] x I64^T
            const uint32_t td = tmem_d + (uint32_t)((kc - p.k_chunks) * 64);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) tc_mma(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc64, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {  // +32 B per UMMA_K step inside the 128 B swizzle row
              if (PAIR) tc_mma_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | k) ? 1u : 0u);
              else tc_mma(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | k) ? 1u : 0u);
            }
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (PAIR) tc_commit_pair(empty_bar(stage), lead); else tc_commit(empty_bar(stage));
          if (kc == chunks - 1) { if (PAIR) tc_commit_pair(tfull_bar(as), lead); else tc_commit(tfull_bar(as)); }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
        if (ARES) {
          // last tile of this row block in the run: once its MMAs retire the producer may overwrite the resident A
          const int nt = tile_of(it + 1);
          if (nt >= tiles || nt / p.n_blocks != ares_mb) { if (PAIR) tc_commit_pair(afree_bar, lead); else tc_commit(afree_bar); }
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================================================= epilogue
    const int ew = warp & 3;                      // TMEM lane quadrant this warp may read
    const int eh = (warp - 4) >> 2;               // 0: warps 4-7, 1: warps 8-11 (other column half; fused epilogues only)
    const int etid = (warp - 4) * 32 + lane;      // 0..255 over both epilogue groups
    if (EPI == 3) {
      // ---- LSTM cell epilogue (BN == 256): a tile holds the i, f, g, o pre-activations of 64 hidden units for 128 rows;
      // each thread owns one row: c = sig(f) c_prev[parent] + sig(i) tanh(g), h = sig(o) tanh(c) (nn.LSTM gate order).
      if (BN == 256) {
        const int row_l = ew * 32 + lane;
        for (int it = 0, tile; (tile = tile_of(it)) < tiles; ++it) {
          int layer = 0, rt = tile;
          if (p.layers > 1) { layer = tile / p.tiles_per_layer; rt = tile - layer * p.tiles_per_layer; }
          const int m0 = tile_m0(rt), n0 = ((rt % p.n_blocks) * p.n_stride + p.n_offset) * BN;
          const bool stack = p.layers > 1;
          const float* bias = stack ? p.bias_l[layer] : p.bias;
          const float* c_prev = stack ? p.cprev_l[layer] : p.c_prev;
          float* c_out = stack ? p.cout_l[layer] : p.c_out;
          void* h0 = stack ? p.h0_l[layer] : p.h0;
          void* h1 = stack ? p.h1_l[layer] : p.h1;
          const long long ldh0 = stack ? p.ldh0_l[layer] : p.ldh0, ldh1 = stack ? p.ldh1_l[layer] : p.ldh1;
          const int as = it & 1;
          // Everything that does not depend on the accumulator is issued BEFORE the wait on it: the bias slice, the
          // beam-parent index and the previous cell values.  Global traffic is COALESCED through a 4 KB per-warp staging
          // area: a thread owns an accumulator row (TMEM lane), so row-wise loads / stores would touch 32 lines per
          // instruction (ncu: the epilogue, not the tensor pipe, bounded the kernel); here lane l moves 16 B of row
          // i * 8 + l / 4, so an instruction covers 8 rows x 64 contiguous bytes, and the row owner exchanges its 64 B
          // with the staging area (XOR-swizzled: both access patterns are bank-conflict free).
          asm volatile("bar.sync 1, 256;" ::: "memory");          // readers of the previous bias slice are done
          for (int i = etid; i < BN; i += 256) bias_s[i] = bias ? __ldg(bias + n0 + i) : 0.f;
          const long long row = (long long)m0 + row_l;
          const bool row_ok = row < p.M;
          const int unit0 = (n0 >> 8) * 64 + eh * 32;             // first of this thread's 32 hidden units
          const long long prow = row_ok ? (p.parent ? (long long)__ldg(p.parent + row) : row) : 0;
          const int cr = lane >> 2, cc = lane & 3;                // cooperative role: row i * 8 + cr, 16-byte chunk cc
          const uint32_t wst = base + C::kRingBytes + (uint32_t)(warp - 4) * 4096u;
          const uint32_t cbuf = wst, hbuf = wst + 2048u;          // [32 rows][64 B] each
          auto swz = [](int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); };
          float4 cpv[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int rr = (g & 3) * 8 + cr;
            const long long pr = __shfl_sync(0xffffffffu, prow, rr);
            const bool ok = (long long)m0 + ew * 32 + rr < p.M;
            cpv[g] = (ok && c_prev) ? __ldg(reinterpret_cast<const float4*>(c_prev + pr * p.H + unit0 + (g >> 2) * 16 + cc * 4))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mbar_wait(tfull_bar(as), (it >> 1) & 1, p.error, 4);
          tc_fence_after();
          const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int u0 = eh * 32 + half * 16;
            uint32_t vi[16], vf[16], vg[16], vo[16];
            tc_ld16(tmem_row + (uint32_t)u0, vi);
            tc_ld16(tmem_row + (uint32_t)(64 + u0), vf);
            tc_ld16(tmem_row + (uint32_t)(128 + u0), vg);
            tc_ld16(tmem_row + (uint32_t)(192 + u0), vo);
            __syncwarp();                                         // the previous half's read-out of cbuf is complete
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 c4 = cpv[half * 4 + i];
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cbuf + swz(i * 8 + cr, cc)), "f"(c4.x), "f"(c4.y),
                           "f"(c4.z), "f"(c4.w) : "memory");
            }
            __syncwarp();
            float cprev[16];
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(cprev[4 * g]), "=f"(cprev[4 * g + 1]), "=f"(cprev[4 * g + 2]), "=f"(cprev[4 * g + 3])
                           : "r"(cbuf + swz(lane, g)) : "memory");
            tc_wait_ld();
            if (half == 1) {
              // the accumulator now lives in registers: hand the TMEM buffer back to the MMA warp before the cell math
              // and the global stores
              tc_fence_before();
              __syncwarp();
              if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_u32(tempty_bar(as), lead)); else mbar_arrive(tempty_bar(as)); }
            }
            float c2[16];
            uint32_t hw[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float h2[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float gi = __uint_as_float(vi[j + e]) + bias_s[u0 + j + e];
                const float gf = __uint_as_float(vf[j + e]) + bias_s[64 + u0 + j + e];
                const float gg = __uint_as_float(vg[j + e]) + bias_s[128 + u0 + j + e];
                const float go = __uint_as_float(vo[j + e]) + bias_s[192 + u0 + j + e];
                // MUFU.TANH (rel. error 2^-11, far inside the bf16 rounding of h): sigmoid(x) = 0.5 + 0.5 tanh(x / 2)
                const float si = fmaf(0.5f, tanh_fast(0.5f * gi), 0.5f), sf = fmaf(0.5f, tanh_fast(0.5f * gf), 0.5f);
                const float so = fmaf(0.5f, tanh_fast(0.5f * go), 0.5f);
                c2[j + e] = fmaf(sf, cprev[j + e], si * tanh_fast(gg));
                h2[e] = so * tanh_fast(c2[j + e]);
              }
              if (p.out_dtype == DH_BF16) {
                __nv_bfloat162 t = __floats2bfloat162_rn(h2[0], h2[1]);
                hw[j >> 1] = *reinterpret_cast<uint32_t*>(&t);
              } else {
                __half2 t = __floats2half2_rn(h2[0], h2[1]);
                hw[j >> 1] = *reinterpret_cast<uint32_t*>(&t);
              }
            }
            // the owner's new cell values replace the old ones in its own row of cbuf; its 16 h values (32 B) go to hbuf
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cbuf + swz(lane, g)), "f"(c2[4 * g]),
                           "f"(c2[4 * g + 1]), "f"(c2[4 * g + 2]), "f"(c2[4 * g + 3]) : "memory");
#pragma unroll
            for (int g = 0; g < 2; ++g)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hbuf + swz(lane, half * 2 + g)), "r"(hw[4 * g]),
                           "r"(hw[4 * g + 1]), "r"(hw[4 * g + 2]), "r"(hw[4 * g + 3]) : "memory");
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = i * 8 + cr;
              const long long grow = (long long)m0 + ew * 32 + rr;
              float4 v;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                           : "r"(cbuf + swz(rr, cc)) : "memory");
              if (grow < p.M) *reinterpret_cast<float4*>(c_out + grow * p.H + unit0 + half * 16 + cc * 4) = v;
            }
          }
          // h of both halves: 64 B per row, again 8 rows per store instruction, to up to two destinations
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + cr;
            const long long grow = (long long)m0 + ew * 32 + rr;
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(hbuf + swz(rr, cc)) : "memory");
            if (grow < p.M) {
              if (h0) *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(h0) + grow * ldh0 + unit0 + cc * 8) = v;
              if (h1) *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(h1) + grow * ldh1 + unit0 + cc * 8) = v;
            }
          }
          if (stack && layer + 1 < p.layers) {
            // publish this tile's h rows to the layer above: the CTA barrier orders every epilogue thread's stores before
            // thread 0, whose GPU-scope fence + release increment is cumulative over them (the pattern of a cooperative grid
            // sync); a fence per thread made all 256 threads wait for their stores to drain (MEMBAR.GPU: 4 % of the kernel's
            // samples).  The consumer orders its async-proxy (TMA) reads after its acquire (wait_ready).
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (etid == 0 && m0 < p.M) {
              __threadfence();
              fence_proxy_async_all();
              red_release_gpu(p.ready + layer * p.mb128 + (m0 >> 7), 1);
            }
          }
        }
      }
    } else if (EPI == 6) {
      // ---- LayerNorm split over a cluster (x = LN(x + sublayer(x)), transformers.py:355-356,365-366,374-375): the two CTAs
      // of a cluster compute the two 256-column halves of the SAME 128-row block (plain tile order, n fastest, even grid), so
      // each needs only 256 tensor-memory columns per block and the accumulator is double-buffered -- the MMAs of block i + 1
      // run under this epilogue (with the whole 512-wide row on one CTA, EPI 5, nothing overlaps: 55 us at 40 960 rows x
      // K = 512 against 15 us of tensor work).  Each epilogue group owns 128 columns: pass 1 gives their (mean, M2), which goes
      // to this CTA's shared memory (plain store + arrive) AND the peer's (st.async: data and barrier transaction in one),
      // and the four partials of a row are merged pairwise (equal counts); pass 2 normalises from tensor memory.
      constexpr int GW = BN / 2;                                 // columns per epilogue group
      const int row_l = ew * 32 + lane;
      float* lnp = reinterpret_cast<float*>(gen_base + C::kLnOff);          // [3][2 BN]: bias | gamma | beta of the whole row
      float* part = lnp + 3 * 2 * BN;                                       // [2 buffers][4 quarters][BM][2]
      const uint32_t part_u = base + C::kLnOff + 3 * 2 * BN * 4;
      const uint32_t sbar = part_u + 2 * 4 * BM * 2 * 4;
      // column half of this CTA (pair) and the CTA that holds the same rows of the other half
      const uint32_t crank = PAIR ? cluster_ctarank() >> 1 : cluster_ctarank();
      const uint32_t peer = cluster_ctarank() ^ (PAIR ? 2u : 1u);
      const uint32_t slab = base + C::kRingBytes + (uint32_t)eh * (BM * 128);
      const uint32_t srow = slab + (uint32_t)row_l * 128u;
      const uint32_t swz = (uint32_t)(row_l & 7);
      const bool elected = (ew == 0 && lane == 0);
      const int bar_half = 2 + eh;
      for (int i = etid; i < 2 * BN; i += 256) {
        lnp[i] = p.bias ? __ldg(p.bias + i) : 0.f;
        lnp[2 * BN + i] = __ldg(p.ln_gamma + i);
        lnp[4 * BN + i] = __ldg(p.ln_beta + i);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int quarter = (int)crank * 2 + eh;                   // which 128 columns of the row this thread reduces
      const int nq = quarter * GW;
      const float2* bias2 = reinterpret_cast<const float2*>(lnp + nq);
      const float2* gam2 = reinterpret_cast<const float2*>(lnp + 2 * BN + nq);
      const float2* bet2 = reinterpret_cast<const float2*>(lnp + 4 * BN + nq);
      bool slab_busy = false;
      for (int it = 0, tile; (tile = tile_of(it)) < tiles; ++it) {
        const int m0 = tile_m0(tile);
        const int as = it & 1;
        mbar_wait(tfull_bar(as), (it >> 1) & 1, p.error, 4);
        tc_fence_after();
        const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN + eh * GW);
        f32x2 s_a = pk2(0.f, 0.f), s_b = s_a, q_a = s_a, q_b = s_a, negp = s_a;
        float pivot = 0.f;
#pragma unroll 1
        for (int c = 0; c < GW / 64; ++c) {
          uint32_t v[64];
          tc_ld32_nw(tmem_row + (uint32_t)(c * 64), v);
          tc_ld32_nw(tmem_row + (uint32_t)(c * 64 + 32), v + 32);
          tc_wait_ld();
          if (c == 0) {
            pivot = __uint_as_float(v[0]) + bias2[0].x;
            negp = pk2(-pivot, -pivot);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 b0 = bias2[c * 32 + j], b1 = bias2[c * 32 + j + 1];
            const f32x2 d0 = add2(add2(pk2u(v[2 * j], v[2 * j + 1]), pk2(b0.x, b0.y)), negp);
            const f32x2 d1 = add2(add2(pk2u(v[2 * j + 2], v[2 * j + 3]), pk2(b1.x, b1.y)), negp);
            s_a = add2(s_a, d0); q_a = fma2(d0, d0, q_a);
            s_b = add2(s_b, d1); q_b = fma2(d1, d1, q_b);
          }
        }
        float s0, s1, q0, q1;
        un2(add2(s_a, s_b), s0, s1);
        un2(add2(q_a, q_b), q0, q1);
        const float sh = s0 + s1, qh = q0 + q1;
        const float mean_q = pivot + sh * (1.f / (float)GW);               // mean of this quarter
        const float m2_q = qh - sh * sh * (1.f / (float)GW);               // sum of squared deviations about it
        const uint32_t slot = (uint32_t)((((it & 1) * 4 + quarter) * BM + row_l) * 8);
        const uint32_t sb_it = sbar + 8u * (uint32_t)(it & 1);
        *reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(part) + slot) = make_float2(mean_q, m2_q);
        st_async_f32x2(mapa_u32(part_u + slot, peer), mean_q, m2_q, mapa_u32(sb_it, peer));
        if (etid == 0) mbar_expect_tx(sb_it, 256 * 8); else mbar_arrive(sb_it);
        mbar_wait_cluster(sb_it, (uint32_t)((it >> 1) & 1), p.error, 7);
        const float2* pp = reinterpret_cast<const float2*>(part) + (size_t)(it & 1) * 4 * BM + row_l;
        const float2 p0 = pp[0], p1 = pp[BM], p2 = pp[2 * BM], p3 = pp[3 * BM];
        const float mean = 0.25f * ((p0.x + p1.x) + (p2.x + p3.x));
        const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
        const float var = ((p0.y + p1.y) + (p2.y + p3.y) + (float)GW * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3))) * (1.f / (float)(4 * GW));
        const float rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
        const f32x2 rr = pk2(rstd, rstd), negm = pk2(-mean, -mean);
#pragma unroll 1
        for (int rd = 0; rd < GW / 64; ++rd) {
          // one slab per group: the previous round's TMA store must have read it before it is rewritten
          if (elected && slab_busy) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          uint32_t v[64];
          tc_ld32_nw(tmem_row + (uint32_t)(rd * 64), v);
          tc_ld32_nw(tmem_row + (uint32_t)(rd * 64 + 32), v + 32);
          tc_wait_ld();
          if (rd == GW / 64 - 1) {
            // the accumulator quarter now lives in registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_u32(tempty_bar(as), lead)); else mbar_arrive(tempty_bar(as)); }
          }
          uint32_t w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 b0 = bias2[rd * 32 + j], g0 = gam2[rd * 32 + j], e0 = bet2[rd * 32 + j];
            const f32x2 a = mul2(rr, pk2(g0.x, g0.y));
            const f32x2 k = fma2(add2(pk2(b0.x, b0.y), negm), a, pk2(e0.x, e0.y));
            float y0, y1;
            un2(fma2(pk2u(v[2 * j], v[2 * j + 1]), a, k), y0, y1);
            if (p.out_dtype == DH_BF16) {
              __nv_bfloat162 t = __floats2bfloat162_rn(y0, y1);
              w[j] = *reinterpret_cast<uint32_t*>(&t);
            } else {
              __half2 t = __floats2half2_rn(y0, y1);
              w[j] = *reinterpret_cast<uint32_t*>(&t);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_half) : "memory");      // the slab is free (elected waited above)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((uint32_t)j ^ swz) << 4)),
                         "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                         : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(bar_half) : "memory");
          if (elected) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&map_c)),
                         "r"(slab), "r"(nq + rd * 64), "r"(m0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          slab_busy = true;
        }
      }
      if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (EPI == 5) {
      // ---- LayerNorm epilogue (transformers.py:355-356,365-366,374-375: x = LN(x + sublayer(x))): accumulator buffer `eh`
      // holds columns [eh * BN, (eh + 1) * BN) of the row block (bias still to add; the residual already rode the tensor
      // core).  A thread owns one row of its half and makes TWO passes over tensor memory, 64 columns per wait, in packed
      // fp32 pairs (FADD2 / FFMA2): (1) sum and squared deviations about a pivot (the half's first element), turned into the
      // half's (mean, M2) and merged with the other half's by the pairwise-variance formula -- as accurate as a two-pass
      // variance; (2) normalise, scale, shift, round, slab, TMA store.  bias / gamma / beta sit in shared memory.
      const int row_l = ew * 32 + lane;
      float* lnp = reinterpret_cast<float*>(gen_base + C::kLnOff);          // [3][2 BN]: bias | gamma | beta
      float* part = lnp + 3 * 2 * BN;                                       // [2 buffers][2 halves][BM][2]
      const uint32_t slab = base + C::kRingBytes + (uint32_t)eh * (BM * 128);
      const uint32_t srow = slab + (uint32_t)row_l * 128u;
      const uint32_t swz = (uint32_t)(row_l & 7);
      const bool elected = (ew == 0 && lane == 0);
      const int bar_half = 2 + eh;                                // named barrier of this half's 128 threads
      const int nh = eh * BN;                                     // first column of this half
      for (int i = etid; i < 2 * BN; i += 256) {
        lnp[i] = p.bias ? __ldg(p.bias + i) : 0.f;
        lnp[2 * BN + i] = __ldg(p.ln_gamma + i);
        lnp[4 * BN + i] = __ldg(p.ln_beta + i);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2* bias2 = reinterpret_cast<const float2*>(lnp + nh);
      const float2* gam2 = reinterpret_cast<const float2*>(lnp + 2 * BN + nh);
      const float2* bet2 = reinterpret_cast<const float2*>(lnp + 4 * BN + nh);
      for (int i = 0, tile; (tile = tile_of(2 * i + eh)) < tiles; ++i) {
        const int m0 = tile_m0(tile);
        mbar_wait(tfull_bar(eh), i & 1, p.error, 4);
        tc_fence_after();
        const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(eh * BN);
        f32x2 s_a = pk2(0.f, 0.f), s_b = s_a, q_a = s_a, q_b = s_a, negp = s_a;
        float pivot = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          uint32_t v[64];
          tc_ld32_nw(tmem_row + (uint32_t)(c * 64), v);
          tc_ld32_nw(tmem_row + (uint32_t)(c * 64 + 32), v + 32);
          tc_wait_ld();
          if (c == 0) {
            pivot = __uint_as_float(v[0]) + bias2[0].x;
            negp = pk2(-pivot, -pivot);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 b0 = bias2[c * 32 + j], b1 = bias2[c * 32 + j + 1];
            const f32x2 d0 = add2(add2(pk2u(v[2 * j], v[2 * j + 1]), pk2(b0.x, b0.y)), negp);
            const f32x2 d1 = add2(add2(pk2u(v[2 * j + 2], v[2 * j + 3]), pk2(b1.x, b1.y)), negp);
            s_a = add2(s_a, d0); q_a = fma2(d0, d0, q_a);
            s_b = add2(s_b, d1); q_b = fma2(d1, d1, q_b);
          }
        }
        float s0, s1, q0, q1;
        un2(add2(s_a, s_b), s0, s1);
        un2(add2(q_a, q_b), q0, q1);
        const float sh = s0 + s1, qh = q0 + q1;
        float* pp = part + (size_t)(i & 1) * (2 * BM * 2);
        pp[(eh * BM + row_l) * 2] = pivot + sh * (1.f / (float)BN);                 // mean of this half
        pp[(eh * BM + row_l) * 2 + 1] = qh - sh * sh * (1.f / (float)BN);           // sum of squared deviations about it
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float m_0 = pp[row_l * 2], M_0 = pp[row_l * 2 + 1], m_1 = pp[(BM + row_l) * 2], M_1 = pp[(BM + row_l) * 2 + 1];
        const float mean = 0.5f * (m_0 + m_1);
        const float var = (M_0 + M_1 + (m_0 - m_1) * (m_0 - m_1) * (0.5f * (float)BN)) * (1.f / (float)(2 * BN));
        const float rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
        const f32x2 rr = pk2(rstd, rstd), negm = pk2(-mean, -mean);
#pragma unroll 1
        for (int rd = 0; rd < BN / 64; ++rd) {
          // one slab per half: the previous round's TMA store must have read it before it is rewritten
          if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          uint32_t v[64];
          tc_ld32_nw(tmem_row + (uint32_t)(rd * 64), v);
          tc_ld32_nw(tmem_row + (uint32_t)(rd * 64 + 32), v + 32);
          tc_wait_ld();
          if (rd == BN / 64 - 1) {
            // the accumulator half now lives in registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_u32(tempty_bar(eh), lead)); else mbar_arrive(tempty_bar(eh)); }
          }
          uint32_t w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 b0 = bias2[rd * 32 + j], g0 = gam2[rd * 32 + j], e0 = bet2[rd * 32 + j];
            const f32x2 a = mul2(rr, pk2(g0.x, g0.y));
            const f32x2 k = fma2(add2(pk2(b0.x, b0.y), negm), a, pk2(e0.x, e0.y));
            float y0, y1;
            un2(fma2(pk2u(v[2 * j], v[2 * j + 1]), a, k), y0, y1);
            if (p.out_dtype == DH_BF16) {
              __nv_bfloat162 t = __floats2bfloat162_rn(y0, y1);
              w[j] = *reinterpret_cast<uint32_t*>(&t);
            } else {
              __half2 t = __floats2half2_rn(y0, y1);
              w[j] = *reinterpret_cast<uint32_t*>(&t);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_half) : "memory");      // the slab is free (elected waited above)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((uint32_t)j ^ swz) << 4)),
                         "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                         : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(bar_half) : "memory");
          if (elected) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&map_c)),
                         "r"(slab), "r"(nh + rd * 64), "r"(m0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (EPI) {
      // ---- selection epilogues: every thread owns one accumulator row; nothing of the [M,N] product is stored.
      // The epilogue, not the tensor pipe, paced these kernels (ncu: 74 % tensor-active at K = 512, no time spent waiting for
      // the accumulator), so everything that is not the scan itself is taken off the per-tile path: the bias slice of the NEXT
      // tile is fetched into a register while this tile is scanned and published through a double-buffered shared-memory
      // slice (one barrier per tile), the row's threshold / target are re-read only when the row block changes, and tensor
      // memory is read 64 columns per wait.
      const int row_l = ew * 32 + lane;
      float* bias_buf[2] = {bias_s, bias_s + BN};
      auto tile_n0 = [&](int tile) { return ((tile % p.n_blocks) * p.n_stride + p.n_offset) * BN; };
      {
        const int t0i = tile_of(0);
        if (t0i < tiles && etid < BN) {
          const int n0 = tile_n0(t0i);
          bias_buf[0][etid] = (p.bias && n0 + etid < p.N) ? __ldg(p.bias + n0 + etid) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      int cur_m0 = -1;
      bool row_ok = false, emit = false;
      long long row = 0;
      float t0 = INFINITY;
      int tcol = -1;
      for (int it = 0, tile; (tile = tile_of(it)) < tiles; ++it) {
        const int m0 = tile_m0(tile), n0 = tile_n0(tile);
        const int as = it & 1;
        const float* bias_cur = bias_buf[it & 1];
        float bias_next = 0.f;
        {
          const int nt = tile_of(it + 1);
          if (nt < tiles && etid < BN) {
            const int nn0 = tile_n0(nt);
            bias_next = (p.bias && nn0 + etid < p.N) ? __ldg(p.bias + nn0 + etid) : 0.f;
          }
        }
        if (m0 != cur_m0) {
          // per-row operands: issued before the wait on the accumulator so that their L2 round trip is hidden
          cur_m0 = m0;
          row = (long long)m0 + row_l;
          row_ok = row < p.M;
          emit = EPI == 2 && row_ok && (!p.redo || __ldg(p.redo + row));
          // candidates are finite logits >= thresh[row]; clamping to -FLT_MAX folds the "> -inf" test into one compare
          t0 = emit ? fmaxf(__ldg(p.thresh + row), -3.402823466e+38f) : INFINITY;
          tcol = (EPI == 4 && row_ok) ? (int)__ldg(p.targets + row) : -1;
        }
        mbar_wait(tfull_bar(as), (it >> 1) & 1, p.error, 4);
        tc_fence_after();
        const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
        unsigned int bits = 0u;                                   // groups of this (row, tile, half) that hold a candidate
        constexpr int kGroups = BN / 64;                          // 32-column groups per thread
#pragma unroll 1
        for (int c2 = 0; c2 < kGroups; c2 += 2) {
          uint32_t vv[2][32];
          const int cbase = eh * kGroups + c2;
          const bool two = c2 + 1 < kGroups;
          tc_ld32_nw(tmem_row + (uint32_t)(cbase * 32), vv[0]);
          if (two) tc_ld32_nw(tmem_row + (uint32_t)((cbase + 1) * 32), vv[1]);
          tc_wait_ld();
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const int c = cbase + u;
            const int col0 = n0 + c * 32;
            if (col0 >= p.N) {                                      // warp-uniform: group past the end of the row
              if (EPI != 2 && row_ok) p.gmax[row * p.ld_gmax + (tile % p.n_blocks) * (BN / 32) + c] = -INFINITY;
              if (EPI == 4 && row_ok) p.gsum[row * p.ld_gmax + (tile % p.n_blocks) * (BN / 32) + c] = 0.f;
              continue;
            }
            const uint32_t* v = vv[u];
            float x[32];
            const float4* bs = reinterpret_cast<const float4*>(bias_cur + c * 32);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b4 = bs[g];
              x[4 * g] = __uint_as_float(v[4 * g]) + b4.x;
              x[4 * g + 1] = __uint_as_float(v[4 * g + 1]) + b4.y;
              x[4 * g + 2] = __uint_as_float(v[4 * g + 2]) + b4.z;
              x[4 * g + 3] = __uint_as_float(v[4 * g + 3]) + b4.w;
            }
            const int nv = p.N - col0;                              // valid columns in this group (>= 1)
            if (nv < 32) {                                          // last group of the row: mask the padding columns
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = j < nv ? x[j] : -INFINITY;
            }
            float m8[4];                                            // maxima of the four 8-column quarters
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              m8[q] = x[8 * q];
#pragma unroll
              for (int j = 1; j < 8; ++j) m8[q] = fmaxf(m8[q], x[8 * q + j]);
            }
            const float mx = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
            if (EPI == 1) {
              if (row_ok) p.gmax[row * p.ld_gmax + (tile % p.n_blocks) * (BN / 32) + c] = mx;
            } else if (EPI == 4) {
              // log-softmax pieces of this group (experiments/metrics.py:5): max, sum of exp(x - max), and the target's logit
              float se = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) se += __expf(x[j] - mx);          // padding columns hold -inf -> 0
              if (row_ok) {
                const long long g = row * p.ld_gmax + (tile % p.n_blocks) * (BN / 32) + c;
                p.gmax[g] = mx;
                p.gsum[g] = se;
                const int tj = tcol - col0;
                if (tj >= 0 && tj < 32) {
                  float tl = 0.f;
#pragma unroll
                  for (int j = 0; j < 32; ++j) tl = (j == tj) ? x[j] : tl;
                  p.tlogit[row] = tl;
                }
              }
            } else if (mx >= t0) {
              // the group holds at least one logit >= thresh[row]: store its 32 logits (one full 128 B line of the row) and
              // leave the element-wise work to the selection kernel -- the epilogue's instruction count does not depend on
              // where the candidates sit (per-lane divergent element scans made unrelated rows 2x slower than collinear ones)
              bits |= 1u << (c - eh * kGroups);
              float4* dst = reinterpret_cast<float4*>(p.sp_logits + row * p.sp_ld + col0);
#pragma unroll
              for (int g = 0; g < 8; ++g) dst[g] = make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
            }
          }
        }
        // the accumulator is drained: hand the TMEM buffer back to the MMA warp BEFORE any list traffic
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_u32(tempty_bar(as), lead)); else mbar_arrive(tempty_bar(as)); }
        if (EPI == 2 && emit) {
          p.hitmap[row * p.hit_ld + (long long)(((tile % p.n_blocks) * p.n_stride + p.n_offset) * 2 + eh)] = (unsigned char)bits;
          if (bits) atomicAdd(p.cand_count + row, __popc(bits));           // result unused: compiles to a fire-and-forget RED
        }
        // publish the next tile's bias slice; the barrier also keeps this tile's slice alive until every reader is done
        if (etid < BN) bias_buf[(it + 1) & 1][etid] = bias_next;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    } else if (G2 == 3) {
      // ---- chained slab epilogue (dh_conv1x1_chain_tc): the slab rounds of the P tiles (BN columns -> map_c) and of the Q
      // tiles (N2 columns -> cm.c2) alternate between the two epilogue groups by a running round count; before a group starts
      // the virtual tile after a P tile its elected thread waits for the group's bulk stores to COMPLETE and arrives on that
      // tile's "stored" barrier, which the producer needs before it reads the tile back for Q.
      const uint32_t slab = base + C::kRingBytes + (uint32_t)eh * (BM * 128);
      const int bar_id = 1 + eh;
      const int row_l = ew * 32 + lane;
      const uint32_t swz = (uint32_t)(row_l & 7);
      const uint32_t srow = slab + (uint32_t)row_l * 128u;
      const bool elected = (ew == 0 && lane == 0);
      float* bias_g = bias_s + eh * BN;                         // this group's copy of the current (sub-)tile's bias slice
      uint32_t rglob = 0;
      bool slab_busy = false, prev_p = false;
      int prev_j = 0, vc = 0, bias_key = -1;
      for (int v = 0; v < 2 * n_own; ++v) {
        bool isq;
        const int j = chain_vt(v, isq);
        const int m0 = (tile0 + j * tstride) * p.bm_rows;
        for (int sub = 0; sub < (isq ? 1 : p.n_blocks); ++sub, ++vc) {
          const int as = vc & 1;
          if (prev_p && elected) {
            // the unit before this (sub-)tile ended a P part: its rows must be complete in L2 before the producer reads them
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            mbar_arrive(stored_bar0 + 8u * (uint32_t)(prev_j & 3));
            slab_busy = false;
          }
          prev_p = !isq && sub == p.n_blocks - 1;
          prev_j = j;
          const int n0 = isq ? 0 : sub * BN;
          const int ncols = isq ? p.N2 : BN;
          const int key = isq ? p.n_blocks : sub;
          if (key != bias_key) {     // (readers of the previous slice are past their last round barrier)
            const float* src = isq ? p.bias2 : p.bias;
            for (int i = row_l; i < ncols; i += 128) bias_g[i] = src ? __ldg(src + n0 + i) : 0.f;
            bias_key = key;
          }
          mbar_wait(tfull_bar(as), (vc >> 1) & 1, p.error, 4);
          tc_fence_after();
          const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
          const int nr = ncols / 64;
          const bool relu = isq ? p.relu2 : p.relu;
#pragma unroll 1
          for (int rd = 0; rd < nr; ++rd) {
            if (((rglob + rd) & 1u) != (uint32_t)eh) continue;
            if (elected && slab_busy) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            uint32_t vv[64];
            tc_ld32_nw(tmem_row + (uint32_t)(rd * 64), vv);
            tc_ld32_nw(tmem_row + (uint32_t)(rd * 64 + 32), vv + 32);
            tc_wait_ld();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // the slab is free (elected waited above); bias_g is set
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float x[32];
#pragma unroll
              for (int q = 0; q < 32; ++q) x[q] = __uint_as_float(vv[h * 32 + q]);
              const float4* bs = reinterpret_cast<const float4*>(bias_g + rd * 64 + h * 32);   // warp-uniform: broadcast
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b4 = bs[g];
                x[4 * g] += b4.x; x[4 * g + 1] += b4.y; x[4 * g + 2] += b4.z; x[4 * g + 3] += b4.w;
              }
              if (relu) {
#pragma unroll
                for (int q = 0; q < 32; ++q) x[q] = fmaxf(x[q], 0.f);
              }
              uint32_t w[16];
              if (p.out_dtype == DH_BF16) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                  __nv_bfloat162 t = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
                  w[q] = *reinterpret_cast<uint32_t*>(&t);
                }
              } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                  __half2 t = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
                  w[q] = *reinterpret_cast<uint32_t*>(&t);
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((uint32_t)(h * 4 + q) ^ swz) << 4)),
                             "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3])
                             : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (elected) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                               reinterpret_cast<uint64_t>(isq ? &cm.c2 : &map_c)),
                           "r"(slab), "r"(n0 + rd * 64), "r"(m0)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              slab_busy = true;
            }
          }
          rglob += (uint32_t)nr;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
      }
      if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (p.tma_store && (eh == 0 || G2)) {
      // ---- slab epilogue: each thread owns one accumulator row; a round covers 128 B of every row (32 fp32 or
      // 64 half columns), written 128B-swizzled into a 16 KB slab and stored by one TMA instruction.  The two epilogue
      // groups (warps 4-7 / 8-11) take ALTERNATE rounds, each with its own two slabs, named barrier, bias copy and bulk groups:
      // a round is one dependent chain (tensor-memory read -> bias / ReLU / rounding -> slab -> fence -> barrier -> store)
      // and a lone warp per scheduler ran it at 15 % issue (ncu on l1.ds: 6.2 k cycles per 128 x 256 tile, which is what
      // paced the contractions with K <= 768 and no residual traffic -- not HBM, not the tensor pipe); two warps per scheduler
      // interleave two chains (template flag G2 == 2; l1.ds 148 -> 108 us, K = 512 decoder contractions +10 %).  The four
      // slabs cost ring depth, which the HBM-bound K = 64 residual convolutions of layer1 need more (l1.c3: 157 us, against
      // 143 us with G2 == 1: both groups, ONE slab each, full ring).  Long K loops hide the epilogue anyway and ran 4-6 %
      // slower with both groups busy: they take G2 == 0, where the first group runs every round over two slabs and the second
      // leaves at once (a warp spinning on the accumulator barrier costs the working warps issue slots).
      constexpr int SPG = G2 == 1 ? 1 : 2;        // slabs per group
      const uint32_t slabs = base + C::kRingBytes + (uint32_t)eh * (SPG * BM * 128);
      float* bias_g = bias_s + (G2 ? eh * BN : 0);
      const int bar_id = 1 + eh;
      const bool out32 = p.out_dtype == DH_F32;
      const int cpr = out32 ? 32 : 64;
      const int rounds = BN / cpr;                // nominal rounds per tile: sets the group that takes round (it, rd)
      const int row_l = ew * 32 + lane;
      const uint32_t swz = (uint32_t)(row_l & 7);
      const bool elected = (ew == 0 && lane == 0);
      uint32_t round_ctr = 0;
      int last_n0 = -1;
      for (int it = 0, tile; (tile = tile_of(it)) < tiles; ++it) {
        const int m0 = tile_m0(tile), n0 = ((tile % p.n_blocks) * p.n_stride + p.n_offset) * BN;
        const int as = it & 1;
        // bias slice of this tile -> smem, only when the N block changed, and before the wait on the accumulator so the
        // global-load latency is off the per-tile critical path (all readers of the previous slice are past their last
        // round barrier)
        if (p.bias && n0 != last_n0) {
          for (int i = row_l; i < BN; i += 128) bias_g[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
          last_n0 = n0;
        }
        mbar_wait(tfull_bar(as), (it >> 1) & 1, p.error, 4);
        tc_fence_after();
        const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
        for (int rd = 0; rd < BN / 32; ++rd) {
          const int col0 = n0 + rd * cpr;
          if (rd * cpr >= BN || col0 >= p.N) break;
          if (G2 && ((it * rounds + rd) & 1) != eh) continue;
          // the slab's previous TMA store (SPG rounds of this group ago) must have read it before it is rewritten
          const uint32_t slab = slabs + (SPG == 2 ? (round_ctr & 1u) : 0u) * (BM * 128);
          if (elected && round_ctr >= SPG) {
            if (SPG == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          uint32_t v[64];
          tc_ld32_nw(tmem_row + (uint32_t)(rd * cpr), v);
          if (!out32) tc_ld32_nw(tmem_row + (uint32_t)(rd * cpr + 32), v + 32);
          tc_wait_ld();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // the slab is free (elected waited above); bias_g is set
          const uint32_t srow = slab + (uint32_t)row_l * 128u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (out32 && h == 1) break;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[h * 32 + j]);
            if (p.bias) {
              const float4* bs = reinterpret_cast<const float4*>(bias_g + rd * cpr + h * 32);   // warp-uniform: broadcast
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b4 = bs[g];
                x[4 * g] += b4.x; x[4 * g + 1] += b4.y; x[4 * g + 2] += b4.z; x[4 * g + 3] += b4.w;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
            }
            if (out32) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((uint32_t)j ^ swz) << 4)),
                             "f"(x[4 * j]), "f"(x[4 * j + 1]), "f"(x[4 * j + 2]), "f"(x[4 * j + 3])
                             : "memory");
            } else {
              uint32_t w[16];
              if (p.out_dtype == DH_BF16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  __nv_bfloat162 t = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
                  w[j] = *reinterpret_cast<uint32_t*>(&t);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  __half2 t = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                  w[j] = *reinterpret_cast<uint32_t*>(&t);
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((uint32_t)(h * 4 + j) ^ swz) << 4)),
                             "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                             : "memory");
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (elected) {
            // split destinations (fused Q | K | V projection): the N tile lies entirely inside one of them
            const CUtensorMap* mc = &map_c;
            int ccol = col0;
            if (p.split_n) {
              const int w = n0 / p.split_n;
              mc = w == 0 ? &map_c : w == 1 ? &map_r : &map_i;
              ccol = col0 - w * p.split_n;
            }
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(mc)),
                         "r"(slab), "r"(ccol), "r"(m0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (p.pool_out) {
            // fused global average pool: the slab holds this round's 64 columns of the tile's whole images exactly as they
            // are stored (already rounded to the output type); one thread per (image, column) adds the image's rows in
            // ascending order -- the summation of dh_avgpool, bit for bit -- while the TMA store drains the same slab
            const int hw = p.pool_hw, gimg = p.bm_rows / hw;
            for (int q = row_l; q < gimg * 64; q += 128) {
              const int i = q >> 6, c = q & 63;
              const long long img = (long long)(m0 / hw) + i;
              if ((img + 1) * hw <= p.M && col0 + c < p.N) {
                float s = 0.f;
                for (int r = i * hw; r < (i + 1) * hw; ++r) {
                  uint16_t raw;
                  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(raw)
                               : "r"(slab + (uint32_t)r * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)c & 7u) * 2u));
                  s += p.out_dtype == DH_BF16 ? __bfloat162float(__ushort_as_bfloat16(raw)) : __half2float(__ushort_as_half(raw));
                }
                p.pool_out[img * p.ld_pool + col0 + c] = s / (float)hw;
              }
            }
          }
          ++round_ctr;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_u32(tempty_bar(as), lead)); else mbar_arrive(tempty_bar(as)); }
      }
      if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (eh != 0) {
      // second epilogue group: idle in the one-group forms
    } else {
    float* st = staging + ew * 32 * kStageLd;
    const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const bool res_vec = p.res && (p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0);
    for (int it = 0, tile; (tile = tile_of(it)) < tiles; ++it) {
      const int m0 = tile_m0(tile), n0 = ((tile % p.n_blocks) * p.n_stride + p.n_offset) * BN;
      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1, p.error, 4);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (n0 + c * 32 >= p.N) break;
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN + c * 32), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + lane * kStageLd + j * 4) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        const int col = n0 + c * 32 + (lane & 7) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + (lane >> 3);
          const long long row = (long long)m0 + ew * 32 + rl;
          if (row >= p.M || col >= p.N) continue;
          float4 a = *reinterpret_cast<const float4*>(st + rl * kStageLd + (lane & 7) * 4);
          float x[4] = {a.x, a.y, a.z, a.w};
          const int nv = min(4, p.N - col);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (e < nv && p.bias) x[e] += __ldg(p.bias + col + e);
          if (p.res) {
            if (p.res_dtype == DH_F32) {
              const float* rp = reinterpret_cast<const float*>(p.res) + row * p.ldr + col;
              if (res_vec && nv == 4) {
                float4 r4 = *reinterpret_cast<const float4*>(rp);
                x[0] += r4.x; x[1] += r4.y; x[2] += r4.z; x[3] += r4.w;
              } else {
                for (int e = 0; e < nv; ++e) x[e] += rp[e];
              }
            } else if (p.res_dtype == DH_F16) {
              const __half* rp = reinterpret_cast<const __half*>(p.res) + row * p.ldr + col;
              if (res_vec && nv == 4) {
                uint2 r2 = *reinterpret_cast<const uint2*>(rp);
                __half2 lo = *reinterpret_cast<__half2*>(&r2.x), hi = *reinterpret_cast<__half2*>(&r2.y);
                x[0] += __low2float(lo); x[1] += __high2float(lo); x[2] += __low2float(hi); x[3] += __high2float(hi);
              } else {
                for (int e = 0; e < nv; ++e) x[e] += __half2float(rp[e]);
              }
            } else {
              const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + row * p.ldr + col;
              if (res_vec && nv == 4) {
                uint2 r2 = *reinterpret_cast<const uint2*>(rp);
                __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&r2.x), hi = *reinterpret_cast<__nv_bfloat162*>(&r2.y);
                x[0] += __low2float(lo); x[1] += __high2float(lo); x[2] += __low2float(hi); x[3] += __high2float(hi);
              } else {
                for (int e = 0; e < nv; ++e) x[e] += __bfloat162float(rp[e]);
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
          }
          if (p.out_dtype == DH_F32) {
            float* op = reinterpret_cast<float*>(p.out) + row * p.ldc + col;
            if (vec_ok && nv == 4) *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
            else for (int e = 0; e < nv; ++e) op[e] = x[e];
          } else if (p.out_dtype == DH_F16) {
            __half* op = reinterpret_cast<__half*>(p.out) + row * p.ldc + col;
            if (vec_ok && nv == 4) {
              __half2 lo = __floats2half2_rn(x[0], x[1]), hi = __floats2half2_rn(x[2], x[3]);
              uint2 o;
              o.x = *reinterpret_cast<uint32_t*>(&lo);
              o.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(op) = o;
            } else {
              for (int e = 0; e < nv; ++e) op[e] = __float2half_rn(x[e]);
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + col;
            if (vec_ok && nv == 4) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]), hi = __floats2bfloat162_rn(x[2], x[3]);
              uint2 o;
              o.x = *reinterpret_cast<uint32_t*>(&lo);
              o.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(op) = o;
            } else {
              for (int e = 0; e < nv; ++e) op[e] = __float2bfloat16_rn(x[e]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_u32(tempty_bar(as), lead)); else mbar_arrive(tempty_bar(as)); }
    }
    }
  }
  tc_fence_before();
  if (PAIR || EPI == 6) cluster_sync_all(); else __syncthreads();   // pair: the peer's MMAs / remote arrives target this CTA until here
  if (warp == 2) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                   : "memory");
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
int g_num_sms = 0;
int* g_error_flag = nullptr;
void* g_identity[3] = {nullptr, nullptr, nullptr};   // [DH_BF16], [DH_F16]: 256 x 256 identity, row-major

template <typename T>
__global__ void identity_kernel(T* I, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * n) I[i] = dh_from_f<T>((i / n) == (i % n) ? 1.f : 0.f);
}

int tc_init() {
  if (g_encode_tiled) return DH_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  DH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return dh_fail(DH_ERR_DEVICE, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  void* fn2 = nullptr;
  DH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn2, cudaEnableDefault, &q));
  if (!fn2 || q != cudaDriverEntryPointSuccess) return dh_fail(DH_ERR_DEVICE, "cuTensorMapEncodeIm2col unavailable", __FILE__, __LINE__);
  int dev = 0;
  DH_CUDA(cudaGetDevice(&dev));
  DH_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  DH_CUDA(cudaMalloc(&g_error_flag, sizeof(int)));
  DH_CUDA(cudaMemset(g_error_flag, 0, sizeof(int)));
  DH_CUDA(cudaMalloc(&g_identity[DH_BF16], 256 * 256 * 2));
  DH_CUDA(cudaMalloc(&g_identity[DH_F16], 256 * 256 * 2));
  identity_kernel<<<256, 256>>>((__nv_bfloat16*)g_identity[DH_BF16], 256);
  identity_kernel<<<256, 256>>>((__half*)g_identity[DH_F16], 256);
  DH_CUDA(cudaDeviceSynchronize());
  g_encode_im2col = (EncodeIm2colFn)fn2;
  g_encode_tiled = (EncodeTiledFn)fn;
  return DH_OK;
}

// 2-D row-major [rows, cols] with leading dimension ld (elements); box = {128 B of columns, box_rows}, 128B swizzle.
int make_map_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows, int dtype) {
  const int esize = dtype == DH_F32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType ty = dtype == DH_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == DH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = g_encode_tiled(map, ty, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dh_fail(DH_ERR_ARG, "cuTensorMapEncodeTiled rejected the operand (alignment / stride)", __FILE__, __LINE__);
  return DH_OK;
}

template <int BN, bool PAIR, int EPI, bool ARES = false, int G2 = 0>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mr, const CUtensorMap& mi,
           TcParams& p, cudaStream_t s) {
  using C = Cfg<BN, PAIR, EPI, ARES, G2>;
  static bool attr = false;
  if (!attr) {
    DH_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, PAIR, EPI, ARES, G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr = true;
  }
  static const ChainMaps no_chain{};
  const ChainMaps& cmaps = (G2 == 3 && tl_chain) ? *tl_chain : no_chain;
  if (p.n_stride < 1) p.n_stride = 1;
  p.n_blocks = dh_cdiv(dh_cdiv(p.N, BN) - p.n_offset, p.n_stride);
  if (p.bm_rows <= 0) p.bm_rows = BM;
  p.m_blocks = dh_cdiv(p.M, PAIR ? 2 * p.bm_rows : p.bm_rows);
  p.tiles_per_layer = p.m_blocks * p.n_blocks;
  p.mb128 = dh_cdiv(p.M, BM);
  // LayerNorm mode schedules ROW BLOCKS (each CTA / pair runs both N halves of a block)
  const int tiles = (EPI == 5 || G2 == 3) ? p.m_blocks : p.tiles_per_layer * ((EPI == 3 && p.layers > 1) ? p.layers : 1);
  if (p.res_chunks) p.res_chunks = BN / BK;
  if (PAIR && EPI == 6) {
    // clusters of two CTA pairs: pair tiles 2 j and 2 j + 1 (the two column halves of a 256-row block) run side by side.
    // The grid is what the device can hold of such clusters at once (GPCs whose SM count is not a multiple of 4 leave SMs out).
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static int max_clusters = -1;
    if (max_clusters < 0) {
      cfg.gridDim = dim3(g_num_sms / 4 * 4);
      int n = 0;
      DH_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, PAIR, EPI, ARES, G2>, &cfg));
      max_clusters = n > 0 ? n : 1;
    }
    const int want = tiles / 2;                              // clusters that have work (tiles is even)
    cfg.gridDim = dim3(4 * (want < max_clusters ? want : max_clusters));
    DH_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, PAIR, EPI, ARES, G2>, ma, mb, mc, mr, mi, p, cmaps));
  } else if (PAIR) {
    const int pairs = g_num_sms / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    DH_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, PAIR, EPI, ARES, G2>, ma, mb, mc, mr, mi, p, cmaps));
  } else if (EPI == 6) {
    // clusters of two plain CTAs: tiles 2 j and 2 j + 1 (the two column halves of row block j) run side by side
    const int even_sms = g_num_sms & ~1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles < even_sms ? tiles : even_sms);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    DH_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, PAIR, EPI, ARES, G2>, ma, mb, mc, mr, mi, p, cmaps));
  } else {
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    gemm_tc_kernel<BN, PAIR, EPI, ARES, G2><<<grid, kThreads, C::kSmemBytes, s>>>(ma, mb, mc, mr, mi, p, cmaps);
  }
  DH_LAUNCH_OK();
  return DH_OK;
}

template <int EPI>
int launch_bn(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mr, const CUtensorMap& mi,
              TcParams& p, int bn, bool pair, int g2, cudaStream_t s) {
  if (EPI == 0 && g2 == 1) return launch<256, false, 0, false, 1>(ma, mb, mc, mr, mi, p, s);   // (dispatch: bn == 256, single CTA)
  if (EPI == 0 && g2 == 3) return launch<256, false, 0, false, 3>(ma, mb, mc, mr, mi, p, s);   // chained second contraction
  if (EPI == 0 && g2 == 2) {                                // plain stores with both epilogue groups at work
    if (bn == 64) return launch<64, false, 0, false, 2>(ma, mb, mc, mr, mi, p, s);
    if (bn == 128) return pair ? launch<128, true, 0, false, 2>(ma, mb, mc, mr, mi, p, s) : launch<128, false, 0, false, 2>(ma, mb, mc, mr, mi, p, s);
    return pair ? launch<256, true, 0, false, 2>(ma, mb, mc, mr, mi, p, s) : launch<256, false, 0, false, 2>(ma, mb, mc, mr, mi, p, s);
  }
  // full vocab-projection pass (143 N tiles per row block, K <= 512): A resident in shared memory, W streamed.  Measured at
  // 40 960 rows: 1109 vs 1159 us; the strided pass 1 (18 tiles per row block) re-loads A too often to gain and stays streamed.
  static const bool ares_ok = !getenv("DH_TC_NO_ARES");
  if (EPI == 2 && ares_ok && bn == 256 && pair && p.k_chunks <= 8 && !p.conv && !p.res_chunks && p.n_stride == 1)
    return launch<256, true, EPI == 2 ? 2 : 1, true>(ma, mb, mc, mr, mi, p, s);
  if (EPI == 3) return pair ? launch<256, true, 3>(ma, mb, mc, mr, mi, p, s) : launch<256, false, 3>(ma, mb, mc, mr, mi, p, s);
  if (bn == 64) return launch<64, false, EPI>(ma, mb, mc, mr, mi, p, s);
  if (bn == 128) return pair ? launch<128, true, EPI>(ma, mb, mc, mr, mi, p, s) : launch<128, false, EPI>(ma, mb, mc, mr, mi, p, s);
  return pair ? launch<256, true, EPI>(ma, mb, mc, mr, mi, p, s) : launch<256, false, EPI>(ma, mb, mc, mr, mi, p, s);
}

int pick_bn(int M, int N) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  // prefer 256-wide tiles when that still gives every SM work
  const long long t256 = (long long)dh_cdiv(M, BM) * dh_cdiv(N, 256);
  return (t256 >= 2ll * (g_num_sms ? g_num_sms : 148) || N % 256 == 0 && t256 >= (g_num_sms ? g_num_sms : 148)) ? 256 : 128;
}

int dispatch(const CUtensorMap& ma, const void* W, long long ldw, TcParams& p, int bn, cudaStream_t s,
             const CUtensorMap* second_a = nullptr, const CUtensorMap* third = nullptr) {
  CUtensorMap mb, mc = ma, mr = second_a ? *second_a : ma, mi = third ? *third : ma;
  // CTA pairs (cta_group::2) for the 128- and 256-wide tiles whenever there are at least two M blocks to pair up
  static const bool pair_ok = !getenv("DH_TC_NO_PAIR");
  // ... and the K loop is long enough (>= 8 chunks incl. residual chunks) to amortise the pair's cross-CTA barrier round
  // trips.  Measured per ResNet-50 / decoder shape (profiles/r01_bench_conv_pairmin.txt): pairs win 6-30 % at K >= 1024,
  // 4-9 % at K = 512 (l3.c1, l3.c3, l3.ds, l4.c3, the vocab projection) now that the remote accumulator hand-off carries no
  // GPU-scope membar, and lose on the 1-2 chunk store-bound tiles of layer1 (l1.c3 +8 % at a threshold of 4).
  static const int pair_min_chunks = getenv("DH_TC_PAIR_MIN_CHUNKS") ? atoi(getenv("DH_TC_PAIR_MIN_CHUNKS")) : 8;
  const int bm_rows = p.bm_rows > 0 ? p.bm_rows : BM;
  // LayerNorm launches split the row over a cluster of two plain CTAs (EPI 6) unless DH_TC_LN_UNSPLIT is set
  static const bool ln_split_ok = !getenv("DH_TC_LN_UNSPLIT");
  const bool ln_split = p.epi_mode == 5 && ln_split_ok && g_num_sms >= 2;
  static const bool ln_pair_ok = !getenv("DH_TC_LN_NO_PAIR");
  const bool pair = pair_ok && bn >= 128 && p.M > bm_rows && p.k_chunks + (p.res ? bn / BK : 0) >= pair_min_chunks &&
                    (!ln_split || (ln_pair_ok && g_num_sms >= 4)) && p.N2 == 0;
  const int b_rows = pair ? bn / 2 : bn;                   // W rows one CTA stages per K chunk
  const long long w_rows = (p.epi_mode == 3 && p.layers > 1) ? (long long)p.layers * p.w_layer_rows : p.N;
  int rc = make_map_2d(&mb, W, w_rows, p.K, ldw, b_rows, p.ab_dtype);
  if (rc) return rc;
  p.error = g_error_flag;
  // TMA-store epilogue when the output rows are 16 B aligned; the residual then rides the tensor core (R x I) if it
  // has the operand dtype and 16 B aligned rows.  Anything else takes the direct-store epilogue.
  const int osize = p.out_dtype == DH_F32 ? 4 : 2;
  const bool out_ok = ((uintptr_t)p.out % 16 == 0) && ((p.ldc * osize) % 16 == 0) && !getenv("DH_TC_DIRECT_EPILOGUE");
  const bool res_ok = !p.res || (p.res_dtype == p.ab_dtype && (uintptr_t)p.res % 16 == 0 && (p.ldr * 2) % 16 == 0);
  p.tma_store = 0;
  p.res_chunks = 0;
  // both epilogue groups for short K loops (measured: K = 512 decoder contractions +10 %, l1.c1 / l1.ds +17-29 %, neutral up
  // to K = 1024; 8192^3 and K = 2048 lose 4-6 %); one slab per group for the single-CTA 256-wide residual launches (l1.c3:
  // ring depth matters more there)
  static const int groups_env = getenv("DH_TC_EPI_GROUPS") ? atoi(getenv("DH_TC_EPI_GROUPS")) : 0;
  int g2 = p.k_chunks + (p.res ? bn / BK : 0) > 16 ? 0 : (p.res && !pair && bn == 256) ? 1 : 2;
  if (groups_env) g2 = groups_env == 1 ? 0 : (groups_env == 3 && !pair && bn == 256) ? 1 : 2;
  if (p.N2 > 0) g2 = 3;
  if (p.epi_mode && p.epi_mode != 5) {
    // selection epilogues store nothing of C
  } else if (out_ok && res_ok) {
    rc = make_map_2d(&mc, p.out, p.M, p.split_n ? p.split_n : p.N, p.ldc, bm_rows, p.out_dtype);
    if (rc) return rc;
    p.tma_store = 1;
    if (p.res) {
      rc = make_map_2d(&mr, p.res, p.M, p.N, p.ldr, bm_rows, p.ab_dtype);
      if (rc) return rc;
      rc = make_map_2d(&mi, g_identity[p.ab_dtype], 256, 256, 256, pair ? b_rows : 64, p.ab_dtype);
      if (rc) return rc;
      p.res_chunks = 1;   // launch<> sets BN / 64
    }
  }
  if (p.split_n && !p.tma_store) return dh_fail(DH_ERR_ARG, "split destinations need the TMA-store epilogue", __FILE__, __LINE__);
  switch (p.epi_mode) {
    case 0: return launch_bn<0>(ma, mb, mc, mr, mi, p, bn, pair, g2, s);
    case 1: return launch_bn<1>(ma, mb, mc, mr, mi, p, bn, pair, g2, s);
    case 2: return launch_bn<2>(ma, mb, mc, mr, mi, p, bn, pair, g2, s);
    case 3: return launch_bn<3>(ma, mb, mc, mr, mi, p, bn, pair, g2, s);
    case 5:
      if (!p.tma_store) return dh_fail(DH_ERR_ARG, "LayerNorm epilogue needs 16-byte aligned output / residual rows", __FILE__, __LINE__);
      if (ln_split) return pair ? launch<256, true, 6>(ma, mb, mc, mr, mi, p, s) : launch<256, false, 6>(ma, mb, mc, mr, mi, p, s);
      return pair ? launch<256, true, 5>(ma, mb, mc, mr, mi, p, s) : launch<256, false, 5>(ma, mb, mc, mr, mi, p, s);
    default: return launch_bn<4>(ma, mb, mc, mr, mi, p, bn, pair, g2, s);
  }
}

}  // namespace

extern "C" int dh_gemm_tc(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                          const void* residual, long long ldr, int res_dtype, void* C, long long ldc, int out_dtype, int M, int N,
                          int K, int relu, int tile_n, cudaStream_t stream) {
  DH_ARG(A && W && C && M >= 0 && N > 0 && K > 0);
  DH_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0);
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0);
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  DH_ARG(out_dtype == DH_F32 || out_dtype == DH_BF16 || out_dtype == DH_F16);
  DH_ARG(!residual || res_dtype == DH_F32 || res_dtype == DH_BF16 || res_dtype == DH_F16);
  DH_ARG(tile_n == 0 || tile_n == 64 || tile_n == 128 || tile_n == 256);
  if (M == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.k_chunks = dh_cdiv(K, BK);
  p.conv = 0;
  p.ab_dtype = ab_dtype;
  p.bias = bias; p.res = residual; p.ldr = ldr; p.res_dtype = res_dtype;
  p.out = C; p.ldc = ldc; p.out_dtype = out_dtype; p.relu = relu;
  CUtensorMap ma;
  rc = make_map_2d(&ma, A, M, K, lda, BM, ab_dtype);
  if (rc) return rc;
  return dispatch(ma, W, ldw, p, tile_n ? tile_n : pick_bn(M, N), stream);
}

// C = act(A W^T + bias + residual) AND pool[i, :] = mean of rows [i * pool_hw, (i + 1) * pool_hw) of C (fp32): the last
// bottleneck's conv3 + bn3 + residual + ReLU with the AdaptiveAvgPool2d of encoders.py:39,60 in its epilogue (a 1x1 / stride 1
// convolution over NHWC is this plain contraction over the pixel rows).  M tiles hold whole images (pool_hw <= 128,
// M % pool_hw == 0); the pooled means equal dh_avgpool on the stored C bit for bit.
extern "C" int dh_gemm_tc_pool(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                               const void* residual, long long ldr, void* C, long long ldc, int M, int N, int K, int relu,
                               int pool_hw, float* pool, long long ld_pool, cudaStream_t stream) {
  DH_ARG(A && W && C && pool && M >= 0 && N > 0 && K > 0);
  DH_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && (!residual || ldr % 8 == 0));
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)C % 16) == 0 && ((uintptr_t)residual % 16) == 0);
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  DH_ARG(pool_hw >= 1 && pool_hw <= BM && M % pool_hw == 0 && ld_pool >= N);
  if (M == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.k_chunks = dh_cdiv(K, BK);
  p.ab_dtype = ab_dtype;
  p.bias = bias; p.res = residual; p.ldr = ldr; p.res_dtype = ab_dtype;
  p.out = C; p.ldc = ldc; p.out_dtype = ab_dtype; p.relu = relu;
  p.bm_rows = BM / pool_hw * pool_hw;
  p.pool_hw = pool_hw; p.pool_out = pool; p.ld_pool = ld_pool;
  CUtensorMap ma;
  rc = make_map_2d(&ma, A, M, K, lda, p.bm_rows, ab_dtype);
  if (rc) return rc;
  rc = dispatch(ma, W, ldw, p, N % 256 == 0 ? 256 : N % 128 == 0 ? 128 : 64, stream);
  if (rc) return rc;
  return p.tma_store ? DH_OK : dh_fail(DH_ERR_ARG, "pooled epilogue needs the TMA-store path", __FILE__, __LINE__);
}

// out = LayerNorm(A W^T + bias + residual) * gamma + beta over rows of N == 512 columns: the post-LN sublayer tail of a decoder
// layer (models/transformers.py:355-356,365-366,374-375 -- fc_o / fc_2, dropout (eval: identity), residual add, nn.LayerNorm)
// in ONE launch: a CTA (pair) holds the whole row block in its two accumulator buffers, so the pre-norm sums are never stored.
// out may alias residual (each row block is read through TMA before its own stores are issued and no other CTA touches it).
extern "C" int dh_gemm_tc_ln(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                             const void* residual, long long ldr, const float* gamma, const float* beta, float eps, void* out,
                             long long ldc, int M, int N, int K, cudaStream_t stream) {
  DH_ARG(A && W && out && gamma && beta && M >= 0 && K > 0 && N == 512 && eps > 0.f);
  DH_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && (!residual || ldr % 8 == 0));
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)residual % 16) == 0);
  DH_ARG(((uintptr_t)bias % 16) == 0 && ((uintptr_t)gamma % 16) == 0 && ((uintptr_t)beta % 16) == 0);
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  if (M == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.k_chunks = dh_cdiv(K, BK);
  p.ab_dtype = ab_dtype;
  p.bias = bias; p.res = residual; p.ldr = ldr; p.res_dtype = ab_dtype;
  p.out = out; p.ldc = ldc; p.out_dtype = ab_dtype;
  p.epi_mode = 5;
  p.ln_gamma = gamma; p.ln_beta = beta; p.ln_eps = eps;
  CUtensorMap ma;
  rc = make_map_2d(&ma, A, M, K, lda, BM, ab_dtype);
  if (rc) return rc;
  return dispatch(ma, W, ldw, p, 256, stream);
}

// One contraction, three destinations: [C0 | C1 | C2][M, 3 * split_n] = A[M,K] * W[3 * split_n, K]^T + bias, block j of
// split_n columns stored to Cj (own leading dimension).  The transformer decode step projects Q, K and V of the new position
// from the same activation row (models/transformers.py:97-99): W = [fc_q | fc_k | fc_v] stacked along N, C0 = the query
// buffer, C1 / C2 = this position's slots of the K / V caches (row stride = slots * positions * D), so the activations are
// read once and the K / V rows land in the cache without a copy.
extern "C" int dh_gemm_tc_split3(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                                 void* C0, long long ldc0, void* C1, long long ldc1, void* C2, long long ldc2, int out_dtype,
                                 int split_n, int M, int K, cudaStream_t stream) {
  DH_ARG(A && W && C0 && C1 && C2 && M >= 0 && K > 0 && split_n > 0 && split_n % 128 == 0);
  DH_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0);
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0);
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  DH_ARG(out_dtype == DH_BF16 || out_dtype == DH_F16);
  DH_ARG(((uintptr_t)C0 % 16) == 0 && ((uintptr_t)C1 % 16) == 0 && ((uintptr_t)C2 % 16) == 0);
  DH_ARG(ldc0 % 8 == 0 && ldc1 % 8 == 0 && ldc2 % 8 == 0);
  if (M == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  p.M = M; p.N = 3 * split_n; p.K = K;
  p.k_chunks = dh_cdiv(K, BK);
  p.ab_dtype = ab_dtype;
  p.bias = bias;
  p.out = C0; p.ldc = ldc0; p.out_dtype = out_dtype;
  p.split_n = split_n;
  CUtensorMap ma, m1, m2;
  rc = make_map_2d(&ma, A, M, K, lda, BM, ab_dtype);
  if (rc) return rc;
  rc = make_map_2d(&m1, C1, M, split_n, ldc1, BM, out_dtype);
  if (rc) return rc;
  rc = make_map_2d(&m2, C2, M, split_n, ldc2, BM, out_dtype);
  if (rc) return rc;
  int bn = pick_bn(M, p.N);
  if (split_n % bn) bn = 128;
  return dispatch(ma, W, ldw, p, bn, stream, &m1, &m2);
}

// Vocab projection with the logits kept on chip (models/rnn_models.py:81,109 and models/transformers.py:488,736 feeding
// models/beam.py:32-37): pass 1 stores the maximum of every 32-column group of logits[M,N] = A W^T + bias, pass 2
// recomputes the identical product and appends (column, logit) of every logit >= thresh[row] to the row's candidate list.
struct VocabFix {          // optional behaviour of a vocab pass (see TcParams::cond_mode)
  int tile_offset = 0;
  int cond_mode = 0, cond_min = 0, cond_max = 0;
  const int* cond_count = nullptr;
  int* cond_flag = nullptr;
  const unsigned char* redo = nullptr;
};

static int vocab_pass(int mode, int tile_stride, const void* A, long long lda, const void* W, long long ldw, int ab_dtype,
                      const float* bias, int M, int N, int K, float* gmax, long long ld_gmax, const float* thresh,
                      int* cand_count, float* sp_logits, long long sp_ld, unsigned char* hitmap, long long hit_ld,
                      cudaStream_t stream, const VocabFix& fx = VocabFix()) {
  DH_ARG(A && W && M >= 0 && N > 0 && K > 0);
  DH_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0);
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0);
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  if (M == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.k_chunks = dh_cdiv(K, BK);
  p.ab_dtype = ab_dtype;
  p.bias = bias;
  p.epi_mode = mode;
  p.n_stride = tile_stride;
  p.n_offset = fx.tile_offset;
  p.cond_mode = fx.cond_mode; p.cond_min = fx.cond_min; p.cond_max = fx.cond_max;
  p.cond_count = fx.cond_count; p.cond_flag = fx.cond_flag; p.redo = fx.redo;
  p.gmax = gmax; p.ld_gmax = ld_gmax;
  p.thresh = thresh; p.cand_count = cand_count;
  p.sp_logits = sp_logits; p.sp_ld = sp_ld; p.hitmap = hitmap; p.hit_ld = hit_ld;
  CUtensorMap ma;
  rc = make_map_2d(&ma, A, M, K, lda, BM, ab_dtype);
  if (rc) return rc;
  return dispatch(ma, W, ldw, p, N <= 64 ? 64 : N <= 128 ? 128 : 256, stream);
}

// One nn.LSTM layer step on tensor cores with the cell update in the epilogue (rnn_models.py:80,108): gates = A [x | h]
// times the gate-packed [W_ih | W_hh] (rows ordered per 64 hidden units as i, f, g, o; bias = b_ih + b_hh packed alike).
extern "C" int dh_lstm_layer_tc(const void* A, long long lda, const void* Wp, long long ldw, int ab_dtype, const float* bias_p,
                                const float* c_prev, const int* parent, float* c_out, void* h_out0, long long ldh0,
                                void* h_out1, long long ldh1, int rows, int H, int K, cudaStream_t stream) {
  DH_ARG(A && Wp && c_out && rows >= 0 && H > 0 && H % 64 == 0 && K > 0);
  DH_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0);
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)Wp % 16) == 0 && ((uintptr_t)c_out % 16) == 0);
  DH_ARG(!c_prev || ((uintptr_t)c_prev % 16) == 0);
  DH_ARG(!h_out0 || (((uintptr_t)h_out0 % 16) == 0 && ldh0 % 8 == 0));
  DH_ARG(!h_out1 || (((uintptr_t)h_out1 % 16) == 0 && ldh1 % 8 == 0));
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  if (rows == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  p.M = rows; p.N = 4 * H; p.K = K;
  p.k_chunks = dh_cdiv(K, BK);
  p.ab_dtype = ab_dtype;
  p.out_dtype = ab_dtype;
  p.bias = bias_p;
  p.epi_mode = 3;
  p.c_prev = c_prev; p.parent = parent; p.c_out = c_out;
  p.h0 = h_out0; p.ldh0 = ldh0; p.h1 = h_out1; p.ldh1 = ldh1; p.H = H;
  CUtensorMap ma;
  rc = make_map_2d(&ma, A, rows, K, lda, BM, ab_dtype);
  if (rc) return rc;
  return dispatch(ma, Wp, ldw, p, 256, stream);
}

// All layers of an nn.LSTM time step in ONE persistent launch (rnn_models.py:80,108 with num_layers > 1): the tiles of the
// layers are queued layer-major on the same CTAs, and a tile of layer l starts its K loop on the recurrent half of its
// operand while the layer below is still running; it reaches the x half once the per-128-row counter says those rows of
// h_{l-1} are stored.  Replaces `layers` dependent launches (each with its own tail wave) by one.
extern "C" int dh_lstm_stack_tc(void* A, long long lda, long long a_layer_rows, const int* in_dims_host, const void* Wp,
                                long long ldw, int ab_dtype, const float* bias_p, const float* c_prev, const int* parent,
                                float* c_out, long long c_layer_stride, void* h_top, long long ld_top, void* hs,
                                long long hs_layer_stride, int* ready, int rows, int H, int layers, int rotate_k,
                                cudaStream_t stream) {
  DH_ARG(A && in_dims_host && Wp && c_out && ready && rows >= 0 && H > 0 && H % 64 == 0);
  DH_ARG(layers >= 1 && layers <= kMaxLayers && a_layer_rows >= rows);
  DH_ARG(lda % 8 == 0 && ldw % 8 == 0 && c_layer_stride % 4 == 0 && hs_layer_stride % 8 == 0);
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)Wp % 16) == 0 && ((uintptr_t)c_out % 16) == 0);
  DH_ARG(!c_prev || ((uintptr_t)c_prev % 16) == 0);
  DH_ARG(h_top && ((uintptr_t)h_top % 16) == 0 && ld_top % 8 == 0);
  DH_ARG(!hs || ((uintptr_t)hs % 16) == 0);
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  DH_ARG((long long)layers * a_layer_rows < (1ll << 31));
  if (rows == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  int kmax = 0;
  for (int l = 0; l < layers; ++l) {
    const int in_l = in_dims_host[l], K_l = in_l + H;
    DH_ARG(in_l > 0 && K_l % 8 == 0 && K_l <= lda && K_l <= ldw);
    DH_ARG(l == 0 || in_l == H);                       // layer l > 0 consumes the H-wide h of the layer below
    p.kch_l[l] = dh_cdiv(K_l, BK);
    // start at the recurrent half when the x | h boundary falls on a chunk boundary
    p.rot_l[l] = (rotate_k && l > 0 && in_l % BK == 0) ? in_l / BK : 0;
    p.bias_l[l] = bias_p ? bias_p + (long long)l * 4 * H : nullptr;
    p.cprev_l[l] = c_prev ? c_prev + l * c_layer_stride : nullptr;
    p.cout_l[l] = c_out + l * c_layer_stride;
    if (l + 1 < layers) {                              // x half of the next layer's operand
      p.h0_l[l] = reinterpret_cast<uint16_t*>(A) + (long long)(l + 1) * a_layer_rows * lda;
      p.ldh0_l[l] = lda;
    } else {
      p.h0_l[l] = h_top;
      p.ldh0_l[l] = ld_top;
    }
    p.h1_l[l] = hs ? reinterpret_cast<uint16_t*>(hs) + l * hs_layer_stride : nullptr;
    p.ldh1_l[l] = H;
    kmax = K_l > kmax ? K_l : kmax;
  }
  p.M = rows; p.N = 4 * H; p.K = kmax;
  p.k_chunks = dh_cdiv(kmax, BK);
  p.ab_dtype = ab_dtype;
  p.out_dtype = ab_dtype;
  p.epi_mode = 3;
  p.parent = parent; p.H = H;
  p.layers = layers; p.a_layer_rows = (int)a_layer_rows; p.w_layer_rows = 4 * H;
  p.ready = ready;
  if (layers == 1) {                                   // plain single-layer launch
    p.bias = p.bias_l[0]; p.c_prev = p.cprev_l[0]; p.c_out = p.cout_l[0];
    p.h0 = p.h0_l[0]; p.ldh0 = p.ldh0_l[0]; p.h1 = p.h1_l[0]; p.ldh1 = p.ldh1_l[0];
  }
  CUtensorMap ma;
  rc = make_map_2d(&ma, A, (long long)(layers - 1) * a_layer_rows + rows, kmax, lda, BM, ab_dtype);
  if (rc) return rc;
  return dispatch(ma, Wp, ldw, p, 256, stream);
}

// log_softmax(A W^T + bias)[row, targets[row]] without storing the logits (experiments/metrics.py:5 on the classifier
// output of rnn_models.py:44 / transformers.py:488,736): the contraction's epilogue keeps (max, sum exp) per 32-column
// group and the target's logit; dh_vocab_logprob_reduce (select.cu) folds the groups of a row.
int dh_vocab_logprob_reduce(const float* gmax, const float* gsum, long long ld, int rows, int n_groups, const float* tlogit,
                            float* out, cudaStream_t s);
extern "C" int dh_vocab_logprob(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                                int M, int N, int K, const long long* targets, float* gmax, float* gsum, long long ld_g,
                                float* tlogit, float* out, cudaStream_t stream) {
  DH_ARG(A && W && targets && gmax && gsum && tlogit && out && M >= 0 && N > 0 && K > 0);
  DH_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0);
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0);
  DH_ARG(ab_dtype == DH_BF16 || ab_dtype == DH_F16);
  const int bn = N <= 64 ? 64 : N <= 128 ? 128 : 256;
  const int n_groups = dh_cdiv(N, bn) * (bn / 32);
  DH_ARG(ld_g >= n_groups);
  if (M == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.k_chunks = dh_cdiv(K, BK);
  p.ab_dtype = ab_dtype;
  p.bias = bias;
  p.epi_mode = 4;
  p.n_stride = 1;
  p.gmax = gmax; p.gsum = gsum; p.ld_gmax = ld_g; p.targets = targets; p.tlogit = tlogit;
  CUtensorMap ma;
  rc = make_map_2d(&ma, A, M, K, lda, BM, ab_dtype);
  if (rc) return rc;
  rc = dispatch(ma, W, ldw, p, bn, stream);
  if (rc) return rc;
  return dh_vocab_logprob_reduce(gmax, gsum, ld_g, M, n_groups, tlogit, out, stream);
}

static int vocab_groups(int N, int tile_stride, int tile_offset) {
  const int bn = N <= 64 ? 64 : N <= 128 ? 128 : 256;
  return dh_cdiv(dh_cdiv(N, bn) - tile_offset, tile_stride) * (bn / 32);
}

extern "C" int dh_vocab_groupmax(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                                 int M, int N, int K, int tile_stride, int tile_offset, float* gmax, long long ld_gmax,
                                 cudaStream_t stream) {
  DH_ARG(gmax && tile_stride >= 1 && tile_offset >= 0 && tile_offset < tile_stride);
  DH_ARG(tile_offset < dh_cdiv(N, N <= 64 ? 64 : N <= 128 ? 128 : 256));
  DH_ARG(ld_gmax >= vocab_groups(N, tile_stride, tile_offset));
  VocabFix fx;
  fx.tile_offset = tile_offset;
  return vocab_pass(1, tile_stride, A, lda, W, ldw, ab_dtype, bias, M, N, K, gmax, ld_gmax, nullptr, nullptr, nullptr, 0, nullptr,
                    0, stream, fx);
}

static int sparse_args_ok(int N, const float* sp_logits, long long sp_ld, const unsigned char* hitmap, long long hit_ld) {
  const int bn = N <= 64 ? 64 : N <= 128 ? 128 : 256;
  const int nb = dh_cdiv(N, bn);
  return sp_logits && hitmap && sp_ld >= (long long)nb * bn && sp_ld % 4 == 0 && ((uintptr_t)sp_logits % 16) == 0 && hit_ld >= 2 * nb;
}

extern "C" int dh_vocab_candidates(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                                   int M, int N, int K, const float* thresh, int* cand_count, float* sp_logits, long long sp_ld,
                                   unsigned char* hitmap, long long hit_ld, cudaStream_t stream) {
  DH_ARG(thresh && cand_count && sparse_args_ok(N, sp_logits, sp_ld, hitmap, hit_ld));
  return vocab_pass(2, 1, A, lda, W, ldw, ab_dtype, bias, M, N, K, nullptr, 0, thresh, cand_count, sp_logits, sp_ld, hitmap,
                    hit_ld, stream);
}

// Fix-up launches behind a SAMPLED pass 1 (see include/deephumor_b200.h): both contractions return at once -- before any
// barrier or tensor-memory set-up -- unless some row's candidate count left [count_min, count_max] / *any_flag is set.
extern "C" int dh_vocab_groupmax_fix(const void* A, long long lda, const void* W, long long ldw, int ab_dtype, const float* bias,
                                     int M, int N, int K, float* gmax, long long ld_gmax, const int* cand_count, int count_min,
                                     int count_max, int* any_flag, cudaStream_t stream) {
  DH_ARG(gmax && cand_count && any_flag && count_min >= 0 && count_max >= count_min);
  DH_ARG(ld_gmax >= vocab_groups(N, 1, 0));
  VocabFix fx;
  fx.cond_mode = 1; fx.cond_min = count_min; fx.cond_max = count_max; fx.cond_count = cand_count; fx.cond_flag = any_flag;
  return vocab_pass(1, 1, A, lda, W, ldw, ab_dtype, bias, M, N, K, gmax, ld_gmax, nullptr, nullptr, nullptr, 0, nullptr, 0,
                    stream, fx);
}

extern "C" int dh_vocab_candidates_fix(const void* A, long long lda, const void* W, long long ldw, int ab_dtype,
                                       const float* bias, int M, int N, int K, const float* thresh, int* cand_count,
                                       float* sp_logits, long long sp_ld, unsigned char* hitmap, long long hit_ld,
                                       const unsigned char* redo, int* any_flag, cudaStream_t stream) {
  DH_ARG(thresh && cand_count && sparse_args_ok(N, sp_logits, sp_ld, hitmap, hit_ld) && redo && any_flag);
  VocabFix fx;
  fx.cond_mode = 2; fx.cond_flag = any_flag; fx.redo = redo;
  return vocab_pass(2, 1, A, lda, W, ldw, ab_dtype, bias, M, N, K, nullptr, 0, thresh, cand_count, sp_logits, sp_ld, hitmap,
                    hit_ld, stream, fx);
}

// im2col-mode tensor map over an NHWC activation tensor: box = {64 channels, 128 output pixels}
static int make_map_im2col(CUtensorMap* map, const void* x, int n, int H, int W, int Cin, int kh, int kw, int stride, int pad,
                           int dtype) {
  cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (kw - 1), pad - (kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(map, dtype == DH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                               const_cast<void*>(x), dims, strides, lower, upper,
                               (cuuint32_t)BK, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dh_fail(DH_ERR_ARG, "cuTensorMapEncodeIm2col rejected the activation tensor", __FILE__, __LINE__);
  return DH_OK;
}

// x [n,H,W,Cin] NHWC (Cin % 64 == 0), w [Cout][kh][kw][Cin] (BN folded), y [n,Ho,Wo,Cout]; all bf16 or all f16.
extern "C" int dh_conv2d_tc(const void* x, const void* w, const float* bias, const void* residual, void* y, int n, int H,
                            int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int relu, int dtype, int tile_n,
                            cudaStream_t stream) {
  DH_ARG(dtype == DH_BF16 || dtype == DH_F16);
  DH_ARG(x && w && y && n >= 0 && Cin > 0 && Cin % 64 == 0 && Cout > 0 && Cout % 4 == 0 && stride > 0 && kh > 0 && kw > 0);
  DH_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0);
  DH_ARG(tile_n == 0 || tile_n == 64 || tile_n == 128 || tile_n == 256);
  if (n == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long M = (long long)n * Ho * Wo;
  DH_ARG(M < (1ll << 31) && Ho > 0 && Wo > 0);
  TcParams p{};
  p.M = (int)M; p.N = Cout; p.K = kh * kw * Cin;
  p.k_chunks = p.K / BK;
  p.conv = 1; p.HoWo = Ho * Wo; p.Wo = Wo; p.c_chunks = Cin / BK; p.kw = kw; p.stride = stride; p.pad = pad;
  p.ab_dtype = dtype;
  p.bias = bias; p.res = residual; p.ldr = Cout; p.res_dtype = dtype;
  p.out = y; p.ldc = Cout; p.out_dtype = dtype; p.relu = relu;
  CUtensorMap ma;
  rc = make_map_im2col(&ma, x, n, H, W, Cin, kh, kw, stride, pad, dtype);
  if (rc) return rc;
  return dispatch(ma, w, p.K, p, tile_n ? tile_n : pick_bn(p.M, Cout), stream);
}

// The first bottleneck of a ResNet stage ends in relu(bn3(conv3(y2)) + bn_d(conv_d(x))) (torchvision resnet.py:154-161 with
// the downsample branch :157-158): two 1x1 convolutions onto the same output grid.  Here they are ONE contraction over
// K = C1 + C2 -- chunks [0, C1/64) from y2 (stride 1), the rest from the block input x (stride2) through a second im2col
// map, weights [W3 | Wd] concatenated along K, bias = b3 + bd -- so the downsample output is neither written nor re-read
// (1.6 GB of HBM traffic per 512 images in layer1) and no residual chunks ride the tensor core.
extern "C" int dh_conv1x1_dual_tc(const void* x1, const void* x2, const void* w_cat, const float* bias, void* y, int n, int Ho,
                                  int Wo, int C1, int H2, int W2, int C2, int stride2, int Cout, int relu, int dtype,
                                  int tile_n, cudaStream_t stream) {
  DH_ARG(dtype == DH_BF16 || dtype == DH_F16);
  DH_ARG(x1 && x2 && w_cat && y && n >= 0 && C1 > 0 && C1 % 64 == 0 && C2 > 0 && C2 % 64 == 0 && Cout > 0 && Cout % 4 == 0);
  DH_ARG(stride2 >= 1 && (H2 - 1) / stride2 + 1 == Ho && (W2 - 1) / stride2 + 1 == Wo && Ho > 0 && Wo > 0);
  DH_ARG(((uintptr_t)x1 % 16) == 0 && ((uintptr_t)x2 % 16) == 0 && ((uintptr_t)w_cat % 16) == 0);
  DH_ARG(tile_n == 0 || tile_n == 64 || tile_n == 128 || tile_n == 256);
  if (n == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  const long long M = (long long)n * Ho * Wo;
  DH_ARG(M < (1ll << 31));
  TcParams p{};
  p.M = (int)M; p.N = Cout; p.K = C1 + C2;
  p.k_chunks = p.K / BK;
  p.conv = 1; p.HoWo = Ho * Wo; p.Wo = Wo; p.c_chunks = C1 / BK; p.kw = 1; p.stride = 1; p.pad = 0;
  p.k1_chunks = C1 / BK; p.stride2 = stride2;
  p.ab_dtype = dtype;
  p.bias = bias;
  p.out = y; p.ldc = Cout; p.out_dtype = dtype; p.relu = relu;
  CUtensorMap ma, ma2;
  rc = make_map_im2col(&ma, x1, n, Ho, Wo, C1, 1, 1, 1, 0, dtype);
  if (rc) return rc;
  rc = make_map_im2col(&ma2, x2, n, H2, W2, C2, 1, 1, stride2, 0, dtype);
  if (rc) return rc;
  return dispatch(ma, w_cat, p.K, p, tile_n ? tile_n : pick_bn(p.M, Cout), stream, &ma2);
}

// conv3 of a bottleneck AND conv1 of the NEXT bottleneck in one launch (torchvision resnet.py:154-161 then :146-148 of the
// following block): out [n,H,W,Cout] = relu(conv1x1(y2; W[:, :C1]) + (x2_is_source ? conv1x1(x2 (stride2); W[:, C1:]) : x2) +
// bias) with Cout a multiple of 256, and z = relu(conv1x1(out; w_next) + bias_next) with N2 = 64 / 128 / 256 output channels.
// A CTA computes tile Q (128 pixels of z) two tiles after the tiles P (the same 128 pixels of out, Cout / 256 of them) it has
// stored, reading P back through TMA while it is still in L2: the next block's conv1 costs no HBM read (layer1: 1.6 GB per
// 1024 images and block boundary).  Bit-identical to the two separate launches.
extern "C" int dh_conv1x1_chain_tc(const void* y2, const void* x2, int x2_is_source, const void* w, const float* bias, void* out,
                                   int n, int H, int W, int C1, int C2, int H2, int W2, int stride2, int Cout, const void* w_next,
                                   const float* bias_next, void* z, int N2, int dtype, cudaStream_t stream) {
  DH_ARG(dtype == DH_BF16 || dtype == DH_F16);
  DH_ARG(y2 && x2 && w && out && w_next && z && n >= 0 && H > 0 && W > 0 && C1 > 0 && C1 % 64 == 0);
  DH_ARG(Cout > 0 && Cout % 256 == 0 && Cout <= 1024);
  DH_ARG(x2_is_source ? (C2 > 0 && C2 % 64 == 0 && stride2 >= 1 && (H2 - 1) / stride2 + 1 == H && (W2 - 1) / stride2 + 1 == W)
                      : C2 == Cout);
  DH_ARG(N2 == 64 || N2 == 128 || N2 == 256);
  DH_ARG(((uintptr_t)y2 % 16) == 0 && ((uintptr_t)x2 % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)out % 16) == 0);
  DH_ARG(((uintptr_t)w_next % 16) == 0 && ((uintptr_t)z % 16) == 0);
  if (n == 0) return DH_OK;
  int rc = tc_init();
  if (rc) return rc;
  const long long M = (long long)n * H * W;
  DH_ARG(M < (1ll << 31));
  TcParams p{};
  p.M = (int)M; p.N = Cout; p.K = x2_is_source ? C1 + C2 : C1;
  p.k_chunks = p.K / BK;
  p.ab_dtype = dtype;
  p.bias = bias;
  p.out = out; p.ldc = Cout; p.out_dtype = dtype; p.relu = 1;
  p.N2 = N2; p.k2_chunks = Cout / BK; p.relu2 = 1; p.bias2 = bias_next;
  CUtensorMap ma, ma2;
  ChainMaps cm;
  rc = make_map_2d(&cm.b2, w_next, N2, Cout, Cout, N2, dtype);
  if (rc) return rc;
  rc = make_map_2d(&cm.c2, z, M, N2, N2, BM, dtype);
  if (rc) return rc;
  if (x2_is_source) {
    // two sources along K through im2col maps (as dh_conv1x1_dual_tc)
    p.conv = 1; p.HoWo = H * W; p.Wo = W; p.c_chunks = C1 / BK; p.kw = 1; p.stride = 1; p.pad = 0;
    p.k1_chunks = C1 / BK; p.stride2 = stride2;
    rc = make_map_im2col(&ma, y2, n, H, W, C1, 1, 1, 1, 0, dtype);
    if (rc) return rc;
    rc = make_map_im2col(&ma2, x2, n, H2, W2, C2, 1, 1, stride2, 0, dtype);
    if (rc) return rc;
  } else {
    p.res = x2; p.ldr = Cout; p.res_dtype = dtype;
    rc = make_map_2d(&ma, y2, M, C1, C1, BM, dtype);
    if (rc) return rc;
  }
  tl_chain = &cm;
  rc = dispatch(ma, w, p.K, p, 256, stream, x2_is_source ? &ma2 : nullptr);
  tl_chain = nullptr;
  if (rc) return rc;
  return p.tma_store ? DH_OK : dh_fail(DH_ERR_ARG, "chained convolution needs the TMA-store path", __FILE__, __LINE__);
}

// Non-zero after a watchdog trap inside gemm_tc_kernel: 1 producer, 2 MMA/accumulator, 3 MMA/operands, 4 epilogue.
extern "C" int dh_tc_error_flag(int* out_host) {
  DH_ARG(out_host);
  *out_host = 0;
  if (!g_error_flag) return DH_OK;
  DH_CUDA(cudaMemcpy(out_host, g_error_flag, sizeof(int), cudaMemcpyDeviceToHost));
  return DH_OK;
}
