// Multi-head attention for the transformer decoders (reference: models/transformers.py:82-129).
//
// One warp per (query row, head).  The same kernel serves
//   * incremental masked self-attention over the KV cache with beam indirection (slot table `src`),
//   * cross-attention over the 49 cached spatial K/V rows of the row's image (shared by its beams),
//   * teacher-forced causal self-/cross-attention over whole sequences (config 3 / forward()).
// Scores: lanes over keys, each lane reads one contiguous head row of K (16 B vector loads);
// output: lanes over head dims, V rows read coalesced.  Masked keys get -1e8 (not -inf) as in the
// reference (:111); keys beyond the causal horizon contribute exactly 0 there and are skipped here.
#include "common.cuh"

namespace {

constexpr int kMaxKeys = 160;
constexpr int kWarps = 8;

struct AttnParams {
  const void* q; long long ldq;
  const void* K; const void* V;     // [n_img * slots, S_alloc, D]
  void* out; long long ldo;
  int R, D, n_heads, rpi, slots, S_alloc;
  const int* src;                   // [n_img, rpi, S_alloc] physical slot of (beam, position), or null
  int slot_shared;                  // 1: slot 0 for every row of the image (cross / full), 0: slot = beam
  int n_keys;                       // fixed key count, or 0 with causal_full
  int causal_full;                  // teacher-forced: keys 0..(r % rpi)
  const int* seq; long long seq_ld; int seq_per_image;   // token history for the pad-key mask (self-attention)
  int pad;
  const unsigned char* enc_mask;    // [n_img, S_alloc], 1 = masked (cross-attention)
  float scale;
};

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) attn_kernel(AttnParams p) {
  __shared__ float s_q[kWarps][128];
  __shared__ float s_p[kWarps][kMaxKeys];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = p.D / p.n_heads;
  const long long item = (long long)blockIdx.x * kWarps + w;
  if (item >= (long long)p.R * p.n_heads) return;
  const int r = (int)(item / p.n_heads), h = (int)(item % p.n_heads);
  const int img = r / p.rpi, b = r % p.rpi;
  const int nk = p.causal_full ? (b + 1) : p.n_keys;
  const T* q = (const T*)p.q + (long long)r * p.ldq + h * hd;
  for (int d = lane; d < hd; d += 32) s_q[w][d] = dh_to_f<T>(q[d]);
  __syncwarp();
  const T* Kb = (const T*)p.K;
  const T* Vb = (const T*)p.V;
  auto kv_row = [&](int t) -> long long {
    int slot = p.slot_shared ? 0 : (p.src ? p.src[((long long)img * p.rpi + b) * p.S_alloc + t] : b);
    return (((long long)img * p.slots + slot) * p.S_alloc + t) * p.D + h * hd;
  };
  const int* seq_row = p.seq ? p.seq + (long long)(p.seq_per_image ? img : r) * p.seq_ld : nullptr;
  float mx = -INFINITY;
  for (int t = lane; t < nk; t += 32) {
    const T* kr = Kb + kv_row(t);
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(s_q[w][d], dh_to_f<T>(kr[d]), acc);
    float e = acc / p.scale;
    bool masked = false;
    if (seq_row && t >= 1) masked = (seq_row[t - 1] == p.pad);
    if (p.enc_mask) masked = p.enc_mask[(long long)img * p.S_alloc + t] != 0;
    if (masked) e = -1e8f;
    s_p[w][t] = e;
    mx = fmaxf(mx, e);
  }
  mx = dh_warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < nk; t += 32) {
    float e = expf(s_p[w][t] - mx);
    s_p[w][t] = e;
    sum += e;
  }
  sum = dh_warp_sum(sum);
  __syncwarp();
  T* o = (T*)p.out + (long long)r * p.ldo + h * hd;
  for (int d = lane; d < hd; d += 32) {
    float acc = 0.f;
    for (int t = 0; t < nk; ++t) acc = fmaf(s_p[w][t] / sum, dh_to_f<T>(Vb[kv_row(t) + d]), acc);
    o[d] = dh_from_f<T>(acc);
  }
}

// enc_mask[n, t] = any(spatial[n, t, :] == 0)   (transformers.py:480-481, Q16)
template <typename T>
__global__ void enc_mask_kernel(const T* __restrict__ x, unsigned char* __restrict__ mask, int rows, int D) {
  int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= rows) return;
  int any0 = 0;
  for (int c = lane; c < D; c += 32) any0 |= (dh_to_f<T>(x[(long long)warp * D + c]) == 0.f);
  any0 = __any_sync(0xffffffffu, any0);
  if (lane == 0) mask[warp] = (unsigned char)any0;
}

}  // namespace

extern "C" int dh_attention(const void* q, long long ldq, const void* K, const void* V, void* out, long long ldo, int rows,
                            int D, int n_heads, int rows_per_image, int slots, int S_alloc, const int* src,
                            int slot_shared, int n_keys, int causal_full, const int* seq, long long seq_ld,
                            int seq_per_image, int pad, const unsigned char* enc_mask, float scale, int dtype,
                            cudaStream_t s) {
  DH_ARG(q && K && V && out && rows >= 0 && n_heads > 0 && D % n_heads == 0 && D / n_heads <= 128);
  DH_ARG(rows_per_image > 0 && slots > 0 && S_alloc > 0);
  DH_ARG((causal_full ? rows_per_image : n_keys) <= kMaxKeys && (causal_full || (n_keys > 0 && n_keys <= S_alloc)));
  if (rows == 0) return DH_OK;
  AttnParams p{q, ldq, K, V, out, ldo, rows, D, n_heads, rows_per_image, slots, S_alloc, src, slot_shared,
               n_keys, causal_full, seq, seq_ld, seq_per_image, pad, enc_mask, scale};
  int grid = dh_cdiv((long long)rows * n_heads, kWarps);
  if (dtype == DH_F32) attn_kernel<float><<<grid, kWarps * 32, 0, s>>>(p);
  else if (dtype == DH_BF16) attn_kernel<__nv_bfloat16><<<grid, kWarps * 32, 0, s>>>(p);
  else return dh_fail(DH_ERR_ARG, "dtype", __FILE__, __LINE__);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_enc_mask(const void* spatial, unsigned char* mask, int rows, int D, int dtype, cudaStream_t s) {
  DH_ARG(spatial && mask && rows >= 0);
  if (rows == 0) return DH_OK;
  int grid = dh_cdiv((long long)rows * 32, 256);
  if (dtype == DH_F32) enc_mask_kernel<float><<<grid, 256, 0, s>>>((const float*)spatial, mask, rows, D);
  else if (dtype == DH_BF16) enc_mask_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)spatial, mask, rows, D);
  else return dh_fail(DH_ERR_ARG, "dtype", __FILE__, __LINE__);
  DH_LAUNCH_OK();
  return DH_OK;
}
