// Multi-head attention for the transformer decoders (reference: models/transformers.py:82-129).
//
// One warp per (query row, head).  The same kernel serves
//   * incremental masked self-attention over the KV cache with beam indirection (slot table `src`),
//   * cross-attention over the 49 cached spatial K/V rows of the row's image (shared by its beams),
//   * teacher-forced causal self-/cross-attention over whole sequences (config 3 / forward()).
// Scores: lanes over keys, each lane reads one contiguous head row of K (16 B vector loads);
// output: lanes over head dims, V rows read coalesced.  Masked keys get -1e8 (not -inf) as in the
// reference (:111); keys beyond the causal horizon contribute exactly 0 there and are skipped here.
#include <stdlib.h>

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

// 16-byte vector of T -> VEC floats
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* out) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* out) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[2 * i] = __uint_as_float(w[i] << 16);
      out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// One warp per (query row, head).  Scores: the warp is split into groups of LPK = hd / VEC lanes; a group reads one
// key's head row with 16-byte loads (a whole 128-byte row per group for hd = 64 bf16, fully coalesced) and reduces
// its dot product with shuffles, 32 / LPK keys per iteration.  Output: lanes over pairs of head dims, one 4/8-byte
// load per key (coalesced).  The physical K/V row of every key (beam slot indirection) is resolved once per key.
template <typename T>
__global__ void __launch_bounds__(kWarps * 32) attn_kernel(AttnParams p) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float s_p[kWarps][kMaxKeys];
  __shared__ long long s_base[kWarps][kMaxKeys];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = p.D / p.n_heads;
  const long long item = (long long)blockIdx.x * kWarps + w;
  if (item >= (long long)p.R * p.n_heads) return;
  const int r = (int)(item / p.n_heads), h = (int)(item % p.n_heads);
  const int img = r / p.rpi, b = r % p.rpi;
  const int nk = p.causal_full ? (b + 1) : p.n_keys;
  const T* Kb = (const T*)p.K;
  const T* Vb = (const T*)p.V;
  for (int t = lane; t < nk; t += 32) {
    const int slot = p.slot_shared ? 0 : (p.src ? p.src[((long long)img * p.rpi + b) * p.S_alloc + t] : b);
    s_base[w][t] = (((long long)img * p.slots + slot) * p.S_alloc + t) * p.D + h * hd;
  }
  __syncwarp();
  const int lpk = hd / VEC;                    // lanes per key (power of two <= 32, checked on the host)
  const int kpi = 32 / lpk;                    // keys per iteration
  const int gl = lane % lpk, gk = lane / lpk;
  float qv[VEC];
  Vec16<T>::load((const T*)p.q + (long long)r * p.ldq + h * hd + gl * VEC, qv);
  const int* seq_row = p.seq ? p.seq + (long long)(p.seq_per_image ? img : r) * p.seq_ld : nullptr;
  float mx = -INFINITY;
  for (int t0 = 0; t0 < nk; t0 += kpi) {
    const int t = t0 + gk;
    float acc = 0.f;
    if (t < nk) {
      float kv[VEC];
      Vec16<T>::load(Kb + s_base[w][t] + gl * VEC, kv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc = fmaf(qv[i], kv[i], acc);
    }
    for (int o = lpk >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (t < nk && gl == 0) {
      float e = acc / p.scale;
      bool masked = false;
      if (seq_row && t >= 1) masked = (seq_row[t - 1] == p.pad);
      if (p.enc_mask) masked = p.enc_mask[(long long)img * p.S_alloc + t] != 0;
      if (masked) e = -1e8f;
      s_p[w][t] = e;
    }
  }
  __syncwarp();
  for (int t = lane; t < nk; t += 32) mx = fmaxf(mx, s_p[w][t]);
  mx = dh_warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < nk; t += 32) {
    const float e = expf(s_p[w][t] - mx);
    s_p[w][t] = e;
    sum += e;
  }
  sum = dh_warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  T* o = (T*)p.out + (long long)r * p.ldo + h * hd;
  for (int d = 2 * lane; d < hd; d += 64) {     // hd is even (multiple of VEC)
    float a0 = 0.f, a1 = 0.f;
    for (int t = 0; t < nk; ++t) {
      const float pt = s_p[w][t];
      const T* vr = Vb + s_base[w][t] + d;
      a0 = fmaf(pt, dh_to_f<T>(vr[0]), a0);
      a1 = fmaf(pt, dh_to_f<T>(vr[1]), a1);
    }
    o[d] = dh_from_f<T>(a0 * inv);
    o[d + 1] = dh_from_f<T>(a1 * inv);
  }
}

// Incremental self-attention over the KV cache at D = 512 (8 heads of 64): one warp per query ROW, all heads at once.
// A cached K / V row is 1 KB contiguous, so the warp reads it with two fully coalesced 512-byte loads (lane l holds dims
// 8 l .. 8 l + 7 of heads l / 8 and 4 + l / 8); four keys are in flight per lane before the first is consumed, and the eight
// partial dot products a lane then holds (4 keys x 2 heads) are reduced across its 8-lane head group with a halving
// exchange (7 shuffles instead of 24).  The rows of a CTA are consecutive beams, mostly of one image, whose slot tables
// share their common ancestors' rows: L1 serves the re-reads.
constexpr int kRowWarps = 8;
constexpr int kRowMaxKeys = 160;
constexpr int kRowPitch = kRowMaxKeys + 1;      // scores [8 heads][keys], odd pitch: the head groups read distinct banks

__device__ __forceinline__ void bf16x8_to_f32(const uint4 v, float* out) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float(w[i] << 16);
    out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(kRowWarps * 32) attn_row_kernel(AttnParams p) {
  __shared__ float s_p[kRowWarps][8 * kRowPitch];
  __shared__ int s_off[kRowWarps][kRowMaxKeys];        // physical K / V row of key t; ~row when the key is masked
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRowWarps + w;
  if (r >= p.R) return;
  const int img = r / p.rpi, b = r % p.rpi;
  const int nk = p.n_keys, D = 512;
  const __nv_bfloat16* Kb = (const __nv_bfloat16*)p.K;
  const __nv_bfloat16* Vb = (const __nv_bfloat16*)p.V;
  float* sp = s_p[w];
  int* soff = s_off[w];
  const int* seq_row = p.seq ? p.seq + (long long)(p.seq_per_image ? img : r) * p.seq_ld : nullptr;
  // physical row of every key (beam slot indirection); the sign bit carries the pad-key / encoder mask
  for (int t = lane; t < nk; t += 32) {
    const int slot = p.slot_shared ? 0 : (p.src ? p.src[((long long)img * p.rpi + b) * p.S_alloc + t] : b);
    const int off = (int)(((long long)img * p.slots + slot) * p.S_alloc + t);      // < 2^31 rows (checked on the host)
    bool masked = false;
    if (seq_row && t >= 1) masked = (seq_row[t - 1] == p.pad);
    if (p.enc_mask) masked = p.enc_mask[(long long)img * p.S_alloc + t] != 0;
    soff[t] = masked ? ~off : off;
  }
  float qa[8], qb[8];
  {
    const __nv_bfloat16* qr = (const __nv_bfloat16*)p.q + (long long)r * p.ldq + lane * 8;
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(qr)), qa);
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(qr + 256)), qb);
  }
  __syncwarp();
  const int grp = lane >> 3;                   // head group: heads grp and 4 + grp
  // ---- scores: 4 keys per iteration
  for (int t0 = 0; t0 < nk; t0 += 4) {
    uint4 ka[4], kb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + j;
      if (t < nk) {
        const int row = soff[t];
        const long long off = (long long)(row < 0 ? ~row : row) * D;
        ka[j] = __ldg(reinterpret_cast<const uint4*>(Kb + off + lane * 8));
        kb[j] = __ldg(reinterpret_cast<const uint4*>(Kb + off + 256 + lane * 8));
      } else {
        ka[j] = kb[j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float v[8];                                // v[2 j] = key j, head grp;  v[2 j + 1] = key j, head 4 + grp
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ka_f[8], kb_f[8];
      bf16x8_to_f32(ka[j], ka_f);
      bf16x8_to_f32(kb[j], kb_f);
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { sa = fmaf(qa[i], ka_f[i], sa); sb = fmaf(qb[i], kb_f[i], sb); }
      v[2 * j] = sa; v[2 * j + 1] = sb;
    }
    // halving exchange over the 8 lanes of the head group: afterwards lane i of the group holds the total of v[i]
    {
      const bool up = lane & 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = up ? v[i] : v[4 + i], keep = up ? v[4 + i] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    }
    {
      const bool up = lane & 2;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = up ? v[i] : v[2 + i], keep = up ? v[2 + i] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
    }
    {
      const bool up = lane & 1;
      const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
      v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    const int idx = lane & 7, t = t0 + (idx >> 1), head = (idx & 1) * 4 + grp;
    if (t < nk) sp[head * kRowPitch + t] = soff[t] < 0 ? -1e8f : v[0] / p.scale;
  }
  __syncwarp();
  // ---- softmax per head: 4 lanes per head
  {
    const int head = lane >> 2, l4 = lane & 3;
    float* ph = sp + head * kRowPitch;
    float mx = -INFINITY;
    for (int t = l4; t < nk; t += 4) mx = fmaxf(mx, ph[t]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
    for (int t = l4; t < nk; t += 4) {
      const float e = expf(ph[t] - mx);
      ph[t] = e;
      sum += e;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (l4 == 0) ph[kRowMaxKeys] = 1.f / sum;           // the pad column of the row holds the normaliser
  }
  __syncwarp();
  // ---- output: lane l accumulates dims 8 l .. 8 l + 7 of head grp and of head 4 + grp
  float oa[8], ob[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) oa[i] = ob[i] = 0.f;
  const float* pa = sp + grp * kRowPitch;
  const float* pb = sp + (4 + grp) * kRowPitch;
  for (int t0 = 0; t0 < nk; t0 += 4) {
    uint4 va[4], vb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + j;
      if (t < nk) {
        const int row = soff[t];
        const long long off = (long long)(row < 0 ? ~row : row) * D;
        va[j] = __ldg(reinterpret_cast<const uint4*>(Vb + off + lane * 8));
        vb[j] = __ldg(reinterpret_cast<const uint4*>(Vb + off + 256 + lane * 8));
      } else {
        va[j] = vb[j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + j;
      const float wa = t < nk ? pa[t] : 0.f, wb = t < nk ? pb[t] : 0.f;
      float fa[8], fb[8];
      bf16x8_to_f32(va[j], fa);
      bf16x8_to_f32(vb[j], fb);
#pragma unroll
      for (int i = 0; i < 8; ++i) { oa[i] = fmaf(wa, fa[i], oa[i]); ob[i] = fmaf(wb, fb[i], ob[i]); }
    }
  }
  const float ia = pa[kRowMaxKeys], ib = pb[kRowMaxKeys];
  uint32_t wa[4], wb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 x = __floats2bfloat162_rn(oa[2 * i] * ia, oa[2 * i + 1] * ia);
    __nv_bfloat162 y = __floats2bfloat162_rn(ob[2 * i] * ib, ob[2 * i + 1] * ib);
    wa[i] = *reinterpret_cast<uint32_t*>(&x);
    wb[i] = *reinterpret_cast<uint32_t*>(&y);
  }
  __nv_bfloat16* o = (__nv_bfloat16*)p.out + (long long)r * p.ldo + lane * 8;
  *reinterpret_cast<uint4*>(o) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
  *reinterpret_cast<uint4*>(o + 256) = make_uint4(wb[0], wb[1], wb[2], wb[3]);
}

// Cross-attention over keys shared by all rows of an image (the 49 spatial tokens, transformers.py:100-121 via :366):
// one CTA per (image, head), one warp per row of the image.  The head slice of K and V ([n_keys, hd] each) is staged
// in shared memory once and reused by the image's rpi beam rows -- rpi times less L2 traffic than one warp per
// (row, head) fetching it alone.
template <typename T>
__global__ void __launch_bounds__(32 * 16) attn_shared_kv_kernel(AttnParams p) {
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ __align__(16) unsigned char smem_attn[];
  const int hd = p.D / p.n_heads, nk = p.n_keys;
  T* sK = reinterpret_cast<T*>(smem_attn);
  T* sV = sK + (size_t)nk * hd;
  float* s_p = reinterpret_cast<float*>(sV + (size_t)nk * hd);     // [rpi][nk]
  const int img = blockIdx.x / p.n_heads, h = blockIdx.x % p.n_heads;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long kv0 = ((long long)img * p.slots * p.S_alloc) * p.D + h * hd;      // slot 0, key 0
  const int cpr = hd / VEC;                                                        // 16-byte chunks per key row
  for (int i = threadIdx.x; i < nk * cpr; i += blockDim.x) {
    const int t = i / cpr, c = i - t * cpr;
    *reinterpret_cast<uint4*>(sK + t * hd + c * VEC) = *reinterpret_cast<const uint4*>((const T*)p.K + kv0 + (long long)t * p.D + c * VEC);
    *reinterpret_cast<uint4*>(sV + t * hd + c * VEC) = *reinterpret_cast<const uint4*>((const T*)p.V + kv0 + (long long)t * p.D + c * VEC);
  }
  __syncthreads();
  const int r = img * p.rpi + w;
  if (w >= p.rpi || r >= p.R) return;
  float* sp = s_p + w * nk;
  const int lpk = cpr, kpi = 32 / lpk;
  const int gl = lane % lpk, gk = lane / lpk;
  float qv[VEC];
  Vec16<T>::load((const T*)p.q + (long long)r * p.ldq + h * hd + gl * VEC, qv);
  for (int t0 = 0; t0 < nk; t0 += kpi) {
    const int t = t0 + gk;
    float acc = 0.f;
    if (t < nk) {
      float kv[VEC];
      Vec16<T>::load(sK + t * hd + gl * VEC, kv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc = fmaf(qv[i], kv[i], acc);
    }
    for (int o = lpk >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (t < nk && gl == 0) {
      float e = acc / p.scale;
      if (p.enc_mask && p.enc_mask[(long long)img * p.S_alloc + t] != 0) e = -1e8f;
      sp[t] = e;
    }
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int t = lane; t < nk; t += 32) mx = fmaxf(mx, sp[t]);
  mx = dh_warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < nk; t += 32) {
    const float e = expf(sp[t] - mx);
    sp[t] = e;
    sum += e;
  }
  sum = dh_warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  T* o = (T*)p.out + (long long)r * p.ldo + h * hd;
  for (int d = 2 * lane; d < hd; d += 64) {
    float a0 = 0.f, a1 = 0.f;
    for (int t = 0; t < nk; ++t) {
      const float pt = sp[t];
      a0 = fmaf(pt, dh_to_f<T>(sV[t * hd + d]), a0);
      a1 = fmaf(pt, dh_to_f<T>(sV[t * hd + d + 1]), a1);
    }
    o[d] = dh_from_f<T>(a0 * inv);
    o[d + 1] = dh_from_f<T>(a1 * inv);
  }
}

// ------------------------------------------------------------------------------------------------ tensor-core path
// Attention whose keys are shared by all query rows of an image -- cross-attention over the 49 spatial tokens (decode:
// the image's beam rows; teacher-forced: its S positions) and teacher-forced causal self-attention -- as two small
// matmuls per (image, head) on mma.sync m16n8k16 (bf16 in, fp32 accumulate): S = Q K^T, softmax on the accumulator
// fragments, O = P V with P re-used from registers as the A operand.  head_dim 64, <= 64 keys, <= 64 query rows per image.
// (tcgen05 needs M >= 64 rows per instruction; a 5-row x 49-key problem per head fits the warp-level MMA instead.)
constexpr int kMmaPitch = 72;        // bf16 elements per staged row (64 + 8: conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int NKT>      // NKT 8-key tiles (even): keys padded to 8 * NKT <= 64
__global__ void __launch_bounds__(128) attn_mma_kernel(AttnParams p) {
  constexpr int NKP = NKT * 8;
  __shared__ __align__(16) __nv_bfloat16 sK[NKP * kMmaPitch];
  __shared__ __align__(16) __nv_bfloat16 sV[NKP * kMmaPitch];
  __shared__ __align__(16) __nv_bfloat16 sQ[64 * kMmaPitch];
  const int img = blockIdx.x / p.n_heads, h = blockIdx.x % p.n_heads;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = p.causal_full ? p.rpi : p.n_keys;
  const int q_rows = ((p.rpi + 15) / 16) * 16;                         // staged query rows (<= 64)
  const __nv_bfloat16* Kg = (const __nv_bfloat16*)p.K + ((long long)img * p.slots * p.S_alloc) * p.D + h * 64;
  const __nv_bfloat16* Vg = (const __nv_bfloat16*)p.V + ((long long)img * p.slots * p.S_alloc) * p.D + h * 64;
  const __nv_bfloat16* Qg = (const __nv_bfloat16*)p.q + ((long long)img * p.rpi) * p.ldq + h * 64;
  // staging by all 128 threads, loads first (all in flight), then the shared-memory stores
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  constexpr int kKV = NKP * 8 / 128;                       // 16-byte chunks per thread per array (NKP * 8 chunks)
  uint4 rk[kKV], rv[kKV];
#pragma unroll
  for (int u = 0; u < kKV; ++u) {
    const int i = threadIdx.x + u * 128, t = i >> 3, c = i & 7;
    rk[u] = t < nk ? __ldg(reinterpret_cast<const uint4*>(Kg + (long long)t * p.D + c * 8)) : zero;
    rv[u] = t < nk ? __ldg(reinterpret_cast<const uint4*>(Vg + (long long)t * p.D + c * 8)) : zero;
  }
  uint4 rq[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {                            // up to 64 query rows x 8 chunks
    const int i = threadIdx.x + u * 128, r = i >> 3, c = i & 7;
    rq[u] = (i < q_rows * 8 && r < p.rpi) ? __ldg(reinterpret_cast<const uint4*>(Qg + (long long)r * p.ldq + c * 8)) : zero;
  }
#pragma unroll
  for (int u = 0; u < kKV; ++u) {
    const int i = threadIdx.x + u * 128, t = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sK + t * kMmaPitch + c * 8) = rk[u];
    *reinterpret_cast<uint4*>(sV + t * kMmaPitch + c * 8) = rv[u];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.x + u * 128, r = i >> 3, c = i & 7;
    if (i < q_rows * 8) *reinterpret_cast<uint4*>(sQ + r * kMmaPitch + c * 8) = rq[u];
  }
  __syncthreads();
  if (w * 16 >= q_rows) return;                            // staging-only warps
  const uint32_t sK_u = (uint32_t)__cvta_generic_to_shared(sK), sV_u = (uint32_t)__cvta_generic_to_shared(sV);
  const uint32_t sQ_u = (uint32_t)__cvta_generic_to_shared(sQ);
  const int g = lane >> 2, t4 = lane & 3;
  const int mat = lane >> 3, l8 = lane & 7;

  // ---- S = Q K^T (16 query rows of this warp x NKP keys)
  float sacc[NKT][4];
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    ldsm_x4(sQ_u + (uint32_t)(((w * 16 + (lane & 15)) * kMmaPitch + kk * 16 + (lane >> 4) * 8) * 2), a);
#pragma unroll
    for (int n2 = 0; n2 < NKT / 2; ++n2) {
      uint32_t b[4];     // (keys 0-7, dims lo), (keys 0-7, dims hi), (keys 8-15, dims lo), (keys 8-15, dims hi)
      ldsm_x4(sK_u + (uint32_t)(((n2 * 16 + (mat >> 1) * 8 + l8) * kMmaPitch + kk * 16 + (mat & 1) * 8) * 2), b);
      mma_bf16(sacc[2 * n2], a, b[0], b[1]);
      mma_bf16(sacc[2 * n2 + 1], a, b[2], b[3]);
    }
  }
  // ---- scale, masks (excluded keys: beyond nk / the causal horizon; masked keys: -1e8 as in the reference), softmax
  const int row0 = w * 16 + g, row1 = row0 + 8;
  const int* seq_row = p.seq ? p.seq + (long long)img * p.seq_ld : nullptr;
  const unsigned char* em = p.enc_mask ? p.enc_mask + (long long)img * p.S_alloc : nullptr;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = nt * 8 + 2 * t4 + (e & 1);
      const int row = e < 2 ? row0 : row1;
      float v = sacc[nt][e] / p.scale;
      if (col < nk) {
        bool masked = false;
        if (seq_row && col >= 1) masked = (seq_row[col - 1] == p.pad);
        if (em) masked = em[col] != 0;
        if (masked) v = -1e8f;
      }
      if (col >= nk || (p.causal_full && col > row)) v = -INFINITY;
      sacc[nt][e] = v;
      if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pv = expf(sacc[nt][e] - (e < 2 ? mx0 : mx1));      // rows past rpi hold zeros: harmless, never stored
      sacc[nt][e] = pv;
      if (e < 2) sum0 += pv; else sum1 += pv;
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  // ---- O = P V: the probability fragments are the A operand as they are (two 8-key tiles per 16-key step)
  float oacc[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) { oacc[nd][0] = oacc[nd][1] = oacc[nd][2] = oacc[nd][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < NKT / 2; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);
    a[1] = pack_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
    a[2] = pack_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
    a[3] = pack_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
    for (int d2 = 0; d2 < 4; ++d2) {
      uint32_t b[4];     // transposed loads: (keys lo, dims 0-7), (keys hi, dims 0-7), (keys lo, dims 8-15), (keys hi, dims 8-15)
      ldsm_x4_trans(sV_u + (uint32_t)(((kk * 16 + (mat & 1) * 8 + l8) * kMmaPitch + d2 * 16 + (mat >> 1) * 8) * 2), b);
      mma_bf16(oacc[2 * d2], a, b[0], b[1]);
      mma_bf16(oacc[2 * d2 + 1], a, b[2], b[3]);
    }
  }
  const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
  __nv_bfloat16* Og = (__nv_bfloat16*)p.out + ((long long)img * p.rpi) * p.ldo + h * 64;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) {
    const int col = nd * 8 + 2 * t4;
    if (row0 < p.rpi) *reinterpret_cast<uint32_t*>(Og + (long long)row0 * p.ldo + col) = pack_bf16(oacc[nd][0] * inv0, oacc[nd][1] * inv0);
    if (row1 < p.rpi) *reinterpret_cast<uint32_t*>(Og + (long long)row1 * p.ldo + col) = pack_bf16(oacc[nd][2] * inv1, oacc[nd][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------------ streaming path
// Cross-attention during generation is a pure HBM stream: every decode step re-reads the 49 cached spatial K/V rows of every
// image (2 x 49 x D bf16 = 100 KB per image and layer, far more than L2 holds for a batch), and does ~1 MFLOP per image on
// them.  One persistent CTA per SM walks the images; a producer warp fetches an image's K and V rows with 1-D bulk copies
// (cp.async.bulk, one per key row, completion on an mbarrier) into a ring of stages while eight consumer warps -- one per
// head -- run the two small mma.sync products of the previous image.  Rows are staged with a pitch of D + 8 elements so
// that ldmatrix is conflict-free; key rows past n_keys read a shared zero row.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();      // a protocol bug must fail the launch, not hang the GPU
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t movm_trans(uint32_t x) {     // 8x8 b16 transpose across the warp
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

constexpr int kStreamWarps = 8;       // consumer warps (one head each)
constexpr int kLoadWarps = 4;         // producer warps: a bulk copy is issued lane by lane (~60 cycles each), so the ~100
                                      // copies of an image are spread over four warps to stay ahead of HBM

template <int NKT>
__global__ void __launch_bounds__((kStreamWarps + kLoadWarps) * 32, 1) attn_stream_kernel(AttnParams p, int n_img, int stages) {
  extern __shared__ __align__(128) unsigned char smem_stream[];
  const int nk = p.n_keys, D = p.D;
  const uint32_t pitch = (uint32_t)(D + 8) * 2u;                 // bytes per staged key row
  const uint32_t half_stage = (uint32_t)nk * pitch;              // K rows, then V rows
  const uint32_t stage_bytes = 2u * half_stage;
  const uint32_t s_base = smem_addr(smem_stream);
  const uint32_t zero_row = s_base + (uint32_t)stages * stage_bytes;                   // pitch bytes of zeros
  const uint32_t bars = zero_row + ((pitch + 15u) & ~15u);                             // full[stages], empty[stages]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x * 4u; i < pitch; i += blockDim.x * 4u)
    *reinterpret_cast<uint32_t*>(smem_stream + (size_t)stages * stage_bytes + i) = 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init_(bars + 8u * s, 1);
      mbar_init_(bars + 8u * (stages + s), kStreamWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long img_stride = (long long)p.slots * p.S_alloc * D;      // elements between images in K / V (slot 0 is used)
  if (w >= kStreamWarps) {
    // ===================================================================== producers: one bulk copy per key row
    const int pw = w - kStreamWarps;
    int stage = 0;
    uint32_t phase = 0;
    for (int img = blockIdx.x; img < n_img; img += gridDim.x) {
      mbar_wait_(bars + 8u * (stages + stage), phase ^ 1u);
      // the transaction count may run ahead of this expect_tx (other producer warps): the phase cannot complete before
      // the single arrival it carries
      if (pw == 0 && lane == 0) mbar_expect_tx_(bars + 8u * stage, 2u * (uint32_t)nk * (uint32_t)D * 2u);
      const __nv_bfloat16* Kg = (const __nv_bfloat16*)p.K + img * img_stride;
      const __nv_bfloat16* Vg = (const __nv_bfloat16*)p.V + img * img_stride;
      const uint32_t dst = s_base + (uint32_t)stage * stage_bytes;
      for (int t = pw + kLoadWarps * lane; t < nk; t += kLoadWarps * 32) {        // rows interleaved over the warps
        bulk_load(dst + (uint32_t)t * pitch, Kg + (long long)t * D, (uint32_t)D * 2u, bars + 8u * stage);
        bulk_load(dst + half_stage + (uint32_t)t * pitch, Vg + (long long)t * D, (uint32_t)D * 2u, bars + 8u * stage);
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    return;
  }
  // ======================================================================= consumers: warp w owns heads w, w + 8, ...
  // The products are taken TRANSPOSED, S^T = K Q^T and O^T = V^T P^T, so that the 16-wide M side of mma.m16n8k16 runs over
  // keys / head dims (always full) and the 8-wide N side over the image's query rows (beam <= 8 in one pass): half the MMAs
  // and half the softmax work of the row-major form, whose 16 query rows would be 2/3 padding at beam 5.
  const int g = lane >> 2, t4 = lane & 3;
  const int mat = lane >> 3, l8 = lane & 7;
  constexpr int MT = NKT / 2;                                     // 16-key tiles
  const int q_groups = (p.rpi + 7) / 8;
  const float inv_scale = 1.f / p.scale;
  int stage = 0;
  uint32_t phase = 0;
  for (int img = blockIdx.x; img < n_img; img += gridDim.x) {
    // the image's encoder-key mask as a 64-bit word (bit t: key t masked), fetched before the wait on the stage
    unsigned long long emask = 0ull;
    if (p.enc_mask) {
      const unsigned char* em = p.enc_mask + (long long)img * p.S_alloc;
      const unsigned lo = __ballot_sync(0xffffffffu, lane < nk && em[lane] != 0);
      const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < nk && em[lane + 32] != 0);
      emask = ((unsigned long long)hi << 32) | lo;
    }
    const uint32_t sK_u = s_base + (uint32_t)stage * stage_bytes, sV_u = sK_u + half_stage;
    bool waited = false;
    for (int h = w; h < p.n_heads; h += kStreamWarps) {
      for (int qg = 0; qg < q_groups; ++qg) {
        // ---- Q^T as the B operand (dims x 8 queries), straight from global memory; queries past rpi are zero
        const int qrow = qg * 8 + g;
        const __nv_bfloat16* Qg = (const __nv_bfloat16*)p.q + ((long long)img * p.rpi + qrow) * p.ldq + h * 64 + 2 * t4;
        uint32_t qb[4][2];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          qb[kk][0] = qrow < p.rpi ? __ldg(reinterpret_cast<const unsigned int*>(Qg + kk * 16)) : 0u;
          qb[kk][1] = qrow < p.rpi ? __ldg(reinterpret_cast<const unsigned int*>(Qg + kk * 16 + 8)) : 0u;
        }
        if (!waited) { mbar_wait_(bars + 8u * stage, phase); waited = true; }
        // ---- S^T = K Q^T: sacc[mt] = {(key g, query 2 t4), (key g, query 2 t4 + 1), (key g + 8, ...), (key g + 8, ...)}
        float sacc[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { sacc[mt][0] = sacc[mt][1] = sacc[mt][2] = sacc[mt][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int key = mt * 16 + (mat & 1) * 8 + l8;
            uint32_t a[4];     // (keys 0-7, dims lo), (keys 8-15, dims lo), (keys 0-7, dims hi), (keys 8-15, dims hi)
            ldsm_x4((key < nk ? sK_u + (uint32_t)key * pitch : zero_row) + (uint32_t)(h * 64 + kk * 16 + (mat >> 1) * 8) * 2u, a);
            mma_bf16(sacc[mt], a, qb[kk][0], qb[kk][1]);
          }
        }
        // ---- scale, any-zero encoder mask (-1e8 as in the reference, transformers.py:111), softmax over the keys: a
        // query's scores sit in the lanes of equal t4 (8 values of g) x MT x 2 registers
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int hi = 0; hi < 2; ++hi) {
            const int key = mt * 16 + hi * 8 + g;
            const bool masked = (emask >> key) & 1ull, oob = key >= nk;
            float v0 = sacc[mt][2 * hi] * inv_scale, v1 = sacc[mt][2 * hi + 1] * inv_scale;
            v0 = oob ? -INFINITY : (masked ? -1e8f : v0);
            v1 = oob ? -INFINITY : (masked ? -1e8f : v1);
            sacc[mt][2 * hi] = v0; sacc[mt][2 * hi + 1] = v1;
            mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
          }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
        }
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int hi = 0; hi < 2; ++hi) {
            const float p0 = __expf(sacc[mt][2 * hi] - mx0), p1 = __expf(sacc[mt][2 * hi + 1] - mx1);
            sacc[mt][2 * hi] = p0; sacc[mt][2 * hi + 1] = p1;
            sum0 += p0; sum1 += p1;
          }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          sum0 += __shfl_xor_sync(0xffffffffu, sum0, o);
          sum1 += __shfl_xor_sync(0xffffffffu, sum1, o);
        }
        // ---- O^T = V^T P^T: the probabilities become the B operand (16 keys x 8 queries) after an 8x8 transpose of the
        // accumulator fragments (movmatrix); V^T fragments come from transposing ldmatrix loads of V[key][dim]
        float oacc[4][4];
#pragma unroll
        for (int md = 0; md < 4; ++md) { oacc[md][0] = oacc[md][1] = oacc[md][2] = oacc[md][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < MT; ++kk) {
          const uint32_t b0 = movm_trans(pack_bf16(sacc[kk][0], sacc[kk][1]));
          const uint32_t b1 = movm_trans(pack_bf16(sacc[kk][2], sacc[kk][3]));
          const int key = kk * 16 + (mat >> 1) * 8 + l8;
          const uint32_t row_u = key < nk ? sV_u + (uint32_t)key * pitch : zero_row;
#pragma unroll
          for (int md = 0; md < 4; ++md) {
            uint32_t a[4];     // (dims 0-7, keys lo), (dims 8-15, keys lo), (dims 0-7, keys hi), (dims 8-15, keys hi)
            ldsm_x4_trans(row_u + (uint32_t)(h * 64 + md * 16 + (mat & 1) * 8) * 2u, a);
            mma_bf16(oacc[md], a, b0, b1);
          }
        }
        // ---- back to [query][dim] (movmatrix again) and out: lane (g, t4) stores query g, dims 16 md + 2 t4 (+ 8)
        const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
        __nv_bfloat16* Og = (__nv_bfloat16*)p.out + ((long long)img * p.rpi + qrow) * p.ldo + h * 64 + 2 * t4;
#pragma unroll
        for (int md = 0; md < 4; ++md) {
          const uint32_t lo = movm_trans(pack_bf16(oacc[md][0] * inv0, oacc[md][1] * inv1));
          const uint32_t hi = movm_trans(pack_bf16(oacc[md][2] * inv0, oacc[md][3] * inv1));
          if (qrow < p.rpi) {
            *reinterpret_cast<uint32_t*>(Og + md * 16) = lo;
            *reinterpret_cast<uint32_t*>(Og + md * 16 + 8) = hi;
          }
        }
      }
    }
    if (!waited) mbar_wait_(bars + 8u * stage, phase);          // a warp without a head still follows the ring
    __syncwarp();
    if (lane == 0) mbar_arrive_(bars + 8u * (stages + stage));    // every ldmatrix of this warp on the stage has retired
    if (++stage == stages) { stage = 0; phase ^= 1u; }
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
  {
    const int vec = dtype == DH_F32 ? 4 : 8, hd = D / n_heads, lpk = hd / vec;
    DH_ARG(hd % vec == 0 && lpk >= 1 && lpk <= 32 && (lpk & (lpk - 1)) == 0 && ldq % vec == 0);   // 16-byte head rows
    DH_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)K % 16) == 0 && ((uintptr_t)V % 16) == 0);
  }
  DH_ARG(rows_per_image > 0 && slots > 0 && S_alloc > 0);
  DH_ARG((causal_full ? rows_per_image : n_keys) <= kMaxKeys && (causal_full || (n_keys > 0 && n_keys <= S_alloc)));
  if (rows == 0) return DH_OK;
  AttnParams p{q, ldq, K, V, out, ldo, rows, D, n_heads, rows_per_image, slots, S_alloc, src, slot_shared,
               n_keys, causal_full, seq, seq_ld, seq_per_image, pad, enc_mask, scale};
  // cross-attention over an image's cached K/V rows (generation and teacher-forced): persistent bulk-copy pipeline
  {
    static const bool stream_ok = !getenv("DH_NO_STREAM_ATTN");
    static int n_sms = 0, smem_max = 0;
    if (stream_ok && dtype == DH_BF16 && D / n_heads == 64 && D % 8 == 0 && slot_shared && !causal_full && !seq && !src &&
        n_keys <= 64 && rows_per_image <= 64 && rows % rows_per_image == 0 && ldo % 2 == 0 && ((uintptr_t)out % 4) == 0 &&
        ((uintptr_t)q % 4) == 0 && ldq % 2 == 0) {
      if (!n_sms) {
        int dev = 0;
        DH_CUDA(cudaGetDevice(&dev));
        DH_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        DH_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
      }
      const int n_img = rows / rows_per_image;
      const size_t pitch = (size_t)(D + 8) * 2, stage_bytes = 2 * (size_t)n_keys * pitch;
      const size_t fixed = ((pitch + 15) & ~(size_t)15) + 16 * 8 + 128;
      int stages = (int)(((size_t)smem_max - fixed) / stage_bytes);
      if (stages > 4) stages = 4;
      if (stages >= 2) {
        const size_t smem = (size_t)stages * stage_bytes + fixed;
        const int grid = n_img < n_sms ? n_img : n_sms;
        const int threads = (kStreamWarps + kLoadWarps) * 32;
#define DH_STREAM_LAUNCH(NKT)                                                                                          \
  do {                                                                                                                 \
    static bool attr = false;                                                                                          \
    if (!attr) {                                                                                                       \
      DH_CUDA(cudaFuncSetAttribute(attn_stream_kernel<NKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));   \
      attr = true;                                                                                                     \
    }                                                                                                                  \
    attn_stream_kernel<NKT><<<grid, threads, smem, s>>>(p, n_img, stages);                                             \
  } while (0)
        if (n_keys <= 16) DH_STREAM_LAUNCH(2);
        else if (n_keys <= 32) DH_STREAM_LAUNCH(4);
        else if (n_keys <= 48) DH_STREAM_LAUNCH(6);
        else DH_STREAM_LAUNCH(8);
#undef DH_STREAM_LAUNCH
        DH_LAUNCH_OK();
        return DH_OK;
      }
    }
  }
  // keys shared by the rows of an image, bf16, head_dim 64, <= 64 keys and query rows per image: tensor-core path
  {
    const int nk = causal_full ? rows_per_image : n_keys;
    static const bool mma_ok = !getenv("DH_NO_MMA_ATTN");
    if (mma_ok && dtype == DH_BF16 && D / n_heads == 64 && slot_shared && nk <= 64 && rows_per_image <= 64 &&
        rows % rows_per_image == 0 && (!seq || seq_per_image) && ldo % 2 == 0 && ((uintptr_t)out % 4) == 0) {
      const int g = (rows / rows_per_image) * n_heads;
      const int threads = 128;                 // 4 warps stage K / V / Q; ceil(rows_per_image / 16) of them compute
      if (nk <= 16) attn_mma_kernel<2><<<g, threads, 0, s>>>(p);
      else if (nk <= 32) attn_mma_kernel<4><<<g, threads, 0, s>>>(p);
      else if (nk <= 48) attn_mma_kernel<6><<<g, threads, 0, s>>>(p);
      else attn_mma_kernel<8><<<g, threads, 0, s>>>(p);
      DH_LAUNCH_OK();
      return DH_OK;
    }
  }
  // keys shared by the rows of an image (cross-attention during generation): stage K/V once per (image, head)
  const int esize = dtype == DH_F32 ? 4 : 2;
  const size_t shared_smem = (size_t)2 * n_keys * (D / n_heads) * esize + (size_t)rows_per_image * n_keys * 4;
  if (slot_shared && !causal_full && !seq && rows_per_image >= 2 && rows_per_image <= 16 && rows % rows_per_image == 0 &&
      shared_smem <= 48 * 1024) {
    const int g = (rows / rows_per_image) * n_heads;
    if (dtype == DH_F32) attn_shared_kv_kernel<float><<<g, 32 * rows_per_image, shared_smem, s>>>(p);
    else if (dtype == DH_BF16) attn_shared_kv_kernel<__nv_bfloat16><<<g, 32 * rows_per_image, shared_smem, s>>>(p);
    else return dh_fail(DH_ERR_ARG, "dtype", __FILE__, __LINE__);
    DH_LAUNCH_OK();
    return DH_OK;
  }
  // incremental bf16 self-attention at D = 512 / 8 heads: one warp per row, whole 1 KB K / V rows per load
  {
    static const bool row_ok = !getenv("DH_NO_ROW_ATTN");
    if (row_ok && dtype == DH_BF16 && D == 512 && n_heads == 8 && !causal_full && n_keys <= kRowMaxKeys && ldq % 8 == 0 &&
        ldo % 8 == 0 && ((uintptr_t)out % 16) == 0 &&
        (long long)(rows / rows_per_image + 1) * slots * S_alloc < (1ll << 31)) {
      attn_row_kernel<<<dh_cdiv(rows, kRowWarps), kRowWarps * 32, 0, s>>>(p);
      DH_LAUNCH_OK();
      return DH_OK;
    }
  }
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
