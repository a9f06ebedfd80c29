// Token selection and beam bookkeeping (reference: models/beam.py:32-108 and the generate() loops in
// models/rnn_models.py:84-143, models/transformers.py:531-579 / 778-825).
//
//  dh_select_tokens : per logits row -- exact k-th-largest threshold (ties kept, <unk> masked but counted,
//                     Q3), softmax(l/T), Exp(1)-race draw of B ids without replacement (Q2), score =
//                     log_softmax over the B picked raw logits (Q4).  One CTA per row; the row is staged
//                     once in shared memory (V*4 B <= 220 KB) so HBM sees exactly one read of the logits.
//  dh_beam_init     : first-step state (sequences, scores, ended flags, KV slot table).
//  dh_beam_step     : one warp per image -- candidate expansion (Q6), stochastic pruning (Q7), sequence /
//                     score / ended write-back, parent indices for recurrent state (LSTM: the reference's
//                     misaligned f // B, Q8; transformer: true parent), "frozen at break" (Q11/Q10).
//  dh_beam_final    : final pick (Q13) and padded [n_img, max_len] output + lengths.
//  dh_token_logprob : log_softmax(logits)[target] per row (experiments/metrics.py:5).
#include "common.cuh"

namespace {

constexpr int kSelThreads = 512;
constexpr int kMaxBeam = 16;
constexpr int kSurvCap = 4096;
constexpr int kMaxSlots = 160;   // max cached positions per beam (S_alloc)

__device__ __forceinline__ unsigned int order_key(float x) {
  unsigned int u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SelParams {
  const float* logits; long long ld;
  int R, V, B, top_k, unk, rpi;
  float T;
  int noise_mode; unsigned long long seed; long long image_base; int step;
  const unsigned char* done;   // [n_img] rows of frozen images are skipped (may be null)
  int* ind; float* val;        // [R, B]
  int* status;
  const long long* dyn;        // nullable: {seed, image_base} overriding the by-value fields
};

// Block-wide helpers (NT threads).
template <int NT>
__device__ __forceinline__ float block_max(float v, float* red, float* bcast) {
  v = dh_warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { float m = red[0]; for (int w = 1; w < NT / 32; ++w) m = fmaxf(m, red[w]); *bcast = m; }
  __syncthreads();
  return *bcast;
}
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* red, float* bcast) {
  v = dh_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < NT / 32; ++w) t += red[w]; *bcast = t; }
  __syncthreads();
  return *bcast;
}

// Shared tail of both selection kernels.  surv_idx / surv_val hold nc candidates that include every logit >= the row's
// top_k-th largest value; this computes that exact value (ties kept), masks <unk>, softmax(l/T), the Exp(1)-race draw of
// B ids and the log_softmax scores (models/beam.py:32-53,79).  All NT threads of the CTA must call it.
template <int NT>
__device__ void select_tail(const SelParams& p, int r, int img, int nc, int* surv_idx, float* surv_val, float* surv_score,
                            int cap) {
  __shared__ int s_nsurv;
  __shared__ float s_kth;
  __shared__ float red_f[NT / 32];
  __shared__ int red_i[NT / 32];
  __shared__ float s_bcast;
  __shared__ int pick_idx[kMaxBeam];
  __shared__ float pick_logit[kMaxBeam];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- exact k-th largest among the candidates: #{> v} < top_k <= #{>= v}
  if (tid == 0) { s_kth = -INFINITY; s_nsurv = 0; }  // fewer than top_k finite values: everything survives
  __syncthreads();
  for (int c = tid; c < nc; c += NT) {
    const float v = surv_val[c];
    int gt = 0, ge = 0;
    for (int j = 0; j < nc; ++j) { const float o = surv_val[j]; gt += o > v; ge += o >= v; }
    if (gt < p.top_k && p.top_k <= ge) s_kth = v;   // all writers hold the same value
  }
  __syncthreads();
  const float kth = s_kth;
  // ---- survivors: value >= k-th largest (ties kept) and id != <unk>; the others are flagged in place (id = -1)
  int mine = 0;
  for (int c = tid; c < nc; c += NT) {
    const bool keep = surv_val[c] >= kth && surv_idx[c] != p.unk;
    if (!keep) surv_idx[c] = -1;
    mine += keep;
  }
  if (mine) atomicAdd(&s_nsurv, mine);
  __syncthreads();
  const int ns = s_nsurv;
  if (ns == 0) {   // whole row filtered: torch.multinomial raises (Q3)
    if (tid == 0) atomicOr(p.status, DH_STATUS_EMPTY_ROW);
    if (tid < p.B) { p.ind[(long long)r * p.B + tid] = 0; p.val[(long long)r * p.B + tid] = 0.f; }
    return;
  }

  // ---- softmax(l / T) over survivors (everything else has p == 0 exactly)
  float m2 = -INFINITY;
  for (int s = tid; s < nc; s += NT)
    if (surv_idx[s] >= 0) m2 = fmaxf(m2, surv_val[s] / p.T);
  m2 = block_max<NT>(m2, red_f, &s_bcast);
  float sum = 0.f;
  for (int s = tid; s < nc; s += NT) {
    const float e = surv_idx[s] >= 0 ? expf(surv_val[s] / p.T - m2) : 0.f;
    surv_score[s] = e;
    sum += e;
  }
  sum = block_sum<NT>(sum, red_f, &s_bcast);
  const unsigned long long seed = p.dyn ? (unsigned long long)p.dyn[0] : p.seed;
  const long long image_base = p.dyn ? p.dyn[1] : p.image_base;
  const unsigned long long rk = dh_noise_row_key(seed, (unsigned long long)(image_base + img), (unsigned long long)p.step,
                                                 DH_CALL_TOKEN, (unsigned long long)(r % p.rpi));
  for (int s = tid; s < nc; s += NT) {
    float pr = -2.f;                                     // not a survivor: never picked
    if (surv_idx[s] >= 0) {
      pr = surv_score[s] / sum;
      if (p.noise_mode == DH_NOISE_INJECTED) pr = pr / dh_exp_noise(rk, (unsigned long long)surv_idx[s]);
    }
    surv_score[s] = pr;
  }
  __syncthreads();

  // ---- B rounds of block arg-max on (score desc, id asc)
  const int npick = ns < p.B ? ns : p.B;
  for (int j = 0; j < npick; ++j) {
    float best = -1.f; int bi = 0x7fffffff, bs = -1;
    for (int s = tid; s < nc; s += NT) {
      float sc = surv_score[s]; int id = surv_idx[s];
      if (sc > best || (sc == best && sc >= 0.f && id < bi)) { best = sc; bi = id; bs = s; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int os = __shfl_xor_sync(0xffffffffu, bs, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; bs = os; }
    }
    if (lane == 0) { red_f[warp] = best; red_i[warp] = bs; }
    __syncthreads();
    if (tid == 0) {
      float b2 = -1.f; int s2 = -1, i2 = 0x7fffffff;
      for (int w = 0; w < NT / 32; ++w) {
        int s = red_i[w];
        if (s < 0) continue;
        int id = surv_idx[s];
        if (red_f[w] > b2 || (red_f[w] == b2 && id < i2)) { b2 = red_f[w]; s2 = s; i2 = id; }
      }
      pick_idx[j] = i2;
      pick_logit[j] = surv_val[s2];
      surv_score[s2] = -2.f;   // taken
    }
    __syncthreads();
  }
  if (tid == 0) {
    // fewer survivors than B: multinomial continues with zero-probability ids; stable order = lowest ids
    int next = 0;
    for (int j = npick; j < p.B; ++j) {
      for (;; ++next) {
        bool used = false;
        for (int q = 0; q < j; ++q) used |= (pick_idx[q] == next);
        if (!used) break;
      }
      pick_idx[j] = next;
      pick_logit[j] = -INFINITY;
      ++next;
    }
    // score = log_softmax over the B picked (filtered, un-tempered) logits (Q4)
    float m3 = -INFINITY;
    for (int j = 0; j < p.B; ++j) m3 = fmaxf(m3, pick_logit[j]);
    float se = 0.f;
    for (int j = 0; j < p.B; ++j) se += expf(pick_logit[j] - m3);
    float lse = logf(se);
    for (int j = 0; j < p.B; ++j) {
      p.ind[(long long)r * p.B + j] = pick_idx[j];
      p.val[(long long)r * p.B + j] = pick_logit[j] - m3 - lse;
    }
  }
}

// One CTA per logits row, two streaming passes over the row (the second one hits L2), no row staging:
//   pass 1  every thread keeps the maximum of its strided slice; the top_k-th largest of those kSelThreads
//           distinct elements is a lower bound t0 of the row's k-th largest value;
//   pass 2  elements >= t0 (a few more than top_k) are compacted into shared memory;
//   then    select_tail on that short list.
__global__ void __launch_bounds__(kSelThreads) select_tokens_kernel(SelParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* surv_idx = reinterpret_cast<int*>(smem_raw);                  // [kSurvCap]
  float* surv_val = reinterpret_cast<float*>(surv_idx + kSurvCap);   // [kSurvCap] raw logits
  float* surv_score = surv_val + kSurvCap;                           // [kSurvCap]
  __shared__ float lm[kSelThreads];
  __shared__ int s_ncand;
  __shared__ float s_t0;

  const int r = blockIdx.x, tid = threadIdx.x;
  const int img = r / p.rpi;
  if (p.done && p.done[img]) return;
  const float* row = p.logits + (long long)r * p.ld;
  const bool vec = (p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.logits) & 15) == 0);
  const int V4 = vec ? (p.V >> 2) : 0;

  // ---- pass 1: per-thread maximum
  float mx = -INFINITY;
  for (int i = tid; i < V4; i += kSelThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int i = V4 * 4 + tid; i < p.V; i += kSelThreads) mx = fmaxf(mx, row[i]);
  lm[tid] = mx;
  if (tid == 0) { s_ncand = 0; s_t0 = -INFINITY; }
  __syncthreads();
  if (p.top_k <= kSelThreads) {
    int rank = 0;
    for (int j = 0; j < kSelThreads; ++j) {
      const float o = lm[j];
      rank += (o > mx) || (o == mx && j < tid);
    }
    if (rank == p.top_k - 1) s_t0 = mx;     // exactly one thread has this rank
  }
  __syncthreads();
  const float t0 = s_t0;                     // -inf when top_k > kSelThreads (every finite element is a candidate)

  // ---- pass 2: compact candidates
  auto push = [&](float x, int i) {
    if (x >= t0 && x > -INFINITY) {
      const int slot = atomicAdd(&s_ncand, 1);
      if (slot < kSurvCap) { surv_idx[slot] = i; surv_val[slot] = x; }
    }
  };
  for (int i = tid; i < V4; i += kSelThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
    push(v.x, 4 * i); push(v.y, 4 * i + 1); push(v.z, 4 * i + 2); push(v.w, 4 * i + 3);
  }
  for (int i = V4 * 4 + tid; i < p.V; i += kSelThreads) push(row[i], i);
  __syncthreads();
  int nc = s_ncand;
  if (nc > kSurvCap) {
    if (tid == 0) atomicOr(p.status, DH_STATUS_TOO_MANY_TIES);
    nc = kSurvCap;
  }
  select_tail<kSelThreads>(p, r, img, nc, surv_idx, surv_val, surv_score, kSurvCap);
}

// ------------------------------------------------------------------------------------------------ fused vocab path
// t0[row] = top_k-th largest of the row's 32-column group maxima (a lower bound of the row's top_k-th largest logit:
// the group maxima are n_groups distinct logits).  At most top_k groups have a maximum > t0, so -- ties aside -- at most
// 32 * top_k logits are >= t0 and the candidate capacity 32 * top_k of dh_vocab_candidates cannot overflow.
// One WARP per row, no block-level synchronisation: the two largest values of every lane's strided slice are 64
// distinct group maxima whose top_k-th largest is a first bound t1 (top_k <= 64); group maxima >= t1 (a few more than
// top_k) are then compacted per warp and ranked exactly.
constexpr int kThrWarps = 2;     // small CTAs: 2560 rows spread evenly over 148 SMs (8-warp CTAs left some SMs 3 CTAs, others 2)
constexpr int kThrCap = 256;

// m-th smallest (1-based, m <= 64) of the 64 order-preserving keys a warp holds two per lane, by m rounds of warp-wide
// minimum extraction (redux.sync): O(8 m) instructions against O(64 * 64 / 32 * 5) for rank counting.  Keys of unused
// slots must be 0xffffffff.
__device__ __forceinline__ unsigned int warp_mth_smallest64(unsigned int a, unsigned int b, int m, int lane) {
  unsigned int res = 0xffffffffu;
  for (int i = 0; i < m; ++i) {
    const unsigned int lo = a < b ? a : b;
    res = __reduce_min_sync(0xffffffffu, lo);
    const unsigned int has = __ballot_sync(0xffffffffu, lo == res);
    if (lane == __ffs(has) - 1) {            // drop ONE instance (the lowest lane holding it)
      if (a == res) a = 0xffffffffu; else b = 0xffffffffu;
    }
  }
  return res;
}
__device__ __forceinline__ float key_to_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// NPL > 0: the whole row (<= 32 * NPL group maxima) is loaded into registers with all loads in flight at once (one L2
// round trip); NPL == 0: generic two-pass version for longer rows.
struct ThrFix {              // fix-up form (dh_vocab_threshold_fix): only rows whose candidate count left [cmin, cmax]
  int on, cmin, cmax;
  unsigned char* redo;       // [rows] 1 = this row's threshold was replaced (its list must be rebuilt), else 0
  int* any_flag;             // set to 1 if any row was replaced
};

template <int NPL>
__global__ void __launch_bounds__(kThrWarps * 32) vocab_threshold_kernel(const float* __restrict__ gmax, long long ld, int rows,
                                                                        int n_groups, int top_k, float* __restrict__ thresh,
                                                                        int* __restrict__ cand_count, ThrFix fx) {
  __shared__ float s_cand[kThrWarps][kThrCap];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kThrWarps + w;
  if (r >= rows) return;
  if (fx.on) {
    const int c = cand_count[r];
    const bool bad = c < fx.cmin || c > fx.cmax;
    if (lane == 0) fx.redo[r] = (unsigned char)bad;
    if (!bad) return;
    if (lane == 0) atomicOr(fx.any_flag, 1);
  }
  const float* row = gmax + (long long)r * ld;
  constexpr int NR = NPL > 0 ? NPL : 1;
  float xr[NR];
  if (NPL > 0) {
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      const int i = u * 32 + lane;
      xr[u] = i < n_groups ? __ldg(row + i) : -INFINITY;
    }
  }
  float t0 = -INFINITY;
  if (top_k <= n_groups) {
    float t1 = -INFINITY;
    if (top_k <= 64) {
      float m1 = -INFINITY, m2 = -INFINITY;
      if (NPL > 0) {
#pragma unroll
        for (int u = 0; u < NR; ++u) {
          const float hi = fmaxf(m1, xr[u]);
          m2 = fmaxf(m2, fminf(m1, xr[u]));
          m1 = hi;
        }
      } else {
        for (int i0 = 0; i0 < n_groups; i0 += 256) {       // 8 independent loads in flight per lane
          float x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 32 + lane;
            x[u] = i < n_groups ? __ldg(row + i) : -INFINITY;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float hi = fmaxf(m1, x[u]);
            m2 = fmaxf(m2, fminf(m1, x[u]));
            m1 = hi;
          }
        }
      }
      // top_k-th largest of the 64 per-lane top-2 values = (64 - top_k + 1)-th smallest
      t1 = key_to_float(warp_mth_smallest64(order_key(m1), order_key(m2), 64 - top_k + 1, lane));
    }
    int n = 0;
    auto offer = [&](float x, bool in_range) {
      const bool keep = in_range && x >= t1;
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = n + __popc(mask & ((1u << lane) - 1u));
        if (pos < kThrCap) s_cand[w][pos] = x;
      }
      n += __popc(mask);
    };
    if (NPL > 0) {
#pragma unroll
      for (int u = 0; u < NR; ++u) offer(xr[u], u * 32 + lane < n_groups);
    } else {
      for (int i0 = 0; i0 < n_groups; i0 += 32) {
        const int i = i0 + lane;
        offer(i < n_groups ? __ldg(row + i) : -INFINITY, i < n_groups);
      }
    }
    __syncwarp();
    if (n <= 64) {
      // exact top_k-th largest of the n compacted maxima = (n - top_k + 1)-th smallest (n >= top_k: the compaction kept at
      // least the top_k values that defined t1, or everything when t1 = -inf)
      const unsigned int a = lane < n ? order_key(s_cand[w][lane]) : 0xffffffffu;
      const unsigned int b = 32 + lane < n ? order_key(s_cand[w][32 + lane]) : 0xffffffffu;
      t0 = key_to_float(warp_mth_smallest64(a, b, n - top_k + 1, lane));
    } else if (n <= kThrCap) {
      // rank counting (independent iterations) for the rare longer lists
      for (int c = lane; c < n; c += 32) {
        const float v = s_cand[w][c];
        int gt = 0, ge = 0;
        for (int j = 0; j < n; ++j) { const float o = s_cand[w][j]; gt += o > v; ge += o >= v; }
        if (gt < top_k && top_k <= ge) t0 = v;
      }
      t0 = dh_warp_max(t0);
    } else {
      t0 = t1;        // too many to rank here (large top_k or massive ties): the looser bound is still valid
    }
  }
  if (lane == 0) { thresh[r] = t0; cand_count[r] = 0; }
}

// Warp-level selection from a row's SPARSELY MATERIALISED logits (written by dh_vocab_candidates): the hit map names the
// 32-column groups that hold at least one logit >= thresh[row]; their 128-byte lines are gathered (one lane per group, eight
// 16-byte loads in flight per lane), filtered against the threshold, and the exact top_k-th largest value (ties kept, Q3) is
// taken from that candidate list.  Candidates and survivors stay in column order, so every floating-point reduction is
// run-to-run deterministic; then softmax(l/T), the Exp(1)-race draw of B ids and the log_softmax scores over the picks
// (models/beam.py:32-53,79).
constexpr int kCandCap = 512;    // candidates (logits >= thresh) a warp can rank
constexpr int kSurvMax = 160;    // survivors (logits >= the exact top_k-th largest): top_k <= 64 plus ties
struct VocabSparse {
  const float* thresh;           // [rows]
  const unsigned char* hitmap;   // [rows, hit_ld]: one byte per (N tile, column half), one bit per stored 32-column group
  long long hit_ld;
  const float* logits;           // [rows, ld]: only the groups named by the hit map hold data
  long long ld;
  int n_bytes, groups_per_byte;
};
struct WarpSel {
  int* c_idx; float* c_val;     // [kCandCap] candidates, later the survivors sorted by column
  int* t_idx; float* t_val;     // [kSurvMax] unsorted survivors   } the group list of the gather phase aliases these
  float* score;                 // [kSurvMax]                       } three arrays (3 * kSurvMax ints)
  int* pick_idx; float* pick_logit;   // [kMaxBeam]
  int* counter;                 // candidate count of the gather phase
};
constexpr int kWarpSelBytes = 2 * kCandCap * 4 + 3 * kSurvMax * 4 + 2 * kMaxBeam * 4 + 16;

template <int KPL>   // keys per lane
__device__ __forceinline__ float warp_kth_largest(const float* vv, int nc, int k, int lane) {
  // bit-wise radix select on the order-preserving keys held in registers: 32 rounds of (KPL compares + one warp-wide add)
  unsigned int key[KPL];
#pragma unroll
  for (int u = 0; u < KPL; ++u) key[u] = u * 32 + lane < nc ? order_key(vv[u * 32 + lane]) : 0u;
  unsigned int ans = 0u;
  for (int b = 31; b >= 0; --b) {
    const unsigned int trial = ans | (1u << b);
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < KPL; ++u) cnt += key[u] >= trial;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt >= k) ans = trial;
  }
  return key_to_float(ans);
}

__device__ void warp_select(const SelParams& p, int r, int img, const VocabSparse& vs, WarpSel w) {
  const int lane = threadIdx.x & 31;
  const float thr = __ldg(vs.thresh + r);
  const float t0 = fmaxf(thr, -3.402823466e+38f);          // candidates are finite
  // ---- phase A: list of stored groups, in column order
  int* glist = w.t_idx;
  constexpr int kGroupCap = 3 * kSurvMax;
  const unsigned char* hm = vs.hitmap + (long long)r * vs.hit_ld;
  int ng = 0;
  for (int b0 = 0; b0 < vs.n_bytes; b0 += 128) {             // four map bytes per lane and round, all loads in flight
    const int bi = b0 + 4 * lane;
    unsigned int word = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (bi + k < vs.n_bytes) word |= (unsigned int)__ldg(hm + bi + k) << (8 * k);
    const int mine = __popc(word);
    int off = mine;                                          // inclusive prefix over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
    const int tot = __shfl_sync(0xffffffffu, off, 31);
    int pos = ng + off - mine;
    while (word) {
      const int bit = __ffs(word) - 1;
      word &= word - 1u;
      if (pos < kGroupCap) glist[pos] = (bi + (bit >> 3)) * vs.groups_per_byte + (bit & 7);
      ++pos;
    }
    ng += tot;
  }
  if (ng > kGroupCap) {
    if (lane == 0) atomicOr(p.status, DH_STATUS_TOO_MANY_TIES);
    ng = kGroupCap;
  }
  __syncwarp();
  // ---- phase B: one lane per group (eight independent 16-byte loads in flight): the 32 logits of a group are compared
  // with the threshold into a bit mask without branching, the masks' population counts are prefix-summed over the warp and
  // every lane appends the COLUMNS of its (few) hits at its own offset -- no atomics, and the list comes out in column
  // order.  The hit values are fetched afterwards, one independent load per candidate (the lines are in L1 / L2).
  const float* lrow = vs.logits + (long long)r * vs.ld;
  int nc = 0;
  for (int i0 = 0; i0 < ng; i0 += 32) {
    const int i = i0 + lane;
    unsigned int mask = 0u;
    int col0 = 0;
    if (i < ng) {
      col0 = glist[i] * 32;
      float4 v[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) v[g] = __ldg(reinterpret_cast<const float4*>(lrow + col0) + g);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        mask |= (v[g].x >= t0 ? 1u : 0u) << (4 * g) | (v[g].y >= t0 ? 1u : 0u) << (4 * g + 1) |
                (v[g].z >= t0 ? 1u : 0u) << (4 * g + 2) | (v[g].w >= t0 ? 1u : 0u) << (4 * g + 3);
      }
    }
    const int mine = __popc(mask);
    int off = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
    const int tot = __shfl_sync(0xffffffffu, off, 31);
    int pos = nc + off - mine;
    while (mask) {
      const int bit = __ffs(mask) - 1;
      mask &= mask - 1u;
      if (pos < kCandCap) w.c_idx[pos] = col0 + bit;
      ++pos;
    }
    nc += tot;
  }
  __syncwarp();
  for (int c = lane; c < min(nc, kCandCap); c += 32) w.c_val[c] = __ldg(lrow + w.c_idx[c]);
  __syncwarp();
  if (nc > kCandCap) {
    if (lane == 0) atomicOr(p.status, DH_STATUS_TOO_MANY_TIES);
    nc = kCandCap;
  }
  const int* ii = w.c_idx;
  const float* vv = w.c_val;
  // ---- exact k-th largest among the candidates: #{> v} < top_k <= #{>= v}; -inf if fewer than top_k candidates
  float kth = -INFINITY;
  if (nc >= p.top_k) {
    if (nc <= 64 && nc - p.top_k < 16) {
      // an exhaustive pass 1 hands over ~top_k + 1 candidates: the top_k-th largest of nc values is the (nc - top_k + 1)-th
      // smallest -- one or two rounds of warp-wide minimum extraction
      const unsigned int a = lane < nc ? order_key(vv[lane]) : 0xffffffffu;
      const unsigned int b = 32 + lane < nc ? order_key(vv[32 + lane]) : 0xffffffffu;
      kth = key_to_float(warp_mth_smallest64(a, b, nc - p.top_k + 1, lane));
    } else if (nc <= 256) {
      kth = warp_kth_largest<8>(vv, nc, p.top_k, lane);      // a sampled pass 1 hands over a few times top_k candidates
    } else {
      kth = warp_kth_largest<kCandCap / 32>(vv, nc, p.top_k, lane);
    }
  }
  // ---- survivors: value >= k-th largest (ties kept) and id != <unk>
  int ns = 0;
  for (int c0 = 0; c0 < nc; c0 += 32) {
    const int c = c0 + lane;
    const bool keep = c < nc && vv[c] >= kth && ii[c] != p.unk;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = ns + __popc(mask & ((1u << lane) - 1u));
      if (pos < kSurvMax) { w.t_idx[pos] = ii[c]; w.t_val[pos] = vv[c]; }
    }
    ns += __popc(mask);
  }
  if (ns > kSurvMax) {
    if (lane == 0) atomicOr(p.status, DH_STATUS_TOO_MANY_TIES);
    ns = kSurvMax;
  }
  __syncwarp();
  if (ns == 0) {   // whole row filtered: torch.multinomial raises (Q3)
    if (lane == 0) atomicOr(p.status, DH_STATUS_EMPTY_ROW);
    if (lane < p.B) { p.ind[(long long)r * p.B + lane] = 0; p.val[(long long)r * p.B + lane] = 0.f; }
    return;
  }
  int* a_idx = w.c_idx;                            // the candidate arrays are free again: sorted survivors
  float* a_val = w.c_val;
  // the gather phase lists the candidates in column order and the compaction keeps it: the survivors are already sorted by
  // column, which is what makes every floating-point reduction below run-to-run deterministic
  for (int c = lane; c < ns; c += 32) {
    a_idx[c] = w.t_idx[c];
    a_val[c] = w.t_val[c];
  }
  __syncwarp();
  // ---- softmax(l / T) over survivors (everything else has p == 0 exactly), noise, B rounds of arg-max
  float m2 = -INFINITY;
  for (int s = lane; s < ns; s += 32) m2 = fmaxf(m2, a_val[s] / p.T);
  m2 = dh_warp_max(m2);
  float sum = 0.f;
  for (int s = lane; s < ns; s += 32) {
    const float e = expf(a_val[s] / p.T - m2);
    w.score[s] = e;
    sum += e;
  }
  sum = dh_warp_sum(sum);
  const unsigned long long seed = p.dyn ? (unsigned long long)p.dyn[0] : p.seed;
  const long long image_base = p.dyn ? p.dyn[1] : p.image_base;
  const unsigned long long rk = dh_noise_row_key(seed, (unsigned long long)(image_base + img), (unsigned long long)p.step,
                                                 DH_CALL_TOKEN, (unsigned long long)(r % p.rpi));
  for (int s = lane; s < ns; s += 32) {
    float pr = w.score[s] / sum;
    if (p.noise_mode == DH_NOISE_INJECTED) pr = pr / dh_exp_noise(rk, (unsigned long long)a_idx[s]);
    w.score[s] = pr;
  }
  __syncwarp();
  const int npick = ns < p.B ? ns : p.B;
  for (int j = 0; j < npick; ++j) {
    float best = -1.f; int bi = 0x7fffffff, bs = -1;
    for (int s = lane; s < ns; s += 32) {
      const float sc = w.score[s]; const int id = a_idx[s];
      if (sc > best || (sc == best && sc >= 0.f && id < bi)) { best = sc; bi = id; bs = s; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int os = __shfl_xor_sync(0xffffffffu, bs, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; bs = os; }
    }
    if (lane == 0) {
      w.pick_idx[j] = bi;
      w.pick_logit[j] = a_val[bs];
      w.score[bs] = -2.f;   // taken
    }
    __syncwarp();
  }
  if (lane == 0) {
    // fewer survivors than B: multinomial continues with zero-probability ids; stable order = lowest ids
    int next = 0;
    for (int j = npick; j < p.B; ++j) {
      for (;; ++next) {
        bool used = false;
        for (int q = 0; q < j; ++q) used |= (w.pick_idx[q] == next);
        if (!used) break;
      }
      w.pick_idx[j] = next;
      w.pick_logit[j] = -INFINITY;
      ++next;
    }
    // score = log_softmax over the B picked (filtered, un-tempered) logits (Q4)
    float m3 = -INFINITY;
    for (int j = 0; j < p.B; ++j) m3 = fmaxf(m3, w.pick_logit[j]);
    float se = 0.f;
    for (int j = 0; j < p.B; ++j) se += expf(w.pick_logit[j] - m3);
    const float lse = logf(se);
    for (int j = 0; j < p.B; ++j) {
      p.ind[(long long)r * p.B + j] = w.pick_idx[j];
      p.val[(long long)r * p.B + j] = w.pick_logit[j] - m3 - lse;
    }
  }
}

// ------------------------------------------------------------------------------------------------ beam state
struct BeamState {
  int* seq;            // [n_img, B, seq_ld]
  long long seq_ld;
  float* val;          // [n_img, B]
  unsigned char* ended;  // [n_img, B]
  unsigned char* done;   // [n_img]
  int* final_len;      // [n_img]
  int* last_tok;       // [n_img * B]
  int* parent_state;   // [n_img * B] row to gather recurrent state from
  int* src;            // [n_img, B, S_alloc] KV slot table (transformer) or null
  int S_alloc;
};

__global__ void beam_init_kernel(BeamState st, const int* __restrict__ ind0, const float* __restrict__ val0,
                                 const int* __restrict__ prefix, long long prefix_ld, int prefix_rows, int p0, int n_img,
                                 int B, int eos, int lstm_semantics) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * B) return;
  int img = i / B, b = i % B;
  int* s = st.seq + (long long)i * st.seq_ld;
  for (int t = 0; t < p0; ++t) s[t] = prefix[(long long)(prefix_rows == 1 ? 0 : img) * prefix_ld + t];
  int tok = ind0[i];
  s[p0] = tok;
  for (long long t = p0 + 1; t < st.seq_ld; ++t) s[t] = 0;
  st.val[i] = val0[i];
  st.ended[i] = (unsigned char)(lstm_semantics ? (tok == eos) : 0);   // Q9
  st.last_tok[i] = tok;
  st.parent_state[i] = img;                                             // expand the 1-row-per-image state
  if (st.src) {
    int* sr = st.src + (long long)i * st.S_alloc;
    for (int t = 0; t < st.S_alloc; ++t) sr[t] = (t <= p0) ? 0 : b;     // positions 0..p0 live in slot 0
  }
  if (b == 0) { st.done[img] = 0; st.final_len[img] = 0; }
}

struct StepParams {
  const int* new_ind; const float* new_val;   // [n_img*B, B]
  int n_img, B, step, max_len, eos, lstm_semantics;
  float T; int noise_mode; unsigned long long seed; long long image_base; const long long* dyn;
};

// One warp per image; s_seq is B * seq_ld (+ B * S_alloc for the KV slot table) ints of DYNAMIC shared memory: a static
// [16][160] slot-table array cost 10 KB per CTA and one resident CTA per SM in the selection launch.
__device__ void beam_step_body(const BeamState& st, const StepParams& p, int img, int* s_seq) {
  __shared__ float s_score[kMaxBeam * kMaxBeam];
  __shared__ float s_cval[kMaxBeam * kMaxBeam];
  __shared__ int s_ctok[kMaxBeam * kMaxBeam];
  __shared__ unsigned char s_cend[kMaxBeam * kMaxBeam], s_cpar[kMaxBeam * kMaxBeam];
  __shared__ int s_f[kMaxBeam];
  const int lane = threadIdx.x & 31, B = p.B;
  if (st.done[img]) return;
  const long long base = (long long)img * B;
  int* s_src = s_seq + (long long)B * st.seq_ld;              // [B][S_alloc]
  // --- candidate layout (beam.py:83-102): ended row -> 1 copy, live row -> B copies
  int e = 0, copies = 0;
  float v = 0.f;
  if (lane < B) { e = st.ended[base + lane]; copies = e ? 1 : B; v = st.val[base + lane]; }
  int off = copies;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
  const int N = __shfl_sync(0xffffffffu, off, 31);
  off -= copies;
  if (lane < B) {
    for (int j = 0; j < copies; ++j) {
      int c = off + j;
      int tok = e ? 0 : p.new_ind[(base + lane) * B + j];
      float dv = e ? 0.f : p.new_val[(base + lane) * B + j];
      s_cval[c] = v + dv;
      s_ctok[c] = tok;
      s_cend[c] = (unsigned char)(e | (tok == p.eos));
      s_cpar[c] = (unsigned char)lane;
    }
  }
  // stage old sequences (and slot table) before they are overwritten
  for (long long i = lane; i < (long long)B * st.seq_ld; i += 32) s_seq[i] = st.seq[base * st.seq_ld + i];
  if (st.src)
    for (int i = lane; i < B * st.S_alloc; i += 32) s_src[i] = st.src[base * st.S_alloc + i];
  __syncwarp();
  // --- pruning draw: sample B of N from softmax(cand_val / T) (Q7)
  float mx = -INFINITY;
  for (int c = lane; c < N; c += 32) mx = fmaxf(mx, s_cval[c] / p.T);
  mx = dh_warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < N; c += 32) { float ex = expf(s_cval[c] / p.T - mx); s_score[c] = ex; sum += ex; }
  sum = dh_warp_sum(sum);
  const unsigned long long seed = p.dyn ? (unsigned long long)p.dyn[0] : p.seed;
  const long long image_base = p.dyn ? p.dyn[1] : p.image_base;
  const unsigned long long rk = dh_noise_row_key(seed, (unsigned long long)(image_base + img), (unsigned long long)p.step,
                                                 DH_CALL_PRUNE, 0ull);
  for (int c = lane; c < N; c += 32) {
    float pr = s_score[c] / sum;
    if (p.noise_mode == DH_NOISE_INJECTED) pr = pr / dh_exp_noise(rk, (unsigned long long)c);
    s_score[c] = pr;
  }
  __syncwarp();
  for (int j = 0; j < B; ++j) {
    float best = -1.f; int bc = 0x7fffffff;
    for (int c = lane; c < N; c += 32) {
      float sc = s_score[c];
      if (sc > best || (sc == best && sc >= 0.f && c < bc)) { best = sc; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
    }
    if (lane == 0) { s_f[j] = bc; s_score[bc] = -2.f; }
    __syncwarp();
  }
  // --- write back
  int all_ended = 1;
  if (lane < B) {
    const int f = s_f[lane];
    const int par = s_cpar[f];
    const int tok = s_ctok[f];
    st.val[base + lane] = s_cval[f];
    st.ended[base + lane] = s_cend[f];
    all_ended = s_cend[f];
    st.last_tok[base + lane] = tok;
    st.parent_state[base + lane] = (int)base + (p.lstm_semantics ? (f / B) : par);   // Q8
  }
  all_ended = __all_sync(0xffffffffu, all_ended);
  for (int j = 0; j < B; ++j) {
    const int f = s_f[j], par = s_cpar[f];
    int* dst = st.seq + (base + j) * st.seq_ld;
    for (int t = lane; t < st.seq_ld; t += 32) {
      int x = s_seq[(long long)par * st.seq_ld + t];
      if (t == p.step && p.step < p.max_len) x = s_ctok[f];      // column `step`; no-op at step == max_len (Q10)
      dst[t] = x;
    }
    if (st.src) {
      int* sr = st.src + (base + j) * st.S_alloc;
      for (int t = lane; t < st.S_alloc; t += 32) sr[t] = (t <= p.step) ? s_src[par * st.S_alloc + t] : j;
    }
  }
  if (lane == 0 && all_ended) {
    st.done[img] = 1;                                              // reference `break` (rnn_models.py:131)
    st.final_len[img] = p.lstm_semantics ? p.step + 1 : p.step;    // Q11 / Q10
  }
}

__global__ void __launch_bounds__(32) beam_step_kernel(BeamState st, StepParams p) {
  extern __shared__ int dyn_smem[];                 // [B, seq_ld]
  beam_step_body(st, p, blockIdx.x, dyn_smem);
}

// One CTA per image, one warp per logits row of the image: warp_select on every row's candidate list, then (do_beam)
// warp 0 runs the beam step of the image on the picks -- selection and beam bookkeeping of one decode step in ONE launch.
// Operand gathers of the next LSTM time step for the image's rows (dh_lstm_operands; L == 0: none)
struct LstmNext {
  const uint16_t* table; long long ldt, n_tok_rows; int E, H, L;
  const uint16_t* hs[8]; uint16_t* A[8]; long long lda[8]; int in_off[8];
};

__global__ void __launch_bounds__(32 * kMaxBeam, 2) select_beam_kernel(SelParams p, VocabSparse vs, int do_beam,
                                                                   BeamState st, StepParams sp, LstmNext nx) {
  extern __shared__ int dyn_smem[];
  const int img = blockIdx.x, warp = threadIdx.x >> 5;
  if (p.done && p.done[img]) return;
  unsigned char* base = reinterpret_cast<unsigned char*>(dyn_smem) + (size_t)warp * kWarpSelBytes;
  WarpSel w;
  w.c_idx = reinterpret_cast<int*>(base);
  w.c_val = reinterpret_cast<float*>(base + kCandCap * 4);
  w.t_idx = reinterpret_cast<int*>(base + 2 * kCandCap * 4);
  w.t_val = reinterpret_cast<float*>(base + 2 * kCandCap * 4 + kSurvMax * 4);
  w.score = reinterpret_cast<float*>(base + 2 * kCandCap * 4 + 2 * kSurvMax * 4);
  w.pick_idx = reinterpret_cast<int*>(base + 2 * kCandCap * 4 + 3 * kSurvMax * 4);
  w.pick_logit = reinterpret_cast<float*>(base + 2 * kCandCap * 4 + 3 * kSurvMax * 4 + kMaxBeam * 4);
  w.counter = reinterpret_cast<int*>(base + 2 * kCandCap * 4 + 3 * kSurvMax * 4 + 2 * kMaxBeam * 4);
  const int r = img * p.rpi + warp;
  warp_select(p, r, img, vs, w);
  if (!do_beam) return;
  __syncthreads();                                   // the picks of all rows (global memory) are visible to warp 0
  if (warp == 0)
    beam_step_body(st, sp, img, reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(dyn_smem) + (size_t)p.rpi * kWarpSelBytes));
  if (nx.L == 0) return;
  // ---- next LSTM step's operands for this image's rows: 16-byte copies by the whole CTA
  __syncthreads();                                   // last_tok / parent_state of the image (written by warp 0) are visible
  // (row token / parent are fetched once into shared memory; then four independent 16-byte copies are in flight per thread)
  __shared__ int s_tok[kMaxBeam], s_par[kMaxBeam];
  if (threadIdx.x < p.rpi) {
    s_tok[threadIdx.x] = st.last_tok[(long long)img * p.rpi + threadIdx.x];
    s_par[threadIdx.x] = st.parent_state[(long long)img * p.rpi + threadIdx.x];
  }
  __syncthreads();
  const int ce = nx.E / 8, ch = nx.H / 8, per_row = ce + nx.L * ch;
  const int total = p.rpi * per_row;
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
    uint4 v[4];
    uint4* dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      dst[u] = nullptr;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i < total) {
        const int b = i / per_row;
        int c = i - b * per_row;
        const long long row = (long long)img * p.rpi + b;
        if (c < ce) {
          const long long t = s_tok[b];
          dst[u] = reinterpret_cast<uint4*>(nx.A[0] + row * nx.lda[0] + c * 8);
          if (t >= 0 && t < nx.n_tok_rows) v[u] = __ldg(reinterpret_cast<const uint4*>(nx.table + t * nx.ldt + c * 8));
        } else {
          c -= ce;
          const int l = c / ch, cc = c - l * ch;
          dst[u] = reinterpret_cast<uint4*>(nx.A[l] + row * nx.lda[l] + nx.in_off[l] + cc * 8);
          v[u] = *reinterpret_cast<const uint4*>(nx.hs[l] + (long long)s_par[b] * nx.H + cc * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (dst[u]) *dst[u] = v[u];
  }
}

__global__ void __launch_bounds__(32) beam_final_kernel(BeamState st, int n_img, int B, float T, int noise_mode,
                                                        unsigned long long seed, long long image_base, int final_step,
                                                        int len_if_running, int pad, int max_len,
                                                        long long* __restrict__ out_ids, long long* __restrict__ out_len,
                                                        const long long* __restrict__ dyn) {
  if (dyn) { seed = (unsigned long long)dyn[0]; image_base = dyn[1]; }
  const int img = blockIdx.x, lane = threadIdx.x;
  const long long base = (long long)img * B;
  float x = lane < B ? st.val[base + lane] / T : -INFINITY;
  float mx = dh_warp_max(x);
  float ex = lane < B ? expf(x - mx) : 0.f;
  float sum = dh_warp_sum(ex);
  float sc = ex / sum;
  if (lane < B && noise_mode == DH_NOISE_INJECTED) {
    unsigned long long rk = dh_noise_row_key(seed, (unsigned long long)(image_base + img), (unsigned long long)final_step,
                                             DH_CALL_FINAL, 0ull);
    sc = sc / dh_exp_noise(rk, (unsigned long long)lane);
  }
  if (lane >= B) sc = -1.f;
  int bi = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, sc, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > sc || (ob == sc && oi < bi)) { sc = ob; bi = oi; }
  }
  const int len = st.done[img] ? st.final_len[img] : len_if_running;
  const int* s = st.seq + (base + bi) * st.seq_ld;
  for (int t = lane; t < max_len; t += 32) out_ids[(long long)img * max_len + t] = (t < len) ? (long long)s[t] : (long long)pad;
  if (lane == 0) out_len[img] = len;
}

// ------------------------------------------------------------------------------------------------ log-prob
__global__ void __launch_bounds__(256) token_logprob_kernel(const float* __restrict__ logits, long long ld, int V,
                                                            const long long* __restrict__ targets, float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float s_b;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (long long)r * ld;
  float mx = -INFINITY;
  for (int i = tid; i < V; i += 256) mx = fmaxf(mx, row[i]);
  mx = dh_warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) { float m = red[0]; for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]); s_b = m; }
  __syncthreads();
  mx = s_b;
  float sum = 0.f;
  for (int i = tid; i < V; i += 256) sum += expf(row[i] - mx);
  sum = dh_warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[r] = row[targets[r]] - mx - logf(t);
  }
}

}  // namespace

extern "C" int dh_select_tokens(const float* logits, long long ld, int rows, int V, int beam, int top_k, float temperature,
                                int unk, int rows_per_image, int noise_mode, unsigned long long seed, long long image_base,
                                int step, const unsigned char* done, int* ind, float* val, int* status, const long long* dyn,
                                cudaStream_t s) {
  DH_ARG(logits && ind && val && status && rows >= 0 && V > 0);
  DH_ARG(beam >= 1 && beam <= kMaxBeam && top_k >= 1 && top_k <= V && beam <= top_k && temperature > 0.f);
  DH_ARG(rows_per_image >= 1 && (noise_mode == DH_NOISE_DETERMINISTIC || noise_mode == DH_NOISE_INJECTED));
  DH_ARG(top_k <= kSelThreads || V <= kSurvCap);   // candidate list is bounded by per-thread maxima
  if (rows == 0) return DH_OK;
  size_t smem = (size_t)kSurvCap * 12;
  static bool attr_set = false;
  if (!attr_set) {
    DH_CUDA(cudaFuncSetAttribute(select_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  SelParams p{logits, ld, rows, V, beam, top_k, unk, rows_per_image, temperature, noise_mode, seed, image_base, step,
              done, ind, val, status, dyn};
  select_tokens_kernel<<<rows, kSelThreads, smem, s>>>(p);
  DH_LAUNCH_OK();
  return DH_OK;
}

static int launch_threshold(const float* gmax, long long ld_gmax, int rows, int n_groups, int top_k, float* thresh,
                            int* cand_count, const ThrFix& fx, cudaStream_t s) {
  const int grid = dh_cdiv(rows, kThrWarps);
  if (n_groups <= 32 * 5)
    vocab_threshold_kernel<5><<<grid, kThrWarps * 32, 0, s>>>(gmax, ld_gmax, rows, n_groups, top_k, thresh, cand_count, fx);
  else if (n_groups <= 32 * 36)
    vocab_threshold_kernel<36><<<grid, kThrWarps * 32, 0, s>>>(gmax, ld_gmax, rows, n_groups, top_k, thresh, cand_count, fx);
  else
    vocab_threshold_kernel<0><<<grid, kThrWarps * 32, 0, s>>>(gmax, ld_gmax, rows, n_groups, top_k, thresh, cand_count, fx);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_vocab_threshold(const float* gmax, long long ld_gmax, int rows, int n_groups, int top_k, float* thresh,
                                  int* cand_count, cudaStream_t s) {
  DH_ARG(gmax && thresh && cand_count && rows >= 0 && n_groups > 0 && ld_gmax >= n_groups && top_k >= 1);
  if (rows == 0) return DH_OK;
  return launch_threshold(gmax, ld_gmax, rows, n_groups, top_k, thresh, cand_count, ThrFix{}, s);
}

extern "C" int dh_vocab_threshold_fix(const float* gmax, long long ld_gmax, int rows, int n_groups, int top_k, float* thresh,
                                      int* cand_count, int count_min, int count_max, unsigned char* redo, int* any_flag,
                                      cudaStream_t s) {
  DH_ARG(gmax && thresh && cand_count && rows >= 0 && n_groups > 0 && ld_gmax >= n_groups && top_k >= 1);
  DH_ARG(redo && any_flag && count_min >= 0 && count_max >= count_min);
  if (rows == 0) return DH_OK;
  return launch_threshold(gmax, ld_gmax, rows, n_groups, top_k, thresh, cand_count, ThrFix{1, count_min, count_max, redo, any_flag},
                          s);
}

static size_t beam_smem_bytes(const dh_beam_state* st, int beam) {      // staged sequences + KV slot table of one image
  return (size_t)beam * (size_t)(st->seq_ld + (st->src ? st->S_alloc : 0)) * sizeof(int);
}

static int check_state(const dh_beam_state* st, int n_img, int beam) {
  if (!st || !st->seq || !st->val || !st->ended || !st->done || !st->final_len || !st->last_tok || !st->parent_state) return 0;
  if (n_img < 0 || beam < 1 || beam > kMaxBeam || st->seq_ld < 1) return 0;
  if (st->src && (st->S_alloc < 1 || st->S_alloc > kMaxSlots)) return 0;
  return 1;
}
static BeamState to_state(const dh_beam_state* st) {
  return BeamState{st->seq, st->seq_ld, st->val, st->ended, st->done, st->final_len, st->last_tok, st->parent_state,
                   st->src, st->S_alloc};
}

extern "C" int dh_beam_init(const dh_beam_state* st, const int* ind0, const float* val0, const int* prefix,
                            long long prefix_ld, int prefix_rows, int prefix_len, int n_img, int beam, int eos,
                            int lstm_semantics, cudaStream_t s) {
  DH_ARG(check_state(st, n_img, beam) && ind0 && val0 && prefix_len >= 0 && prefix_len < st->seq_ld);
  DH_ARG(prefix_len == 0 || (prefix && (prefix_rows == 1 || prefix_rows == n_img)));
  if (n_img == 0) return DH_OK;
  int total = n_img * beam;
  beam_init_kernel<<<dh_cdiv(total, 128), 128, 0, s>>>(to_state(st), ind0, val0, prefix, prefix_ld, prefix_rows, prefix_len,
                                                      n_img, beam, eos, lstm_semantics);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_beam_step(const dh_beam_state* st, const int* new_ind, const float* new_val, int n_img, int beam,
                            int step, int max_len, int eos, int lstm_semantics, float temperature, int noise_mode,
                            unsigned long long seed, long long image_base, const long long* dyn, cudaStream_t s) {
  DH_ARG(check_state(st, n_img, beam) && new_ind && new_val && temperature > 0.f && step >= 1);
  DH_ARG(st->seq_ld >= max_len && beam_smem_bytes(st, beam) <= 40 * 1024);
  if (n_img == 0) return DH_OK;
  StepParams p{new_ind, new_val, n_img, beam, step, max_len, eos, lstm_semantics, temperature, noise_mode, seed, image_base, dyn};
  beam_step_kernel<<<n_img, 32, beam_smem_bytes(st, beam), s>>>(to_state(st), p);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_beam_final(const dh_beam_state* st, int n_img, int beam, float temperature, int noise_mode,
                             unsigned long long seed, long long image_base, int final_step, int len_if_running, int pad,
                             int max_len, long long* out_ids, long long* out_len, const long long* dyn, cudaStream_t s) {
  DH_ARG(check_state(st, n_img, beam) && out_ids && out_len && temperature > 0.f && max_len <= st->seq_ld);
  if (n_img == 0) return DH_OK;
  beam_final_kernel<<<n_img, 32, 0, s>>>(to_state(st), n_img, beam, temperature, noise_mode, seed, image_base, final_step,
                                        len_if_running, pad, max_len, out_ids, out_len, dyn);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_token_logprob(const float* logits, long long ld, int rows, int V, const long long* targets, float* out,
                                cudaStream_t s) {
  DH_ARG(logits && targets && out && rows >= 0 && V > 0);
  if (rows == 0) return DH_OK;
  token_logprob_kernel<<<rows, 256, 0, s>>>(logits, ld, V, targets, out);
  DH_LAUNCH_OK();
  return DH_OK;
}

static int to_sparse(const dh_vocab_sparse* v, VocabSparse* out) {
  if (!v || !v->thresh || !v->hitmap || !v->logits || v->n_cols <= 0) return 0;
  const int bn = v->n_cols <= 64 ? 64 : v->n_cols <= 128 ? 128 : 256;
  const int nb = dh_cdiv(v->n_cols, bn);
  if (v->hit_ld < 2 * nb || v->ld < (long long)nb * bn || v->ld % 4 != 0 || ((uintptr_t)v->logits % 16) != 0) return 0;
  *out = VocabSparse{v->thresh, v->hitmap, v->hit_ld, v->logits, v->ld, 2 * nb, bn / 64};
  return 1;
}

static int launch_select_beam(const SelParams& p, const VocabSparse& vs, int n_img, int do_beam, const BeamState& st,
                              const StepParams& sp, size_t seq_bytes, cudaStream_t s, const LstmNext& nx = LstmNext{}) {
  static bool attr_set = false;
  if (!attr_set) {
    DH_CUDA(cudaFuncSetAttribute(select_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kMaxBeam * kWarpSelBytes + 40 * 1024));
    attr_set = true;
  }
  const size_t smem = (size_t)p.rpi * kWarpSelBytes + (do_beam ? seq_bytes : 0);
  select_beam_kernel<<<n_img, 32 * p.rpi, smem, s>>>(p, vs, do_beam, st, sp, nx);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_select_candidates(const dh_vocab_sparse* cand, int rows, int beam, int top_k, float temperature, int unk,
                                    int rows_per_image, int noise_mode, unsigned long long seed, long long image_base, int step,
                                    const unsigned char* done, int* ind, float* val, int* status, const long long* dyn,
                                    cudaStream_t s) {
  VocabSparse vs{};
  DH_ARG(to_sparse(cand, &vs) && ind && val && status && rows >= 0);
  DH_ARG(beam >= 1 && beam <= kMaxBeam && top_k >= 1 && beam <= top_k && temperature > 0.f);
  DH_ARG(rows_per_image >= 1 && rows_per_image <= kMaxBeam && rows % rows_per_image == 0);
  DH_ARG(noise_mode == DH_NOISE_DETERMINISTIC || noise_mode == DH_NOISE_INJECTED);
  if (rows == 0) return DH_OK;
  SelParams p{nullptr, 0, rows, 0, beam, top_k, unk, rows_per_image, temperature, noise_mode, seed, image_base, step,
              done, ind, val, status, dyn};
  return launch_select_beam(p, vs, rows / rows_per_image, 0, BeamState{}, StepParams{}, 0, s);
}

static int select_beam_step(const dh_vocab_sparse* cand, const dh_beam_state* st, int* ind, float* val, int* status, int n_img,
                            int beam, int top_k, float temperature, int unk, int step, int max_len, int eos, int lstm_semantics,
                            int noise_mode, unsigned long long seed, long long image_base, const long long* dyn,
                            const dh_lstm_operands* next, cudaStream_t s) {
  VocabSparse vs{};
  DH_ARG(to_sparse(cand, &vs) && ind && val && status);
  DH_ARG(check_state(st, n_img, beam) && top_k >= 1 && beam <= top_k && temperature > 0.f && step >= 1);
  DH_ARG(st->seq_ld >= max_len && beam_smem_bytes(st, beam) <= 40 * 1024);
  DH_ARG(noise_mode == DH_NOISE_DETERMINISTIC || noise_mode == DH_NOISE_INJECTED);
  LstmNext nx{};
  if (next) {
    DH_ARG(next->table && next->L >= 1 && next->L <= 8 && next->E % 8 == 0 && next->H % 8 == 0 && next->ldt % 8 == 0);
    nx.table = (const uint16_t*)next->table; nx.ldt = next->ldt; nx.n_tok_rows = next->n_tok_rows;
    nx.E = next->E; nx.H = next->H; nx.L = next->L;
    for (int l = 0; l < next->L; ++l) {
      DH_ARG(next->hs[l] && next->A[l] && next->lda[l] % 8 == 0 && next->in_off[l] % 8 == 0);
      DH_ARG(((uintptr_t)next->hs[l] % 16) == 0 && ((uintptr_t)next->A[l] % 16) == 0);
      nx.hs[l] = (const uint16_t*)next->hs[l]; nx.A[l] = (uint16_t*)next->A[l]; nx.lda[l] = next->lda[l];
      nx.in_off[l] = next->in_off[l];
    }
  }
  if (n_img == 0) return DH_OK;
  SelParams p{nullptr, 0, n_img * beam, 0, beam, top_k, unk, beam, temperature, noise_mode, seed, image_base, step,
              st->done, ind, val, status, dyn};
  StepParams sp{ind, val, n_img, beam, step, max_len, eos, lstm_semantics, temperature, noise_mode, seed, image_base, dyn};
  return launch_select_beam(p, vs, n_img, 1, to_state(st), sp, beam_smem_bytes(st, beam), s, nx);
}

extern "C" int dh_select_beam_step(const dh_vocab_sparse* cand, const dh_beam_state* st, int* ind, float* val, int* status,
                                   int n_img, int beam, int top_k, float temperature, int unk, int step, int max_len, int eos,
                                   int lstm_semantics, int noise_mode, unsigned long long seed, long long image_base,
                                   const long long* dyn, cudaStream_t s) {
  return select_beam_step(cand, st, ind, val, status, n_img, beam, top_k, temperature, unk, step, max_len, eos, lstm_semantics,
                          noise_mode, seed, image_base, dyn, nullptr, s);
}

extern "C" int dh_select_beam_step_lstm(const dh_vocab_sparse* cand, const dh_beam_state* st, int* ind, float* val, int* status,
                                        int n_img, int beam, int top_k, float temperature, int unk, int step, int max_len, int eos,
                                        int lstm_semantics, int noise_mode, unsigned long long seed, long long image_base,
                                        const long long* dyn, const dh_lstm_operands* next, cudaStream_t s) {
  DH_ARG(next);
  return select_beam_step(cand, st, ind, val, status, n_img, beam, top_k, temperature, unk, step, max_len, eos, lstm_semantics,
                          noise_mode, seed, image_base, dyn, next, s);
}

// out[row] = tlogit[row] - logsumexp(row) from the per-group (max, sum exp) pairs written by the contraction's epilogue.
namespace {
__global__ void __launch_bounds__(256) vocab_logprob_reduce_kernel(const float* __restrict__ gmax, const float* __restrict__ gsum,
                                                                  long long ld, int rows, int n_groups,
                                                                  const float* __restrict__ tlogit, float* __restrict__ out) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + w;
  if (r >= rows) return;
  const float* m = gmax + (long long)r * ld;
  const float* sg = gsum + (long long)r * ld;
  float mx = -INFINITY;
  for (int i = lane; i < n_groups; i += 32) mx = fmaxf(mx, __ldg(m + i));
  mx = dh_warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < n_groups; i += 32) {
    const float mi = __ldg(m + i);
    if (mi > -INFINITY) sum += __ldg(sg + i) * expf(mi - mx);
  }
  sum = dh_warp_sum(sum);
  if (lane == 0) out[r] = tlogit[r] - mx - logf(sum);
}
}  // namespace

int dh_vocab_logprob_reduce(const float* gmax, const float* gsum, long long ld, int rows, int n_groups, const float* tlogit,
                            float* out, cudaStream_t s) {
  if (rows == 0) return DH_OK;
  vocab_logprob_reduce_kernel<<<dh_cdiv(rows, 8), 256, 0, s>>>(gmax, gsum, ld, rows, n_groups, tlogit, out);
  DH_LAUNCH_OK();
  return DH_OK;
}
