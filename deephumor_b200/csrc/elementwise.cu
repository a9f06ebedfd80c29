// Bandwidth-bound kernels of the caption path: layout conversion, pooling, embedding gathers, LSTM cell,
// residual+LayerNorm, state gathers.  All are templated on the storage type (fp32 check mode / bf16), use
// 64-bit indexing, vectorised (16 B) accesses where the layout guarantees alignment, and grid-stride loops
// sized in multiples of the SM count.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
inline int grid_for(long long work_items, int per_block = kThreads) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = 148ll * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------ images NCHW fp32 -> NHWC4 (4th channel 0)
// With halo > 0 the output is [N, H+2*halo, W+2*halo, 4] with a zero border (stem conv TMA layout).
template <typename T>
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ in, T* __restrict__ out, int N, int H, int W, int halo) {
  const int Hp = H + 2 * halo, Wp = W + 2 * halo;
  long long total = (long long)N * Hp * Wp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int xw = (int)(i % Wp);
    long long t = i / Wp;
    int yh = (int)(t % Hp);
    int n = (int)(t / Hp);
    int y = yh - halo, x = xw - halo;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const float* p = in + ((long long)n * 3 * H + y) * W + x;
      v[0] = p[0];
      v[1] = p[(long long)H * W];
      v[2] = p[2ll * H * W];
    }
    T* o = out + i * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[c] = dh_from_f<T>(v[c]);
  }
}

// ------------------------------------------------------------------ maxpool 3x3 / 2, pad 1 (NHWC)
template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo) {
  long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int ow = (int)(t % Wo);
    t /= Wo;
    int oh = (int)(t % Ho);
    int n = (int)(t / Ho);
    float m = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      int ih = oh * 2 + dy - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        int iw = ow * 2 + dx - 1;
        if (iw < 0 || iw >= W) continue;
        m = fmaxf(m, dh_to_f<T>(x[(((long long)n * H + ih) * W + iw) * C + c]));
      }
    }
    y[i] = dh_from_f<T>(m);
  }
}

// ------------------------------------------------------------------ mean over HW: x [N,HW,C] -> out [N,C]
template <typename T, typename TO>
__global__ void avgpool_kernel(const T* __restrict__ x, TO* __restrict__ out, int N, int HW, int C, long long ldo) {
  long long total = (long long)N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int n = (int)(i / C);
    const T* p = x + (long long)n * HW * C + c;
    float s = 0.f;
    for (int j = 0; j < HW; ++j) s += dh_to_f<T>(p[(long long)j * C]);
    out[(long long)n * ldo + c] = dh_from_f<TO>(s / (float)HW);
  }
}

// 2-byte activations: 8 channels per thread with 16-byte loads, 7 rows in flight; per channel the same ascending-row
// summation as the scalar kernel (bit-identical results).
template <typename T, typename TO>
__global__ void __launch_bounds__(256) avgpool_vec_kernel(const T* __restrict__ x, TO* __restrict__ out, int N, int HW, int C,
                                                          long long ldo) {
  const int C8 = C >> 3;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)N * C8) return;
  const int c = (int)(i % C8) * 8, n = (int)(i / C8);
  const T* p = x + (long long)n * HW * C + c;
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  for (int j0 = 0; j0 < HW; j0 += 7) {
    uint4 v[7];
#pragma unroll
    for (int u = 0; u < 7; ++u)
      v[u] = j0 + u < HW ? __ldg(reinterpret_cast<const uint4*>(p + (long long)(j0 + u) * C)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      if (j0 + u < HW) {
        const T* e8 = reinterpret_cast<const T*>(&v[u]);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += dh_to_f<T>(e8[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) out[(long long)n * ldo + c + e] = dh_from_f<TO>(s[e] / (float)HW);
}

// ------------------------------------------------------------------ row gathers
// dst[r, 0:width] = src[idx ? idx[r] : r, 0:width]   (optionally scaled and with a second addend row)
template <typename TS, typename TD, typename TI>
__global__ void gather_rows_kernel(const TS* __restrict__ src, long long lds, const TI* __restrict__ idx, long long n_src_rows,
                                   TD* __restrict__ dst, long long ldd, int R, int width) {
  long long total = (long long)R * width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % width);
    int r = (int)(i / width);
    long long s = idx ? (long long)idx[r] : r;
    float v = (s >= 0 && s < n_src_rows) ? dh_to_f<TS>(src[s * lds + c]) : 0.f;
    dst[(long long)r * ldd + c] = dh_from_f<TD>(v);
  }
}

// label encoder: out[n] = mean_j table[ids[n, j]]   (models/encoders.py:104; mean over the full width, Q23)
template <typename T, typename TO>
__global__ void embed_mean_kernel(const T* __restrict__ table, long long ldt, const long long* __restrict__ ids, int L,
                                  TO* __restrict__ out, long long ldo, int N, int E) {
  long long total = (long long)N * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % E);
    int n = (int)(i / E);
    float s = 0.f;
    for (int j = 0; j < L; ++j) s += dh_to_f<T>(table[ids[(long long)n * L + j] * ldt + c]);
    out[(long long)n * ldo + c] = dh_from_f<TO>(s / (float)L);
  }
}

// ------------------------------------------------------------------ LSTM step operands in one launch
// Per row r: A[0][r, 0:E] = table[tok[r]] (next input embedding) and, for every layer l, A[l][r, in_l : in_l + H] =
// hs[l][parent[r]] (recurrent h through the beam parent; rnn_models.py:107,135-137).  16-byte copies, 2-byte elements.
constexpr int kMaxLstmLayers = 8;
struct LstmPrepParams {
  const uint16_t* table; long long ldt; const int* tok; int E; long long n_tok_rows;
  const int* parent; int H, L, rows;
  const uint16_t* hs[kMaxLstmLayers];       // [*, H] contiguous
  uint16_t* A[kMaxLstmLayers]; long long lda[kMaxLstmLayers]; int in_off[kMaxLstmLayers];
};
__global__ void lstm_prepare_kernel(LstmPrepParams p) {
  const int ce = p.E / 8, ch = p.H / 8;
  const int per_row = ce + p.L * ch;
  const long long total = (long long)p.rows * per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / per_row);
    int c = (int)(i - (long long)r * per_row);
    if (c < ce) {
      const long long t = p.tok[r];
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (t >= 0 && t < p.n_tok_rows) v = *reinterpret_cast<const uint4*>(p.table + t * p.ldt + c * 8);
      *reinterpret_cast<uint4*>(p.A[0] + (long long)r * p.lda[0] + c * 8) = v;
    } else {
      c -= ce;
      const int l = c / ch, cc = c - l * ch;
      const long long pr = p.parent ? p.parent[r] : r;
      *reinterpret_cast<uint4*>(p.A[l] + (long long)r * p.lda[l] + p.in_off[l] + cc * 8) =
          *reinterpret_cast<const uint4*>(p.hs[l] + pr * p.H + cc * 8);
    }
  }
}

// ------------------------------------------------------------------ LSTM cell (gate order i,f,g,o)
// gates [R,4H] fp32 (bias already added); c_prev rows gathered through parent[] (beam reorder folded into
// the read); h goes to up to two destinations (next layer's input slot and the
// recurrent staging buffer).
template <typename T>
__global__ void lstm_cell_kernel(const float* __restrict__ gates, long long ldg, const float* __restrict__ c_prev,
                                 const int* __restrict__ parent, float* __restrict__ c_out, T* __restrict__ h_out0,
                                 long long ldh0, T* __restrict__ h_out1, long long ldh1, int R, int H) {
  long long total = (long long)R * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % H);
    int r = (int)(i / H);
    const float* g = gates + (long long)r * ldg;
    float gi = g[j], gf = g[H + j], gg = g[2 * H + j], go = g[3 * H + j];
    long long pr = parent ? parent[r] : r;
    float cp = c_prev ? c_prev[pr * H + j] : 0.f;
    float si = 1.f / (1.f + expf(-gi)), sf = 1.f / (1.f + expf(-gf)), so = 1.f / (1.f + expf(-go));
    float c2 = sf * cp + si * tanhf(gg);
    float h2 = so * tanhf(c2);
    c_out[(long long)r * H + j] = c2;
    T hv = dh_from_f<T>(h2);
    if (h_out0) h_out0[(long long)r * ldh0 + j] = hv;
    if (h_out1) h_out1[(long long)r * ldh1 + j] = hv;
  }
}

// ------------------------------------------------------------------ out = LayerNorm(x + y) * gamma + beta (eps 1e-5)
// one warp per row; two-pass (mean, then centred variance) in fp32 like torch's native_layer_norm.
template <typename T>
__global__ void add_layernorm_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ y, long long ldy,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     T* __restrict__ out, long long ldo, int R, int D, float eps) {
  int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r < R; r += nwarps) {
    const T* xr = x + (long long)r * ldx;
    const T* yr = y ? y + (long long)r * ldy : nullptr;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += dh_to_f<T>(xr[c]) + (yr ? dh_to_f<T>(yr[c]) : 0.f);
    float mean = dh_warp_sum(s) / (float)D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) {
      float d = dh_to_f<T>(xr[c]) + (yr ? dh_to_f<T>(yr[c]) : 0.f) - mean;
      v += d * d;
    }
    float rstd = rsqrtf(dh_warp_sum(v) / (float)D + eps);
    for (int c = lane; c < D; c += 32) {
      float d = dh_to_f<T>(xr[c]) + (yr ? dh_to_f<T>(yr[c]) : 0.f) - mean;
      out[(long long)r * ldo + c] = dh_from_f<T>(d * rstd * gamma[c] + beta[c]);
    }
  }
}

// 2-byte types with D = 256 * NCH: every element is read ONCE (16-byte loads, 8 elements per lane and 256-column chunk),
// mean / centred variance / normalisation run on registers, 16-byte stores.  One warp per row.
template <typename T, int NCH>
__global__ void __launch_bounds__(256) add_layernorm_vec_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ y,
                                                                long long ldy, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, T* __restrict__ out,
                                                                long long ldo, int R, float eps) {
  constexpr int D = 256 * NCH;
  const int lane = threadIdx.x & 31;
  const int r = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  if (r >= R) return;
  float v[NCH * 8];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = k * 256 + lane * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(x + (long long)r * ldx + c);
    const T* ap = reinterpret_cast<const T*>(&a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[k * 8 + i] = dh_to_f<T>(ap[i]);
    if (y) {
      const uint4 b = *reinterpret_cast<const uint4*>(y + (long long)r * ldy + c);
      const T* bp = reinterpret_cast<const T*>(&b);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[k * 8 + i] += dh_to_f<T>(bp[i]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) s += v[i];
  const float mean = dh_warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) { v[i] -= mean; q += v[i] * v[i]; }
  const float rstd = rsqrtf(dh_warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = k * 256 + lane * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint4 o;
    T* op = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int i = 0; i < 8; ++i) op[i] = dh_from_f<T>(v[k * 8 + i] * rstd * gg[i] + bb[i]);
    *reinterpret_cast<uint4*>(out + (long long)r * ldo + c) = o;
  }
}

// ------------------------------------------------------------------ transformer input embedding
// x[r] = (pos==0 ? start[r / rows_per_start] : tok_table[tokens[r]]) / scale + pos_table[pos]
// (transformers.py:455-470; the image slot is scaled too, Q18).  tokens may be null when pos_of_row==0 everywhere.
template <typename T>
__global__ void xfmr_embed_kernel(const T* __restrict__ tok_table, const T* __restrict__ pos_table, long long ldt,
                                  const float* __restrict__ start, long long lds, int rows_per_start,
                                  const int* __restrict__ tokens, const int* __restrict__ positions, int pos_const,
                                  float scale, T* __restrict__ out, long long ldo, int R, int D) {
  long long total = (long long)R * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % D);
    int r = (int)(i / D);
    int pos = positions ? positions[r] : pos_const;
    float e = (pos == 0) ? start[(long long)(r / rows_per_start) * lds + c]
                         : dh_to_f<T>(tok_table[(long long)tokens[r] * ldt + c]);
    out[(long long)r * ldo + c] = dh_from_f<T>(e / scale + dh_to_f<T>(pos_table[(long long)pos * ldt + c]));
  }
}

// 2-byte tables: eight columns per thread with 16-byte loads / stores, same arithmetic per element as the scalar kernel
template <typename T>
__global__ void __launch_bounds__(256) xfmr_embed_vec_kernel(const T* __restrict__ tok_table, const T* __restrict__ pos_table,
                                                             long long ldt, const float* __restrict__ start, long long lds,
                                                             int rows_per_start, const int* __restrict__ tokens,
                                                             const int* __restrict__ positions, int pos_const, float scale,
                                                             T* __restrict__ out, long long ldo, int R, int D) {
  const int d8 = D >> 3;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)R * d8) return;
  const int r = (int)(i / d8), c = (int)(i - (long long)r * d8) * 8;
  const int pos = positions ? positions[r] : pos_const;
  float e[8];
  if (pos == 0) {
    const float4* sp = reinterpret_cast<const float4*>(start + (long long)(r / rows_per_start) * lds + c);
    const float4 a = __ldg(sp), b = __ldg(sp + 1);
    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
  } else {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(tok_table + (long long)tokens[r] * ldt + c));
    const T* t8 = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = dh_to_f<T>(t8[j]);
  }
  const uint4 pv = __ldg(reinterpret_cast<const uint4*>(pos_table + (long long)pos * ldt + c));
  const T* p8 = reinterpret_cast<const T*>(&pv);
  uint4 o;
  T* o8 = reinterpret_cast<T*>(&o);
#pragma unroll
  for (int j = 0; j < 8; ++j) o8[j] = dh_from_f<T>(e[j] / scale + dh_to_f<T>(p8[j]));
  *reinterpret_cast<uint4*>(out + (long long)r * ldo + c) = o;
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = dh_from_f<TD>(dh_to_f<TS>(src[i]));
}

}  // namespace

#define DH_DISPATCH(dtype, ...)                                         \
  do {                                                                  \
    if ((dtype) == DH_F32) { using T = float; __VA_ARGS__; }            \
    else if ((dtype) == DH_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else if ((dtype) == DH_F16) { using T = __half; __VA_ARGS__; }      \
    else return dh_fail(DH_ERR_ARG, "dtype", __FILE__, __LINE__);       \
  } while (0)

extern "C" int dh_nchw_to_nhwc4(const float* images, void* out, int n, int H, int W, int halo, int dtype, cudaStream_t s) {
  DH_ARG(images && out && n >= 0 && halo >= 0);
  if (n == 0) return DH_OK;
  long long total = (long long)n * (H + 2 * halo) * (W + 2 * halo);
  DH_DISPATCH(dtype, (nchw_to_nhwc4_kernel<T><<<grid_for(total), kThreads, 0, s>>>(images, (T*)out, n, H, W, halo)));
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_maxpool3x3s2(const void* x, void* y, int n, int H, int W, int C, int dtype, cudaStream_t s) {
  DH_ARG(x && y && n >= 0);
  if (n == 0) return DH_OK;
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)n * Ho * Wo * C;
  DH_DISPATCH(dtype, (maxpool3x3s2_kernel<T><<<grid_for(total), kThreads, 0, s>>>((const T*)x, (T*)y, n, H, W, C, Ho, Wo)));
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_avgpool(const void* x, void* out, long long ldo, int n, int HW, int C, int dtype, int out_dtype,
                          cudaStream_t s) {
  DH_ARG(x && out && n >= 0 && HW > 0);
  if (n == 0) return DH_OK;
  long long total = (long long)n * C;
  if (dtype != DH_F32 && out_dtype == DH_F32 && C % 8 == 0 && ((uintptr_t)x % 16) == 0) {
    const int gv = dh_cdiv(total / 8, 256);
    if (dtype == DH_BF16) avgpool_vec_kernel<__nv_bfloat16, float><<<gv, 256, 0, s>>>((const __nv_bfloat16*)x, (float*)out, n, HW, C, ldo);
    else avgpool_vec_kernel<__half, float><<<gv, 256, 0, s>>>((const __half*)x, (float*)out, n, HW, C, ldo);
    DH_LAUNCH_OK();
    return DH_OK;
  }
  if (out_dtype == DH_F32)
    DH_DISPATCH(dtype, (avgpool_kernel<T, float><<<grid_for(total), kThreads, 0, s>>>((const T*)x, (float*)out, n, HW, C, ldo)));
  else
    DH_DISPATCH(dtype, (avgpool_kernel<T, __nv_bfloat16><<<grid_for(total), kThreads, 0, s>>>((const T*)x, (__nv_bfloat16*)out, n, HW, C, ldo)));
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_gather_rows(const void* src, long long lds, long long n_src_rows, const int* idx, void* dst, long long ldd,
                              int rows, int width, int src_dtype, int dst_dtype, cudaStream_t s) {
  DH_ARG(src && dst && rows >= 0 && width > 0);
  if (rows == 0) return DH_OK;
  long long total = (long long)rows * width;
  int g = grid_for(total);
  if (src_dtype == DH_F32 && dst_dtype == DH_F32)
    gather_rows_kernel<float, float, int><<<g, kThreads, 0, s>>>((const float*)src, lds, idx, n_src_rows, (float*)dst, ldd, rows, width);
  else if (src_dtype == DH_F32 && dst_dtype == DH_BF16)
    gather_rows_kernel<float, __nv_bfloat16, int><<<g, kThreads, 0, s>>>((const float*)src, lds, idx, n_src_rows, (__nv_bfloat16*)dst, ldd, rows, width);
  else if (src_dtype == DH_BF16 && dst_dtype == DH_BF16)
    gather_rows_kernel<__nv_bfloat16, __nv_bfloat16, int><<<g, kThreads, 0, s>>>((const __nv_bfloat16*)src, lds, idx, n_src_rows, (__nv_bfloat16*)dst, ldd, rows, width);
  else if (src_dtype == DH_BF16 && dst_dtype == DH_F32)
    gather_rows_kernel<__nv_bfloat16, float, int><<<g, kThreads, 0, s>>>((const __nv_bfloat16*)src, lds, idx, n_src_rows, (float*)dst, ldd, rows, width);
  else
    return dh_fail(DH_ERR_ARG, "dtype", __FILE__, __LINE__);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_embed_mean(const void* table, long long ldt, const long long* ids, int L, void* out, long long ldo, int n,
                             int E, int dtype, int out_dtype, cudaStream_t s) {
  DH_ARG(table && ids && out && n >= 0 && L > 0);
  if (n == 0) return DH_OK;
  long long total = (long long)n * E;
  if (out_dtype == DH_F32)
    DH_DISPATCH(dtype, (embed_mean_kernel<T, float><<<grid_for(total), kThreads, 0, s>>>((const T*)table, ldt, ids, L, (float*)out, ldo, n, E)));
  else
    DH_DISPATCH(dtype, (embed_mean_kernel<T, __nv_bfloat16><<<grid_for(total), kThreads, 0, s>>>((const T*)table, ldt, ids, L, (__nv_bfloat16*)out, ldo, n, E)));
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_lstm_cell(const float* gates, long long ldg, const float* c_prev, const int* parent, float* c_out,
                            void* h_out0, long long ldh0, void* h_out1, long long ldh1, int rows, int H, int dtype,
                            cudaStream_t s) {
  DH_ARG(gates && c_out && rows >= 0 && H > 0);
  if (rows == 0) return DH_OK;
  long long total = (long long)rows * H;
  DH_DISPATCH(dtype, (lstm_cell_kernel<T><<<grid_for(total), kThreads, 0, s>>>(gates, ldg, c_prev, parent, c_out, (T*)h_out0, ldh0,
                                                                             (T*)h_out1, ldh1, rows, H)));
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_add_layernorm(const void* x, long long ldx, const void* y, long long ldy, const float* gamma,
                                const float* beta, void* out, long long ldo, int rows, int D, int dtype, cudaStream_t s) {
  DH_ARG(x && gamma && beta && out && rows >= 0 && D > 0);
  if (rows == 0) return DH_OK;
  // 2-byte activations with 256 / 512 / 1024 columns and 16-byte aligned rows: single-read vectorised kernel
  if (dtype != DH_F32 && (D == 256 || D == 512 || D == 1024) && ldx % 8 == 0 && ldo % 8 == 0 && (!y || ldy % 8 == 0) &&
      ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && (!y || ((uintptr_t)y % 16) == 0) &&
      ((uintptr_t)gamma % 16) == 0 && ((uintptr_t)beta % 16) == 0) {
    const int gv = dh_cdiv(rows, 8);
#define DH_LN_VEC(T, NCH) add_layernorm_vec_kernel<T, NCH><<<gv, 256, 0, s>>>((const T*)x, ldx, (const T*)y, ldy, gamma, beta, (T*)out, ldo, rows, 1e-5f)
    if (dtype == DH_BF16) { if (D == 256) DH_LN_VEC(__nv_bfloat16, 1); else if (D == 512) DH_LN_VEC(__nv_bfloat16, 2); else DH_LN_VEC(__nv_bfloat16, 4); }
    else { if (D == 256) DH_LN_VEC(__half, 1); else if (D == 512) DH_LN_VEC(__half, 2); else DH_LN_VEC(__half, 4); }
#undef DH_LN_VEC
    DH_LAUNCH_OK();
    return DH_OK;
  }
  int g = grid_for((long long)rows * 32);
  DH_DISPATCH(dtype, (add_layernorm_kernel<T><<<g, kThreads, 0, s>>>((const T*)x, ldx, (const T*)y, ldy, gamma, beta, (T*)out, ldo,
                                                                  rows, D, 1e-5f)));
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_xfmr_embed(const void* tok_table, const void* pos_table, long long ldt, const float* start, long long lds,
                             int rows_per_start, const int* tokens, const int* positions, int pos_const, float scale,
                             void* out, long long ldo, int rows, int D, int dtype, cudaStream_t s) {
  DH_ARG(tok_table && pos_table && start && out && rows >= 0 && rows_per_start > 0);
  DH_ARG(tokens || (!positions && pos_const == 0));
  if (rows == 0) return DH_OK;
  long long total = (long long)rows * D;
  if (dtype != DH_F32 && D % 8 == 0 && ldt % 8 == 0 && ldo % 8 == 0 && lds % 4 == 0 && ((uintptr_t)tok_table % 16) == 0 &&
      ((uintptr_t)pos_table % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)start % 16) == 0) {
    const int g = dh_cdiv(total / 8, 256);
    if (dtype == DH_BF16)
      xfmr_embed_vec_kernel<__nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)tok_table, (const __nv_bfloat16*)pos_table, ldt,
                                                             start, lds, rows_per_start, tokens, positions, pos_const, scale,
                                                             (__nv_bfloat16*)out, ldo, rows, D);
    else
      xfmr_embed_vec_kernel<__half><<<g, 256, 0, s>>>((const __half*)tok_table, (const __half*)pos_table, ldt, start, lds,
                                                      rows_per_start, tokens, positions, pos_const, scale, (__half*)out, ldo, rows, D);
    DH_LAUNCH_OK();
    return DH_OK;
  }
  DH_DISPATCH(dtype, (xfmr_embed_kernel<T><<<grid_for(total), kThreads, 0, s>>>((const T*)tok_table, (const T*)pos_table, ldt, start,
                                                                              lds, rows_per_start, tokens, positions, pos_const,
                                                                              scale, (T*)out, ldo, rows, D)));
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_cast(const void* src, void* dst, long long n, int src_dtype, int dst_dtype, cudaStream_t s) {
  DH_ARG(src && dst && n >= 0);
  if (n == 0) return DH_OK;
  int g = grid_for(n);
  if (src_dtype == DH_F32 && dst_dtype == DH_BF16)
    cast_kernel<float, __nv_bfloat16><<<g, kThreads, 0, s>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == DH_BF16 && dst_dtype == DH_F32)
    cast_kernel<__nv_bfloat16, float><<<g, kThreads, 0, s>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == DH_F32 && dst_dtype == DH_F16)
    cast_kernel<float, __half><<<g, kThreads, 0, s>>>((const float*)src, (__half*)dst, n);
  else if (src_dtype == DH_F16 && dst_dtype == DH_F32)
    cast_kernel<__half, float><<<g, kThreads, 0, s>>>((const __half*)src, (float*)dst, n);
  else if (src_dtype == DH_F32 && dst_dtype == DH_F32)
    cast_kernel<float, float><<<g, kThreads, 0, s>>>((const float*)src, (float*)dst, n);
  else
    return dh_fail(DH_ERR_ARG, "dtype", __FILE__, __LINE__);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_lstm_prepare(const void* table, long long ldt, long long n_tok_rows, const int* tok, int E, const int* parent,
                               const void* const* hs, void* const* A, const long long* lda, const int* in_off, int L, int H,
                               int rows, cudaStream_t s) {
  DH_ARG(table && tok && hs && A && lda && in_off && L >= 1 && L <= kMaxLstmLayers && rows >= 0);
  DH_ARG(E % 8 == 0 && H % 8 == 0 && ldt % 8 == 0);
  if (rows == 0) return DH_OK;
  LstmPrepParams p{};
  p.table = (const uint16_t*)table; p.ldt = ldt; p.tok = tok; p.E = E; p.n_tok_rows = n_tok_rows;
  p.parent = parent; p.H = H; p.L = L; p.rows = rows;
  for (int l = 0; l < L; ++l) {
    DH_ARG(hs[l] && A[l] && lda[l] % 8 == 0 && in_off[l] % 8 == 0);
    p.hs[l] = (const uint16_t*)hs[l]; p.A[l] = (uint16_t*)A[l]; p.lda[l] = lda[l]; p.in_off[l] = in_off[l];
  }
  const long long total = (long long)rows * (E / 8 + L * (H / 8));
  lstm_prepare_kernel<<<grid_for(total), kThreads, 0, s>>>(p);
  DH_LAUNCH_OK();
  return DH_OK;
}
