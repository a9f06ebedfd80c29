// Explicit im2col gathers feeding the tensor-core contraction kernel where im2col-mode TMA does not apply:
//  * the 7x7/2 stem (C_in = 3: a pixel is 6 B, below TMA's 16 B granule) -- gathered straight from the
//    NCHW fp32 input image, fusing the fp32 -> bf16 cast and the NCHW -> (r,s,c) re-layout;
//  * a generic NHWC bf16 gather used as the checked alternative to the im2col-TMA path (tests compare the two).
#include "common.cuh"

namespace {

// A[m, k] with m = (img, oh, ow), k = (r*KW + s)*3 + c for k < 147, zero for 147 <= k < Kp.  One thread = 8 k's (16 B).
template <typename T>
__global__ void im2col_stem_kernel(const float* __restrict__ img, T* __restrict__ A, int n, int H, int W, int Ho,
                                   int Wo, int KH, int KW, int stride, int pad, int Kp) {
  const int chunks = Kp / 8;
  const long long total = (long long)n * Ho * Wo * chunks;
  const int Kreal = KH * KW * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    long long m = i / chunks;
    const int ow = (int)(m % Wo);
    long long t = m / Wo;
    const int oh = (int)(t % Ho);
    const int im = (int)(t / Ho);
    __align__(16) T v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ch * 8 + e;
      float x = 0.f;
      if (k < Kreal) {
        const int c = k % 3, tap = k / 3;
        const int r = tap / KW, s = tap - r * KW;
        const int ih = oh * stride + r - pad, iw = ow * stride + s - pad;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) x = __ldg(img + (((long long)im * 3 + c) * H + ih) * W + iw);
      }
      v[e] = dh_from_f<T>(x);
    }
    *reinterpret_cast<uint4*>(A + m * Kp + ch * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// x [n,H,W,C] bf16 (C % 8 == 0) -> A[m, (r*KW + s)*C + c]
__global__ void im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ A, int n, int H, int W, int C,
                                   int Ho, int Wo, int KH, int KW, int stride, int pad) {
  const int cc = C / 8;
  const long long total = (long long)n * Ho * Wo * KH * KW * cc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cc);
    long long t = i / cc;
    const int tap = (int)(t % (KH * KW));
    const long long m = t / (KH * KW);
    const int ow = (int)(m % Wo);
    long long u = m / Wo;
    const int oh = (int)(u % Ho);
    const int im = (int)(u / Ho);
    const int r = tap / KW, s = tap - r * KW;
    const int ih = oh * stride + r - pad, iw = ow * stride + s - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W)
      v = *reinterpret_cast<const uint4*>(x + (((long long)im * H + ih) * W + iw) * C + c8 * 8);
    *reinterpret_cast<uint4*>(A + (m * KH * KW + tap) * C + c8 * 8) = v;
  }
}

inline int grid_for(long long items) {
  long long b = (items + 255) / 256;
  const long long cap = 148ll * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int dh_im2col_stem(const float* images_nchw, void* A, int n, int H, int W, int kh, int kw, int stride, int pad,
                              int k_padded, int out_dtype, cudaStream_t s) {
  DH_ARG(images_nchw && A && n >= 0 && k_padded % 8 == 0 && k_padded >= kh * kw * 3 && ((uintptr_t)A % 16) == 0);
  if (n == 0) return DH_OK;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long total = (long long)n * Ho * Wo * (k_padded / 8);
  if (out_dtype == DH_BF16)
    im2col_stem_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>(images_nchw, (__nv_bfloat16*)A, n, H, W, Ho, Wo, kh, kw, stride, pad, k_padded);
  else if (out_dtype == DH_F16)
    im2col_stem_kernel<__half><<<grid_for(total), 256, 0, s>>>(images_nchw, (__half*)A, n, H, W, Ho, Wo, kh, kw, stride, pad, k_padded);
  else
    return dh_fail(DH_ERR_ARG, "out_dtype", __FILE__, __LINE__);
  DH_LAUNCH_OK();
  return DH_OK;
}

extern "C" int dh_im2col_nhwc(const void* x, void* A, int n, int H, int W, int C, int kh, int kw, int stride, int pad,
                              cudaStream_t s) {
  DH_ARG(x && A && n >= 0 && C % 8 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)x % 16) == 0);
  if (n == 0) return DH_OK;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long total = (long long)n * Ho * Wo * kh * kw * (C / 8);
  im2col_nhwc_kernel<<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)A, n, H, W, C, Ho, Wo, kh, kw, stride, pad);
  DH_LAUNCH_OK();
  return DH_OK;
}
