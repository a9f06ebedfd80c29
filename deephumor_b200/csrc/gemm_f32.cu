// fp32 "check mode" contraction kernel: C[M,N] = act(A[M,K] * W[N,K]^T + bias[N] + residual[M,N]).
//
// One SIMT (FFMA, true fp32 operands and accumulation) tiled kernel serves both the plain linear
// layers and -- with the A-tile loader switched to an on-the-fly NHWC im2col gather -- every
// convolution of the ResNet-50 trunk.  It exists because bit-exact token parity with the fp32
// reference needs true fp32 products (TF32/bf16 tensor-core math is not accurate enough,
// SURVEY.md 7.3 item 1); the throughput path is the tcgen05 kernel in gemm_tc.cu.
//
// Tile TxTx16 (T = 128: 8x8 outputs per thread; T = 64: 4x4, picked when 128-tiles would leave SMs idle, e.g. the
// [N,2048]x[2048,E] global head), 256 threads, register-prefetch double buffering.
#include "common.cuh"

namespace {

constexpr int BK = 16;

struct ConvGeom {
  int H, W, C, Ho, Wo, kh, kw, stride, pad;  // C = input channels as stored (multiple of 4)
};

template <bool CONV>
__device__ __forceinline__ float4 load_a(const float* __restrict__ A, long long lda, int M, int K, int row, int k,
                                         const ConvGeom& g, int n_img, int oh, int ow) {
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= M || k >= K) return z;
  if (!CONV) return *reinterpret_cast<const float4*>(A + (long long)row * lda + k);
  int c = k % g.C;
  int t = k / g.C;
  int kw = t % g.kw, kh = t / g.kw;
  int ih = oh * g.stride + kh - g.pad, iw = ow * g.stride + kw - g.pad;
  if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return z;
  return *reinterpret_cast<const float4*>(A + (((long long)n_img * g.H + ih) * g.W + iw) * g.C + c);
}

template <bool CONV, int T>
__global__ void __launch_bounds__(256)
igemm_f32_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Wt, long long ldw,
                 const float* __restrict__ bias, const float* __restrict__ res, long long ldr,
                 float* __restrict__ C, long long ldc, int M, int N, int K, int relu, ConvGeom g) {
  constexpr int BM = T, BN = T, LDS = T + 4, R = T / 16, H = T / 2, Q = R / 2;
  constexpr int NL = T / 64;               // float4 (along k) per thread per operand
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow0 = tid >> 2, lrow1 = lrow0 + 64, lk = (tid & 3) * 4;
  int nimg[2] = {0, 0}, oh[2] = {0, 0}, ow[2] = {0, 0};
  if (CONV) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = m0 + (i ? lrow1 : lrow0);
      int hw = g.Ho * g.Wo;
      nimg[i] = r / hw;
      int rem = r - nimg[i] * hw;
      oh[i] = rem / g.Wo;
      ow[i] = rem - oh[i] * g.Wo;
    }
  }
  auto fetch = [&](int k0, float4* ra, float4* rb) {
    ConvGeom dummy{};
    ra[0] = load_a<CONV>(A, lda, M, K, m0 + lrow0, k0 + lk, g, nimg[0], oh[0], ow[0]);
    rb[0] = load_a<false>(Wt, ldw, N, K, n0 + lrow0, k0 + lk, dummy, 0, 0, 0);
    if (NL == 2) {
      ra[1] = load_a<CONV>(A, lda, M, K, m0 + lrow1, k0 + lk, g, nimg[1], oh[1], ow[1]);
      rb[1] = load_a<false>(Wt, ldw, N, K, n0 + lrow1, k0 + lk, dummy, 0, 0, 0);
    }
  };
  auto stash = [&](int buf, const float4* ra, const float4* rb) {
    const float* a0 = reinterpret_cast<const float*>(&ra[0]);
    const float* a1 = reinterpret_cast<const float*>(&ra[1]);
    const float* b0 = reinterpret_cast<const float*>(&rb[0]);
    const float* b1 = reinterpret_cast<const float*>(&rb[1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][lk + i][lrow0] = a0[i];
      Bs[buf][lk + i][lrow0] = b0[i];
      if (NL == 2) {
        As[buf][lk + i][lrow1] = a1[i];
        Bs[buf][lk + i][lrow1] = b1[i];
      }
    }
  };
  const int ty = tid >> 4, tx = tid & 15;
  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  fetch(0, ra, rb);
  stash(0, ra, rb);
  __syncthreads();
  const int nk = (K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK, ra, rb);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[R], b[R];
      if (R == 8) {
        float4 a_lo = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 a_hi = *reinterpret_cast<const float4*>(&As[buf][k][H + ty * 4]);
        float4 b_lo = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        float4 b_hi = *reinterpret_cast<const float4*>(&Bs[buf][k][H + tx * 4]);
        a[0] = a_lo.x; a[1] = a_lo.y; a[2] = a_lo.z; a[3] = a_lo.w;
        a[R - 4] = a_hi.x; a[R - 3] = a_hi.y; a[R - 2] = a_hi.z; a[R - 1] = a_hi.w;
        b[0] = b_lo.x; b[1] = b_lo.y; b[2] = b_lo.z; b[3] = b_lo.w;
        b[R - 4] = b_hi.x; b[R - 3] = b_hi.y; b[R - 2] = b_hi.z; b[R - 1] = b_hi.w;
      } else {
        float2 a_lo = *reinterpret_cast<const float2*>(&As[buf][k][ty * 2]);
        float2 a_hi = *reinterpret_cast<const float2*>(&As[buf][k][H + ty * 2]);
        float2 b_lo = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
        float2 b_hi = *reinterpret_cast<const float2*>(&Bs[buf][k][H + tx * 2]);
        a[0] = a_lo.x; a[1] = a_lo.y; a[R - 2] = a_hi.x; a[R - 1] = a_hi.y;
        b[0] = b_lo.x; b[1] = b_lo.y; b[R - 2] = b_hi.x; b[R - 1] = b_hi.y;
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1, ra, rb);
      __syncthreads();
    }
  }
  // epilogue: bias + residual + relu
#pragma unroll
  for (int i = 0; i < R; ++i) {
    int row = m0 + (i < Q ? ty * Q + i : H + ty * Q + (i - Q));
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int col = n0 + (j < Q ? tx * Q + j : H + tx * Q + (j - Q));
      if (col >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[col];
      if (res) v += res[(long long)row * ldr + col];
      if (relu) v = fmaxf(v, 0.f);
      C[(long long)row * ldc + col] = v;
    }
  }
}

}  // namespace

extern "C" int dh_gemm_f32(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                           const float* residual, long long ldr, float* C, long long ldc, int M, int N, int K,
                           int relu, cudaStream_t stream) {
  DH_ARG(A && W && C && M >= 0 && N > 0 && K > 0);
  DH_ARG(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0);
  DH_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0);
  if (M == 0) return DH_OK;
  ConvGeom g{};
  if ((long long)dh_cdiv(N, 128) * dh_cdiv(M, 128) < 148) {
    dim3 grid(dh_cdiv(N, 64), dh_cdiv(M, 64));
    igemm_f32_kernel<false, 64><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, relu, g);
  } else {
    dim3 grid(dh_cdiv(N, 128), dh_cdiv(M, 128));
    igemm_f32_kernel<false, 128><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, relu, g);
  }
  DH_LAUNCH_OK();
  return DH_OK;
}

// x [n,H,W,Cin] NHWC fp32 (Cin % 4 == 0), w [Cout][kh][kw][Cin] (BN folded), y [n,Ho,Wo,Cout].
extern "C" int dh_conv2d_f32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                             int n, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int relu,
                             cudaStream_t stream) {
  DH_ARG(x && w && y && n >= 0 && Cin % 4 == 0 && Cout > 0 && stride > 0);
  if (n == 0) return DH_OK;
  ConvGeom g;
  g.H = H; g.W = W; g.C = Cin; g.kh = kh; g.kw = kw; g.stride = stride; g.pad = pad;
  g.Ho = (H + 2 * pad - kh) / stride + 1;
  g.Wo = (W + 2 * pad - kw) / stride + 1;
  long long M = (long long)n * g.Ho * g.Wo;
  DH_ARG(M < (1ll << 31));
  int K = kh * kw * Cin;
  dim3 grid(dh_cdiv(Cout, 128), dh_cdiv(M, 128));
  igemm_f32_kernel<true, 128><<<grid, 256, 0, stream>>>(x, 0, w, K, bias, residual, Cout, y, Cout, (int)M, Cout, K, relu, g);
  DH_LAUNCH_OK();
  return DH_OK;
}
