// Shared helpers for libdeephumor_sm100.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/deephumor_b200.h"

int dh_fail(int code, const char* what, const char* file, int line);

#define DH_ARG(cond)                                                      \
  do {                                                                    \
    if (!(cond)) return dh_fail(DH_ERR_ARG, #cond, __FILE__, __LINE__);   \
  } while (0)

#define DH_LAUNCH_OK()                                                                  \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return dh_fail((int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define DH_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) return dh_fail((int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

static inline int dh_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- element type helpers
template <typename T> __device__ __forceinline__ float dh_to_f(T v);
template <> __device__ __forceinline__ float dh_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float dh_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float dh_to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T dh_from_f(float v);
template <> __device__ __forceinline__ __half dh_from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float dh_from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 dh_from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- counter-based hash (mirrors utils/synth.py)
#define DH_GOLD 0x9E3779B97F4A7C15ull
__host__ __device__ __forceinline__ uint64_t dh_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t dh_fold(uint64_t h, uint64_t field) {
  return dh_mix64(h + (field + 1ull) * DH_GOLD);
}
__host__ __device__ __forceinline__ uint64_t dh_key0(uint64_t seed) { return dh_mix64(seed + DH_GOLD); }
__host__ __device__ __forceinline__ uint32_t dh_bits24(uint64_t key, uint64_t idx) {
  return (uint32_t)(dh_fold(key, idx) >> 40);
}
// float32((m + 0.5) * 2^-23 - 1) * scale   (exactly the numpy expression)
__device__ __forceinline__ float dh_sym_uniform(uint64_t key, uint64_t idx, float scale) {
  double m = (double)dh_bits24(key, idx);
  float v = (float)((m + 0.5) * (1.0 / 8388608.0) - 1.0);
  return v * scale;
}
// Exp(1) race noise: float32(-log((m + 0.5) * 2^-24)) in float64 (oracle/noise.py)
__device__ __forceinline__ float dh_exp_noise(uint64_t row_key, uint64_t col) {
  double u = ((double)dh_bits24(row_key, col) + 0.5) * (1.0 / 16777216.0);
  return (float)(-log(u));
}
#define DH_NOISE_STREAM 0x4E5Aull
__host__ __device__ __forceinline__ uint64_t dh_noise_row_key(uint64_t seed, uint64_t image, uint64_t step,
                                                              uint64_t call, uint64_t row) {
  uint64_t h = dh_key0(seed);
  h = dh_fold(h, DH_NOISE_STREAM);
  h = dh_fold(h, image);
  h = dh_fold(h, step * 4ull + call);
  return dh_fold(h, row);
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float dh_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dh_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
