// Probe: latency and throughput of legacy warp-level mma.sync m16n8k16 (bf16, fp32 accumulate) and ldmatrix on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate scripts/probes/hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int ILP>
__global__ void probe(float* out, long long* cyc, int iters) {
  float acc[ILP][4];
  uint32_t a[4] = {threadIdx.x, 2u, 3u, 4u};
  for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma_bf16(acc[i], a, 0x3f803f80u + i, 0x3f803f80u);
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  for (int warps = 1; warps <= 16; warps *= 2) {
    probe<1><<<148, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  ILP 1: %.1f cyc per dependent mma\n", warps, (double)h / iters);
    probe<8><<<148, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  ILP 8: %.1f cyc per mma per warp -> %.1f mma/cyc/SM = %.0f flop/cyc/SM\n", warps, (double)h / iters / 8,
           warps * 8.0 * iters / h, warps * 8.0 * iters / h * 4096);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
