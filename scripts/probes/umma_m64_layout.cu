// Probe: where do the rows of a tcgen05.mma cta_group::1 M = 64 accumulator land in TMEM?
// A[64][16] = row index (bf16, K-major SW128 tile with only K chunk 0..15 used), B[N=64][16] with B[n][0] = 1, others 0
// => D[m][n] = m for every n.  All 128 TMEM lanes, columns 0..63 are dumped.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__global__ void probe(float* out, int M) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - smem_u32(raw));
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(g);              // 128 rows x 128 B
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(g + 16384);      // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(g + 16384 + 8192);
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(g + 16384 + 8192 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0u;
  __syncthreads();
  // element (row, k) of a SW128 K-major tile: byte row*128 + (((k/8) ^ (row&7)) << 4) + (k%8)*2
  if (tid < 128) {
    const int row = tid;
    for (int k = 0; k < 16; ++k)
      *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(A) + row * 128 + ((((k >> 3) ^ (row & 7))) << 4) + (k & 7) * 2) =
          __float2bfloat16(k == 0 ? (float)row : 0.f);
    if (row < 64)
      *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(B) + row * 128 + (((0 ^ (row & 7))) << 4)) = __float2bfloat16(1.f);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_word);
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                 "l"(umma_desc(base)), "l"(umma_desc(base + 16384)), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) out[tid * 64 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int M : {128, 64}) {
    cudaMemset(d, 0xff, 128 * 64 * 4);
    probe<<<1, 128, 40000>>>(d, M);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[128 * 64];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%d status=%s\n", M, cudaGetErrorString(e));
    for (int lane = 0; lane < 128; ++lane) printf("%s%g/%g", lane % 16 ? " " : "\n  ", h[lane * 64], h[lane * 64 + 63]);
    printf("\n");
  }
  return 0;
}
