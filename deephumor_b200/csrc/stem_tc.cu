// Fused ResNet-50 stem for sm_100a: conv 7x7/2 (3 -> 64, BN folded) + ReLU + maxpool 3x3/2, straight from the NCHW fp32
// image to the NHWC 56x56x64 tensor that layer1 reads (torchvision resnet.py:197-200,268-271 <- encoders.py:56).
//
// The stem's A operand cannot come from im2col-mode TMA (a pixel is 3 channels = 6 B, below TMA's 16 B granule) and
// materialising it in HBM costs 4.8 MB/image.  Here the CTA builds it in shared memory itself:
//
//   work item   = one image x two pooled rows (112 pooled pixels = the 128 TMEM lanes of one accumulator, 16 idle)
//   input band  = the 15 input rows those pixels depend on, fp32 -> half, pixel-interleaved [row][col][rgb] in smem
//   tap (dy,dx) = the conv output at (2py-1+dy, 2px-1+dx) of EVERY lane's pooled pixel: one 128 x 64 x 160 UMMA
//                 (tcgen05.mma kind::f16, fp32 accumulators in TMEM) whose A operand lives in TENSOR MEMORY: the 256
//                 threads gather their pixel's patch from the band with 4-byte loads and tcgen05.st it into their own
//                 TMEM lane (K layout: 7 kernel rows x 22 slots = 21 (s,c) values + one junk value that meets a zero
//                 weight; padded to 160 with zeros).  With A in shared memory every N = 64 MMA was bound by re-reading
//                 its 128 x 16 A slice (128 cycles instead of 32) and the tile cost an extra smem write pass.
//   pooling     = max over the taps' accumulators of a lane, in registers (+bias, ReLU after the max: both monotone);
//                 taps that fall into the pool's padding are skipped per lane.  Only the six taps dx = 1, 2 are computed:
//                 tap (dy, 0) of a pooled pixel is tap (dy, 2) of its left neighbour = the neighbouring lane's accumulator.
//
// Computing each conv pixel once per pooling window costs 2.25x the conv FLOPs (0.53 GFLOP/image on a >1 PFLOP/s pipe)
// and removes every intermediate from HBM: the kernel reads 602 KB and writes 401 KB per image.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPitch = 696;                 // halves per band row: (5 + 224 + 2 pad pixels) * 3 channels + junk, even
constexpr int kBandRows = 15;
constexpr int kKp = 192;                    // padded K: 7 * 22 = 154 real slots
constexpr int kWBytes = 3 * 64 * 128;       // W tile: 3 K-chunks of [64 rows x 128 B]
constexpr int kBandBytes = kBandRows * kPitch * 2;
constexpr int kSmemBytes = 1024 + kWBytes + ((kBandBytes + 127) / 128) * 128 + 256 + 128 + 8 * 32 * 4;
constexpr int kACols = 80;                  // A operand in TMEM: 160 halves = 80 32-bit columns per lane
constexpr uint32_t kTmemCols = 256;         // 2 accumulator slots x 64 + A (80), power of two: two CTAs share an SM's 512

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000ll) __trap();   // a protocol bug must fail the launch, never hang the GPU
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
// A from tensor memory (lane = row, 32-bit column j = K elements 2 j, 2 j + 1), B from shared memory
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tc_st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// K-major, 128B-swizzled operand tile: start >> 4 | SBO = 1024 B | descriptor version 1 | SWIZZLE_128B (see gemm_tc.cu).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <typename T> struct Pack2;
template <> struct Pack2<__half> {
  static __device__ __forceinline__ uint32_t f(float a, float b) { __half2 t = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&t); }
};
template <> struct Pack2<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t f(float a, float b) { __nv_bfloat162 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&t); }
};

// Per-channel input normalisation of the uint8 path: (x / 255 - mean) / std, the exact fp32 operations of torchvision's
// ToTensor + Normalize (deephumor_demo.ipynb cell 11), so a uint8 batch gives bit-identical features to the float one.
struct PixelNorm { float mean[3], std[3]; };

template <typename TIN> struct BandLoad;
template <> struct BandLoad<float> {
  using Vec = float4;
  static __device__ __forceinline__ Vec zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void unpack(const Vec& v, int, const PixelNorm&, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct BandLoad<unsigned char> {
  using Vec = uchar4;
  static __device__ __forceinline__ Vec zero() { return make_uchar4(0, 0, 0, 0); }
  static __device__ __forceinline__ void unpack(const Vec& v, int c, const PixelNorm& nm, float* o) {
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v.x, 255.f), nm.mean[c]), nm.std[c]);
    o[1] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v.y, 255.f), nm.mean[c]), nm.std[c]);
    o[2] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v.z, 255.f), nm.mean[c]), nm.std[c]);
    o[3] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v.w, 255.f), nm.mean[c]), nm.std[c]);
  }
};

// images [n,3,224,224] fp32 (or uint8 + PixelNorm); wp [64][192] T (k = r*22 + s*3 + c, zero elsewhere); bias [64];
// out [n,56,56,64] T.
template <typename T, typename TIN>
__global__ void __launch_bounds__(kThreads, 2)
stem_pool_kernel(const TIN* __restrict__ images, const T* __restrict__ wp, const float* __restrict__ bias,
                 T* __restrict__ out, int n_img, const PixelNorm nm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_buf = base;
  T* band = reinterpret_cast<T*>(gbase + kWBytes);
  constexpr int kBandPad = ((kBandBytes + 127) / 128) * 128;
  float* bias_s = reinterpret_cast<float*>(gbase + kWBytes + kBandPad);
  const uint32_t bars = base + kWBytes + kBandPad + 256u;     // mma_done[2]: the MMAs of the taps of each parity
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(gbase + kWBytes + kBandPad + 256 + 64);
  float* xch = reinterpret_cast<float*>(gbase + kWBytes + kBandPad + 256 + 128);   // [8 warps][32]: last lane's accumulator
  auto mma_done = [&](uint32_t b) { return bars + 8u * b; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- one-time setup: barriers, TMEM, zeroed band (pad columns stay zero), weights
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(bars + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < (kWBytes + kBandPad) / 16; i += kThreads)
    *reinterpret_cast<uint4*>(gbase + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 64) bias_s[tid] = bias[tid];
  __syncthreads();
  for (int i = tid; i < 64 * (kKp / 8); i += kThreads) {     // 16 B chunks of wp -> swizzled K-major tiles
    const int row = i / (kKp / 8), c = i % (kKp / 8);
    const uint4 v = *reinterpret_cast<const uint4*>(wp + row * kKp + c * 8);
    sts128(w_buf + (uint32_t)((c >> 3) * 64 * 128 + row * 128 + (((c & 7) ^ (row & 7)) << 4)), v.x, v.y, v.z, v.w);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_word);
  const uint32_t tmem_a = tmem_base + 128u;                 // columns 128 .. 207: the A operand
  // kind::f16 instruction descriptor: D fp32, A/B half (0) or bf16 (1), K-major, N = 64, M = 128
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  // builder role: row m of the A operand = TMEM lane m (warps w and w + 4 share lane quadrant w % 4); half 0 writes the
  // columns of kernel rows 0..3 (words 0..43), half 1 those of rows 4..6 (words 44..76) and the zero padding (77..79)
  const int m = tid & 127, half = tid >> 7;
  const bool m_ok = m < 112;
  const int mc = m_ok ? m : 0;                 // idle lanes gather lane 0's patch: finite values, rows never read back
  const int m_py = mc / 56, m_px = mc - m_py * 56;
  const uint32_t a_lane = tmem_a + ((uint32_t)((warp & 3) * 32) << 16);
  // epilogue role: TMEM lane quadrant q, channel half ch0
  const int q = warp & 3, ch0 = (warp >> 2) * 32;
  const int e_m = q * 32 + lane;
  const int e_py = e_m / 56, e_px = e_m - e_py * 56;
  const bool e_ok = e_m < 112;

  const int items = n_img * 28;
  uint32_t n_tap = 0;                           // taps issued so far by this CTA (uniform): slot / barrier parity bookkeeping
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int img = item / 28, py0 = (item - img * 28) * 2;
    // ---- input band: rows 4*py0-5 .. 4*py0+9, fp32 NCHW -> T [row][5 + col][rgb]   (the previous item's last gather
    //      is behind the __syncthreads that closed it, so the band may be overwritten)
    const TIN* src = images + (long long)img * 3 * 224 * 224;
    using BV = typename BandLoad<TIN>::Vec;
    // all of a thread's loads are issued before the first conversion / store: one memory round trip per item, not ten
    constexpr int kBandVec = 3 * kBandRows * 56, kBandIter = (kBandVec + kThreads - 1) / kThreads;
    BV bv[kBandIter];
    bool bz[kBandIter];                         // rows outside the image are zero AFTER normalisation (conv padding)
#pragma unroll
    for (int u = 0; u < kBandIter; ++u) {
      const int i = tid + u * kThreads;
      const int c4 = i % 56, rr = (i / 56) % kBandRows, c = i / (56 * kBandRows);
      const int gr = 4 * py0 - 5 + rr;
      bv[u] = BandLoad<TIN>::zero();
      bz[u] = !(i < kBandVec && gr >= 0 && gr < 224);
      if (!bz[u]) bv[u] = __ldg(reinterpret_cast<const BV*>(src + ((long long)c * 224 + gr) * 224) + c4);
    }
#pragma unroll
    for (int u = 0; u < kBandIter; ++u) {
      const int i = tid + u * kThreads;
      if (i < kBandVec) {
        const int c4 = i % 56, rr = (i / 56) % kBandRows, c = i / (56 * kBandRows);
        T* d = band + rr * kPitch + (5 + 4 * c4) * 3 + c;
        float px[4] = {0.f, 0.f, 0.f, 0.f};
        if (!bz[u]) BandLoad<TIN>::unpack(bv[u], c, nm, px);
        d[0] = dh_from_f<T>(px[0]); d[3] = dh_from_f<T>(px[1]); d[6] = dh_from_f<T>(px[2]); d[9] = dh_from_f<T>(px[3]);
      }
    }
    __syncthreads();

    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = -INFINITY;
    // max-pool the accumulator of tap (dy, dx), issued as this CTA's tap number n, into the thread's 32 channels
    // Only the taps dx = 1, 2 are computed: tap (dy, 0) of pooled pixel px is conv column 2 px - 1 = tap (dy, 2) of pooled
    // pixel px - 1, i.e. the accumulator of the NEIGHBOURING lane (one shuffle per channel; the first lane of a warp takes it
    // from the last lane of the warp below through shared memory).  Six taps instead of nine, the same values.
    auto drain = [&](uint32_t n, int dy, int dx) {
      mbar_wait(mma_done(n & 1u), (n >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (n & 1u) * 64u + (uint32_t)ch0, v);
      const bool valid = !(dy == 0 && py0 + e_py == 0);         // pool padding (-inf) is skipped
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], __uint_as_float(v[j]));
      }
      if (dx == 2) {
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < 32; ++j) xch[warp * 32 + j] = __uint_as_float(v[j]);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const bool shift_ok = valid && e_ok && e_px >= 1;       // px = 0: tap dx = 0 lies in the pool's padding
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float u = __shfl_up_sync(0xffffffffu, __uint_as_float(v[j]), 1);
          if (lane == 0 && q > 0) u = xch[(warp - 1) * 32 + j];
          if (shift_ok) acc[j] = fmaxf(acc[j], u);
        }
      }
    };

#pragma unroll 1
    for (int tap = 0; tap < 6; ++tap, ++n_tap) {
      const int dy = tap >> 1, dx = 1 + (tap & 1);
      // gather this thread's part of its pixel's patch (LSU work that overlaps the previous tap's MMAs)
      const uint32_t* rowp = reinterpret_cast<const uint32_t*>(band + (4 * m_py + 2 * dy) * kPitch + 12 * m_px + 6 * dx);
      // 8-byte loads: adjacent lanes are 24 B apart, so the 16 lanes of a half-warp cover 32 distinct banks with LDS.64
      // (4-byte loads are 2-way conflicted); the patch row starts on an even word for dx = 0, 2 and on an odd one for dx = 1
      uint32_t w[44];
      auto gather_row = [&](const uint32_t* rp, uint32_t* d, bool odd) {
        if (!odd) {
#pragma unroll
          for (int q2 = 0; q2 < 5; ++q2) {
            const uint2 v = *reinterpret_cast<const uint2*>(rp + 2 * q2);
            d[2 * q2] = v.x; d[2 * q2 + 1] = v.y;
          }
          d[10] = rp[10];
        } else {
          d[0] = rp[0];
#pragma unroll
          for (int q2 = 0; q2 < 5; ++q2) {
            const uint2 v = *reinterpret_cast<const uint2*>(rp + 1 + 2 * q2);
            d[1 + 2 * q2] = v.x; d[2 + 2 * q2] = v.y;
          }
        }
      };
      if (half == 0) {
        if (dx & 1) {
#pragma unroll
          for (int r = 0; r < 4; ++r) gather_row(rowp + r * (kPitch / 2), w + r * 11, true);
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r) gather_row(rowp + r * (kPitch / 2), w + r * 11, false);
        }
      } else {
        if (dx & 1) {
#pragma unroll
          for (int r = 0; r < 3; ++r) gather_row(rowp + (4 + r) * (kPitch / 2), w + r * 11, true);
        } else {
#pragma unroll
          for (int r = 0; r < 3; ++r) gather_row(rowp + (4 + r) * (kPitch / 2), w + r * 11, false);
        }
#pragma unroll
        for (int j = 33; j < 44; ++j) w[j] = 0u;
      }
      // the previous tap's MMAs have read the A operand (and its accumulator is complete)
      if (n_tap > 0) { mbar_wait(mma_done((n_tap - 1u) & 1u), ((n_tap - 1u) >> 1) & 1u); tc_fence_after(); }
      if (half == 0) {                          // warp-uniform: warps 0..3
        tc_st32(a_lane, w);
        tc_st8(a_lane + 32u, w + 32);
        tc_st4(a_lane + 40u, w + 40);
      } else {
        tc_st32(a_lane + 44u, w);
        tc_st4(a_lane + 76u, w + 32);
      }
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (n_tap & 1u) * 64u;
        // K slots 0..159 (154 real + zero padding) = 10 K steps of 16; B chunk kc holds K slots 64 kc .. 64 kc + 63
#pragma unroll
        for (int ks = 0; ks < 10; ++ks) {
          const uint64_t db = umma_desc(w_buf + (uint32_t)((ks >> 2) * 64 * 128)) + (uint64_t)(2 * (ks & 3));
          tc_mma_ts(tmem_d, tmem_a + (uint32_t)(ks * 8), db, idesc, ks ? 1u : 0u);
        }
        tc_commit(mma_done(n_tap & 1u));
      }
      // pool the previous tap while this one runs on the tensor core
      if (tap > 0) { drain(n_tap - 1u, (tap - 1) >> 1, 1 + ((tap - 1) & 1)); tc_fence_before(); }
    }
    drain(n_tap - 1u, 2, 2);
    tc_fence_before();
    // ---- bias + ReLU + store (32 channels = 64 B per thread)
    if (e_ok) {
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        o[j] = Pack2<T>::f(fmaxf(acc[2 * j] + bias_s[ch0 + 2 * j], 0.f), fmaxf(acc[2 * j + 1] + bias_s[ch0 + 2 * j + 1], 0.f));
      T* dst = out + (((long long)img * 56 + py0 + e_py) * 56 + e_px) * 64 + ch0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
    __syncthreads();     // every warp is past its TMEM reads and band reads before the next item overwrites them
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

int g_sms = 0;

}  // namespace

template <typename TIN>
static int stem_launch(const TIN* images, const void* w_packed, const float* bias, void* out, int n, int dtype,
                       const PixelNorm& nm, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    int dev = 0;
    DH_CUDA(cudaGetDevice(&dev));
    DH_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    DH_CUDA(cudaFuncSetAttribute(stem_pool_kernel<__half, TIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    DH_CUDA(cudaFuncSetAttribute(stem_pool_kernel<__nv_bfloat16, TIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr = true;
  }
  const int items = n * 28;
  const int grid = items < 2 * g_sms ? items : 2 * g_sms;       // two CTAs per SM (256 TMEM columns and ~47 KB each)
  if (dtype == DH_F16)
    stem_pool_kernel<__half, TIN><<<grid, kThreads, kSmemBytes, stream>>>(images, (const __half*)w_packed, bias, (__half*)out, n, nm);
  else
    stem_pool_kernel<__nv_bfloat16, TIN><<<grid, kThreads, kSmemBytes, stream>>>(images, (const __nv_bfloat16*)w_packed, bias,
                                                                                (__nv_bfloat16*)out, n, nm);
  DH_LAUNCH_OK();
  return DH_OK;
}

// images [n,3,224,224] fp32 NCHW; w_packed [64][192] (k = r*22 + s*3 + c, BN folded, zeros elsewhere) and out
// [n,56,56,64] NHWC of dtype (DH_F16 / DH_BF16); bias fp32 [64].
extern "C" int dh_stem_pool_tc(const float* images_nchw, const void* w_packed, const float* bias, void* out, int n, int H,
                               int W, int dtype, cudaStream_t stream) {
  DH_ARG(images_nchw && w_packed && bias && out && n >= 0);
  DH_ARG(H == 224 && W == 224);
  DH_ARG(dtype == DH_F16 || dtype == DH_BF16);
  DH_ARG(((uintptr_t)images_nchw % 16) == 0 && ((uintptr_t)w_packed % 16) == 0 && ((uintptr_t)out % 16) == 0);
  if (n == 0) return DH_OK;
  return stem_launch<float>(images_nchw, w_packed, bias, out, n, dtype, PixelNorm{}, stream);
}

// Same from raw uint8 pixels [n,3,224,224] (NCHW, 0..255): (x / 255 - mean[c]) / std[c] is applied while the input band
// is staged (torchvision ToTensor + Normalize, deephumor_demo.ipynb cell 11).
extern "C" int dh_stem_pool_tc_u8(const unsigned char* images_nchw_u8, const float* mean3_host, const float* std3_host,
                                  const void* w_packed, const float* bias, void* out, int n, int H, int W, int dtype,
                                  cudaStream_t stream) {
  DH_ARG(images_nchw_u8 && mean3_host && std3_host && w_packed && bias && out && n >= 0);
  DH_ARG(H == 224 && W == 224);
  DH_ARG(dtype == DH_F16 || dtype == DH_BF16);
  DH_ARG(((uintptr_t)images_nchw_u8 % 4) == 0 && ((uintptr_t)w_packed % 16) == 0 && ((uintptr_t)out % 16) == 0);
  if (n == 0) return DH_OK;
  PixelNorm nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3_host[c]; nm.std[c] = std3_host[c]; DH_ARG(std3_host[c] != 0.f); }
  return stem_launch<unsigned char>(images_nchw_u8, w_packed, bias, out, n, dtype, nm, stream);
}
