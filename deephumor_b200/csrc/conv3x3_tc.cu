// 3x3 / stride 1 / pad 1 convolution (+ folded BN bias + ReLU) for sm_100a that reads its input ONCE per tile
// (torchvision resnet.py:146-148, the conv2 of every Bottleneck; layer1 / layer2 shapes).
//
// The generic implicit-GEMM path (gemm_tc.cu) fetches the A operand with im2col-mode TMA: nine taps = nine loads of the
// same 128 pixels, and that L2->SM operand traffic (~9-12 TB/s chip-wide ceiling) is what bounds the layer1/2 3x3 convs.
// Here a tile is an 8 x 16 rectangle of output pixels of one image.  Per 64-channel chunk ONE tiled TMA load brings the
// 10 x 18 pixel halo (zero-filled outside the image = the conv padding) into shared memory, 128 B per pixel, 128B-swizzled.
// Tap (r, s) is then just a shifted view of that halo: the UMMA descriptor starts at pixel (r, s) and steps 10 pixels
// (1280 B) per group of 8 rows, so nine tcgen05.mma groups reuse the same bytes.  The swizzle is a function of the shared
// memory address, so a view that starts at any 128 B multiple stays consistent with what TMA wrote.
//
//   warp 0 : TMA producer (halo per chunk, W tile per (chunk, tap); the whole W stays resident when it fits the ring)
//   warp 1 : MMA issuer (128 x BN x 16, fp32 accumulators double-buffered in TMEM)
//   warp 2 : TMEM allocator
//   warps 4-7 : epilogue (bias, ReLU, 16-bit pack, swizzled slab, 4-D TMA store of the 8 x 16 x 64-channel box)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kHaloW = 10, kHaloH = 18;
constexpr int kHaloTx = kHaloW * kHaloH * 128;          // bytes one halo load delivers
constexpr int kHaloBytes = 23552;                       // rounded to 1024 (swizzle atom alignment of the second buffer)

struct HaloParams {
  int n, H, W, Cin, Cout;
  int tiles_x, tiles_y, n_blocks, c_chunks;
  const float* bias;
  int relu, dtype;
  int w_resident;
  int* error;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {     // a protocol bug must fail the launch, never hang the GPU
      if (error) atomicExch(error, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
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
// Packed fp32 pairs (FADD2 on sm_100a): two IEEE fp32 additions per issued instruction
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void un2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// K-major, 128B-swizzled operand: start >> 4 | stride between 8-row groups (SBO) | descriptor version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

template <int BN>
struct HCfg {
  static constexpr int kHaloBufs = BN <= 64 ? 4 : 2;     // halo tiles in flight (the loads are latency-bound)
  static constexpr int kWStageBytes = BN * 128;
  static constexpr int kSlabBytes = 2 * 128 * 128;
  static constexpr int kFixed = 1024 + kHaloBufs * kHaloBytes + kSlabBytes + 1024 + 512;
  static constexpr int kWStagesMax = (232448 - kFixed) / kWStageBytes;
  static constexpr int kWStages = kWStagesMax > 12 ? 12 : kWStagesMax;
  static constexpr int kSmemBytes = kFixed + kWStages * kWStageBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_y, const HaloParams p) {
  using C = HCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  constexpr int HB = C::kHaloBufs;
  const uint32_t halo = base;                                              // HB x kHaloBytes
  const uint32_t wring = base + HB * kHaloBytes;                            // kWStages x kWStageBytes
  const uint32_t slabs = wring + C::kWStages * C::kWStageBytes;            // 2 x 16 KB
  float* bias_s = reinterpret_cast<float*>(gen + HB * kHaloBytes + C::kWStages * C::kWStageBytes + C::kSlabBytes);
  const uint32_t bars = slabs + C::kSlabBytes + 1024;
  auto hfull = [&](int s) { return bars + 8u * s; };
  auto hempty = [&](int s) { return bars + 8u * (HB + s); };
  auto tfull = [&](int s) { return bars + 8u * (2 * HB + s); };
  auto tempty = [&](int s) { return bars + 8u * (2 * HB + 2 + s); };
  auto wfull = [&](int s) { return bars + 8u * (2 * HB + 4 + s); };
  auto wempty = [&](int s) { return bars + 8u * (2 * HB + 4 + C::kWStages + s); };
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(gen + HB * kHaloBytes + C::kWStages * C::kWStageBytes + C::kSlabBytes + 1024 +
                                                    8 * (2 * HB + 4 + 2 * C::kWStages));
  static_assert(8 * (2 * HB + 4 + 2 * C::kWStages) + 4 <= 512, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HB; ++s) {
      mbar_init(hfull(s), 1);
      mbar_init(hempty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 4);
    }
    for (int s = 0; s < C::kWStages; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_word);

  const int sp_tiles = p.n * p.tiles_y * p.tiles_x;
  const int tiles = sp_tiles * p.n_blocks;
  const int wsteps = 9 * p.c_chunks;                 // W loads per tile
  auto decode = [&](int tile, int& img, int& y0, int& x0, int& n0) {
    n0 = (tile % p.n_blocks) * BN;
    int sp = tile / p.n_blocks;
    x0 = (sp % p.tiles_x) * 8;
    sp /= p.tiles_x;
    y0 = (sp % p.tiles_y) * 16;
    img = sp / p.tiles_y;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================================================================== TMA producer
      int hs = 0, ws = 0;
      uint32_t hphase = 0, wphase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int img, y0, x0, n0;
        decode(tile, img, y0, x0, n0);
        for (int c = 0; c < p.c_chunks; ++c) {
          mbar_wait(hempty(hs), hphase ^ 1u, p.error, 11);
          mbar_expect_tx(hfull(hs), kHaloTx);
          tma_load_4d(halo + hs * kHaloBytes, &map_x, hfull(hs), c * 64, x0 - 1, y0 - 1, img);
          if (++hs == HB) { hs = 0; hphase ^= 1u; }
          if (p.w_resident) {
            if (first)
              for (int t = 0; t < 9; ++t) {
                const int slot = c * 9 + t;
                mbar_expect_tx(wfull(slot), C::kWStageBytes);
                tma_load_2d(wring + slot * C::kWStageBytes, &map_w, wfull(slot), t * p.Cin + c * 64, n0);
              }
          } else {
            for (int t = 0; t < 9; ++t) {
              mbar_wait(wempty(ws), wphase ^ 1u, p.error, 12);
              mbar_expect_tx(wfull(ws), C::kWStageBytes);
              tma_load_2d(wring + ws * C::kWStageBytes, &map_w, wfull(ws), t * p.Cin + c * 64, n0);
              if (++ws == C::kWStages) { ws = 0; wphase ^= 1u; }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================================================================== MMA issuer
      const uint32_t fmt = p.dtype == DH_BF16 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int hs = 0, ws = 0, it = 0;
      uint32_t hphase = 0, wphase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(tempty(as), ((it >> 1) & 1) ^ 1u, p.error, 13);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int c = 0; c < p.c_chunks; ++c) {
          mbar_wait(hfull(hs), hphase, p.error, 14);
          tc_fence_after();
          const uint32_t hb = halo + hs * kHaloBytes;
          for (int t = 0; t < 9; ++t) {
            uint32_t wb;
            if (p.w_resident) {
              const int slot = c * 9 + t;
              if (first) { mbar_wait(wfull(slot), 0u, p.error, 15); tc_fence_after(); }
              wb = wring + slot * C::kWStageBytes;
            } else {
              mbar_wait(wfull(ws), wphase, p.error, 15);
              tc_fence_after();
              wb = wring + ws * C::kWStageBytes;
            }
            // tap (r, s): the same halo bytes, viewed from pixel (r, s); 8-row groups are one halo row (10 pixels) apart
            const int r = t / 3, s = t - r * 3;
            const uint64_t da = umma_desc(hb + (uint32_t)((r * kHaloW + s) * 128), kHaloW * 128);
            const uint64_t db = umma_desc(wb, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (c | t | k) ? 1u : 0u);
            if (!p.w_resident) {
              tc_commit(wempty(ws));
              if (++ws == C::kWStages) { ws = 0; wphase ^= 1u; }
            }
          }
          tc_commit(hempty(hs));
          if (++hs == HB) { hs = 0; hphase ^= 1u; }
        }
        tc_commit(tfull(as));
        first = false;
      }
    }
  } else if (warp >= 4) {
    // ======================================================================= epilogue
    const int ew = warp - 4;
    const int row_l = ew * 32 + lane;
    const uint32_t swz = (uint32_t)(row_l & 7);
    const bool elected = (warp == 4 && lane == 0);
    uint32_t round_ctr = 0;
    int it = 0, last_n0 = -1;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      int img, y0, x0, n0;
      decode(tile, img, y0, x0, n0);
      const int as = it & 1;
      // bias slice, reloaded only when the N block changes and before the wait on the accumulator so that the global-load
      // latency is off the per-tile critical path (readers of the previous slice are past their last round barrier)
      if (n0 != last_n0) {
        for (int i = row_l; i < BN; i += 128) bias_s[i] = (p.bias && n0 + i < p.Cout) ? __ldg(p.bias + n0 + i) : 0.f;
        last_n0 = n0;
      }
      mbar_wait(tfull(as), (it >> 1) & 1, p.error, 16);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int rd = 0; rd < BN / 64; ++rd) {
        const int col0 = n0 + rd * 64;
        if (col0 >= p.Cout) break;
        const uint32_t slab = slabs + (round_ctr & 1u) * (128 * 128);
        if (elected && round_ctr >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t srow = slab + (uint32_t)row_l * 128u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tc_ld32(tmem_row + (uint32_t)(rd * 64 + h * 32), v);
          float x[32];
          const float4* bs = reinterpret_cast<const float4*>(bias_s + rd * 64 + h * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b4 = bs[g];
            x[4 * g] = __uint_as_float(v[4 * g]) + b4.x;
            x[4 * g + 1] = __uint_as_float(v[4 * g + 1]) + b4.y;
            x[4 * g + 2] = __uint_as_float(v[4 * g + 2]) + b4.z;
            x[4 * g + 3] = __uint_as_float(v[4 * g + 3]) + b4.w;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          uint32_t w[16];
          if (p.dtype == DH_BF16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __nv_bfloat162 t2 = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
              w[j] = *reinterpret_cast<uint32_t*>(&t2);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __half2 t2 = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
              w[j] = *reinterpret_cast<uint32_t*>(&t2);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((uint32_t)(h * 4 + j) ^ swz) << 4)),
                         "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                         : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (elected) {
          // rows of the slab are the tile's pixels in (y, x) order: one 4-D box {64 channels, 8, 16, 1}; pixels outside the
          // image are clipped by TMA
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&map_y)),
                       "r"(slab), "r"(col0), "r"(x0), "r"(y0), "r"(img)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++round_ctr;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
    }
    if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}


// ------------------------------------------------------------------------------------------- Cout == 64: transposed
// With N = 64 output channels the M = 128 x N = 64 instruction is bound by re-reading its 128 x 16 A slice from shared
// memory (~128 cycles per K = 16 step whatever N is).  Swapping the operands halves that: D^T[co, pixel] = W[co, k] X[pixel, k]^T
// with the 64 channels as the UMMA M (A = the resident weight tile, 64 rows) and 256 pixels -- an 8 x 32 rectangle, the
// same shifted halo views -- as N.  A tcgen05.mma of M = 64 puts accumulator row r in TMEM lane 32 * (r / 16) + r % 16
// (probed: scripts/probes/umma_m64_layout.cu), so each of the four epilogue warps owns 16 channels; it transposes its
// 16 x 256 slice through the swizzled slab back to NHWC for the TMA store.
constexpr int kTHaloH = 34;                                   // 32 output rows + 2
constexpr int kTHaloTx = kHaloW * kTHaloH * 128;              // 43 520 B
constexpr int kTHaloBytes = 44032;                            // rounded to 1024
constexpr int kTWStages = 9;                                  // one 8 KB slot per tap: resident when Cin == 64
constexpr int kTSmemBytes = 1024 + 2 * kTHaloBytes + kTWStages * 8192 + 256 * 128 + 512;

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_halo_t_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_y, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t halo = base;                                        // 2 x kTHaloBytes
  const uint32_t wring = base + 2 * kTHaloBytes;                     // 9 x 8 KB
  const uint32_t slab = wring + kTWStages * 8192;                    // 256 pixels x 128 B
  const uint32_t bars = slab + 256 * 128;
  auto hfull = [&](int s) { return bars + 8u * s; };
  auto hempty = [&](int s) { return bars + 8u * (2 + s); };
  auto tfull = [&](int s) { return bars + 8u * (4 + s); };
  auto tempty = [&](int s) { return bars + 8u * (6 + s); };
  auto wfull = [&](int s) { return bars + 8u * (8 + s); };
  auto wempty = [&](int s) { return bars + 8u * (8 + kTWStages + s); };
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(gen + 2 * kTHaloBytes + kTWStages * 8192 + 256 * 128 + 8 * (8 + 2 * kTWStages));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(hfull(s), 1);
      mbar_init(hempty(s), 1);
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 4);
    }
    for (int s = 0; s < kTWStages; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_word);

  const int tiles = p.n * p.tiles_y * p.tiles_x;            // tiles_y counts 32-row tiles here; one N block (Cout == 64)
  auto decode = [&](int tile, int& img, int& y0, int& x0) {
    x0 = (tile % p.tiles_x) * 8;
    const int sp = tile / p.tiles_x;
    y0 = (sp % p.tiles_y) * 32;
    img = sp / p.tiles_y;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================================================================== TMA producer
      int hs = 0, ws = 0;
      uint32_t hphase = 0, wphase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int img, y0, x0;
        decode(tile, img, y0, x0);
        for (int c = 0; c < p.c_chunks; ++c) {
          mbar_wait(hempty(hs), hphase ^ 1u, p.error, 21);
          mbar_expect_tx(hfull(hs), kTHaloTx);
          tma_load_4d(halo + hs * kTHaloBytes, &map_x, hfull(hs), c * 64, x0 - 1, y0 - 1, img);
          if (++hs == 2) { hs = 0; hphase ^= 1u; }
          for (int t = 0; t < 9; ++t) {
            if (p.w_resident) {
              if (first) {
                mbar_expect_tx(wfull(t), 8192);
                tma_load_2d(wring + t * 8192, &map_w, wfull(t), t * p.Cin, 0);
              }
            } else {
              mbar_wait(wempty(ws), wphase ^ 1u, p.error, 22);
              mbar_expect_tx(wfull(ws), 8192);
              tma_load_2d(wring + ws * 8192, &map_w, wfull(ws), t * p.Cin + c * 64, 0);
              if (++ws == kTWStages) { ws = 0; wphase ^= 1u; }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================================================================== MMA issuer: M = 64 channels, N = 256 pixels
      const uint32_t fmt = p.dtype == DH_BF16 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      int hs = 0, ws = 0, it = 0;
      uint32_t hphase = 0, wphase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(tempty(as), ((it >> 1) & 1) ^ 1u, p.error, 23);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
        for (int c = 0; c < p.c_chunks; ++c) {
          mbar_wait(hfull(hs), hphase, p.error, 24);
          tc_fence_after();
          const uint32_t hb = halo + hs * kTHaloBytes;
          for (int t = 0; t < 9; ++t) {
            uint32_t wb;
            if (p.w_resident) {
              if (first) { mbar_wait(wfull(t), 0u, p.error, 25); tc_fence_after(); }
              wb = wring + t * 8192;
            } else {
              mbar_wait(wfull(ws), wphase, p.error, 25);
              tc_fence_after();
              wb = wring + ws * 8192;
            }
            const int r = t / 3, s = t - r * 3;
            const uint64_t da = umma_desc(wb, 1024);                                              // 64 channels x 64 K
            const uint64_t db = umma_desc(hb + (uint32_t)((r * kHaloW + s) * 128), kHaloW * 128);   // 256 pixels, shifted view
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (c | t | k) ? 1u : 0u);
            if (!p.w_resident) {
              tc_commit(wempty(ws));
              if (++ws == kTWStages) { ws = 0; wphase ^= 1u; }
            }
          }
          tc_commit(hempty(hs));
          if (++hs == 2) { hs = 0; hphase ^= 1u; }
        }
        tc_commit(tfull(as));
        first = false;
      }
    }
  } else if (warp >= 4) {
    // ======================================================================= epilogue: warp ew owns channels 16 ew .. 16 ew + 15
    const int ew = warp - 4;
    const int ch = ew * 16 + (lane & 15);
    const bool active = lane < 16;
    const float bias_v = (p.bias && active) ? __ldg(p.bias + ch) : 0.f;
    const bool elected = (warp == 4 && lane == 0);
    const uint32_t ch_chunk = (uint32_t)(ch >> 3), ch_off = (uint32_t)((ch & 7) * 2);
    bool pending = false;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      int img, y0, x0;
      decode(tile, img, y0, x0);
      const int as = it & 1;
      mbar_wait(tfull(as), (it >> 1) & 1, p.error, 26);
      tc_fence_after();
      if (elected && pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // slab free again
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * 256);
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {                         // 32 pixels per TMEM load
        uint32_t v[32];
        tc_ld32(tmem_row + (uint32_t)(c * 32), v);
        if (active) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]) + bias_v;
            if (p.relu) x = fmaxf(x, 0.f);
            const uint32_t pix = (uint32_t)(c * 32 + j);
            const uint32_t addr = slab + pix * 128u + ((ch_chunk ^ (pix & 7u)) << 4) + ch_off;
            if (p.dtype == DH_BF16) {
              const __nv_bfloat16 hv = __float2bfloat16_rn(x);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
            } else {
              const __half hv = __float2half_rn(x);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));               // the accumulator is drained
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (elected) {
        // slab rows are the tile's pixels in (y, x) order: two boxes {64 channels, 8, 16, 1}
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&map_y)),
                     "r"(slab), "r"(0), "r"(x0), "r"(y0), "r"(img)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&map_y)),
                     "r"(slab + 128u * 128u), "r"(0), "r"(x0), "r"(y0 + 16), "r"(img)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        pending = true;
      }
    }
    if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ------------------------------------------------------------------------------------------- conv2 -> conv3 in one kernel
// The tail of a layer1 identity bottleneck (torchvision resnet.py:150-161): relu(bn2(conv2 3x3 (64 -> 64))) feeding
// relu(bn3(conv3 1x1 (64 -> 256)) + identity).  Both run at the HBM roofline one layer at a time, and conv3 is POINTWISE on
// conv2's output, so the 8 x 32 pixel tile that conv2 leaves in the swizzled slab (256 pixel rows x 64 channels, K-major) is
// exactly the A operand of conv3: it never goes to HBM (-410 MB per block and 512 images).  Per tile:
//   MMA warp    conv2 as in conv3x3_halo_t_kernel (weights = M, 256 pixels = N, nine shifted halo views) into TMEM columns
//               [0, 256); conv3 for each 128-pixel half as 4 tcgen05.mma of 128 x 256 x 16 (A = the slab half, B = the resident
//               64 x 256 weight tile) into columns [256, 512).  The issue order is software-pipelined -- conv3(half 0),
//               conv2 of the NEXT tile, conv3(half 1) -- so the tensor core works while the epilogue drains.
//   epilogue    (1) conv2 accumulator -> bias, ReLU, 16-bit, transposed into the slab; (2) per half: conv3 accumulator + bias
//               + identity (read straight from the block input, 128 contiguous bytes per pixel and round) -> ReLU -> slab ->
//               4-D TMA store of 64-channel boxes.
constexpr int kFW3Bytes = 256 * 128;                              // conv3 weights [256 cout][64] K-major
constexpr int kFHaloBytes = 2 * kHaloBytes;                        // two 10 x 18 halos: rows 0-15 and 16-31 of the 8 x 32 tile
constexpr int kFSmemBytes = 1024 + kFHaloBytes + 9 * 8192 + kFW3Bytes + 256 * 128 + 2 * 128 * 128 + 1024 + 512;

struct FusedTailParams {
  int n, H, W, tiles_x, tiles_y;
  const float* bias2; const float* bias3;
  const void* residual;                                         // block input [n, H, W, 256]
  int dtype;
  int* error;
};

constexpr int kFThreads = 384;                                   // producer / MMA / TMEM warps + eight epilogue warps
__global__ void __launch_bounds__(kFThreads, 1)
conv3x3_c3_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w2,
                        const __grid_constant__ CUtensorMap map_w3, const __grid_constant__ CUtensorMap map_out,
                        const FusedTailParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t halo = base;                                        // 2 x (10 x 18 pixels x 128 B): one per 16-row half
  const uint32_t w2s = base + kFHaloBytes;                           // 9 taps x 8 KB
  const uint32_t w3s = w2s + 9 * 8192;                               // 256 x 128 B
  const uint32_t y2s = w3s + kFW3Bytes;                              // conv2 output: 256 pixels x 128 B = conv3's A operand
  const uint32_t outs = y2s + 256 * 128;                             // two 128-pixel x 128 B store slabs
  float* bias3_s = reinterpret_cast<float*>(gen + kFHaloBytes + 9 * 8192 + kFW3Bytes + 256 * 128 + 2 * 128 * 128);
  const uint32_t bars = outs + 2 * 128 * 128 + 1024;
  const uint32_t wfull = bars + 16, t2full = bars + 24, t2empty = bars + 32, y2full = bars + 40, y2free = bars + 48,
                 t3full = bars + 56, t3empty = bars + 64;
  auto hfull = [&](int s) { return bars + 72u + 8u * s; };
  auto hempty = [&](int s) { return bars + 88u + 8u * s; };
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(gen + kFHaloBytes + 9 * 8192 + kFW3Bytes + 256 * 128 + 2 * 128 * 128 + 1024 + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(hfull(s), 1); mbar_init(hempty(s), 1); }
    mbar_init(wfull, 1);
    mbar_init(t2full, 1); mbar_init(t2empty, 8);
    mbar_init(y2full, 1); mbar_init(y2free, 1);
    mbar_init(t3full, 1); mbar_init(t3empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += kFThreads) bias3_s[i] = p.bias3 ? __ldg(p.bias3 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_word);

  const int tiles = p.n * p.tiles_y * p.tiles_x;
  const int n_my = tiles > (int)blockIdx.x ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto decode = [&](int it, int& img, int& y0, int& x0) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    x0 = (tile % p.tiles_x) * 8;
    const int sp = tile / p.tiles_x;
    y0 = (sp % p.tiles_y) * 32;
    img = sp / p.tiles_y;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================================================================== TMA producer: weights once, one halo per tile
      mbar_expect_tx(wfull, 9 * 8192 + kFW3Bytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(w2s + t * 8192, &map_w2, wfull, t * 64, 0);
      tma_load_2d(w3s, &map_w3, wfull, 0, 0);
      for (int it = 0; it < n_my; ++it) {
        int img, y0, x0;
        decode(it, img, y0, x0);
        // the tile's two 16-row halves have a halo buffer each: the next tile's upper half loads while this tile's lower
        // half is still being multiplied
        for (int sub = 0; sub < 2; ++sub) {
          mbar_wait(hempty(sub), (uint32_t)((it & 1) ^ 1), p.error, 31);
          mbar_expect_tx(hfull(sub), kHaloTx);
          tma_load_4d(halo + sub * kHaloBytes, &map_x, hfull(sub), 0, x0 - 1, y0 + 16 * sub - 1, img);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_my > 0) {
      // ===================================================================== MMA issuer
      const uint32_t fmt = p.dtype == DH_BF16 ? 1u : 0u;
      const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      const uint32_t idesc3 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(wfull, 0u, p.error, 32);
      tc_fence_after();
      auto conv2 = [&](int it) {
        mbar_wait(t2empty, (uint32_t)((it & 1) ^ 1), p.error, 34);
        for (int sub = 0; sub < 2; ++sub) {                       // 128 pixels (16 rows) per halo buffer
          mbar_wait(hfull(sub), (uint32_t)(it & 1), p.error, 33);
          tc_fence_after();
          const uint32_t hb = halo + sub * kHaloBytes;
          for (int t = 0; t < 9; ++t) {
            const int r = t / 3, s = t - r * 3;
            const uint64_t da = umma_desc(w2s + t * 8192, 1024);                                  // 64 channels x 64 K
            const uint64_t db = umma_desc(hb + (uint32_t)((r * kHaloW + s) * 128), kHaloW * 128);   // 128 pixels, shifted view
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma(tmem_base + (uint32_t)(sub * 128), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (t | k) ? 1u : 0u);
          }
          tc_commit(hempty(sub));
        }
        tc_commit(t2full);
      };
      auto conv3 = [&](int it, int h) {
        const int q = 2 * it + h;
        if (h == 0) mbar_wait(y2full, (uint32_t)(it & 1), p.error, 35);
        mbar_wait(t3empty, (uint32_t)((q & 1) ^ 1), p.error, 36);
        tc_fence_after();
        const uint64_t da = umma_desc(y2s + (uint32_t)h * (128 * 128), 1024);                     // 128 pixels x 64 K
        const uint64_t db = umma_desc(w3s, 1024);                                                 // 256 channels x 64 K
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma(tmem_base + 256u, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc3, k ? 1u : 0u);
        tc_commit(t3full);
        if (h == 1) tc_commit(y2free);                            // the slab may be rewritten once these MMAs have read it
      };
      conv2(0);
      for (int it = 0; it < n_my; ++it) {
        conv3(it, 0);
        if (it + 1 < n_my) conv2(it + 1);
        conv3(it, 1);
      }
    }
  } else if (warp >= 4) {
    // ======================================================================= epilogue: EIGHT warps.  A warp may only read the
    // TMEM lane quadrant warp % 4, so warps 4-7 and 8-11 cover the same rows and split the COLUMNS: with one thread per
    // accumulator row and four warps the epilogue (512 outputs per thread and tile: unpack identity, two adds, max, pack),
    // not HBM, paced the kernel (ncu: 30 k cycles per tile, 22 % tensor-active).
    const int ew = warp & 3, eg = (warp - 4) >> 2;
    const int ch = ew * 16 + (lane & 15);                       // conv2 accumulator row (M = 64 layout) of this lane
    const bool active = lane < 16;
    const float bias2_v = (p.bias2 && active) ? __ldg(p.bias2 + ch) : 0.f;
    const bool elected = (warp == 4 && lane == 0);
    const uint32_t ch_chunk = (uint32_t)(ch >> 3), ch_off = (uint32_t)((ch & 7) * 2);
    const int row_l = ew * 32 + lane;                           // conv3 accumulator row = pixel of the half
    const uint32_t swz = (uint32_t)(row_l & 7);
    uint32_t round_ctr = 0;
    for (int it = 0; it < n_my; ++it) {
      int img, y0, x0;
      decode(it, img, y0, x0);
      // identity rows of this thread's two pixels (one per half): this group's 32 channels of every 64-channel round are 64
      // contiguous bytes.  They are needed a few thousand cycles from now: pull the lines into L2 first, and fetch each round's
      // chunk one round ahead into registers
      const uint4* res_h[2];
      bool inb_h[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pix = h * 128 + row_l, gy = y0 + (pix >> 3), gx = x0 + (pix & 7);
        inb_h[h] = gy < p.H && gx < p.W;
        res_h[h] = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) +
                                                  (((long long)img * p.H + gy) * p.W + gx) * 256) + eg * 4;
        if (inb_h[h]) {
#pragma unroll
          for (int l = 0; l < 2; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(res_h[h] + (2 * l + eg) * 8 - eg * 4));
        }
      }
      // ---- (1) conv2 accumulator -> slab (conv3's A operand); group eg takes pixels [128 eg, 128 eg + 128)
      mbar_wait(t2full, (uint32_t)(it & 1), p.error, 37);
      mbar_wait(y2free, (uint32_t)((it & 1) ^ 1), p.error, 38);  // conv3 of the previous tile has read the slab
      tc_fence_after();
      {
        const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
        for (int c = eg * 4; c < eg * 4 + 4; ++c) {             // 32 pixels per TMEM load
          uint32_t v[32];
          tc_ld32(tmem_row + (uint32_t)(c * 32), v);
          if (active) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = fmaxf(__uint_as_float(v[j]) + bias2_v, 0.f);
              const uint32_t pix = (uint32_t)(c * 32 + j);
              const uint32_t addr = y2s + pix * 128u + ((ch_chunk ^ (pix & 7u)) << 4) + ch_off;
              if (p.dtype == DH_BF16) {
                const __nv_bfloat16 hv = __float2bfloat16_rn(x);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
              } else {
                const __half hv = __float2half_rn(x);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t2empty);                      // conv2's accumulator is drained
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (elected) mbar_arrive(y2full);
      // ---- (2) conv3 accumulator of each 128-pixel half + bias + identity -> ReLU -> store; group eg takes channels
      // [64 rd + 32 eg, 64 rd + 32 eg + 32) of round rd
      uint4 rnext[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) rnext[j] = inb_h[0] ? __ldg(res_h[0] + j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int q = 2 * it + h;
        mbar_wait(t3full, (uint32_t)(q & 1), p.error, 39);
        tc_fence_after();
        const uint32_t tmem_row = tmem_base + ((uint32_t)(ew * 32) << 16) + 256u + (uint32_t)(eg * 32);
#pragma unroll 1
        for (int rd = 0; rd < 4; ++rd) {
          uint4 rv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) rv[j] = rnext[j];
          {
            // next round's identity chunk (the other half's first chunk after this half's last)
            const bool more = !(h == 1 && rd == 3);
            const uint4* np = rd == 3 ? res_h[1] : res_h[h] + (rd + 1) * 8;
            const bool nin = rd == 3 ? inb_h[1] : inb_h[h];
#pragma unroll
            for (int j = 0; j < 4; ++j) rnext[j] = (more && nin) ? __ldg(np + j) : make_uint4(0u, 0u, 0u, 0u);
          }
          const uint32_t slab = outs + (round_ctr & 1u) * (128 * 128);
          if (elected && round_ctr >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          uint32_t v[32];
          tc_ld32(tmem_row + (uint32_t)(rd * 64), v);
          if (rd == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t3empty);                // conv3's accumulator now lives in registers
          }
          const float2* b2p = reinterpret_cast<const float2*>(bias3_s + rd * 64 + eg * 32);
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {                         // 8 channels per 16-byte chunk
            const uint32_t rr[4] = {rv[j].x, rv[j].y, rv[j].z, rv[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float r0, r1;
              if (p.dtype == DH_BF16) {
                const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&rr[e]);
                r0 = __low2float(t); r1 = __high2float(t);
              } else {
                const __half2 t = *reinterpret_cast<const __half2*>(&rr[e]);
                r0 = __low2float(t); r1 = __high2float(t);
              }
              const float2 bb = b2p[4 * j + e];
              // packed fp32 pair: (acc + bias) + identity, two IEEE additions per issued instruction
              float a0, a1;
              un2(add2(add2(pk2u(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]), pk2(bb.x, bb.y)), pk2(r0, r1)), a0, a1);
              a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f);
              if (p.dtype == DH_BF16) {
                __nv_bfloat162 t = __floats2bfloat162_rn(a0, a1);
                w[4 * j + e] = *reinterpret_cast<uint32_t*>(&t);
              } else {
                __half2 t = __floats2half2_rn(a0, a1);
                w[4 * j + e] = *reinterpret_cast<uint32_t*>(&t);
              }
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");          // the slab is free (the elected thread waited above)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + (uint32_t)row_l * 128u + (((uint32_t)(eg * 4 + j) ^ swz) << 4)),
                         "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (elected) {
            // slab rows are the half's pixels in (y, x) order: one box {64 channels, 8, 16, 1}; pixels outside the image
            // are clipped by TMA
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&map_out)),
                         "r"(slab), "r"(rd * 64), "r"(x0), "r"(y0 + h * 16), "r"(img)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++round_ctr;
        }
      }
    }
    if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_sms = 0;
int* g_error = nullptr;

int halo_init() {
  if (g_encode) return DH_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  DH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return dh_fail(DH_ERR_DEVICE, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  int dev = 0;
  DH_CUDA(cudaGetDevice(&dev));
  DH_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  DH_CUDA(cudaMalloc(&g_error, sizeof(int)));
  DH_CUDA(cudaMemset(g_error, 0, sizeof(int)));
  g_encode = (EncodeTiledFn)fn;
  return DH_OK;
}

int map_nhwc(CUtensorMap* m, const void* ptr, int n, int H, int W, int C, int box_w, int box_h, int dtype) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, dtype == DH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dh_fail(DH_ERR_ARG, "cuTensorMapEncodeTiled rejected the NHWC tensor", __FILE__, __LINE__);
  return DH_OK;
}

template <int BN>
int launch_halo(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, HaloParams& p, cudaStream_t s) {
  using C = HCfg<BN>;
  static bool attr = false;
  if (!attr) {
    DH_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr = true;
  }
  p.n_blocks = dh_cdiv(p.Cout, BN);
  p.w_resident = (p.n_blocks == 1 && 9 * p.c_chunks <= C::kWStages) ? 1 : 0;
  const int tiles = p.n * p.tiles_y * p.tiles_x * p.n_blocks;
  const int grid = tiles < g_sms ? tiles : g_sms;
  conv3x3_halo_kernel<BN><<<grid, kThreads, C::kSmemBytes, s>>>(mx, mw, my, p);
  DH_LAUNCH_OK();
  return DH_OK;
}

int launch_halo_t(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, HaloParams& p, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    DH_CUDA(cudaFuncSetAttribute(conv3x3_halo_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmemBytes));
    attr = true;
  }
  p.n_blocks = 1;
  p.tiles_y = dh_cdiv(p.H, 32);
  p.w_resident = p.c_chunks == 1 ? 1 : 0;
  const int tiles = p.n * p.tiles_y * p.tiles_x;
  const int grid = tiles < g_sms ? tiles : g_sms;
  conv3x3_halo_t_kernel<<<grid, kThreads, kTSmemBytes, s>>>(mx, mw, my, p);
  DH_LAUNCH_OK();
  return DH_OK;
}

}  // namespace

// x [n,H,W,Cin] NHWC, w [Cout][3][3][Cin] (BN folded), y [n,H,W,Cout]; all dtype (DH_F16 / DH_BF16); Cin % 64 == 0,
// Cout % 64 == 0 and Cout <= 128 per N tile choice (64 or 128).  y = act(conv3x3(x, w, stride 1, pad 1) + bias).
extern "C" int dh_conv3x3_halo_tc(const void* x, const void* w, const float* bias, void* y, int n, int H, int W, int Cin,
                                  int Cout, int relu, int dtype, cudaStream_t stream) {
  DH_ARG(x && w && y && n >= 0 && H > 0 && W > 0);
  DH_ARG(Cin > 0 && Cin % 64 == 0 && Cout > 0 && Cout % 64 == 0);
  DH_ARG(dtype == DH_BF16 || dtype == DH_F16);
  DH_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0);
  if (n == 0) return DH_OK;
  int rc = halo_init();
  if (rc) return rc;
  HaloParams p{};
  p.n = n; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_x = dh_cdiv(W, 8); p.tiles_y = dh_cdiv(H, 16); p.c_chunks = Cin / 64;
  p.bias = bias; p.relu = relu; p.dtype = dtype; p.error = g_error;
  CUtensorMap mx, mw, my;
  static const bool transposed_ok = !getenv("DH_NO_HALO_T");
  const bool transposed = transposed_ok && Cout == 64;          // channels as the UMMA M, 256 pixels as N
  rc = map_nhwc(&mx, x, n, H, W, Cin, kHaloW, transposed ? kTHaloH : kHaloH, dtype);
  if (rc) return rc;
  rc = map_nhwc(&my, y, n, H, W, Cout, 8, 16, dtype);
  if (rc) return rc;
  const int bn = Cout % 128 == 0 ? 128 : 64;
  {
    // W as a 2-D [Cout, 9 * Cin] K-major matrix, box {64, bn}
    cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * Cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&mw, dtype == DH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                          const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return dh_fail(DH_ERR_ARG, "cuTensorMapEncodeTiled rejected the weights", __FILE__, __LINE__);
  }
  if (transposed) return launch_halo_t(mx, mw, my, p, stream);
  return bn == 128 ? launch_halo<128>(mx, mw, my, p, stream) : launch_halo<64>(mx, mw, my, p, stream);
}

// Tail of a layer1-shaped identity bottleneck in ONE launch (torchvision resnet.py:150-161): y = relu(bn3(conv3(relu(bn2(
// conv2(y1))))) + x) with conv2 3x3 / stride 1 / pad 1 (64 -> 64) and conv3 1x1 (64 -> 256); y1 [n,H,W,64], x and y
// [n,H,W,256] NHWC, w2 [64][3][3][64], w3 [256][64] (BN folded), all `dtype`.  conv2's output never reaches HBM.
extern "C" int dh_bottleneck_tail_tc(const void* y1, const void* w2, const float* bias2, const void* w3, const float* bias3,
                                     const void* x, void* y, int n, int H, int W, int dtype, cudaStream_t stream) {
  DH_ARG(y1 && w2 && w3 && x && y && n >= 0 && H > 0 && W > 0);
  DH_ARG(dtype == DH_BF16 || dtype == DH_F16);
  DH_ARG(((uintptr_t)y1 % 16) == 0 && ((uintptr_t)w2 % 16) == 0 && ((uintptr_t)w3 % 16) == 0 && ((uintptr_t)x % 16) == 0 &&
         ((uintptr_t)y % 16) == 0);
  if (n == 0) return DH_OK;
  int rc = halo_init();
  if (rc) return rc;
  FusedTailParams p{};
  p.n = n; p.H = H; p.W = W; p.tiles_x = dh_cdiv(W, 8); p.tiles_y = dh_cdiv(H, 32);
  p.bias2 = bias2; p.bias3 = bias3; p.residual = x; p.dtype = dtype; p.error = g_error;
  CUtensorMap mx, mw2, mw3, mo;
  rc = map_nhwc(&mx, y1, n, H, W, 64, kHaloW, kHaloH, dtype);
  if (rc) return rc;
  rc = map_nhwc(&mo, y, n, H, W, 256, 8, 16, dtype);
  if (rc) return rc;
  const CUtensorMapDataType ty = dtype == DH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * 64, 64};
    cuuint64_t strides[1] = {(cuuint64_t)9 * 64 * 2};
    cuuint32_t box[2] = {64, 64};
    if (g_encode(&mw2, ty, 2, const_cast<void*>(w2), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return dh_fail(DH_ERR_ARG, "cuTensorMapEncodeTiled rejected the conv2 weights", __FILE__, __LINE__);
  }
  {
    cuuint64_t dims[2] = {64, 256};
    cuuint64_t strides[1] = {64 * 2};
    cuuint32_t box[2] = {64, 256};
    if (g_encode(&mw3, ty, 2, const_cast<void*>(w3), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return dh_fail(DH_ERR_ARG, "cuTensorMapEncodeTiled rejected the conv3 weights", __FILE__, __LINE__);
  }
  static bool attr = false;
  if (!attr) {
    DH_CUDA(cudaFuncSetAttribute(conv3x3_c3_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes));
    attr = true;
  }
  const int tiles = n * p.tiles_y * p.tiles_x;
  const int grid = tiles < g_sms ? tiles : g_sms;
  conv3x3_c3_fused_kernel<<<grid, kFThreads, kFSmemBytes, stream>>>(mx, mw2, mw3, mo, p);
  DH_LAUNCH_OK();
  return DH_OK;
}
