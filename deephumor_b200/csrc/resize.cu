// torchvision `Resize((224, 224))` on PIL images, on the device (reference: deephumor_demo.ipynb cell 11 and
// deephumor/data/datasets.py:48-53,94-98 -- `Image.open` + `image_transform`).  For a PIL input torchvision calls
// `Image.resize(size, BILINEAR)`, i.e. Pillow's ImagingResample (third-party, un-vendored; Pillow 12.2 in this image):
//
//   * separable two-pass convolution, HORIZONTAL pass first, then vertical, with a uint8 (rounded, clipped) intermediate;
//   * triangle filter whose support is scaled by the reduction factor (an antialiased reduction, not 2-tap bilinear):
//       scale = in / out, filterscale = max(scale, 1), support = filterscale, center = (xx + 0.5) * scale,
//       xmin = max(0, (int)(center - support + 0.5)), xmax = min(in, (int)(center + support + 0.5)) - xmin,
//       w[x] = tri((x + xmin - center + 0.5) / filterscale), normalised to sum 1 (all in float64);
//   * coefficients in 22-bit fixed point, (int)(0.5 + w * 2^22); accumulator starts at 2^21; result clip8(acc >> 22);
//   * the horizontal pass only covers the source rows the vertical pass reads;
//   * images taller than 100 x their width whose height is being reduced run the VERTICAL pass first (observed on Pillow
//     12.2: exactly `height > 100 * width && height > out_height` on every size probed; the uint8 intermediate makes the
//     order visible in the last bit; tests/test_resize.py pins both sides of the threshold against PIL).
//
// Everything is integer / exactly-rounded float64 arithmetic, so the result equals Pillow's bit for bit (oracle/resize.py
// restates the algorithm in numpy and tests/ pin both to PIL.Image.resize itself).  float64 expressions use the explicit
// round-to-nearest intrinsics: the compiler must not contract a*b+c into an FMA, which rounds once instead of twice.
//
// Images of a batch have different sizes: they arrive packed HWC uint8 (RGB) in one buffer with per-image offsets / sizes.
// Three launches: coefficients (one block per image and axis), first pass into a workspace, second pass into
// uint8 [n,3,out,out] NCHW -- the input format of dh_stem_pool_tc_u8, which fuses ToTensor + Normalize into the stem.
#include "common.cuh"

namespace {

constexpr int kMaxTaps = 160;        // ceil(in / out) * 2 + 1 taps: in / out up to 79 (224 <- 17 696 pixels)

struct ImageDesc {
  long long src_off;                 // byte offset of the image in the packed buffer
  long long tmp_off;                 // byte offset of its horizontal-pass rows in the workspace
  int h, w;
};

// coefficient table of one (image, axis): bounds[o] = {first source index, tap count}, taps[o][kMaxTaps]
struct AxisTable {
  int first_row, n_rows;             // vertical axis only: source rows [first_row, first_row + n_rows) are ever read
  int bounds[2 * 256];
};

__device__ __forceinline__ double tri(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for one axis; one thread per output index.
__global__ void resize_coeffs_kernel(const ImageDesc* __restrict__ desc, AxisTable* __restrict__ tables, int* __restrict__ taps,
                                     int out_size) {
  const int img = blockIdx.x >> 1, axis = blockIdx.x & 1;          // axis 0: horizontal (width), 1: vertical (height)
  const int in_size = axis == 0 ? desc[img].w : desc[img].h;
  AxisTable* t = tables + blockIdx.x;
  int* tp = taps + (long long)blockIdx.x * out_size * kMaxTaps;
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                               // bilinear: filter support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x)
      ww = __dadd_rn(ww, tri(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)));
    for (int x = 0; x < xmax; ++x) {
      double w = tri(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      tp[xx * kMaxTaps + x] = w < 0.0 ? (int)__dadd_rn(-0.5, __dmul_rn(w, 4194304.0)) : (int)__dadd_rn(0.5, __dmul_rn(w, 4194304.0));
    }
    t->bounds[2 * xx] = xmin;
    t->bounds[2 * xx + 1] = xmax;
  }
  __syncthreads();
  if (axis == 1 && threadIdx.x == 0) {
    t->first_row = t->bounds[0];
    t->n_rows = t->bounds[2 * (out_size - 1)] + t->bounds[2 * (out_size - 1) + 1] - t->bounds[0];
  }
}

__device__ __forceinline__ unsigned char clip8(int acc) {
  const int v = acc >> 22;
  return (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : v);
}

// One separable pass of one image.  ALONG_X: taps run along a source row, else down a source column.  FINAL: the pass writes
// the NCHW uint8 planes of the output, else the HWC intermediate in the workspace.
//   horizontal-first (Pillow's normal order): pass 1 ALONG_X over source rows [first_row, first_row + n_rows) -> tmp
//   [n_rows, out, 3]; pass 2 down the columns of tmp -> out.
//   vertical-first (height > 100 * width and height > out): pass 1 down the source columns -> tmp [out, w, 3];
//   pass 2 ALONG_X over tmp -> out.
struct PassView {
  const unsigned char* src; int src_w;      // HWC uint8 source of this pass and its row length in pixels
  int rows, cols;                           // output extent of this pass (rows x cols pixels)
  int row0;                                 // ALONG_X: first source row; !ALONG_X: subtracted from the tap origin
  unsigned char* dst;
};

template <bool ALONG_X, bool FINAL>
__device__ __forceinline__ void resize_pass(const PassView& v, const int* __restrict__ bounds, const int* __restrict__ taps,
                                            int out_size) {
  const long long total = (long long)v.rows * v.cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % v.cols), r = (int)(i / v.cols);
    const int o = ALONG_X ? c : r;                                  // output index along the resampled axis
    const int first = bounds[2 * o], n = bounds[2 * o + 1];
    const int* k = taps + o * kMaxTaps;
    const unsigned char* p = ALONG_X ? v.src + ((long long)(v.row0 + r) * v.src_w + first) * 3
                                     : v.src + ((long long)(first - v.row0) * v.src_w + c) * 3;
    const long long step = ALONG_X ? 3 : (long long)v.src_w * 3;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int t = 0; t < n; ++t) {
      const int kv = k[t];
      s0 += (int)p[0] * kv;
      s1 += (int)p[1] * kv;
      s2 += (int)p[2] * kv;
      p += step;
    }
    if (FINAL) {
      const long long plane = (long long)out_size * out_size;
      unsigned char* q = v.dst + i;
      q[0] = clip8(s0); q[plane] = clip8(s1); q[2 * plane] = clip8(s2);
    } else {
      unsigned char* q = v.dst + i * 3;
      q[0] = clip8(s0); q[1] = clip8(s1); q[2] = clip8(s2);
    }
  }
}

__device__ __forceinline__ bool vertical_first(const ImageDesc& d, int out_size) {
  return (long long)d.h > 100ll * d.w && d.h > out_size;
}

__global__ void __launch_bounds__(256) resize_pass1_kernel(const unsigned char* __restrict__ packed,
                                                           const ImageDesc* __restrict__ desc,
                                                           const AxisTable* __restrict__ tables, const int* __restrict__ taps,
                                                           unsigned char* __restrict__ tmp, int out_size) {
  const int img = blockIdx.y;
  const ImageDesc d = desc[img];
  const AxisTable* th = tables + 2 * img;
  const AxisTable* tv = tables + 2 * img + 1;
  const int* tph = taps + (long long)(2 * img) * out_size * kMaxTaps;
  PassView v;
  v.src = packed + d.src_off; v.src_w = d.w; v.dst = tmp + d.tmp_off;
  if (vertical_first(d, out_size)) {
    v.rows = out_size; v.cols = d.w; v.row0 = 0;
    resize_pass<false, false>(v, tv->bounds, tph + (long long)out_size * kMaxTaps, out_size);
  } else {
    v.rows = tv->n_rows; v.cols = out_size; v.row0 = tv->first_row;
    resize_pass<true, false>(v, th->bounds, tph, out_size);
  }
}

__global__ void __launch_bounds__(256) resize_pass2_kernel(const ImageDesc* __restrict__ desc,
                                                           const AxisTable* __restrict__ tables, const int* __restrict__ taps,
                                                           const unsigned char* __restrict__ tmp, unsigned char* __restrict__ out,
                                                           int out_size) {
  const int img = blockIdx.y;
  const ImageDesc d = desc[img];
  const AxisTable* th = tables + 2 * img;
  const AxisTable* tv = tables + 2 * img + 1;
  const int* tph = taps + (long long)(2 * img) * out_size * kMaxTaps;
  PassView v;
  v.src = tmp + d.tmp_off; v.rows = out_size; v.cols = out_size;
  v.dst = out + (long long)img * 3 * out_size * out_size;
  if (vertical_first(d, out_size)) {
    v.src_w = d.w; v.row0 = 0;
    resize_pass<true, true>(v, th->bounds, tph, out_size);
  } else {
    v.src_w = out_size; v.row0 = tv->first_row;
    resize_pass<false, true>(v, tv->bounds, tph + (long long)out_size * kMaxTaps, out_size);
  }
}

long long align256(long long v) { return (v + 255) / 256 * 256; }

struct Layout {
  long long desc_off, tables_off, taps_off, tmp_off, total;
};

int layout_of(int n, const int* heights, const int* widths, int out_size, Layout* L, long long* tmp_offsets, int* max_rows) {
  long long tmp = 0;
  int mr = 0;
  for (int i = 0; i < n; ++i) {
    const int h = heights[i], w = widths[i];
    if (h < 1 || w < 1) return dh_fail(DH_ERR_ARG, "image sizes must be positive", __FILE__, __LINE__);
    // taps per output index: ceil(max(in / out, 1)) * 2 + 1 (Pillow's ksize)
    const int kh = (int)((h + out_size - 1) / out_size) * 2 + 1, kw = (int)((w + out_size - 1) / out_size) * 2 + 1;
    if (kh > kMaxTaps || kw > kMaxTaps) return dh_fail(DH_ERR_ARG, "image too large for the resize tap table", __FILE__, __LINE__);
    if (tmp_offsets) tmp_offsets[i] = tmp;
    // intermediate: at most h rows x out_size (horizontal pass first) or out_size rows x w (vertical pass first)
    tmp += align256((long long)(h > w ? h : w) * out_size * 3);
    mr = (h > w ? h : w) > mr ? (h > w ? h : w) : mr;
  }
  L->desc_off = 0;
  L->tables_off = align256((long long)n * sizeof(ImageDesc));
  L->taps_off = L->tables_off + align256((long long)n * 2 * sizeof(AxisTable));
  L->tmp_off = L->taps_off + align256((long long)n * 2 * out_size * kMaxTaps * sizeof(int));
  L->total = L->tmp_off + tmp;
  if (max_rows) *max_rows = mr;
  return DH_OK;
}

}  // namespace

extern "C" int dh_resize_workspace_bytes(int n, const int* heights_host, const int* widths_host, int out_size,
                                         long long* bytes_out) {
  DH_ARG(n >= 0 && bytes_out && out_size >= 1 && out_size <= 256 && (n == 0 || (heights_host && widths_host)));
  Layout L;
  int rc = layout_of(n, heights_host, widths_host, out_size, &L, nullptr, nullptr);
  if (rc) return rc;
  *bytes_out = L.total;
  return DH_OK;
}

extern "C" int dh_resize_bilinear_u8(const unsigned char* packed_hwc, const long long* offsets_host, const int* heights_host,
                                     const int* widths_host, int n, int out_size, unsigned char* out_nchw, void* workspace,
                                     long long workspace_bytes, cudaStream_t stream) {
  DH_ARG(n >= 0 && out_size >= 1 && out_size <= 256);
  if (n == 0) return DH_OK;
  DH_ARG(packed_hwc && offsets_host && heights_host && widths_host && out_nchw && workspace);
  DH_ARG(((uintptr_t)workspace % 256) == 0);
  Layout L;
  int max_rows = 0;
  ImageDesc* hd = (ImageDesc*)malloc((size_t)n * sizeof(ImageDesc));
  long long* toff = (long long*)malloc((size_t)n * sizeof(long long));
  if (!hd || !toff) { free(hd); free(toff); return dh_fail(DH_ERR_ARG, "host allocation failed", __FILE__, __LINE__); }
  int rc = layout_of(n, heights_host, widths_host, out_size, &L, toff, &max_rows);
  if (!rc && workspace_bytes < L.total) rc = dh_fail(DH_ERR_ARG, "workspace smaller than dh_resize_workspace_bytes", __FILE__, __LINE__);
  if (rc) { free(hd); free(toff); return rc; }
  for (int i = 0; i < n; ++i) {
    hd[i].src_off = offsets_host[i];
    hd[i].tmp_off = toff[i];
    hd[i].h = heights_host[i];
    hd[i].w = widths_host[i];
  }
  unsigned char* ws = (unsigned char*)workspace;
  ImageDesc* desc = (ImageDesc*)(ws + L.desc_off);
  AxisTable* tables = (AxisTable*)(ws + L.tables_off);
  int* taps = (int*)(ws + L.taps_off);
  unsigned char* tmp = ws + L.tmp_off;
  // pageable host source: the runtime stages the bytes before returning, so hd may be freed right away
  cudaError_t e = cudaMemcpyAsync(desc, hd, (size_t)n * sizeof(ImageDesc), cudaMemcpyHostToDevice, stream);
  free(hd);
  free(toff);
  if (e != cudaSuccess) return dh_fail((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
  resize_coeffs_kernel<<<2 * n, 256, 0, stream>>>(desc, tables, taps, out_size);
  DH_LAUNCH_OK();
  const int gx = dh_cdiv((long long)max_rows * out_size, 256);
  resize_pass1_kernel<<<dim3(gx < 1024 ? gx : 1024, n), 256, 0, stream>>>(packed_hwc, desc, tables, taps, tmp, out_size);
  DH_LAUNCH_OK();
  resize_pass2_kernel<<<dim3(dh_cdiv((long long)out_size * out_size, 256), n), 256, 0, stream>>>(desc, tables, taps, tmp, out_nchw,
                                                                                                out_size);
  DH_LAUNCH_OK();
  return DH_OK;
}
