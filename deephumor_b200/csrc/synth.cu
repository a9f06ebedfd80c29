// Device-side synthetic image generator: bit-identical to deephumor_b200/utils/synth.py::images (hash of
// (seed, 'IMG', global image index, element index) -> uniform(-sqrt3, sqrt3)), so benchmark inputs are
// independent of batch split and world size.
#include "common.cuh"

namespace {
__global__ void synth_images_kernel(float* __restrict__ out, unsigned long long seed, long long first_index, int count,
                                    long long per_image) {
  long long total = (long long)count * per_image;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long img = i / per_image, e = i % per_image;
    uint64_t key = dh_fold(dh_fold(dh_key0(seed), 0x494D47ull), (uint64_t)(first_index + img));
    // per-image, per-channel style (synth.py::image_style): gain = 1 + 0.75*s, offset = 0.5*s'
    const uint64_t c = (uint64_t)(e / (per_image / 3));
    const uint64_t skey = dh_fold(key, 0x5354594Cull);
    const float gain = __fadd_rn(1.0f, __fmul_rn(0.75f, dh_sym_uniform(skey, c, 1.0f)));
    const float offset = __fmul_rn(0.5f, dh_sym_uniform(skey, 3ull + c, 1.0f));
    out[i] = __fadd_rn(__fmul_rn(dh_sym_uniform(key, (uint64_t)e, 1.7320508f), gain), offset);
  }
}
}  // namespace

extern "C" int dh_synth_images(float* out_nchw, unsigned long long seed, long long first_index, int count, int size,
                               cudaStream_t s) {
  DH_ARG(out_nchw && count >= 0 && size > 0);
  if (count == 0) return DH_OK;
  long long per = 3ll * size * size;
  long long blocks = ((long long)count * per + 255) / 256;
  int grid = (int)(blocks > 148 * 32 ? 148 * 32 : blocks);
  synth_images_kernel<<<grid, 256, 0, s>>>(out_nchw, seed, first_index, count, per);
  DH_LAUNCH_OK();
  return DH_OK;
}
