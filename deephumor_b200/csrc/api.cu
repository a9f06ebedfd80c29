// Library-level entry points: version, error text, device check.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

int dh_fail(int code, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s:%d)", what, base ? base + 1 : file, line);
  return code;
}

extern "C" const char* dh_last_error(void) { return g_last_error; }
extern "C" int dh_version(void) { return DH_VERSION; }

// 0 when device `dev` is an sm_100 part the library's sm_100a cubins can run on.
extern "C" int dh_check_device(int dev) {
  cudaDeviceProp prop;
  DH_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return dh_fail(DH_ERR_DEVICE, "libdeephumor_sm100 needs an sm_100 (B200) device", __FILE__, __LINE__);
  return DH_OK;
}
