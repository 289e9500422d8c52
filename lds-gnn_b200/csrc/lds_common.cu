// lds_common.cu — error slot, version, device check, host Philox restatement.
#include <stdarg.h>
#include <string.h>
#include "lds_common.cuh"
#include "lds_philox.cuh"

namespace lds {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return LDS_ERR_CUDA;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    else { (void)cudaGetLastError(); return kNumSMsB200; }
  }
  return cached;
}

}  // namespace lds

extern "C" int32_t lds_version(void) { return 100; }   // 0.1.0

extern "C" const char* lds_last_error(void) { return lds::g_error; }

extern "C" int32_t lds_device_check(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return lds::cuda_fail(e, "lds_device_check");
  if (major != 10) { lds::set_error("liblds_b200 is built for sm_100a only; current device has compute capability major %d", major); return LDS_ERR_UNSUPPORTED; }
  return LDS_OK;
}

extern "C" int64_t lds_padded_ld(int32_t n) { return lds::round_up(n, lds::kLdAlign); }

extern "C" float lds_philox_uniform(uint64_t seed, uint64_t step, uint32_t stream, uint32_t sample, uint32_t i, uint32_t j) {
  const lds::PhiloxKey key = lds::philox_key(seed, step, stream, sample);
  uint32_t w[4];
  if (stream == LDS_STREAM_EDGES) {
    const uint32_t a = i < j ? i : j, b = i < j ? j : i;
    lds::philox4x32_10(b / 2, a / 2, key, w);
    return lds::philox_to_uniform(w[2 * (a % 2) + (b % 2)]);
  }
  lds::philox4x32_10(j / 4, i, key, w);
  return lds::philox_to_uniform(w[j % 4]);
}
