#include <stdlib.h>
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

constexpr int kMaxMarks = 64;
static bool g_profiling = false;
static int g_marks = 0;
static cudaEvent_t g_events[kMaxMarks];
static int g_ids[kMaxMarks];
static bool g_events_created = false;

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

bool profile_active() { return g_profiling; }

bool pdl_enabled() {
  static const bool on = getenv("LDS_NO_PDL") == nullptr;
  return on;
}

void profile_mark(cudaStream_t stream, int id) {
  if (!g_profiling || g_marks >= kMaxMarks) return;
  cudaEventRecord(g_events[g_marks], stream);
  g_ids[g_marks++] = id;
}

}  // namespace lds

extern "C" int32_t lds_profile_begin(void) {
  if (!lds::g_events_created) {
    for (int i = 0; i < lds::kMaxMarks; ++i) LDS_CHECK_CUDA(cudaEventCreate(&lds::g_events[i]));
    lds::g_events_created = true;
  }
  lds::g_marks = 0;
  lds::g_profiling = true;
  return LDS_OK;
}

extern "C" int32_t lds_profile_end(float* ms_out, int32_t* ids_out, int32_t cap) {
  lds::g_profiling = false;
  if (lds::g_marks == 0) return 0;
  if (cudaEventSynchronize(lds::g_events[lds::g_marks - 1]) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
  int n = 0;
  for (int i = 1; i < lds::g_marks && n < cap; ++i, ++n) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, lds::g_events[i - 1], lds::g_events[i]);
    ms_out[n] = ms;
    ids_out[n] = lds::g_ids[i];
  }
  return n;
}

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


// ------------------------------------------------------------------------------------------------
// NVLink peer-memory all-gather push (sharded step): every rank stores its block straight into all
// peers' gather buffers (P2P-mapped pointers, e.g. torch symmetric memory) with 128-bit stores; the caller follows
// it with a cross-GPU barrier. Replaces an NCCL all-gather whose launch + protocol latency dominated at these sizes
// (0.5 - 5 MB per rank).
// ------------------------------------------------------------------------------------------------
namespace lds {
__global__ void __launch_bounds__(256)
peer_push_kernel(const uint4* __restrict__ src, uint8_t* const* __restrict__ dst_bases, int world, int64_t dst_offset_bytes, int64_t n16) {
  const int peer = blockIdx.y;
  if (peer >= world) return;
  uint4* dst = reinterpret_cast<uint4*>(dst_bases[peer] + dst_offset_bytes);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace lds

// Upload of a step's inputs (the GCN weights of an outer step that arrive from the host) on a stream of the library's own:
// the copy overlaps whatever is still executing on `stream` (the backward half and the theta update of the previous step),
// and work enqueued on `stream` after this call sees the data. Three driver calls, no host synchronisation.
extern "C" int32_t lds_upload_async(void* dst_device, const void* src_pinned_host, int64_t bytes, void* stream) {
  LDS_CHECK_ARG(dst_device && src_pinned_host && bytes > 0, "lds_upload_async: bad arguments");
  static cudaStream_t side[lds::kMaxDevices];
  static cudaEvent_t ev[lds::kMaxDevices][8];
  static unsigned turn[lds::kMaxDevices];
  static lds::PerDeviceOnce once;
  const int d = lds::current_device();
  if (lds::first_use(once)) {
    LDS_CHECK_CUDA(cudaStreamCreateWithFlags(&side[d], cudaStreamNonBlocking));
    for (int k = 0; k < 8; ++k) LDS_CHECK_CUDA(cudaEventCreateWithFlags(&ev[d][k], cudaEventDisableTiming));
    turn[d] = 0;
  }
  cudaEvent_t e = ev[d][turn[d]++ & 7];
  LDS_CHECK_CUDA(cudaMemcpyAsync(dst_device, src_pinned_host, (size_t)bytes, cudaMemcpyHostToDevice, side[d]));
  LDS_CHECK_CUDA(cudaEventRecord(e, side[d]));
  LDS_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, e, 0));
  return LDS_OK;
}

extern "C" int32_t lds_peer_push(const void* src, void* const* dst_bases, int32_t world, int64_t dst_offset_bytes, int64_t bytes, void* stream) {
  LDS_CHECK_ARG(src && dst_bases && world > 0 && world <= 64, "lds_peer_push: bad arguments");
  LDS_CHECK_ARG(bytes > 0 && bytes % 16 == 0 && dst_offset_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
                "lds_peer_push: src, offset and size must be 16-byte aligned");
  const int64_t n16 = bytes / 16;
  int bx = (int)((n16 + 255) / 256);
  const int cap = (4 * lds::num_sms() + world - 1) / world;      // ~4 CTAs per SM in total
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  lds::peer_push_kernel<<<dim3((unsigned)bx, (unsigned)world), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint8_t* const*>(dst_bases), world, dst_offset_bytes, n16);
  LDS_CHECK_LAUNCH("peer_push_kernel");
  return LDS_OK;
}
