// lds_common.cuh — shared helpers for liblds_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/lds_b200.h"

namespace lds {

// thread-local error slot behind lds_last_error()
void set_error(const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what);

#define LDS_CHECK_ARG(cond, ...)                      \
  do { if (!(cond)) { lds::set_error(__VA_ARGS__); return LDS_ERR_ARG; } } while (0)
#define LDS_CHECK_CUDA(expr)                          \
  do { cudaError_t _e = (expr); if (_e != cudaSuccess) return lds::cuda_fail(_e, #expr); } while (0)
#define LDS_CHECK_LAUNCH(name)                        \
  do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return lds::cuda_fail(_e, name); } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

constexpr int kLdAlign = 64;          // theta / A_tilde row stride multiple (elements)
constexpr int kNumSMsB200 = 148;

int num_sms();                        // cached cudaDevAttrMultiProcessorCount of the current device
int current_device();                 // cudaGetDevice (0 on failure), < kMaxDevices
constexpr int kMaxDevices = 64;
// One-time per-DEVICE initialisation flags (function attributes such as the dynamic shared-memory limit are per device, so a
// process that drives several GPUs must set them on each): `if (first_use(flags)) { ...set attributes... }`.
struct PerDeviceOnce { bool done[kMaxDevices] = {}; };
inline bool first_use(PerDeviceOnce& o) { const int d = current_device(); if (o.done[d]) return false; o.done[d] = true; return true; }

// Per-kernel timing of lds_outer_step for bench.py (lds_profile_begin/end): records a CUDA event after each launch.
void profile_mark(cudaStream_t stream, int id);
bool profile_active();               // between lds_profile_begin / _end (an event is recorded between the launches: no dependent launch)

// Programmatic dependent launch for the chains of small kernels of the unrolled inner steps (a captured bilevel block is ~450
// kernels of a few us each, back to back): a kernel launched with `launch_dependent` may be scheduled while its predecessor on
// the stream is still running, provided that predecessor has issued griddepcontrol.launch_dependents — every kernel launched this
// way does so in its first statement (`pdl_prologue`) and then waits until the predecessor has completed and its writes are
// visible. What is hidden is the dependent's launch latency; after a kernel that never triggers (ATen's) the launch is an
// ordinary one. LDS_NO_PDL=1 disables it (A/B switch).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// bf16 hi/lo split of an fp32 value: x ~= hi + lo with |err| <= 2^-17 |x|
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

}  // namespace lds
