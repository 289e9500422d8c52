// lds_k3.cuh — internal interface of the K3+K4 theta update (shared with lds_outer_step.cu).
#pragma once
#include "lds_common.cuh"

namespace lds {

static inline int k3_padded_k(int d) { return (int)round_up(3 * (int64_t)d, 64); }

// Element k of the packed bf16 operand rows: Pm = [fa_hi | fa_hi | fa_lo | 0], Qm = [fb_hi | fb_lo | fb_hi | 0].
// (No __restrict__/read-only path on the rows: epi_bwd1 packs rows its own warp has just written.)
template <typename FloatPtr>
__device__ __forceinline__ void k3_pack_element(FloatPtr fa_row, FloatPtr fb_row, int d, int k,
                                                __nv_bfloat16& pm, __nv_bfloat16& qm) {
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  pm = zero; qm = zero;
  if (k >= 3 * d) return;
  const int seg = k / d, c = k - seg * d;
  __nv_bfloat16 ah, al, bh, bl;
  split_bf16(fa_row[c], ah, al);
  split_bf16(fb_row[c], bh, bl);
  pm = (seg == 2) ? al : ah;
  qm = (seg == 1) ? bl : bh;
}

// fa, fb fp32 [n][ldf] -> packed bf16 operands Pm, Qm [n][kp] (coalesced, one thread per element).
int32_t k3_launch_pack(const float* fa, const float* fb, int64_t ldf, int n, int d, int kp, void* pm, void* qm, cudaStream_t stream);

// Tensor-core SGD update of rows [row0, row0+rows): theta <- clamp(theta - lr g), g from the packed operands.
int32_t k3_launch_tc(float* theta, int64_t ldt, int n, int row0, int rows, const void* pm, const void* qm, int kp, int d,
                     const float* cvec, float lr, cudaStream_t stream);

}  // namespace lds
