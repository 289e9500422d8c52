// lds_k3.cuh — internal interface of the K3+K4 theta update (shared with lds_outer_step.cu).
#pragma once
#include "lds_common.cuh"

namespace lds {

// Packed bf16 operand rows of the tensor-core update: per 16-column step q of the factors, one 64-column group
//   F[i][64 q + 16 t + k] = block_t(i)[16 q + k],  t = 0 a_hi, 1 a_lo, 2 b_hi, 3 b_lo   (a = fa row, b = fb row, zero past d)
static inline int k3_packed_k(int d) { return 4 * (int)round_up(d, 16); }

// (No __restrict__/read-only path on the rows: callers may pack rows their own warp has just written.)
template <typename FloatPtr>
__device__ __forceinline__ __nv_bfloat16 k3_pack_element(FloatPtr fa_row, FloatPtr fb_row, int d, int k) {
  const int q = k >> 6, t = (k >> 4) & 3, c = 16 * q + (k & 15);
  if (c >= d) return __float2bfloat16_rn(0.f);
  __nv_bfloat16 hi, lo;
  split_bf16((t < 2) ? fa_row[c] : fb_row[c], hi, lo);
  return (t & 1) ? lo : hi;
}

// fa, fb fp32 [n][ldf] -> packed operand rows F [n][k3_packed_k(d)] (coalesced, one thread per element).
int32_t k3_launch_pack(const float* fa, const float* fb, int64_t ldf, int n, int d, void* f, cudaStream_t stream);

// Tensor-core SGD update of rows [row0, row0+rows): theta <- clamp(theta - lr g), g from the packed operand rows.
int32_t k3_launch_tc(float* theta, int64_t ldt, int n, int row0, int rows, const void* f, int d,
                     const float* cvec, float lr, cudaStream_t stream);

}  // namespace lds
