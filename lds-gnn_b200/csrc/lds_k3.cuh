// lds_k3.cuh — internal interface of the K3+K4 theta update (shared with lds_outer_step.cu).
#pragma once
#include "lds_common.cuh"

namespace lds {

// Packed bf16 operand rows of the tensor-core update. The factor vectors a = fa row, b = fb row are laid out as
// [hidden part (h columns, zero padded to h16 = round_up(h, 16)) | class part (c columns, padded to 16)] — zero columns
// add nothing to fa.fb — and every 16-column step q becomes one 64-column group (= one 128-byte swizzle atom):
//   F[i][64 q + 16 t + k] = block_t(i)[16 q + k],  t = 0 a_hi, 1 a_lo, 2 b_hi, 3 b_lo
// A generic [n][d] factor pair is the case h = d, c = 0.
static inline int k3_packed_k(int h, int c) { return 4 * (int)(round_up(h, 16) + round_up(c, 16)); }

// (No __restrict__/read-only path on the rows: callers may pack rows their own warp has just written.)
template <typename FloatPtr>
__device__ __forceinline__ __nv_bfloat16 k3_pack_element(FloatPtr fa_row, FloatPtr fb_row, int h, int c, int k) {
  const int h16 = (h + 15) & ~15;
  const int e = 16 * (k >> 6) + (k & 15), t = (k >> 4) & 3;
  int col = -1;
  if (e < h16) { if (e < h) col = e; }
  else if (e - h16 < c) col = h + e - h16;
  if (col < 0) return __float2bfloat16_rn(0.f);
  __nv_bfloat16 hi, lo;
  split_bf16((t < 2) ? fa_row[col] : fb_row[col], hi, lo);
  return (t & 1) ? lo : hi;
}

// fa, fb fp32 [n][ldf] (columns [0,h) hidden part, [h,h+c) class part) -> packed rows F [n][k3_packed_k(h, c)].
int32_t k3_launch_pack(const float* fa, const float* fb, int64_t ldf, int n, int h, int c, void* f, cudaStream_t stream);

// Tensor-core SGD update of rows [row0, row0+rows): theta <- clamp(theta - lr g), g from the packed rows F [n][kf].
// dependent_launch: the previous kernel on `stream` issues griddepcontrol.launch_dependents (programmatic dependent launch).
int32_t k3_launch_tc(float* theta, int64_t ldt, int n, int row0, int rows, const void* f, int kf,
                     const float* cvec, float lr, cudaStream_t stream, bool dependent_launch = false);

}  // namespace lds
