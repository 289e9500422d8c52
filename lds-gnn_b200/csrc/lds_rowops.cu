// lds_rowops.cu — the row-local pieces of the unrolled inner steps that autograd would otherwise run as chains of 1-2 us
// elementwise / reduction kernels on [N, <= 128] tensors (profiles/r01n_launches_graph_block.md: two thirds of a block's kernels).
//
//  * masked NLL of log-softmax on the logits (src/models/gcn.py:34 + src/trainers/inner.py:63-66, outer.py:65-67:
//    F.nll_loss(log_softmax(Z)[mask], y[mask]) and the accuracy), its gradient w.r.t. the logits and the gradient of THAT
//    (what the hyper step's double backward takes through the inner optimiser's create_graph gradient): three kernels instead
//    of ~28 per inner step.  With s = softmax(Z_i), m = |mask|, g = upstream of the loss:
//        loss = (1/m) sum_{i in mask} (logsumexp(Z_i) - Z_i[y_i]);   dZ_i = (g/m) (s - e_{y_i}) on mask rows, 0 elsewhere
//        given U = upstream of dZ:   dZ'_ik = (g/m) s_k (U_ik - <s, U_i>);   dg' = (1/m) sum_i <U_i, s - e_{y_i}>
//  * dr = (sum_c dZ Z + sum_c dQ Q) / r, the gradient of the normalised propagation w.r.t. r = deg^-1/2
//    (models/sampling.py:_FactoredPropagate; src/utils/graph.py:148-152): one kernel instead of mul, addcmul, sum, div.
// `slot[i]` = position of row i among the masked rows or -1; labels y are indexed by the global row.
#include "lds_common.cuh"

namespace lds {

constexpr int RO_THREADS = 256;
constexpr int RO_MAXC = 128;

__device__ __forceinline__ float block_sum(float v, float* scratch) {      // fixed order: deterministic
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
  return s;                                                                 // valid in thread 0
}

// One CTA: loss and accuracy over the m masked rows (rows[k] = global row), written to out[0], out[1].
__global__ void __launch_bounds__(1024)
masked_nll_forward_kernel(const float* __restrict__ z, int64_t ldz, int c, const int64_t* __restrict__ rows, const int64_t* __restrict__ y,
                          int m, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float scratch[32];
  float loss = 0.f, correct = 0.f;
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    const int64_t i = rows[k];
    const float* zi = z + i * ldz;
    const int yi = (int)y[i];
    float mx = zi[0]; int arg = 0;
    for (int j = 1; j < c; ++j) { const float v = zi[j]; if (v > mx) { mx = v; arg = j; } }     // first maximum, like torch.argmax
    float se = 0.f;
    for (int j = 0; j < c; ++j) se += expf(zi[j] - mx);
    loss += (mx + logf(se)) - zi[yi];
    correct += (arg == yi) ? 1.f : 0.f;
  }
  const float l = block_sum(loss, scratch);
  const float a = block_sum(correct, scratch);
  if (threadIdx.x == 0) { out[0] = l / (float)m; out[1] = a / (float)m; }
}

// dZ [n][c] (dense, zero rows outside the mask) = (g / m) (softmax(Z_i) - e_y)
__global__ void __launch_bounds__(RO_THREADS)
masked_nll_grad_kernel(const float* __restrict__ z, int64_t ldz, int c, const int32_t* __restrict__ slot, const int64_t* __restrict__ y,
                       int n, int m, const float* __restrict__ g, float* __restrict__ dz, int64_t ldd) {
  pdl_prologue();
  const float scale = *g / (float)m;
  for (int i = blockIdx.x * RO_THREADS + threadIdx.x; i < n; i += gridDim.x * RO_THREADS) {
    float* di = dz + (int64_t)i * ldd;
    if (slot[i] < 0) { for (int j = 0; j < c; ++j) di[j] = 0.f; continue; }
    const float* zi = z + (int64_t)i * ldz;
    const int yi = (int)y[i];
    float mx = zi[0];
    for (int j = 1; j < c; ++j) mx = fmaxf(mx, zi[j]);
    float se = 0.f;
    for (int j = 0; j < c; ++j) se += expf(zi[j] - mx);
    const float inv = 1.f / se;
    for (int j = 0; j < c; ++j) di[j] = scale * (expf(zi[j] - mx) * inv - (j == yi ? 1.f : 0.f));
  }
}

// Backward of the kernel above. out_z [n][c] = (g/m) s_k (U_ik - <s, U_i>) on mask rows, 0 elsewhere; if out_g != NULL,
// per-row terms <U_i, s - e_y> / m go to out_g_rows[n] (the caller sums them: deterministic).
__global__ void __launch_bounds__(RO_THREADS)
masked_nll_grad_grad_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ z, int64_t ldz, int c,
                            const int32_t* __restrict__ slot, const int64_t* __restrict__ y, int n, int m, const float* __restrict__ g,
                            float* __restrict__ out_z, int64_t ldo, float* __restrict__ out_g_rows) {
  pdl_prologue();
  const float scale = *g / (float)m;
  for (int i = blockIdx.x * RO_THREADS + threadIdx.x; i < n; i += gridDim.x * RO_THREADS) {
    float* oi = out_z ? out_z + (int64_t)i * ldo : nullptr;
    if (slot[i] < 0) {
      if (oi) for (int j = 0; j < c; ++j) oi[j] = 0.f;
      if (out_g_rows) out_g_rows[i] = 0.f;
      continue;
    }
    const float* zi = z + (int64_t)i * ldz;
    const float* ui = u + (int64_t)i * ldu;
    const int yi = (int)y[i];
    float mx = zi[0];
    for (int j = 1; j < c; ++j) mx = fmaxf(mx, zi[j]);
    float se = 0.f;
    for (int j = 0; j < c; ++j) se += expf(zi[j] - mx);
    const float inv = 1.f / se;
    float su = 0.f;
    for (int j = 0; j < c; ++j) su += expf(zi[j] - mx) * inv * ui[j];
    if (oi) for (int j = 0; j < c; ++j) oi[j] = scale * (expf(zi[j] - mx) * inv) * (ui[j] - su);
    if (out_g_rows) out_g_rows[i] = (su - ui[yi]) / (float)m;
  }
}

// out[i] = (sum_j a1[i][j] b1[i][j] + a2[i][j] b2[i][j]) / r[i]
__global__ void __launch_bounds__(RO_THREADS)
row_dot2_kernel(const float* __restrict__ a1, const float* __restrict__ b1, const float* __restrict__ a2, const float* __restrict__ b2,
                int64_t ld, int w, const float* __restrict__ r, int n, float* __restrict__ out) {
  pdl_prologue();
  for (int i = blockIdx.x * RO_THREADS + threadIdx.x; i < n; i += gridDim.x * RO_THREADS) {
    const int64_t o = (int64_t)i * ld;
    float s = 0.f;
    for (int j = 0; j < w; ++j) s = fmaf(a1[o + j], b1[o + j], s);
    float t = 0.f;
    for (int j = 0; j < w; ++j) t = fmaf(a2[o + j], b2[o + j], t);
    out[i] = (s + t) / r[i];
  }
}

static int ro_grid(int n) {
  int64_t g = ceil_div(n, RO_THREADS);
  const int64_t cap = (int64_t)num_sms() * 4;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace lds

using namespace lds;

extern "C" int32_t lds_masked_nll_forward(const float* z, int64_t ld_z, int32_t c, const int64_t* rows, const int64_t* y, int32_t m,
                                          float* out_loss_acc, void* stream) {
  LDS_CHECK_ARG(z && rows && y && out_loss_acc, "lds_masked_nll_forward: null pointer");
  LDS_CHECK_ARG(c > 0 && c <= RO_MAXC && ld_z >= c && m > 0, "lds_masked_nll_forward: need 0 < c <= 128, ld_z >= c, m > 0");
  LDS_CHECK_CUDA(launch_dependent(masked_nll_forward_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, z, ld_z, c, rows, y, m, out_loss_acc));
  return LDS_OK;
}

extern "C" int32_t lds_masked_nll_grad(const float* z, int64_t ld_z, int32_t c, const int32_t* slot, const int64_t* y, int32_t n, int32_t m,
                                       const float* grad_loss, float* dz, int64_t ld_dz, void* stream) {
  LDS_CHECK_ARG(z && slot && y && grad_loss && dz, "lds_masked_nll_grad: null pointer");
  LDS_CHECK_ARG(c > 0 && c <= RO_MAXC && ld_z >= c && ld_dz >= c && n > 0 && m > 0, "lds_masked_nll_grad: need 0 < c <= 128, ld >= c, n, m > 0");
  LDS_CHECK_CUDA(launch_dependent(masked_nll_grad_kernel, dim3((unsigned)ro_grid(n)), dim3(RO_THREADS), 0, (cudaStream_t)stream, z, ld_z, c, slot, y, n, m, grad_loss, dz, ld_dz));
  return LDS_OK;
}

extern "C" int32_t lds_masked_nll_grad_grad(const float* u, int64_t ld_u, const float* z, int64_t ld_z, int32_t c, const int32_t* slot,
                                            const int64_t* y, int32_t n, int32_t m, const float* grad_loss,
                                            float* out_z, int64_t ld_out, float* out_g_rows, void* stream) {
  LDS_CHECK_ARG(u && z && slot && y && grad_loss && (out_z || out_g_rows), "lds_masked_nll_grad_grad: null pointer");
  LDS_CHECK_ARG(c > 0 && c <= RO_MAXC && ld_z >= c && ld_u >= c && (!out_z || ld_out >= c) && n > 0 && m > 0,
                "lds_masked_nll_grad_grad: need 0 < c <= 128, ld >= c, n, m > 0");
  LDS_CHECK_CUDA(launch_dependent(masked_nll_grad_grad_kernel, dim3((unsigned)ro_grid(n)), dim3(RO_THREADS), 0, (cudaStream_t)stream, u, ld_u, z, ld_z, c, slot, y, n, m, grad_loss, out_z, ld_out, out_g_rows));
  return LDS_OK;
}

extern "C" int32_t lds_row_dot2(const float* a1, const float* b1, const float* a2, const float* b2, int64_t ld, int32_t w,
                                const float* r, int32_t n, float* out, void* stream) {
  LDS_CHECK_ARG(a1 && b1 && a2 && b2 && r && out, "lds_row_dot2: null pointer");
  LDS_CHECK_ARG(w > 0 && ld >= w && n > 0, "lds_row_dot2: need w > 0, ld >= w, n > 0");
  LDS_CHECK_CUDA(launch_dependent(row_dot2_kernel, dim3((unsigned)ro_grid(n)), dim3(RO_THREADS), 0, (cudaStream_t)stream, a1, b1, a2, b2, ld, w, r, n, out));
  return LDS_OK;
}
