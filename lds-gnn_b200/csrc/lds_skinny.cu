// lds_skinny.cu — the skinny dense products around the second GCN layer in the unrolled inner steps.
//
// Reference: MetaLinear of layer_out, P2 = dropout(H1) W1^T + b1 (src/models/layers.py:43, src/models/gcn.py:29-30), and what
// autograd derives from it in the inner step's backward and the hypergradient's double backward: X W^T with a tiny W
// (hidden x classes) and its weight / bias gradients A^T B, a reduction over all N rows with a result of a few dozen numbers.
// As library calls these were a 30 us "large-K" SGEMM or 8-20 us single-block column reductions each, ~50 times per bilevel
// block (profiles/r01m_launches_graph_block.md); here each is one ~3 us launch. Both are closed under differentiation
// (d(XW^T) needs dY W and dY^T X; d(A^T B) needs B dG^T and A dG), which is what makes them usable under a double backward.
#include "lds_common.cuh"

namespace lds {

constexpr int SK_THREADS = 256;
constexpr int SK_CHUNK = 64;          // rows staged per pass of the Gram kernel

// Y[n][j] = sum_k X[n][k] * W[j*sw0 + k*sw1] (+ bias[j]);  k < kdim <= 128, j < m <= 128.  W is staged in shared memory.
__global__ void __launch_bounds__(SK_THREADS)
row_linear_kernel(const float* __restrict__ X, int64_t ldx, int kdim, const float* __restrict__ W, int64_t sw0, int64_t sw1, int m,
                  const float* __restrict__ bias, float* __restrict__ Y, int64_t ldy, int64_t n_rows) {
  extern __shared__ float w_s[];                                    // [m][kdim + 1]
  const int kp = kdim + 1;
  for (int e = threadIdx.x; e < m * kdim; e += SK_THREADS) { const int j = e / kdim, k = e % kdim; w_s[j * kp + k] = W[(int64_t)j * sw0 + (int64_t)k * sw1]; }
  __syncthreads();
  const int64_t total = n_rows * m;
  for (int64_t e = (int64_t)blockIdx.x * SK_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * SK_THREADS) {
    const int64_t n = e / m; const int j = (int)(e % m);
    const float* x = X + n * ldx;
    const float* w = w_s + j * kp;
    float acc = bias ? bias[j] : 0.f;
    for (int k = 0; k < kdim; ++k) acc = fmaf(x[k], w[k], acc);
    Y[n * ldy + j] = acc;
  }
}

// out[i][j] = sum_n A[n][i] * B[n][j];  i < a, j < b, a, b <= 128.  Each CTA reduces a contiguous row range (staged through
// shared memory in chunks), writes its partial tile, and the LAST CTA to arrive sums the partials in CTA order: deterministic.
__global__ void __launch_bounds__(SK_THREADS)
gram_tn_kernel(const float* __restrict__ A, int64_t lda, int a, const float* __restrict__ B, int64_t ldb, int b, int64_t n_rows,
               float* __restrict__ partial, unsigned int* __restrict__ counter, float* __restrict__ out, int64_t ldo) {
  extern __shared__ float sm[];                                     // A chunk [SK_CHUNK][a], B chunk [SK_CHUNK][b]
  float* a_s = sm; float* b_s = sm + SK_CHUNK * a;
  __shared__ bool last;
  const int pairs = a * b;
  const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(n_rows, r0 + per);
  constexpr int MAXP = 4;                                            // pairs per thread kept in registers (a*b <= 1024), else loop
  float acc[MAXP] = {0.f, 0.f, 0.f, 0.f};
  const bool in_regs = pairs <= MAXP * SK_THREADS;
  for (int64_t base = r0; base < r1; base += SK_CHUNK) {
    const int rows = (int)min((int64_t)SK_CHUNK, r1 - base);
    for (int e = threadIdx.x; e < rows * a; e += SK_THREADS) a_s[e] = A[(base + e / a) * lda + e % a];
    for (int e = threadIdx.x; e < rows * b; e += SK_THREADS) b_s[e] = B[(base + e / b) * ldb + e % b];
    __syncthreads();
    if (in_regs) {
#pragma unroll
      for (int q = 0; q < MAXP; ++q) {
        const int p = threadIdx.x + q * SK_THREADS;
        if (p < pairs) { const int i = p / b, j = p % b; float s = acc[q]; for (int r = 0; r < rows; ++r) s = fmaf(a_s[r * a + i], b_s[r * b + j], s); acc[q] = s; }
      }
    } else {
      for (int p = threadIdx.x; p < pairs; p += SK_THREADS) {
        const int i = p / b, j = p % b; float s = 0.f;
        for (int r = 0; r < rows; ++r) s = fmaf(a_s[r * a + i], b_s[r * b + j], s);
        float* dst = partial + (int64_t)blockIdx.x * pairs + p;
        *dst = (base == r0 ? 0.f : *dst) + s;
      }
    }
    __syncthreads();
  }
  if (in_regs) {
#pragma unroll
    for (int q = 0; q < MAXP; ++q) { const int p = threadIdx.x + q * SK_THREADS; if (p < pairs) partial[(int64_t)blockIdx.x * pairs + p] = acc[q]; }
  } else if (r0 >= r1) {
    for (int p = threadIdx.x; p < pairs; p += SK_THREADS) partial[(int64_t)blockIdx.x * pairs + p] = 0.f;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int p = threadIdx.x; p < pairs; p += SK_THREADS) {
    // four independent chains so the L2 loads overlap; the association ((c0 + c4 + ..) + (c1 + c5 + ..)) + .. is fixed
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    unsigned c = 0;
    for (; c + 3 < gridDim.x; c += 4) {
      s0 += __ldcg(partial + (int64_t)c * pairs + p);       s1 += __ldcg(partial + (int64_t)(c + 1) * pairs + p);
      s2 += __ldcg(partial + (int64_t)(c + 2) * pairs + p); s3 += __ldcg(partial + (int64_t)(c + 3) * pairs + p);
    }
    for (; c < gridDim.x; ++c) s0 += __ldcg(partial + (int64_t)c * pairs + p);
    out[(int64_t)(p / b) * ldo + p % b] = (s0 + s1) + (s2 + s3);
  }
  if (threadIdx.x == 0) *counter = 0u;                               // re-armed for the next call
}

static int gram_ctas(int64_t n_rows) {
  int64_t g = ceil_div(n_rows, 128);                // few, fatter CTAs: the last CTA's pass over the partial tiles is the serial part
  if (g > num_sms()) g = num_sms();
  return (int)(g < 1 ? 1 : g);
}

}  // namespace lds

extern "C" int32_t lds_row_linear(const float* x, int64_t ldx, int32_t k, const float* w, int64_t sw0, int64_t sw1, int32_t m,
                                  const float* bias, float* y, int64_t ldy, int64_t n_rows, void* stream) {
  using namespace lds;
  LDS_CHECK_ARG(x && w && y, "lds_row_linear: null pointer");
  LDS_CHECK_ARG(n_rows > 0 && k > 0 && k <= 128 && m > 0 && m <= 128 && ldx >= k && ldy >= m, "lds_row_linear: need n_rows > 0, 0 < k, m <= 128, ldx >= k, ldy >= m");
  const size_t smem = (size_t)m * (k + 1) * sizeof(float);
  if (smem > 48 * 1024) LDS_CHECK_CUDA(cudaFuncSetAttribute(row_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = ceil_div(n_rows * m, SK_THREADS);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  row_linear_kernel<<<(unsigned)blocks, SK_THREADS, smem, (cudaStream_t)stream>>>(x, ldx, k, w, sw0, sw1, m, bias, y, ldy, n_rows);
  LDS_CHECK_LAUNCH("row_linear_kernel");
  return LDS_OK;
}

extern "C" int64_t lds_gram_tn_workspace_bytes(int32_t a, int32_t b) {
  if (a <= 0 || b <= 0 || a > 128 || b > 128) return -1;
  return lds::round_up((int64_t)lds::kNumSMsB200 * 2 * a * b * sizeof(float) + 256, 256);
}

extern "C" int32_t lds_gram_tn(const float* a_mat, int64_t lda, int32_t a, const float* b_mat, int64_t ldb, int32_t b, int64_t n_rows,
                               float* out, int64_t ldo, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace lds;
  LDS_CHECK_ARG(a_mat && b_mat && out && workspace, "lds_gram_tn: null pointer");
  LDS_CHECK_ARG(n_rows > 0 && a > 0 && a <= 128 && b > 0 && b <= 128 && lda >= a && ldb >= b && ldo >= b, "lds_gram_tn: need n_rows > 0, 0 < a, b <= 128, lda >= a, ldb >= b, ldo >= b");
  const int ctas = gram_ctas(n_rows);
  const int64_t need = (int64_t)ctas * a * b * sizeof(float) + 256;
  if (workspace_bytes < need) { set_error("lds_gram_tn: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need); return LDS_ERR_WORKSPACE; }
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "lds_gram_tn: workspace must be 256-byte aligned");
  // layout: [counter (256 B, zero before the first use: the kernel leaves it re-armed)] [ctas][a*b] partial tiles
  auto* counter = reinterpret_cast<unsigned int*>(workspace);
  auto* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  const size_t smem = (size_t)SK_CHUNK * (a + b) * sizeof(float);
  if (smem > 48 * 1024) LDS_CHECK_CUDA(cudaFuncSetAttribute(gram_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_tn_kernel<<<ctas, SK_THREADS, smem, (cudaStream_t)stream>>>(a_mat, lda, a, b_mat, ldb, b, n_rows, partial, counter, out, ldo);
  LDS_CHECK_LAUNCH("gram_tn_kernel");
  return LDS_OK;
}
