// lds_skinny.cu — the skinny dense products around the second GCN layer in the unrolled inner steps.
//
// Reference: MetaLinear of layer_out, P2 = dropout(H1) W1^T + b1 (src/models/layers.py:43, src/models/gcn.py:29-30), and what
// autograd derives from it in the inner step's backward and the hypergradient's double backward: X W^T with a tiny W
// (hidden x classes) and its weight / bias gradients A^T B, a reduction over all N rows with a result of a few dozen numbers.
// As library calls these were a 30 us "large-K" SGEMM or 8-20 us single-block column reductions each, ~50 times per bilevel
// block (profiles/r01m_launches_graph_block.md); here each is one ~3 us launch. Both are closed under differentiation
// (d(XW^T) needs dY W and dY^T X; d(A^T B) needs B dG^T and A dG), which is what makes them usable under a double backward.
#include <stdlib.h>
#include "lds_common.cuh"

namespace lds {

constexpr int SK_THREADS = 256;
constexpr int SK_CHUNK = 64;          // rows staged per pass of the Gram kernel

// Y[n][j] = sum_k X[n][k] * W[j*sw0 + k*sw1] (+ bias[j]);  k < kdim <= 128, j < m <= 128.  W is staged in shared memory.
__global__ void __launch_bounds__(SK_THREADS)
row_linear_kernel(const float* __restrict__ X, int64_t ldx, int kdim, const float* __restrict__ W, int64_t sw0, int64_t sw1, int m,
                  const float* __restrict__ bias, float* __restrict__ Y, int64_t ldy, int64_t n_rows) {
  pdl_prologue();
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
  pdl_prologue();
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

// Small problems (the unrolled inner steps at Cora / Citeseer shape: N ~ 3000 rows, <= 512 outputs): ONE thread-block cluster.
// The kernel above is bound by its arrival protocol — partial tiles through L2, fence, atomic counter, last-arriver pass: 7.5 us
// per call, 38 calls per bilevel block — not by arithmetic. Here every CTA of the cluster stages its row range in shared memory
// once, thread = (output pair, row slice), the slices are summed in shared memory and the CTAs' partial tiles by the rank-0 CTA
// over distributed shared memory, both in a fixed order: deterministic, no global scratch, no atomics.
constexpr int GC_THREADS = 512;
__global__ void __launch_bounds__(GC_THREADS)
gram_tn_cluster_kernel(const float* __restrict__ A, int64_t lda, int a, const float* __restrict__ B, int64_t ldb, int b, int n_rows,
                       int per, float* __restrict__ out, int64_t ldo) {
  pdl_prologue();
  extern __shared__ float gsm[];
  uint32_t rank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int pairs = a * b, nsl = GC_THREADS / pairs;
  float* a_s = gsm;                         // [per][a]
  float* b_s = a_s + (size_t)per * a;       // [per][b]
  float* sl_s = b_s + (size_t)per * b;      // [nsl][pairs]
  float* part = sl_s + (size_t)nsl * pairs; // [pairs]: this CTA's partial tile, read by rank 0
  const int r0 = (int)rank * per, rows = max(0, min(n_rows, r0 + per) - r0);
  for (int e = threadIdx.x; e < rows * a; e += GC_THREADS) a_s[e] = A[(int64_t)(r0 + e / a) * lda + e % a];
  for (int e = threadIdx.x; e < rows * b; e += GC_THREADS) b_s[e] = B[(int64_t)(r0 + e / b) * ldb + e % b];
  __syncthreads();
  const int p = threadIdx.x % pairs, sl = threadIdx.x / pairs;
  if (sl < nsl) {
    const int i = p / b, j = p - i * b;
    float s0 = 0.f, s1 = 0.f;
    int r = sl;
    for (; r + nsl < rows; r += 2 * nsl) {
      s0 = fmaf(a_s[r * a + i], b_s[r * b + j], s0);
      s1 = fmaf(a_s[(r + nsl) * a + i], b_s[(r + nsl) * b + j], s1);
    }
    if (r < rows) s0 = fmaf(a_s[r * a + i], b_s[r * b + j], s0);
    sl_s[sl * pairs + p] = s0 + s1;
  }
  __syncthreads();
  if ((int)threadIdx.x < pairs) {
    float s = 0.f;
    for (int k = 0; k < nsl; ++k) s += sl_s[k * pairs + threadIdx.x];
    part[threadIdx.x] = s;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && (int)threadIdx.x < pairs) {
    const uint32_t mine = (uint32_t)__cvta_generic_to_shared(part + threadIdx.x);
    float s = 0.f;
    for (uint32_t k = 0; k < csize; ++k) {
      uint32_t remote; float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(mine), "r"(k));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
      s += v;
    }
    out[(int64_t)(threadIdx.x / b) * ldo + threadIdx.x % b] = s;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // nobody exits while rank 0 still reads its shared memory
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
  LDS_CHECK_CUDA(launch_dependent(row_linear_kernel, dim3((unsigned)blocks), dim3(SK_THREADS), smem, (cudaStream_t)stream, x, ldx, k, w, sw0, sw1, m, bias, y, ldy, n_rows));
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
  static const bool no_cluster = getenv("LDS_GRAM_NO_CLUSTER") != nullptr;      // A/B switch
  if (!no_cluster && a * b <= GC_THREADS && n_rows <= 8 * 1024) {
    constexpr int CS = 8;                                      // portable cluster size
    const int per = (int)ceil_div(n_rows, CS);
    const size_t smem = ((size_t)per * (a + b) + (size_t)(GC_THREADS / (a * b)) * a * b + (size_t)a * b) * sizeof(float);
    if (smem <= 200 * 1024) {
      static PerDeviceOnce once;
      if (smem > 48 * 1024 && first_use(once))
        LDS_CHECK_CUDA(cudaFuncSetAttribute(gram_tn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CS); cfg.blockDim = dim3(GC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
      const int n32 = (int)n_rows;
      LDS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gram_tn_cluster_kernel, a_mat, lda, (int)a, b_mat, ldb, (int)b, n32, per, out, ldo));
      return LDS_OK;
    }
  }
  const int ctas = gram_ctas(n_rows);
  const int64_t need = (int64_t)ctas * a * b * sizeof(float) + 256;
  if (workspace_bytes < need) { set_error("lds_gram_tn: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need); return LDS_ERR_WORKSPACE; }
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "lds_gram_tn: workspace must be 256-byte aligned");
  // layout: [counter (256 B, zero before the first use: the kernel leaves it re-armed)] [ctas][a*b] partial tiles
  auto* counter = reinterpret_cast<unsigned int*>(workspace);
  auto* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  const size_t smem = (size_t)SK_CHUNK * (a + b) * sizeof(float);
  if (smem > 48 * 1024) LDS_CHECK_CUDA(cudaFuncSetAttribute(gram_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LDS_CHECK_CUDA(launch_dependent(gram_tn_kernel, dim3((unsigned)ctas), dim3(SK_THREADS), smem, (cudaStream_t)stream, a_mat, lda, a, b_mat, ldb, b, n_rows, partial, counter, out, ldo));
  return LDS_OK;
}
