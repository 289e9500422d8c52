// lds_spmm.cu — sparse-feature products of the unrolled inner steps.
//
// Reference: MetaLinear = F.linear(dropout(X), W0, b0) inside MetaDenseGraphConvolution.forward
// (src/models/layers.py:43, src/models/gcn.py:27-28) and its autograd transposes (dW0 = dP1^T X', and the same two
// products again in the double backward of the hypergradient). X is a row-normalised bag of words, ~1 % dense
// (NormalizeFeatures, src/data/dataloader.py:100-101): as dense fp32 SGEMMs these products were 40 % of the GPU time
// of a bilevel block (profiles/r01m_graph_block_kernels.md); from CSR they touch nnz * w values.
//
//   Y[i][c] = sum_{k in [ptr[i], ptr[i+1])} val[perm ? perm[k] : k] * B[idx[k] * ldb_row + c * ldb_col]
//
// One kernel serves S*B (CSR of X) and S^T*B (CSR of X^T = CSC of X, `perm` mapping its entries to X's value order, so a
// per-step dropout of the values needs no second value array). Fixed summation order: bitwise reproducible.
#include "lds_common.cuh"

namespace lds {

// blockDim = (WP, 256 / WP): threadIdx.x = output column, threadIdx.y = row within the block. The WP lanes of a row read
// one value (broadcast) and WP consecutive (or ldb_col-strided) elements of B per non-zero.
template <int WP>
__global__ void __launch_bounds__(256)
spmm_csr_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                const int32_t* __restrict__ perm, int rows, const float* __restrict__ B, int64_t ldb_row, int64_t ldb_col, int w,
                float* __restrict__ Y, int64_t ldy) {
  pdl_prologue();
  const int row = blockIdx.x * (256 / WP) + threadIdx.y;
  if (row >= rows) return;
  const int k0 = ptr[row], k1 = ptr[row + 1];
  for (int c = threadIdx.x; c < w; c += WP) {
    const float* bc = B + (int64_t)c * ldb_col;
    float acc0 = 0.f, acc1 = 0.f;
    int k = k0;
    for (; k + 1 < k1; k += 2) {                     // two independent chains: the loop is latency-bound (short rows)
      const float v0 = val[perm ? perm[k] : k], v1 = val[perm ? perm[k + 1] : k + 1];
      acc0 = fmaf(v0, bc[(int64_t)idx[k] * ldb_row], acc0);
      acc1 = fmaf(v1, bc[(int64_t)idx[k + 1] * ldb_row], acc1);
    }
    if (k < k1) acc0 = fmaf(val[perm ? perm[k] : k], bc[(int64_t)idx[k] * ldb_row], acc0);
    Y[(int64_t)row * ldy + c] = acc0 + acc1;
  }
}

}  // namespace lds

extern "C" int32_t lds_spmm_csr(const int32_t* ptr, const int32_t* idx, const float* val, const int32_t* perm, int32_t rows,
                                const float* b, int64_t ldb_row, int64_t ldb_col, int32_t w,
                                float* y, int64_t ldy, void* stream) {
  using namespace lds;
  LDS_CHECK_ARG(ptr && idx && val && b && y, "lds_spmm_csr: null pointer");
  LDS_CHECK_ARG(rows > 0 && w > 0 && ldy >= w, "lds_spmm_csr: need rows > 0, w > 0, ldy >= w");
  const cudaStream_t s = (cudaStream_t)stream;
  if (w <= 8) {
    LDS_CHECK_CUDA(launch_dependent(spmm_csr_kernel<8>, dim3((unsigned)ceil_div(rows, 32)), dim3(8, 32), 0, s, ptr, idx, val, perm, rows, b, ldb_row, ldb_col, w, y, ldy));
  } else if (w <= 16) {
    LDS_CHECK_CUDA(launch_dependent(spmm_csr_kernel<16>, dim3((unsigned)ceil_div(rows, 16)), dim3(16, 16), 0, s, ptr, idx, val, perm, rows, b, ldb_row, ldb_col, w, y, ldy));
  } else {
    LDS_CHECK_CUDA(launch_dependent(spmm_csr_kernel<32>, dim3((unsigned)ceil_div(rows, 8)), dim3(32, 8), 0, s, ptr, idx, val, perm, rows, b, ldb_row, ldb_col, w, y, ldy));
  }
  return LDS_OK;
}
