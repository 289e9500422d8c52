// lds_theta0.cu — construction of theta_0, the initial edge-probability matrix, on the device.
//
// The reference builds theta_0 on the host, once per run (paths relative to the reference repo):
//   dense adjacency from an edge list        to_dense_adj                     src/utils/graph.py:80-116
//   kNN graph (k = 10 / 20, cosine / p = 2)  sklearn kneighbors_graph(mode="connectivity", include_self=loop)
//                                            src/data/utils.py:165-183, KNNGraph src/data/transforms.py:15-28
//   symmetrisation                           MakeUndirected src/data/transforms.py:31-38 (max(A, A^T) on a 0/1 matrix)
//   random edge removal                      remove_edges* src/data/utils.py:186-227, RemoveEdges transforms.py:41-55
// One-off per run but O(N^2 F) on the CPU. Here: one tiled fp32 distance pass + a warp-per-row k-selection, a scatter,
// and an order-preserving (row-major, like Tensor.nonzero()) compaction that consumes the SAME permutation the
// reference draws with torch.randperm, so the retained edge set is identical.
#include "lds_common.cuh"

namespace lds {

// ------------------------------------------------------------------------------------------------
// kNN
// ------------------------------------------------------------------------------------------------
// sq[i] = |x_i|^2, inv[i] = 1 / |x_i| (0 for a zero row: sklearn's normalize leaves it zero => cosine distance 1)
__global__ void knn_norm_kernel(const float* __restrict__ x, int64_t ldx, int n, int f, float* __restrict__ sq, float* __restrict__ inv) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int k = lane; k < f; k += 32) { const float v = x[(int64_t)row * ldx + k]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if (lane == 0) { sq[row] = s; inv[row] = s > 0.f ? 1.0f / sqrtf(s) : 0.f; }
}

constexpr int KT = 64;     // output tile
constexpr int KK = 16;     // feature chunk

// dist[i][j] for a 64 x 64 tile: cosine 1 - <x_i, x_j> / (|x_i| |x_j|), or squared euclidean |x_i|^2 + |x_j|^2 - 2 <x_i, x_j>
// (monotone in the p = 2 Minkowski distance). The point itself gets -1 (always its own first neighbour, include_self) or +inf.
__global__ void __launch_bounds__(256)
knn_dist_kernel(const float* __restrict__ x, int64_t ldx, int n, int f, const float* __restrict__ sq, const float* __restrict__ inv,
                int metric, int loop, float* __restrict__ dist, int64_t ldd) {
  __shared__ float sa[KK][KT + 4], sb[KK][KT + 4];
  const int i0 = blockIdx.y * KT, j0 = blockIdx.x * KT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < f; k0 += KK) {
    for (int idx = threadIdx.x; idx < KT * KK; idx += 256) {
      const int r = idx / KK, kk = idx - r * KK;
      const int k = k0 + kk;
      sa[kk][r] = (i0 + r < n && k < f) ? x[(int64_t)(i0 + r) * ldx + k] : 0.f;
      sb[kk][r] = (j0 + r < n && k < f) ? x[(int64_t)(j0 + r) * ldx + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sa[kk][4 * ty]);
      const float4 bv = *reinterpret_cast<const float4*>(&sb[kk][4 * tx]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + 4 * ty + a;
    if (i >= n) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + 4 * tx + b;
      if (j >= (int)ldd) continue;
      float d;
      if (j >= n) d = __int_as_float(0x7f800000);
      else if (i == j) d = loop ? -1.0f : __int_as_float(0x7f800000);
      else if (metric == 0) d = 1.0f - acc[a][b] * inv[i] * inv[j];
      else d = fmaxf(sq[i] + sq[j] - 2.0f * acc[a][b], 0.f);
      dist[(int64_t)i * ldd + j] = d;
    }
  }
}

// One warp per row: k rounds of "smallest (distance, column) pair lexicographically above the previous pick" — ties go to
// the smaller column index, deterministic. Then the row of distances is overwritten in place by the 0/1 connectivity row.
__global__ void knn_select_kernel(float* __restrict__ dist, int64_t ldd, int n, int k) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float* d = dist + (int64_t)row * ldd;
  float last_v = -__int_as_float(0x7f800000); int last_j = -1;
  int mine0 = -1, mine1 = -1;                                  // lane l keeps picks l and l + 32 (k <= 64)
  for (int t = 0; t < k; ++t) {
    float bv = __int_as_float(0x7f800000); int bj = 0x7fffffff;
    for (int j = lane; j < n; j += 32) {
      const float v = d[j];
      const bool after = (v > last_v) || (v == last_v && j > last_j);
      if (after && (v < bv || (v == bv && j < bj))) { bv = v; bj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
    }
    last_v = bv; last_j = bj;
    if ((t & 31) == lane) { if (t < 32) mine0 = bj; else mine1 = bj; }
  }
  __syncwarp();
  for (int j = lane; j < (int)ldd; j += 32) d[j] = 0.f;
  __syncwarp();                                                // orders the zero fill before the ones (same warp)
  if (mine0 >= 0 && mine0 < n) d[mine0] = 1.0f;
  if (mine1 >= 0 && mine1 < n) d[mine1] = 1.0f;
}

// In place A <- max(A, A^T) (MakeUndirected on a 0/1 matrix), one CTA per unordered pair of 32 x 32 tiles.
__global__ void symmetrize_max_kernel(float* __restrict__ a, int64_t ld, int n) {
  __shared__ float ta[32][33], tb[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 256 threads: ty in [0, 8)
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    ta[r][tx] = (i < n && j < n) ? a[(int64_t)i * ld + j] : 0.f;
    const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
    tb[r][tx] = (i2 < n && j2 < n) ? a[(int64_t)i2 * ld + j2] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < n && j < n) a[(int64_t)i * ld + j] = fmaxf(ta[r][tx], tb[tx][r]);
    const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
    if (bi != bj && i2 < n && j2 < n) a[(int64_t)i2 * ld + j2] = fmaxf(tb[r][tx], ta[tx][r]);
  }
}

// ------------------------------------------------------------------------------------------------
// edge list -> dense
// ------------------------------------------------------------------------------------------------
__global__ void edges_scatter_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t e, int n,
                                     float* __restrict__ adj, int64_t ld, int symmetric, int* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= e) return;
  const int64_t i = src[t], j = dst[t];
  if (i < 0 || j < 0 || i >= n || j >= n) { atomicAdd(bad, 1); return; }
  adj[i * ld + j] = 1.0f;
  if (symmetric) adj[j * ld + i] = 1.0f;
}

// ------------------------------------------------------------------------------------------------
// random edge removal
// ------------------------------------------------------------------------------------------------
// count[i] = number of non-zeros of row i (columns >= i only when triu)
__global__ void edge_count_kernel(const float* __restrict__ adj, int64_t ld, int n, int triu, int64_t* __restrict__ count) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int c = 0;
  for (int j = (triu ? row : 0) + lane; j < n; j += 32) c += adj[(int64_t)row * ld + j] != 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) count[row + 1] = c;
}

// in-place inclusive scan of off[1..n] (off[0] = 0): one block, sequential 1024-element chunks
__global__ void edge_scan_kernel(int64_t* __restrict__ off, int n) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) { carry = 0; off[0] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int64_t v = (i < n) ? off[i + 1] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    if (w == 0) {
      int64_t t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int64_t before = carry + (w > 0 ? warp_tot[w - 1] : 0);
    if (i < n) off[i + 1] = before + v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
}

__global__ void keep_flags_kernel(const int64_t* __restrict__ perm, int64_t num_keep, int64_t nnz, uint8_t* __restrict__ keep) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_keep) return;
  const int64_t e = perm[t];
  if (e >= 0 && e < nnz) keep[e] = 1;
}

// Edge e = position of (i, j) in the row-major list of non-zeros (Tensor.nonzero() order): kept iff keep[e].
// out must be zero-filled. triu: the kept upper-triangle entries are mirrored, the diagonal kept once
// (to_undirected(from_triu_only=True), src/utils/graph.py:35-37).
__global__ void remove_edges_apply_kernel(const float* __restrict__ adj, int64_t ld, int n, int triu, const int64_t* __restrict__ off,
                                          const uint8_t* __restrict__ keep, float* __restrict__ out, int64_t ldo) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int64_t e = off[row];
  const int start = triu ? row : 0;
  for (int j0 = start; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    const float v = (j < n) ? adj[(int64_t)row * ld + j] : 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
    if (v != 0.f) {
      const int64_t mine = e + __popc(m & ((1u << lane) - 1u));
      if (keep[mine]) {
        out[(int64_t)row * ldo + j] = v;
        if (triu && j != row) out[(int64_t)j * ldo + row] = v;
      }
    }
    e += __popc(m);
  }
}

// ------------------------------------------------------------------------------------------------
// empirical_mean_loss (src/utils/evaluation.py:51-84): masked NLL and accuracy of S x N x C log-probabilities for TWO
// node masks (validation, test) in one pass. Deterministic: per-block partials, summed in a fixed order by the finisher.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eval_metrics_kernel(const float* __restrict__ logp, int n, int c, const int64_t* __restrict__ y, const uint8_t* __restrict__ mask_a,
                    const uint8_t* __restrict__ mask_b, float* __restrict__ partial) {
  __shared__ float red[8][4];
  const int i = blockIdx.x * 256 + threadIdx.x, s = blockIdx.y;
  float v[4] = {0.f, 0.f, 0.f, 0.f};                          // loss a, correct a, loss b, correct b
  if (i < n) {
    const bool ma = mask_a[i] != 0, mb = mask_b[i] != 0;
    if (ma || mb) {
      const float* row = logp + ((int64_t)s * n + i) * c;
      const int yi = (int)y[i];
      float best = row[0]; int arg = 0;
      for (int o = 1; o < c; ++o) { const float t = row[o]; if (t > best) { best = t; arg = o; } }    // first maximum wins (torch.argmax)
      const float nll = -row[yi], hit = (arg == yi) ? 1.f : 0.f;
      if (ma) { v[0] = nll; v[1] = hit; }
      if (mb) { v[2] = nll; v[3] = hit; }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w][0] = v[0]; red[w][1] = v[1]; red[w][2] = v[2]; red[w][3] = v[3]; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    partial[((int64_t)s * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = t;
  }
}

__global__ void eval_metrics_finish_kernel(const float* __restrict__ partial, int blocks, int samples, float inv_a, float inv_b, float* __restrict__ out4) {
  if (threadIdx.x >= 4) return;
  double t = 0.0;                                             // mean over samples of the per-sample masked means
  for (int k = 0; k < blocks * samples; ++k) t += (double)partial[(int64_t)k * 4 + threadIdx.x];
  out4[threadIdx.x] = (float)(t * (double)((threadIdx.x < 2) ? inv_a : inv_b) / (double)samples);
}

}  // namespace lds

using namespace lds;

extern "C" int64_t lds_eval_metrics_workspace_bytes(int32_t n, int32_t samples) {
  return (n > 0 && samples > 0) ? (int64_t)samples * ceil_div(n, 256) * 16 : -1;
}

extern "C" int32_t lds_eval_metrics(const float* logp, int32_t samples, int32_t n, int32_t c, const int64_t* y,
                                    const uint8_t* mask_a, int32_t count_a, const uint8_t* mask_b, int32_t count_b,
                                    float* out4, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(logp && y && mask_a && mask_b && out4 && workspace, "lds_eval_metrics: null pointer");
  LDS_CHECK_ARG(samples > 0 && n > 0 && c > 0 && count_a > 0 && count_b > 0, "lds_eval_metrics: need positive sizes and mask counts");
  if (workspace_bytes < lds_eval_metrics_workspace_bytes(n, samples)) { set_error("lds_eval_metrics: workspace too small"); return LDS_ERR_WORKSPACE; }
  const int blocks = (int)ceil_div(n, 256);
  eval_metrics_kernel<<<dim3((unsigned)blocks, (unsigned)samples), 256, 0, stream>>>(logp, n, c, y, mask_a, mask_b, reinterpret_cast<float*>(workspace));
  LDS_CHECK_LAUNCH("eval_metrics_kernel");
  eval_metrics_finish_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const float*>(workspace), blocks, samples, 1.0f / (float)count_a, 1.0f / (float)count_b, out4);
  LDS_CHECK_LAUNCH("eval_metrics_finish_kernel");
  return LDS_OK;
}

extern "C" int64_t lds_knn_workspace_bytes(int32_t n) { return n > 0 ? round_up((int64_t)n * 8, 256) : -1; }

extern "C" int32_t lds_knn_graph(const float* x, int64_t ld_x, int32_t n, int32_t f, int32_t k, int32_t metric, int32_t loop,
                                 int32_t symmetrize, float* adj_out, int64_t ld_adj, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(x && adj_out && workspace, "lds_knn_graph: null pointer");
  LDS_CHECK_ARG(n > 0 && f > 0 && ld_x >= f && ld_adj >= n, "lds_knn_graph: need n, f > 0, ld_x >= f, ld_adj >= n");
  LDS_CHECK_ARG(metric == 0 || metric == 1, "lds_knn_graph: metric must be 0 (cosine) or 1 (euclidean / minkowski p = 2)");
  LDS_CHECK_ARG(k > 0 && k <= 64 && k <= n - (loop ? 0 : 1), "lds_knn_graph: k = %d outside [1, min(64, %d)]", k, n - (loop ? 0 : 1));
  if (workspace_bytes < lds_knn_workspace_bytes(n)) { set_error("lds_knn_graph: workspace too small"); return LDS_ERR_WORKSPACE; }
  float* sq = reinterpret_cast<float*>(workspace);
  float* inv = sq + n;
  knn_norm_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(x, ld_x, n, f, sq, inv);
  LDS_CHECK_LAUNCH("knn_norm_kernel");
  dim3 grid((unsigned)ceil_div(ld_adj, KT), (unsigned)ceil_div(n, KT));
  knn_dist_kernel<<<grid, 256, 0, stream>>>(x, ld_x, n, f, sq, inv, metric, loop, adj_out, ld_adj);
  LDS_CHECK_LAUNCH("knn_dist_kernel");
  knn_select_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(adj_out, ld_adj, n, k);
  LDS_CHECK_LAUNCH("knn_select_kernel");
  if (symmetrize) {
    const unsigned t = (unsigned)ceil_div(n, 32);
    symmetrize_max_kernel<<<dim3(t, t), 256, 0, stream>>>(adj_out, ld_adj, n);
    LDS_CHECK_LAUNCH("symmetrize_max_kernel");
  }
  return LDS_OK;
}

extern "C" int32_t lds_symmetrize_max(float* adj, int64_t ld, int32_t n, void* stream_) {
  LDS_CHECK_ARG(adj && n > 0 && ld >= n, "lds_symmetrize_max: bad arguments");
  const unsigned t = (unsigned)ceil_div(n, 32);
  symmetrize_max_kernel<<<dim3(t, t), 256, 0, (cudaStream_t)stream_>>>(adj, ld, n);
  LDS_CHECK_LAUNCH("symmetrize_max_kernel");
  return LDS_OK;
}

extern "C" int32_t lds_edges_to_dense(const int64_t* edge_index, int64_t num_edges, int32_t n, int32_t symmetric,
                                      float* adj_out, int64_t ld_adj, int32_t* bad_count, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(adj_out && bad_count && n > 0 && ld_adj >= n && num_edges >= 0, "lds_edges_to_dense: bad arguments");
  LDS_CHECK_ARG(edge_index || num_edges == 0, "lds_edges_to_dense: null edge_index");
  LDS_CHECK_CUDA(cudaMemsetAsync(adj_out, 0, (size_t)n * ld_adj * sizeof(float), stream));
  LDS_CHECK_CUDA(cudaMemsetAsync(bad_count, 0, sizeof(int32_t), stream));
  if (num_edges > 0) {
    edges_scatter_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(edge_index, edge_index + num_edges, num_edges, n, adj_out, ld_adj, symmetric, bad_count);
    LDS_CHECK_LAUNCH("edges_scatter_kernel");
  }
  return LDS_OK;
}

extern "C" int32_t lds_edge_offsets(const float* adj, int64_t ld, int32_t n, int32_t triu, int64_t* offsets, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(adj && offsets && n > 0 && ld >= n, "lds_edge_offsets: bad arguments");
  edge_count_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(adj, ld, n, triu, offsets);
  LDS_CHECK_LAUNCH("edge_count_kernel");
  edge_scan_kernel<<<1, 1024, 0, stream>>>(offsets, n);
  LDS_CHECK_LAUNCH("edge_scan_kernel");
  return LDS_OK;
}

extern "C" int32_t lds_remove_edges_apply(const float* adj, int64_t ld, int32_t n, int32_t triu, const int64_t* offsets,
                                          const int64_t* perm, int64_t nnz, int64_t num_keep,
                                          float* out, int64_t ld_out, uint8_t* keep_flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(adj && offsets && out && n > 0 && ld >= n && ld_out >= n, "lds_remove_edges_apply: bad arguments");
  LDS_CHECK_ARG(nnz >= 0 && num_keep >= 0 && num_keep <= nnz && (nnz == 0 || (perm && keep_flags)), "lds_remove_edges_apply: need 0 <= num_keep <= nnz, perm and keep_flags");
  LDS_CHECK_ARG(out != adj, "lds_remove_edges_apply: out must not alias adj");
  LDS_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)n * ld_out * sizeof(float), stream));
  if (nnz == 0) return LDS_OK;
  LDS_CHECK_CUDA(cudaMemsetAsync(keep_flags, 0, (size_t)nnz, stream));
  if (num_keep > 0) {
    keep_flags_kernel<<<(unsigned)ceil_div(num_keep, 256), 256, 0, stream>>>(perm, num_keep, nnz, keep_flags);
    LDS_CHECK_LAUNCH("keep_flags_kernel");
  }
  remove_edges_apply_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(adj, ld, n, triu, offsets, keep_flags, out, ld_out);
  LDS_CHECK_LAUNCH("remove_edges_apply_kernel");
  return LDS_OK;
}
