// lds_k1_sample.cu — theta layouts (a1/a2), projection, statistics (a14) and K1: the fused
// Bernoulli-sample / mirror / self-loop / degree / D^-1/2 pass (a2,a3,a5,a6,a7 of SURVEY.md §8).
//
// Reference semantics restated (paths relative to the reference repo):
//   theta_ij = clamp(probs[idx(min,max)], 0, 1)                      src/utils/graph.py:166-181
//   s_ij = [u_ij < theta_ij], upper-triangle draw wins, diag kept    src/models/sampling.py:68,76
//   A_tilde = s with diag := 1; deg_i = sum_j A_tilde_ij; r = 1/sqrt  src/utils/graph.py:123-153
// HBM-bound: reads theta once (4 B/elem, 128-bit loads), writes bf16 A_tilde once (2 B/elem).
#include "lds_common.cuh"
#include "lds_philox.cuh"

namespace lds {

// ------------------------------------------------------------------------------------------------
// theta layout kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t triu_index(int64_t i, int64_t j, int64_t n) {
  return i * n - (i * (i - 1)) / 2 + (j - i);          // row-major upper triangle incl. diagonal
}

__global__ void triu_to_full_kernel(const float* __restrict__ triu, float* __restrict__ full, int64_t ld, int n, int clamp01) {
  const int i = blockIdx.x;
  float* row = full + (int64_t)i * ld;
  for (int j = threadIdx.x; j < (int)ld; j += blockDim.x) {
    float v = 0.f;
    if (j < n) {
      v = (j >= i) ? triu[triu_index(i, j, n)] : triu[triu_index(j, i, n)];
      if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    }
    row[j] = v;
  }
}

__global__ void full_to_triu_kernel(const float* __restrict__ full, int64_t ld, float* __restrict__ triu, int n, int sym_sum) {
  const int i = blockIdx.x;
  const float* row = full + (int64_t)i * ld;
  float* out = triu + triu_index(i, i, n);
  for (int j = i + threadIdx.x; j < n; j += blockDim.x) {
    float v = row[j];
    if (sym_sum && j != i) v += full[(int64_t)j * ld + i];
    out[j - i] = v;
  }
}

__global__ void clamp_kernel(float* __restrict__ full, int64_t ld, int n) {
  const int i = blockIdx.x;
  float4* row = reinterpret_cast<float4*>(full + (int64_t)i * ld);
  for (int q = threadIdx.x; q < (int)(ld / 4); q += blockDim.x) {
    float4 v = row[q];
    v.x = fminf(fmaxf(v.x, 0.f), 1.f); v.y = fminf(fmaxf(v.y, 0.f), 1.f);
    v.z = fminf(fmaxf(v.z, 0.f), 1.f); v.w = fminf(fmaxf(v.w, 0.f), 1.f);
    row[q] = v;
  }
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do { assumed = old; if (__longlong_as_double(assumed) <= v) break;
       old = atomicCAS(a, assumed, __double_as_longlong(v)); } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do { assumed = old; if (__longlong_as_double(assumed) >= v) break;
       old = atomicCAS(a, assumed, __double_as_longlong(v)); } while (assumed != old);
}

__global__ void stats_init_kernel(double* out4) {
  out4[0] = 0.0; out4[1] = 0.0; out4[2] = 1e300; out4[3] = -1e300;
}

__global__ void stats_kernel(const float* __restrict__ full, int64_t ld, int n, double* out4) {
  const int i = blockIdx.x;
  const float* row = full + (int64_t)i * ld;
  double s_full = 0.0, s_triu = 0.0; float mn = 3.4e38f, mx = -3.4e38f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float v = row[j];
    s_full += (double)fminf(fmaxf(v, 0.f), 1.f);
    if (j >= i) { s_triu += (double)v; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
  __shared__ double sh[4][32];
  for (int o = 16; o > 0; o >>= 1) {
    s_full += __shfl_xor_sync(0xffffffffu, s_full, o);
    s_triu += __shfl_xor_sync(0xffffffffu, s_triu, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) { sh[0][w] = s_full; sh[1][w] = s_triu; sh[2][w] = mn; sh[3][w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 1e300, d = -1e300;
    for (int k = 0; k < nw; ++k) { a += sh[0][k]; b += sh[1][k]; c = fmin(c, sh[2][k]); d = fmax(d, sh[3][k]); }
    atomicAdd(&out4[0], a); atomicAdd(&out4[1], b); atomic_min_double(&out4[2], c); atomic_max_double(&out4[3], d);
  }
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
constexpr int K1_THREADS = 256;
constexpr int K1_UNROLL = 4;      // column quads per thread per trip: 8 x 16-byte loads in flight per thread

// Fast path of a quad: all four columns in range, no diagonal element, Philox mode, only A_tilde wanted.
// `u < clamp(theta, 0, 1)` with u = (w >> 8) * 2^-24 is evaluated exactly in integers: (w >> 8) < ceil(theta * 2^24)
// (the saturating round-up conversion gives 0 for theta <= 0 or NaN and > 2^24 for theta > 1, which is the clamp).
__device__ __forceinline__ void k1_quad_fast(int j0, const float (&th0)[4], const float (&th1)[4], int p, const PhiloxRounds& R,
                                             uint32_t c2, uint32_t c3, __nv_bfloat16* __restrict__ a0, __nv_bfloat16* __restrict__ a1,
                                             uint32_t& cnt0, uint32_t& cnt1) {
  uint32_t b0[4], b1[4];
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const int q = (j0 >> 1) + b;
    const bool upper = p < q;                       // p == q cannot happen here (no diagonal element in the quad)
    uint32_t w[4];
    philox4x32_10_rk((uint32_t)(upper ? q : p), (uint32_t)(upper ? p : q), R, c2, c3, w);
    const uint32_t w01 = upper ? w[1] : w[2], w10 = upper ? w[2] : w[1];
    b0[2 * b]     = ((w[0] >> 8) < __float2uint_ru(th0[2 * b] * 16777216.f)) ? 1u : 0u;
    b0[2 * b + 1] = ((w01 >> 8)  < __float2uint_ru(th0[2 * b + 1] * 16777216.f)) ? 1u : 0u;
    b1[2 * b]     = ((w10 >> 8)  < __float2uint_ru(th1[2 * b] * 16777216.f)) ? 1u : 0u;
    b1[2 * b + 1] = ((w[3] >> 8) < __float2uint_ru(th1[2 * b + 1] * 16777216.f)) ? 1u : 0u;
  }
  cnt0 += b0[0] + b0[1] + b0[2] + b0[3];
  cnt1 += b1[0] + b1[1] + b1[2] + b1[3];
  uint2 r0, r1;                                       // bf16 1.0 = 0x3F80
  r0.x = b0[0] * 0x3F80u + b0[1] * 0x3F800000u; r0.y = b0[2] * 0x3F80u + b0[3] * 0x3F800000u;
  r1.x = b1[0] * 0x3F80u + b1[1] * 0x3F800000u; r1.y = b1[2] * 0x3F80u + b1[3] * 0x3F800000u;
  *reinterpret_cast<uint2*>(a0) = r0;
  *reinterpret_cast<uint2*>(a1) = r1;
}

// Processes one column quad (4 columns) of the row pair (2p, 2p+1): draws, compares, self loops, A_tilde store.
template <bool EXPLICIT_U>
__device__ __forceinline__ void k1_quad(int j0, const float (&th0)[4], const float (&th1)[4], int p, int gi0, int gi1, int li0, int li1,
                                        bool has1, int n, const PhiloxKey& key, const float* __restrict__ U, int64_t ldu,
                                        __nv_bfloat16* __restrict__ A, int64_t lda, float* __restrict__ S, int64_t lds_,
                                        float& sum0, float& sum1) {
  float u0[4], u1[4];
  if (EXPLICIT_U) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0 + e;
      u0[e] = 2.f; u1[e] = 2.f;
      if (j < n) {
        u0[e] = (gi0 <= j) ? U[(int64_t)gi0 * ldu + j] : U[(int64_t)j * ldu + gi0];
        if (has1) u1[e] = (gi1 <= j) ? U[(int64_t)gi1 * ldu + j] : U[(int64_t)j * ldu + gi1];
      }
    }
  } else {
#pragma unroll
    for (int b = 0; b < 2; ++b) {                     // two 2x2 blocks: column blocks q = j0/2 + b
      const int q = (j0 >> 1) + b;
      const bool upper = p <= q;                      // canonical block = (min, max): counter (c0 = max block, c1 = min block)
      uint32_t w[4];
      philox4x32_10((uint32_t)(upper ? q : p), (uint32_t)(upper ? p : q), key, w);
      // word = 2*(a%2) + (b%2) of the canonical pair (a, b) = (min, max): [di][dj] for rows (2p, 2p+1) x cols (2q, 2q+1)
      const uint32_t w01 = (p < q) ? w[1] : ((p > q) ? w[2] : w[1]);
      const uint32_t w10 = (p < q) ? w[2] : ((p > q) ? w[1] : w[1]);
      u0[2 * b] = philox_to_uniform(w[0]); u0[2 * b + 1] = philox_to_uniform(w01);
      u1[2 * b] = philox_to_uniform(w10);  u1[2 * b + 1] = philox_to_uniform(w[3]);
    }
  }
  float s0[4], s1[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const bool in = (j0 + e) < n;
    s0[e] = (in && u0[e] < fminf(fmaxf(th0[e], 0.f), 1.f)) ? 1.f : 0.f;
    s1[e] = (in && has1 && u1[e] < fminf(fmaxf(th1[e], 0.f), 1.f)) ? 1.f : 0.f;
  }
  if (S != nullptr) {                               // raw sample incl. the sampled diagonal
#pragma unroll
    for (int e = 0; e < 4; ++e) if (j0 + e < n) {
      S[(int64_t)li0 * lds_ + j0 + e] = s0[e];
      if (has1) S[(int64_t)li1 * lds_ + j0 + e] = s1[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {                     // self loops: diag := 1 (src/utils/graph.py:131-132)
    if (j0 + e == gi0) s0[e] = 1.f;
    if (j0 + e == gi1 && has1) s1[e] = 1.f;
    sum0 += s0[e]; sum1 += s1[e];
  }
  if (A != nullptr) {
    __nv_bfloat162 a01 = __floats2bfloat162_rn(s0[0], s0[1]), a23 = __floats2bfloat162_rn(s0[2], s0[3]);
    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&a01); pk.y = *reinterpret_cast<uint32_t*>(&a23);
    *reinterpret_cast<uint2*>(A + (int64_t)li0 * lda + j0) = pk;
    if (has1) {
      __nv_bfloat162 b01 = __floats2bfloat162_rn(s1[0], s1[1]), b23 = __floats2bfloat162_rn(s1[2], s1[3]);
      uint2 pk1; pk1.x = *reinterpret_cast<uint32_t*>(&b01); pk1.y = *reinterpret_cast<uint32_t*>(&b23);
      *reinterpret_cast<uint2*>(A + (int64_t)li1 * lda + j0) = pk1;
    }
  }
}

// One CTA per global row pair (2p, 2p+1). A thread owns column quads; each quad = two 2x2 Philox blocks.
// All of a trip's 128-bit loads are issued before the first is consumed: the kernel is bound by memory-level
// parallelism (each quad costs ~200 ALU instructions after its load), not by issue slots.
// STEP_DEV: the Philox step is *step_base + step_offset, read here from device memory (a captured CUDA graph then draws a
// fresh graph on every replay); otherwise the host-prepared counter words are used and the code is unchanged.
template <bool EXPLICIT_U, bool STEP_DEV>
__global__ void __launch_bounds__(K1_THREADS, 3)
k1_sample_kernel(const float* __restrict__ theta, int64_t ldt, int n, int row0, int rows,
                 const PhiloxKey key_in, const __grid_constant__ PhiloxRounds rounds, const float* __restrict__ U, int64_t ldu,
                 __nv_bfloat16* __restrict__ A, int64_t lda, float* __restrict__ S, int64_t lds_,
                 float* __restrict__ deg, float* __restrict__ rs,
                 const unsigned long long* __restrict__ step_base, unsigned long long step_offset) {
  PhiloxKey key = key_in;
  uint32_t c2 = rounds.c2, c3 = rounds.c3;
  if (STEP_DEV) {
    const unsigned long long step = *step_base + step_offset;
    c2 = (uint32_t)(step & 0xffffffffull);
    c3 = (c3 & 0xffffu) | (uint32_t)(((step >> 32) & 0xffffull) << 16);
    key.c2 = c2; key.c3 = c3;
  }
  const int p = (row0 >> 1) + blockIdx.x;             // global row-pair index
  const int gi0 = 2 * p, gi1 = 2 * p + 1;             // global rows
  const int li0 = gi0 - row0, li1 = gi1 - row0;       // local rows in this shard
  const bool has1 = (gi1 < n) && (li1 < rows);
  const float* t0 = theta + (int64_t)li0 * ldt;
  const float* t1 = theta + (int64_t)li1 * ldt;
  const int ncols = (A != nullptr) ? (int)lda : ((n + 3) & ~3);      // cover A's padding so it is zeroed
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t cnt0 = 0, cnt1 = 0;

  for (int base = 4 * threadIdx.x; base < ncols; base += 4 * K1_THREADS * K1_UNROLL) {
    float th0[K1_UNROLL][4], th1[K1_UNROLL][4];
#pragma unroll
    for (int u = 0; u < K1_UNROLL; ++u) {
      const int j0 = base + u * 4 * K1_THREADS;
#pragma unroll
      for (int e = 0; e < 4; ++e) { th0[u][e] = 0.f; th1[u][e] = 0.f; }
      if (j0 + 3 < n) {
        const float4 v0 = *reinterpret_cast<const float4*>(t0 + j0);
        th0[u][0] = v0.x; th0[u][1] = v0.y; th0[u][2] = v0.z; th0[u][3] = v0.w;
        if (has1) { const float4 v1 = *reinterpret_cast<const float4*>(t1 + j0);
                    th1[u][0] = v1.x; th1[u][1] = v1.y; th1[u][2] = v1.z; th1[u][3] = v1.w; }
      } else if (j0 < n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (j0 + e < n) { th0[u][e] = t0[j0 + e]; if (has1) th1[u][e] = t1[j0 + e]; }
      }
    }
#pragma unroll
    for (int u = 0; u < K1_UNROLL; ++u) {
      const int j0 = base + u * 4 * K1_THREADS;
      if (j0 >= ncols) continue;
      const bool fast = !EXPLICIT_U && has1 && S == nullptr && A != nullptr && (j0 + 3 < n) && (gi1 < j0 || gi0 > j0 + 3);
      if (fast) k1_quad_fast(j0, th0[u], th1[u], p, rounds, c2, c3, A + (int64_t)li0 * lda + j0, A + (int64_t)li1 * lda + j0, cnt0, cnt1);
      else k1_quad<EXPLICIT_U>(j0, th0[u], th1[u], p, gi0, gi1, li0, li1, has1, n, key, U, ldu, A, lda, S, lds_, sum0, sum1);
    }
  }
  // block reduction of the two row sums (integers in fp32: exact, order-independent for N < 2^24)
  __shared__ float sh0[K1_THREADS / 32], sh1[K1_THREADS / 32];
  sum0 += (float)cnt0; sum1 += (float)cnt1;          // per-thread counts are far below 2^24: exact
  sum0 = warp_sum(sum0); sum1 = warp_sum(sum1);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh0[w] = sum0; sh1[w] = sum1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int k = 0; k < K1_THREADS / 32; ++k) { d0 += sh0[k]; d1 += sh1[k]; }
    if (li0 < rows) { deg[li0] = d0; rs[li0] = 1.0f / sqrtf(d0); }
    if (has1)       { deg[li1] = d1; rs[li1] = 1.0f / sqrtf(d1); }
  }
}

}  // namespace lds

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace lds;

extern "C" int32_t lds_theta_triu_to_full(const float* triu, float* theta_full, int64_t ld, int32_t n, int32_t clamp01, void* stream) {
  LDS_CHECK_ARG(triu && theta_full, "lds_theta_triu_to_full: null pointer");
  LDS_CHECK_ARG(n > 0 && ld >= n && ld % 4 == 0, "lds_theta_triu_to_full: need n > 0, ld >= n, ld %% 4 == 0 (n=%d ld=%lld)", n, (long long)ld);
  triu_to_full_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(triu, theta_full, ld, n, clamp01);
  LDS_CHECK_LAUNCH("triu_to_full_kernel");
  return LDS_OK;
}

extern "C" int32_t lds_theta_full_to_triu(const float* full, int64_t ld, float* triu, int32_t n, int32_t sym_sum, void* stream) {
  LDS_CHECK_ARG(triu && full, "lds_theta_full_to_triu: null pointer");
  LDS_CHECK_ARG(n > 0 && ld >= n, "lds_theta_full_to_triu: need n > 0, ld >= n");
  full_to_triu_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(full, ld, triu, n, sym_sum);
  LDS_CHECK_LAUNCH("full_to_triu_kernel");
  return LDS_OK;
}

extern "C" int32_t lds_theta_clamp(float* theta_full, int64_t ld, int32_t n, void* stream) {
  LDS_CHECK_ARG(theta_full, "lds_theta_clamp: null pointer");
  LDS_CHECK_ARG(n > 0 && ld >= n && ld % 4 == 0, "lds_theta_clamp: need n > 0, ld >= n, ld %% 4 == 0");
  clamp_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(theta_full, ld, n);
  LDS_CHECK_LAUNCH("clamp_kernel");
  return LDS_OK;
}

extern "C" int32_t lds_theta_stats(const float* theta_full, int64_t ld, int32_t n, double* out4, void* stream) {
  LDS_CHECK_ARG(theta_full && out4, "lds_theta_stats: null pointer");
  LDS_CHECK_ARG(n > 0 && ld >= n, "lds_theta_stats: need n > 0, ld >= n");
  stats_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out4);
  stats_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(theta_full, ld, n, out4);
  LDS_CHECK_LAUNCH("stats_kernel");
  return LDS_OK;
}

static int32_t k1_launch(const float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                         uint64_t seed, uint64_t step, const uint64_t* step_base, uint32_t sample,
                         const float* u_explicit, int64_t ld_u,
                         void* a_out, int64_t ld_a, float* sample_out, int64_t ld_s,
                         float* deg_out, float* rsqrt_out, uint32_t flags, void* stream) {
  using namespace lds;
  LDS_CHECK_ARG(theta_full && deg_out && rsqrt_out, "lds_k1_sample_normalize: null pointer");
  LDS_CHECK_ARG(n > 0 && rows > 0 && row0 >= 0 && row0 + rows <= n, "lds_k1_sample_normalize: rows [%d, %d) outside [0, %d)", row0, row0 + rows, n);
  LDS_CHECK_ARG((row0 & 1) == 0, "lds_k1_sample_normalize: row0 must be even (got %d)", row0);
  LDS_CHECK_ARG(ld_theta >= n && ld_theta % 4 == 0, "lds_k1_sample_normalize: ld_theta must be >= n and a multiple of 4");
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(theta_full) & 15) == 0, "lds_k1_sample_normalize: theta must be 16-byte aligned");
  if (a_out) {
    LDS_CHECK_ARG(ld_a >= n && ld_a % 8 == 0, "lds_k1_sample_normalize: ld_a must be >= n and a multiple of 8 (TMA stride)");
    LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(a_out) & 15) == 0, "lds_k1_sample_normalize: a_out must be 16-byte aligned");
  }
  if (sample_out) LDS_CHECK_ARG(ld_s >= n, "lds_k1_sample_normalize: ld_s must be >= n");
  const bool explicit_u = (flags & LDS_K1_EXPLICIT_U) != 0;
  if (explicit_u) LDS_CHECK_ARG(u_explicit && ld_u >= n, "lds_k1_sample_normalize: LDS_K1_EXPLICIT_U needs u_explicit with ld_u >= n");
  const PhiloxKey key = philox_key(seed, step_base ? 0 : step, LDS_STREAM_EDGES, sample);
  const PhiloxRounds rounds = philox_rounds(key);
  const int pairs = (rows + 1) / 2;
  auto* A = reinterpret_cast<__nv_bfloat16*>(a_out);
  auto* sb = reinterpret_cast<const unsigned long long*>(step_base);
  if (explicit_u)
    k1_sample_kernel<true, false><<<pairs, K1_THREADS, 0, (cudaStream_t)stream>>>(theta_full, ld_theta, n, row0, rows, key, rounds, u_explicit, ld_u, A, ld_a, sample_out, ld_s, deg_out, rsqrt_out, nullptr, 0ull);
  else if (step_base)
    k1_sample_kernel<false, true><<<pairs, K1_THREADS, 0, (cudaStream_t)stream>>>(theta_full, ld_theta, n, row0, rows, key, rounds, nullptr, 0, A, ld_a, sample_out, ld_s, deg_out, rsqrt_out, sb, (unsigned long long)step);
  else
    k1_sample_kernel<false, false><<<pairs, K1_THREADS, 0, (cudaStream_t)stream>>>(theta_full, ld_theta, n, row0, rows, key, rounds, nullptr, 0, A, ld_a, sample_out, ld_s, deg_out, rsqrt_out, nullptr, 0ull);
  LDS_CHECK_LAUNCH("k1_sample_kernel");
  return LDS_OK;
}

extern "C" int32_t lds_k1_sample_normalize(const float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                           uint64_t seed, uint64_t step, uint32_t sample,
                                           const float* u_explicit, int64_t ld_u,
                                           void* a_out, int64_t ld_a, float* sample_out, int64_t ld_s,
                                           float* deg_out, float* rsqrt_out, uint32_t flags, void* stream) {
  return k1_launch(theta_full, ld_theta, n, row0, rows, seed, step, nullptr, sample, u_explicit, ld_u, a_out, ld_a, sample_out, ld_s,
                   deg_out, rsqrt_out, flags, stream);
}

extern "C" int32_t lds_k1_sample_normalize_dstep(const float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                                 uint64_t seed, const uint64_t* step_base, uint64_t step_offset, uint32_t sample,
                                                 void* a_out, int64_t ld_a, float* deg_out, float* rsqrt_out, void* stream) {
  LDS_CHECK_ARG(step_base, "lds_k1_sample_normalize_dstep: null step_base");
  return k1_launch(theta_full, ld_theta, n, row0, rows, seed, step_offset, step_base, sample, nullptr, 0, a_out, ld_a, nullptr, 0,
                   deg_out, rsqrt_out, 0u, stream);
}
