// lds_k3_theta_update.cu — K3+K4: closed-form straight-through hypergradient on theta fused with the
// projected optimiser step (a10's theta part + a11 of SURVEY.md §8).
//
// The reference reaches dL/dtheta by autograd through ~25 N x N fp32 temporaries and four N^3 SGEMMs
// (src/trainers/outer.py:77 through src/utils/graph.py:136-181, src/models/sampling.py:76-85), then
// SGD.step + clamp (src/models/factory.py:66-69, src/trainers/outer.py:78-83, src/models/graph.py:16-20).
// In closed form (checked against the reference's autograd in oracle/):
//     dL/dtheta_ij = fa_i.fb_j + fb_i.fa_j + c_i + c_j   (i != j),   0 on the diagonal,
//     fa = r * (dZ1 | dZ2),  fb = r * (P1 | P2),  c_i = -(rho_i + kappa_i) / (2 deg_i)
// so one pass reads theta, forms the rank-2(h+C) term from two skinny factor matrices held in shared
// memory, applies theta <- clamp(theta - lr g, 0, 1) and writes theta back: 8 B/element of HBM traffic.
//
// Symmetry is exact by construction: element (i,j) computes t1 = sum_k fa_i[k] fb_j[k], t2 = sum_k fb_i[k] fa_j[k]
// with sequential FMAs in k order; element (j,i) computes the same two chains swapped, and fp32 add/mul/fma are
// commutative in their multiplicands, so theta_ij == theta_ji bit for bit (the row-block shards of a multi-GPU
// run therefore stay consistent without any exchange).
#include "lds_common.cuh"

namespace lds {

constexpr int K3_TILE = 64;
constexpr int K3_THREADS = 256;
constexpr int K3_DCHUNK = 32;
constexpr int K3_PAD = K3_TILE + 4;

template <bool DENSE_GRAD>
__global__ void __launch_bounds__(K3_THREADS)
k3_update_kernel(float* __restrict__ theta, int64_t ldt, int n, int row0, int rows,
                 const float* __restrict__ fa, const float* __restrict__ fb, int64_t ldf, int d,
                 const float* __restrict__ cvec, float lr, int opt_kind,
                 float* __restrict__ adam_m, float* __restrict__ adam_v, float beta1, float beta2, float eps,
                 float bc1, float bc2, float* __restrict__ grad_out, int64_t ldg, int accumulate) {
  __shared__ __align__(16) float sAi[K3_DCHUNK][K3_PAD];
  __shared__ __align__(16) float sBi[K3_DCHUNK][K3_PAD];
  __shared__ __align__(16) float sAj[K3_DCHUNK][K3_PAD];
  __shared__ __align__(16) float sBj[K3_DCHUNK][K3_PAD];

  const int j0 = blockIdx.x * K3_TILE;
  const int li0 = blockIdx.y * K3_TILE;            // local row of the tile
  const int gi0 = row0 + li0;                      // global row
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float t1[4][4], t2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { t1[a][b] = 0.f; t2[a][b] = 0.f; }

  for (int k0 = 0; k0 < d; k0 += K3_DCHUNK) {
    // cooperative transposed load of the factor chunks: [row][k] in global -> [k][row] in smem
    for (int idx = threadIdx.x; idx < K3_TILE * K3_DCHUNK; idx += K3_THREADS) {
      const int r = idx / K3_DCHUNK, kk = idx - r * K3_DCHUNK;
      const int k = k0 + kk;
      const int gi = gi0 + r, gj = j0 + r;
      const bool ki = (k < d) && (gi < n), kj = (k < d) && (gj < n);
      sAi[kk][r] = ki ? fa[(int64_t)gi * ldf + k] : 0.f;
      sBj[kk][r] = kj ? fb[(int64_t)gj * ldf + k] : 0.f;
      if (!DENSE_GRAD) {
        sBi[kk][r] = ki ? fb[(int64_t)gi * ldf + k] : 0.f;
        sAj[kk][r] = kj ? fa[(int64_t)gj * ldf + k] : 0.f;
      }
    }
    __syncthreads();
    const int kmax = min(K3_DCHUNK, d - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      const float4 ai = *reinterpret_cast<const float4*>(&sAi[kk][4 * ty]);
      const float4 bj = *reinterpret_cast<const float4*>(&sBj[kk][4 * tx]);
      const float av[4] = {ai.x, ai.y, ai.z, ai.w}, bv[4] = {bj.x, bj.y, bj.z, bj.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) t1[a][b] = fmaf(av[a], bv[b], t1[a][b]);
      if (!DENSE_GRAD) {
        const float4 bi = *reinterpret_cast<const float4*>(&sBi[kk][4 * ty]);
        const float4 aj = *reinterpret_cast<const float4*>(&sAj[kk][4 * tx]);
        const float bw[4] = {bi.x, bi.y, bi.z, bi.w}, aw[4] = {aj.x, aj.y, aj.z, aj.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) t2[a][b] = fmaf(bw[a], aw[b], t2[a][b]);
      }
    }
    __syncthreads();
  }

  float cj[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) { const int gj = j0 + 4 * tx + b; cj[b] = (gj < n) ? cvec[gj] : 0.f; }

#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int li = li0 + 4 * ty + a, gi = row0 + li;
    if (li >= rows || gi >= n) continue;
    const float ci = cvec[gi];
    const int jb = j0 + 4 * tx;
    if (jb >= n) continue;
    if (DENSE_GRAD) {
      float* gp = grad_out + (int64_t)li * ldg + jb;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (jb + b >= n) break;
        const float g = (gi == jb + b) ? 0.f : (t1[a][b] + ci);
        gp[b] = accumulate ? gp[b] + g : g;
      }
      continue;
    }
    float* tp = theta + (int64_t)li * ldt + jb;
    const bool full4 = (jb + 3 < n);
    float th[4];
    if (full4) { const float4 v = *reinterpret_cast<const float4*>(tp); th[0] = v.x; th[1] = v.y; th[2] = v.z; th[3] = v.w; }
    else { for (int b = 0; b < 4; ++b) th[b] = (jb + b < n) ? tp[b] : 0.f; }
    float mo[4] = {0.f, 0.f, 0.f, 0.f}, vo[4] = {0.f, 0.f, 0.f, 0.f};
    float* mp = nullptr; float* vp = nullptr;
    if (opt_kind == LDS_OPT_ADAM) {
      mp = adam_m + (int64_t)li * ldt + jb; vp = adam_v + (int64_t)li * ldt + jb;
      for (int b = 0; b < 4; ++b) if (jb + b < n) { mo[b] = mp[b]; vo[b] = vp[b]; }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      float g = (t1[a][b] + t2[a][b]) + (ci + cj[b]);                    // symmetric in (i, j) bit for bit
      if (gi == jb + b) g = 0.f;                                         // fill_diagonal_ backward (utils/graph.py:131-132)
      if (th[b] < 0.f || th[b] > 1.f) g = 0.f;                           // clamp backward (utils/graph.py:180)
      float nt;
      if (opt_kind == LDS_OPT_ADAM) {
        mo[b] = beta1 * mo[b] + (1.f - beta1) * g;
        vo[b] = beta2 * vo[b] + (1.f - beta2) * g * g;
        nt = th[b] - lr * (mo[b] / bc1) / (sqrtf(vo[b] / bc2) + eps);
      } else {
        nt = fmaf(-lr, g, th[b]);                                        // SGD (models/factory.py:66-69)
      }
      th[b] = fminf(fmaxf(nt, 0.f), 1.f);                                // project_parameters (models/graph.py:16-20)
    }
    if (full4) *reinterpret_cast<float4*>(tp) = make_float4(th[0], th[1], th[2], th[3]);
    else { for (int b = 0; b < 4; ++b) if (jb + b < n) tp[b] = th[b]; }
    if (opt_kind == LDS_OPT_ADAM)
      for (int b = 0; b < 4; ++b) if (jb + b < n) { mp[b] = mo[b]; vp[b] = vo[b]; }
  }
}

}  // namespace lds

using namespace lds;

extern "C" int32_t lds_k3k4_theta_update(float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                         const float* fa, const float* fb, int64_t ld_f, int32_t d, const float* cvec,
                                         float lr, int32_t opt_kind, float* adam_m, float* adam_v,
                                         float beta1, float beta2, float eps, int32_t t,
                                         float* grad_out, int64_t ld_g, uint32_t flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(fa && fb && cvec, "lds_k3k4_theta_update: null factor pointer");
  LDS_CHECK_ARG(n > 0 && rows > 0 && row0 >= 0 && row0 + rows <= n, "lds_k3k4_theta_update: rows [%d, %d) outside [0, %d)", row0, row0 + rows, n);
  LDS_CHECK_ARG(d > 0 && d <= 4096 && ld_f >= d, "lds_k3k4_theta_update: need 0 < d <= 4096 and ld_f >= d");
  dim3 grid((unsigned)ceil_div(n, K3_TILE), (unsigned)ceil_div(rows, K3_TILE));
  if (flags & LDS_K3_DENSE_GRAD) {
    LDS_CHECK_ARG(grad_out && ld_g >= n, "lds_k3k4_theta_update: LDS_K3_DENSE_GRAD needs grad_out with ld_g >= n");
    k3_update_kernel<true><<<grid, K3_THREADS, 0, stream>>>(nullptr, 0, n, row0, rows, fa, fb, ld_f, d, cvec, 0.f, 0, nullptr, nullptr,
                                                            0.f, 0.f, 0.f, 1.f, 1.f, grad_out, ld_g, (flags & LDS_K3_ACCUMULATE) ? 1 : 0);
    LDS_CHECK_LAUNCH("k3_update_kernel<dense>");
    return LDS_OK;
  }
  LDS_CHECK_ARG(theta_full && ld_theta >= n && ld_theta % 4 == 0 && (reinterpret_cast<uintptr_t>(theta_full) & 15) == 0,
                "lds_k3k4_theta_update: theta must be 16-byte aligned with ld_theta >= n, ld_theta %% 4 == 0");
  LDS_CHECK_ARG(opt_kind == LDS_OPT_SGD || opt_kind == LDS_OPT_ADAM, "lds_k3k4_theta_update: unknown optimiser kind %d", opt_kind);
  float bc1 = 1.f, bc2 = 1.f;
  if (opt_kind == LDS_OPT_ADAM) {
    LDS_CHECK_ARG(adam_m && adam_v && t >= 1, "lds_k3k4_theta_update: Adam needs m, v and t >= 1");
    bc1 = (float)(1.0 - pow((double)beta1, (double)t));
    bc2 = (float)(1.0 - pow((double)beta2, (double)t));
  }
  k3_update_kernel<false><<<grid, K3_THREADS, 0, stream>>>(theta_full, ld_theta, n, row0, rows, fa, fb, ld_f, d, cvec, lr, opt_kind,
                                                           adam_m, adam_v, beta1, beta2, eps, bc1, bc2, nullptr, 0, 0);
  LDS_CHECK_LAUNCH("k3_update_kernel");
  return LDS_OK;
}
