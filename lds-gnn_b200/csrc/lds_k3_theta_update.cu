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
#include <stdlib.h>
#include "lds_k3.cuh"
#include "lds_tc.cuh"

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

// ================================================================================================
// Tensor-core variant (SGD): the rank-2d term is a GEMM, so at large d (h = 64 => d = 71, 142 FMA per element
// on CUDA cores) it must run on tcgen05 or the update stops being an HBM stream. Operands: the bf16 hi/lo split
// of the factor rows, packed K-major and INTERLEAVED per 16-column step so that one 128-byte swizzle atom holds
// everything a k-step needs:
//     F[i][64 q + 16 t + k] = block_t(i)[16 q + k],   blocks t = 0 a_hi, 1 a_lo, 2 b_hi, 3 b_lo   (a = fa, b = fb)
//     ([n][kf] bf16, kf = 4 * round_up(d, 16); 8 B per factor element instead of the 12 B of a K-concatenated
//      [hi|hi|lo] x [hi|lo|hi] packing — the operand re-reads from L2, not HBM, were the bound at d = 71)
// Per k-step q the issuer runs six 128x128x16 MMAs on the two boxes F_i[q], F_j[q]:
//     D1 += a_hi b_hi' + a_hi b_lo' + a_lo b_hi'        D2 += b_hi a_hi' + b_lo a_hi' + b_hi a_lo'
// D1(i,j) and D2(j,i) are the same exact bf16 products accumulated in the same order, so g_ij == g_ji bit for
// bit (fp32 addition is commutative) — asserted by the tests; row-block shards stay consistent with no exchange.
// One persistent CTA per SM walks a contiguous range of 128 x 128 tiles:
//   warp 0    TMA producer of the operand boxes (128 x 64 bf16, SWIZZLE_128B), K3T_OSTAGES stages of (F_i[q], F_j[q])
//   warp 1    tcgen05.mma issuer, two double-buffered accumulator pairs in TMEM (512 columns)
//   warps 2-9 epilogue: tcgen05.ld D1/D2, theta from smem, g = (D1 + D2) + (c_i + c_j), SGD step + clamp written
//             back into the same smem slab, then one TMA store per slab (coalesced, asynchronous)
//   warp 10   TMA producer of the theta slabs (128 rows x 32 fp32, SWIZZLE_128B) into a K3T_RING-slab ring. The ring
//             depth is what hides HBM latency: measured at N = 20 000, 4 / 6 / 8 slabs in flight = 875 / 705 / 630 us.
// ================================================================================================
constexpr int K3T_TILE = 128;
constexpr int K3T_KB = 64;                                   // packed columns per k-step = one 128-byte swizzle atom
constexpr int K3T_SLAB_COLS = 32;
constexpr int K3T_RING = 8;                                  // theta slabs in flight (2 tiles)
constexpr int K3T_OSTAGES = 3;                               // operand stages
constexpr int K3T_OP_BYTES = K3T_TILE * K3T_KB * 2;          // 16 KB
constexpr int K3T_STAGE_BYTES = 2 * K3T_OP_BYTES;            // F_i[q], F_j[q]
constexpr int K3T_SLAB_BYTES = K3T_TILE * K3T_SLAB_COLS * 4; // 16 KB
constexpr int K3T_THREADS = 352;
constexpr int K3T_SMEM = K3T_OSTAGES * K3T_STAGE_BYTES + K3T_RING * K3T_SLAB_BYTES + 1024 + 512;

// One thread's row of one theta slab (32 columns, 128-byte swizzled in smem): g = (D1 + D2) + (c_i + c_j) with the
// diagonal and clamp-backward masks, SGD step + projection written back in place (src/models/factory.py:66-69,
// src/trainers/outer.py:78-83, src/models/graph.py:16-20).
// `cj` = c_j of the slab's 32 columns (zero past n), staged through shared memory once per tile: the 224 KB carve-out
// leaves almost no L1, and a cvec load issued per cell cost an L2 round trip each (45 % of the stall samples at N = 20 000).
// MIRROR (tile-symmetric launch, tile strictly above the diagonal): the new values also go to the transposed tile,
// theta[jb + c][gi] — for a fixed column c the 32 lanes of a warp hold 32 consecutive rows gi, i.e. one full 128-byte
// line of the mirrored row: plain coalesced stores, no staging buffer (tdst = &theta[jb][gi], row stride ldt).
template <bool MIRROR>
__device__ __forceinline__ void k3_update_slab_row(uint8_t* slab, int row, const uint32_t (&d1)[32], const uint32_t (&d2)[32],
                                                   const float* cj_s, bool interior, float ci, int gi, int jb, int n, float lr,
                                                   float* __restrict__ tdst = nullptr, int64_t ldt = 0) {
  if (interior) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      float4* cell = reinterpret_cast<float4*>(slab + ((c4 ^ (row & 7)) << 4));
      const float4 th = *cell;
      const float4 cv = *reinterpret_cast<const float4*>(cj_s + 4 * c4);
      const float tv[4] = {th.x, th.y, th.z, th.w}, cj[4] = {cv.x, cv.y, cv.z, cv.w};
      float nv[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float g = (__uint_as_float(d1[4 * c4 + b]) + __uint_as_float(d2[4 * c4 + b])) + (ci + cj[b]);
        if (tv[b] < 0.f || tv[b] > 1.f) g = 0.f;                 // clamp backward
        nv[b] = fminf(fmaxf(fmaf(-lr, g, tv[b]), 0.f), 1.f);
        if (MIRROR) __stcs(tdst + (int64_t)(4 * c4 + b) * ldt, nv[b]);
      }
      *cell = make_float4(nv[0], nv[1], nv[2], nv[3]);
    }
  } else {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      float4* cell = reinterpret_cast<float4*>(slab + ((c4 ^ (row & 7)) << 4));
      float4 th = *cell;
      const int gj = jb + 4 * c4;
      const float4 cv = *reinterpret_cast<const float4*>(cj_s + 4 * c4);
      const float cj[4] = {cv.x, cv.y, cv.z, cv.w};
      float tv[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float g = (__uint_as_float(d1[4 * c4 + b]) + __uint_as_float(d2[4 * c4 + b])) + (ci + cj[b]);
        if (gi == gj + b) g = 0.f;
        if (tv[b] < 0.f || tv[b] > 1.f) g = 0.f;
        const float nv = fminf(fmaxf(fmaf(-lr, g, tv[b]), 0.f), 1.f);
        if (gj + b < n) tv[b] = nv;               // the TMA store clips at 16-byte granularity: leave padding as loaded
        if (MIRROR && gj + b < n) __stcs(tdst + (int64_t)(4 * c4 + b) * ldt, nv);
      }
      *cell = make_float4(tv[0], tv[1], tv[2], tv[3]);
    }
  }
}

// Tile (bi, bj) of linear index t, row by row. SYM: the upper triangle incl. the diagonal (bj >= bi).
// (Tried: bands of 32 tile rows walked column by column, so that every CTA re-reads a few MB of F_i and uses each F_j block
// 32 times in a row — ncu at N = 65 536 shows 20.5 GB of DRAM reads for 17.2 GB of theta, i.e. factor rows missing L2.
// Measured: 5.84 -> 5.76 ms tile-symmetric, 7.39 -> 7.62 ms full at N = 65 536, no change at N = 20 000: not kept.
// Also tried for the same reason: F pinned in the persisting part of L2 through an access-policy window on the stream
// (hitProp persisting, missProp streaming) for the duration of the launch — 5.8 -> 12.4 ms and 7.5 -> 17.1 ms: not kept.)
template <bool SYM>
__device__ __forceinline__ void k3_tile_of(int t, int tiles_j, int& bi, int& bj) {
  if (!SYM) { bi = t / tiles_j; bj = t - bi * tiles_j; return; }
  const float bb = 2.0f * tiles_j + 1.0f;
  bi = (int)((bb - sqrtf(bb * bb - 8.0f * (float)t)) * 0.5f);              // float estimate, then exact correction
  bi = max(0, min(bi, tiles_j - 1));
  while (bi > 0 && bi * tiles_j - bi * (bi - 1) / 2 > t) --bi;
  while (bi + 1 < tiles_j && (bi + 1) * tiles_j - (bi + 1) * bi / 2 <= t) ++bi;
  bj = bi + (t - (bi * tiles_j - bi * (bi - 1) / 2));
}

// SYM (unsharded update of a SYMMETRIC theta — the expansion of the reference's (T,) `probs`): only the tiles on and above
// the diagonal are read and computed; a tile strictly above it is stored twice, as it is and transposed (k3_update_slab_row),
// so the launch moves 6 N^2 bytes (2 read + 4 written) instead of 8 N^2. theta_ji is then a copy of theta_ij — the same bits
// the full launch computes for it (D1(i,j) and D2(j,i) are the same products in the same order).
template <bool SYM>
__global__ void __launch_bounds__(K3T_THREADS, 1)
k3_tc_kernel(const __grid_constant__ CUtensorMap tm_theta, const __grid_constant__ CUtensorMap tm_f,
             const float* __restrict__ cvec, int n, int row0, int rows, int ksteps, float lr, float* __restrict__ theta, int64_t ldt) {
  __shared__ __align__(16) float cj_stage[2][64];            // c_j of the tile's columns, one half per epilogue warp group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ops = smem;                                        // operand stages
  uint8_t* slabs = smem + K3T_OSTAGES * K3T_STAGE_BYTES;      // theta ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabs + K3T_RING * K3T_SLAB_BYTES);
  uint64_t* ofull = bars;                      // [OSTAGES]
  uint64_t* oempty = bars + K3T_OSTAGES;       // [OSTAGES]
  uint64_t* tfull = bars + 2 * K3T_OSTAGES;    // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint64_t* thfull = tempty + 2;               // [RING]
  uint64_t* thempty = thfull + K3T_RING;       // [RING]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(thempty + K3T_RING);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_j = (n + K3T_TILE - 1) / K3T_TILE;
  const int tiles_i = (rows + K3T_TILE - 1) / K3T_TILE;
  const int num_tiles = SYM ? tiles_j * (tiles_j + 1) / 2 : tiles_i * tiles_j;
  const int t_lo = (int)(((int64_t)blockIdx.x * num_tiles) / gridDim.x);          // contiguous range: F_i stays hot in L2/L1
  const int t_hi = (int)(((int64_t)(blockIdx.x + 1) * num_tiles) / gridDim.x);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_theta); tma_prefetch_desc(&tm_f); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < K3T_OSTAGES; ++i) { mbar_init(&ofull[i], 1); mbar_init(&oempty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
      for (int i = 0; i < K3T_RING; ++i) { mbar_init(&thfull[i], 1); mbar_init(&thempty[i], 1); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (lds_outer_step after the fused small-graph kernel): everything above — and this kernel's first
  // instruction fetches — may run while the producer of F / c / theta is still in its last phase; nothing below may. Without
  // the launch attribute this returns at once.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer: operand boxes =====
    if (lane == 0) {
      int oq = 0;
      for (int t = t_lo; t < t_hi; ++t) {
        int bi, bj;
        k3_tile_of<SYM>(t, tiles_j, bi, bj);
        for (int q = 0; q < ksteps; ++q, ++oq) {
          const int os = oq % K3T_OSTAGES;
          uint8_t* st = ops + os * K3T_STAGE_BYTES;
          mbar_wait(&oempty[os], (uint32_t)(((oq / K3T_OSTAGES) & 1) ^ 1));
          mbar_expect_tx(&ofull[os], K3T_STAGE_BYTES);
          tma_load_2d(st, &tm_f, &ofull[os], q * K3T_KB, row0 + bi * K3T_TILE);          // F_i[q]
          tma_load_2d(st + K3T_OP_BYTES, &tm_f, &ofull[os], q * K3T_KB, bj * K3T_TILE);   // F_j[q]
        }
      }
    }
  } else if (warp == 10) {
    // ===== TMA producer: theta slabs (never blocked by an operand wait) =====
    if (lane == 0) {
      int q = 0;
      for (int t = t_lo; t < t_hi; ++t) {
        int bi, bj;
        k3_tile_of<SYM>(t, tiles_j, bi, bj);
        for (int s = 0; s < 4; ++s, ++q) {
          const int slot = q % K3T_RING;
          mbar_wait(&thempty[slot], (uint32_t)(((q / K3T_RING) & 1) ^ 1));
          mbar_expect_tx(&thfull[slot], K3T_SLAB_BYTES);
          tma_load_2d(slabs + slot * K3T_SLAB_BYTES, &tm_theta, &thfull[slot], bj * K3T_TILE + s * K3T_SLAB_COLS, bi * K3T_TILE);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(K3T_TILE, K3T_TILE);
      int oq = 0, tt = 0;
      for (int t = t_lo; t < t_hi; ++t, ++tt) {
        const int acc = tt & 1;
        mbar_wait(&tempty[acc], (uint32_t)(((tt >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d1 = tmem_base + (uint32_t)(acc * 256), d2 = d1 + 128;
        for (int q = 0; q < ksteps; ++q, ++oq) {
          const int os = oq % K3T_OSTAGES;
          mbar_wait(&ofull[os], (uint32_t)((oq / K3T_OSTAGES) & 1));
          tc_fence_after();
          const uint32_t base = smem_u32(ops + os * K3T_STAGE_BYTES);
          const uint64_t fi = umma_desc_k_sw128(base), fj = umma_desc_k_sw128(base + K3T_OP_BYTES);
          // 16-column block t of the atom sits at +32 t bytes = +2 t in the descriptor's address field
          const uint64_t i_ah = fi, i_al = fi + 2, i_bh = fi + 4, i_bl = fi + 6;
          const uint64_t j_ah = fj, j_al = fj + 2, j_bh = fj + 4, j_bl = fj + 6;
          const uint32_t accum = q != 0;
          tc_mma_bf16(d1, i_ah, j_bh, idesc, accum);
          tc_mma_bf16(d1, i_ah, j_bl, idesc, 1u);
          tc_mma_bf16(d1, i_al, j_bh, idesc, 1u);
          tc_mma_bf16(d2, i_bh, j_ah, idesc, accum);
          tc_mma_bf16(d2, i_bl, j_ah, idesc, 1u);
          tc_mma_bf16(d2, i_bh, j_al, idesc, 1u);
          tc_commit(&oempty[os]);
        }
        tc_commit(&tfull[acc]);
      }
    }
  } else {
    // ===== epilogue =====
    const int e = warp - 2;
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access (warp id mod 4)
    const int hf = e >> 2;                           // which pair of slabs (column half of the tile)
    const int row = quarter * 32 + lane;
    const bool storer = ((e & 3) == 0) && (lane == 0);
    int pending = -1;
    int tt = 0;
    for (int t = t_lo; t < t_hi; ++t, ++tt) {
      int bi, bj;
        k3_tile_of<SYM>(t, tiles_j, bi, bj);
      const int li0 = bi * K3T_TILE, gi0 = row0 + li0, j0 = bj * K3T_TILE;
      const int acc = tt & 1;
      const int gi = gi0 + row;
      const float ci = (gi < n) ? cvec[gi] : 0.f;
      const int gt = (e & 3) * 32 + lane;            // 0..127 inside this warp group
      const int cjcol = j0 + 64 * hf + gt;
      const float cj_mine = (gt < 64 && cjcol < n) ? cvec[cjcol] : 0.f;                // in flight across the accumulator wait
      const bool interior = (gi0 + K3T_TILE <= j0 || j0 + K3T_TILE <= gi0) && (j0 + K3T_TILE <= n);   // no diagonal element, every column in range
      mbar_wait(&tfull[acc], (uint32_t)((tt >> 1) & 1));
      tc_fence_after();
      if (gt < 64) cj_stage[hf][gt] = cj_mine;       // the group's reads of the previous tile ended before its last named barrier
      named_bar_sync(1 + hf, 128);
      for (int sl = 0; sl < 2; ++sl) {
        const int s = 2 * hf + sl;
        const int q = tt * 4 + s, slot = q % K3T_RING;
        uint32_t d1[32], d2[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + s * K3T_SLAB_COLS);
        {
          uint32_t (&lo1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&d1[0]);
          uint32_t (&hi1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&d1[16]);
          uint32_t (&lo2)[16] = *reinterpret_cast<uint32_t (*)[16]>(&d2[0]);
          uint32_t (&hi2)[16] = *reinterpret_cast<uint32_t (*)[16]>(&d2[16]);
          tc_ld16(taddr, lo1); tc_ld16(taddr + 16, hi1);
          tc_ld16(taddr + 128, lo2); tc_ld16(taddr + 128 + 16, hi2);
        }
        const int jb = j0 + s * K3T_SLAB_COLS;
        mbar_wait(&thfull[slot], (uint32_t)((q / K3T_RING) & 1));
        tc_wait_ld();
        if (SYM && bj > bi)
          k3_update_slab_row<true>(slabs + slot * K3T_SLAB_BYTES + row * 128, row, d1, d2, &cj_stage[hf][32 * sl], interior, ci, gi, jb, n, lr,
                                   theta + (int64_t)jb * ldt + gi, ldt);
        else
          k3_update_slab_row<false>(slabs + slot * K3T_SLAB_BYTES + row * 128, row, d1, d2, &cj_stage[hf][32 * sl], interior, ci, gi, jb, n, lr);
        fence_proxy_async_smem();                      // generic-proxy writes -> visible to the TMA store
        named_bar_sync(1 + hf, 128);                   // the four warps that own this slab
        if (storer) {
          if (pending >= 0) { tma_store_wait_read(); mbar_arrive(&thempty[pending]); }
          tma_store_2d(slabs + slot * K3T_SLAB_BYTES, &tm_theta, jb, li0);
          tma_store_commit();
          pending = slot;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (storer) {
      if (pending >= 0) { tma_store_wait_read(); mbar_arrive(&thempty[pending]); }
      tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// fa, fb fp32 [n][ldf] -> packed rows F bf16 [n][kf] (layout: lds_k3.cuh)
__global__ void k3_pack_kernel(const float* __restrict__ fa, const float* __restrict__ fb, int64_t ldf, int n, int h, int c, int kf,
                               __nv_bfloat16* __restrict__ f) {
  pdl_prologue();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * kf) return;
  const int i = (int)(idx / kf), k = (int)(idx - (int64_t)i * kf);
  f[idx] = k3_pack_element(fa + (int64_t)i * ldf, fb + (int64_t)i * ldf, h, c, k);
}

int32_t k3_launch_pack(const float* fa, const float* fb, int64_t ldf, int n, int h, int c, void* f, cudaStream_t stream) {
  const int kf = k3_packed_k(h, c);
  const int64_t total = (int64_t)n * kf;
  LDS_CHECK_CUDA(launch_dependent(k3_pack_kernel, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, stream, fa, fb, ldf, n, h, c, kf, reinterpret_cast<__nv_bfloat16*>(f)));
  LDS_CHECK_LAUNCH("k3_pack_kernel");
  return LDS_OK;
}

int32_t k3_launch_tc(float* theta, int64_t ldt, int n, int row0, int rows, const void* f, int kf,
                     const float* cvec, float lr, cudaStream_t stream, bool dependent_launch) {
  CUtensorMap tth, tf;
  int32_t rc;
  if ((rc = make_tmap_2d(&tth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, theta, n, rows, ldt, K3T_SLAB_COLS, K3T_TILE)) != LDS_OK) return rc;
  if ((rc = make_tmap_2d(&tf, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, f, kf, n, kf, K3T_KB, K3T_TILE)) != LDS_OK) return rc;
  static PerDeviceOnce once;
  if (first_use(once)) {
    LDS_CHECK_CUDA(cudaFuncSetAttribute(k3_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K3T_SMEM));
    LDS_CHECK_CUDA(cudaFuncSetAttribute(k3_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K3T_SMEM));
  }
  // Unsharded: the tile-symmetric launch (theta is symmetric). LDS_K3_FULL=1: A/B switch, every tile computed and stored on its own.
  static const bool force_full = getenv("LDS_K3_FULL") != nullptr;
  const bool sym = !force_full && row0 == 0 && rows == n;
  const int tj = (int)ceil_div(n, K3T_TILE);
  const int tiles = sym ? tj * (tj + 1) / 2 : (int)(ceil_div(rows, K3T_TILE) * tj);
  const int grid = tiles < kNumSMsB200 ? tiles : kNumSMsB200;
  static const bool no_pdl = getenv("LDS_NO_PDL") != nullptr;        // A/B switch
  if (dependent_launch && !no_pdl) {
    // the previous kernel on this stream triggers its dependents early (griddepcontrol.launch_dependents): this launch may start
    // its prologue before that kernel has finished; the kernel waits (griddepcontrol.wait) before it touches memory
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(K3T_THREADS); cfg.dynamicSmemBytes = K3T_SMEM; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int ks = kf / K3T_KB;
    cudaError_t e = sym ? cudaLaunchKernelEx(&cfg, k3_tc_kernel<true>, tth, tf, cvec, n, row0, rows, ks, lr, theta, ldt)
                        : cudaLaunchKernelEx(&cfg, k3_tc_kernel<false>, tth, tf, cvec, n, row0, rows, ks, lr, theta, ldt);
    if (e == cudaSuccess) return LDS_OK;
    (void)cudaGetLastError();                                  // fall through to the plain launch
  }
  if (sym) k3_tc_kernel<true><<<grid, K3T_THREADS, K3T_SMEM, stream>>>(tth, tf, cvec, n, row0, rows, kf / K3T_KB, lr, theta, ldt);
  else k3_tc_kernel<false><<<grid, K3T_THREADS, K3T_SMEM, stream>>>(tth, tf, cvec, n, row0, rows, kf / K3T_KB, lr, theta, ldt);
  LDS_CHECK_LAUNCH("k3_tc_kernel");
  return LDS_OK;
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

extern "C" int64_t lds_k3_workspace_bytes(int32_t n, int32_t d) {
  if (n <= 0 || d <= 0) return -1;
  return round_up((int64_t)n * k3_packed_k(d, 0) * 2, 1024);
}

extern "C" int32_t lds_k3k4_theta_update_tc(float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                            const float* fa, const float* fb, int64_t ld_f, int32_t d, const float* cvec,
                                            float lr, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(theta_full && fa && fb && cvec, "lds_k3k4_theta_update_tc: null pointer");
  LDS_CHECK_ARG(n > 0 && rows > 0 && row0 >= 0 && row0 + rows <= n, "lds_k3k4_theta_update_tc: rows [%d, %d) outside [0, %d)", row0, row0 + rows, n);
  LDS_CHECK_ARG(d > 0 && ld_f >= d, "lds_k3k4_theta_update_tc: need d > 0 and ld_f >= d");
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(cvec) & 15) == 0, "lds_k3k4_theta_update_tc: cvec must be 16-byte aligned");
  LDS_CHECK_ARG(ld_theta >= n && ld_theta % 4 == 0 && (reinterpret_cast<uintptr_t>(theta_full) & 15) == 0,
                "lds_k3k4_theta_update_tc: theta must be 16-byte aligned with ld_theta >= n, ld_theta %% 4 == 0");
  const int64_t need = lds_k3_workspace_bytes(n, d);
  if (!workspace || workspace_bytes < need) { set_error("lds_k3k4_theta_update_tc: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need); return LDS_ERR_WORKSPACE; }
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "lds_k3k4_theta_update_tc: workspace must be 1024-byte aligned");
  int32_t rc = k3_launch_pack(fa, fb, ld_f, n, d, 0, workspace, stream);
  if (rc != LDS_OK) return rc;
  return k3_launch_tc(theta_full, ld_theta, n, row0, rows, workspace, k3_packed_k(d, 0), cvec, lr, stream);
}
