// lds_outer_step.cu — the fused direct outer step: one OuterProblemTrainer.train_step
// (reference src/trainers/outer.py:57-87) with gcn_predict_fct = InnerProblemTrainer.model_forward
// (src/trainers/inner.py:76-78) evaluated at fixed GCN weights, regularize=False.
//
// Launch sequence (all on the caller's stream, no host sync, graph-capturable):
//   K1            theta -> A_tilde (bf16), deg, r                                   lds_k1_sample.cu
//   feat_linear   P1 = dropout(X) W0^T + b0 (Philox dropout fused, X read once); operand (r*P1)^T     gcn.py:27-28, layers.py:43
//   K2            A_tilde (r*P1)                                                    lds_k2_propagate.cu
//   epi_layer1    Z1 = r*., H1 = relu, dropout, P2 = H1' W1^T + b1; operand (r*P2)^T                 gcn.py:28-30
//   K2            A_tilde (r*P2)
//   epi_layer2    Z2 = r*., log_softmax, masked NLL + accuracy, dZ2; operand (r*dZ2)^T              gcn.py:34, outer.py:65-67
//   K2            A_tilde (r*dZ2)       (A_hat is symmetric: dP = A_hat dZ)
//   epi_bwd2      dP2 = r*., dH1' = dP2 W1, dZ1 = dropout'/relu' ; operand (r*dZ1)^T
//   K2            A_tilde (r*dZ1)
//   epi_bwd1      dP1 = r*., rho, kappa, c = -(rho+kappa)/(2 deg); factors fa = r(dZ1|dZ2), fb = r(P1|P2); loss/acc finalised
//   K3+K4         theta <- clamp(theta - lr * g(fa, fb, c))                         lds_k3_theta_update.cu
// The outer step needs no weight gradients (the reference computes and discards them), so X is read once.
#include <stdlib.h>
#include <string.h>
#include "lds_epilogue.cuh"
#include "lds_fused_small.cuh"
#include "lds_k2_packed.cuh"

namespace lds {

constexpr int EPI_ROWS = 32;        // rows per CTA in the row epilogues
constexpr int EPI_THREADS = 1024;   // 32 warps, one row each: the epilogues are latency-bound, so maximise rows in flight
constexpr int FEAT_THREADS = 256;   // dense feature GEMM: 8 warps x 4 rows
constexpr int EPI_MAXW = 128;       // widest operand

enum { B_A = 0, B_DEG, B_RS, B_P1, B_Z1, B_P2, B_Z2, B_DZ2, B_DP2, B_DZ1, B_DP1, B_FA, B_FB, B_C, B_BTHI, B_BTLO, B_BT2HI, B_BT2LO, B_PARTIAL, B_LOSSP, B_CORRP, B_F, B_W0S, B_CNT, B_OPND, B_DEGP, B_GBAR, B_BITS, B_ROWCNT, B_EVALTMP, B_END };
// The bf16 copy of A_tilde (2 N^2 bytes: the pre-packed launch plan, kept for the stream-K / A-B flags and the tests that
// compare the two plans) is only provisioned up to this many nodes; larger problems run on the bit-packed plan alone.
constexpr int kBf16AdjMaxN = 8192;
struct OuterLayout {
  int n, rows, f, h, c, hp1, hp2, hpmax, nblk, panels;
  int64_t lda, ldb, ldf, ldr;
  int kf;
  K2Sched s1, s2, sf;       // sf: panel-aligned schedule of the fused small-graph kernel
  K2PSched p1, p2;          // schedules of the packed propagation (lds_k2_packed.cu)
  bool fused; int kb_real;
  bool has_bf16_adj;
  int64_t off[B_END];
  int64_t total;
};

static bool make_layout(int n, int rows, int f, int h, int c, OuterLayout& L, bool streamk = false) {
  L.n = n; L.rows = rows; L.f = f; L.h = h; L.c = c;
  L.hp1 = k2_padded_width(h); L.hp2 = k2_padded_width(c);
  if (n <= 0 || rows <= 0 || rows > n || f <= 0 || L.hp1 < 0 || L.hp2 < 0) return false;
  L.hpmax = L.hp1 > L.hp2 ? L.hp1 : L.hp2;
  L.lda = round_up(n, kLdAlign); L.ldb = k2_operand_ld(n); L.ldf = round_up(h + c, 4);
  L.ldr = round_up(rows, 32);                               // row stride of the transposed row-local state ([w][ldr])
  L.kf = k3_packed_k(h, c);
  L.s1 = k2_make_schedule(n, rows, L.hp1, streamk); L.s2 = k2_make_schedule(n, rows, L.hp2, streamk);
  L.nblk = (int)ceil_div(rows, EPI_ROWS);
  L.panels = (int)ceil_div(rows, K2_BLOCK_M);
  int64_t bytes[B_END];
  L.has_bf16_adj = n <= kBf16AdjMaxN;
  bytes[B_A] = L.has_bf16_adj ? (int64_t)rows * L.lda * 2 : 0;
  bytes[B_BITS] = pk_bytes(n, rows);                          // bit-packed A_tilde (lds_packed.cuh)
  bytes[B_EVALTMP] = 16;
  bytes[B_ROWCNT] = (int64_t)(rows + 1) * 4;                        // integer row sums of the packed sampling pass (zero between calls)
  L.p1 = k2p_make_schedule(n, rows, L.hp1); L.p2 = k2p_make_schedule(n, rows, L.hp2);
  bytes[B_DEG] = bytes[B_RS] = bytes[B_C] = (int64_t)rows * 4;
  bytes[B_P1] = bytes[B_Z1] = bytes[B_DZ1] = bytes[B_DP1] = L.ldr * h * 4;
  bytes[B_P2] = bytes[B_Z2] = bytes[B_DZ2] = bytes[B_DP2] = L.ldr * c * 4;
  bytes[B_FA] = bytes[B_FB] = (int64_t)rows * L.ldf * 4;
  bytes[B_OPND] = (rows < n) ? (int64_t)rows * (h > c ? h : c) * 4 : 0;        // sharded: operand rows for the all-gather
  bytes[B_BTHI] = bytes[B_BTLO] = bytes[B_BT2HI] = bytes[B_BT2LO] = k2_operand_bytes(n, L.hpmax);   // operand ping-pong (see propagate)
  // partial tiles are sized for the stream-K schedule whichever schedule runs (workspace size must not depend on flags)
  const int64_t p1 = k2_partial_bytes(k2_make_schedule(n, rows, L.hp1, true)), p2 = k2_partial_bytes(k2_make_schedule(n, rows, L.hp2, true));
  bytes[B_PARTIAL] = p1 > p2 ? p1 : p2;
  if (k2p_partial_bytes(L.p1) > bytes[B_PARTIAL]) bytes[B_PARTIAL] = k2p_partial_bytes(L.p1);
  if (k2p_partial_bytes(L.p2) > bytes[B_PARTIAL]) bytes[B_PARTIAL] = k2p_partial_bytes(L.p2);
  L.fused = (rows == n) && fused_small_schedule(n, L.hp1, L.hp2, L.sf, L.kb_real);
  if (L.fused && k2_partial_bytes(L.sf) > bytes[B_PARTIAL]) bytes[B_PARTIAL] = k2_partial_bytes(L.sf);
  bytes[B_LOSSP] = bytes[B_CORRP] = (int64_t)ceil_div(rows, K2_BLOCK_M) * 4;
  bytes[B_F] = (int64_t)n * L.kf * 2;                       // packed bf16 factor rows of the tensor-core update (all n rows: K3 needs F_j of every column)
  bytes[B_CNT] = (int64_t)ceil_div(rows, K2_BLOCK_M) * 4;     // per-panel arrival counters of the stream-K reduction
  bytes[B_DEGP] = 0;                                        // (was: per-tile row sums of the fused small-graph path)
  bytes[B_GBAR] = 16;                                       // grid barrier of the fused small-graph kernel {count, generation}
  bytes[B_W0S] = (int64_t)h * round_up(f, 4) * 4;          // staged layer_in weight: transposed [f][h] (CSR path) or padded [h][ldx]
  int64_t o = 0;
  for (int b = 0; b < B_END; ++b) { L.off[b] = o; o += round_up(bytes[b], 1024); }
  L.total = o;
  return true;
}

// Transposed operand store: tile[c][r] (already scaled by r_i) -> bt_hi/lo[c][i0 + r], bf16 hi/lo split.
__device__ __forceinline__ void store_operand_tile(const float (*tile)[EPI_ROWS + 1], int hp, int i0, int64_t ldb,
                                                   __nv_bfloat16* __restrict__ bt_hi, __nv_bfloat16* __restrict__ bt_lo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = i0 + lane;
  if (i >= (int)ldb) return;
  for (int c = warp; c < hp; c += (int)(blockDim.x >> 5)) {
    __nv_bfloat16 hi, lo;
    split_bf16(tile[c][lane], hi, lo);
    bt_hi[(int64_t)c * ldb + i] = hi;
    bt_lo[(int64_t)c * ldb + i] = lo;
  }
}

// Stage layer_in.fc.weight [h][ldw] for the feature kernels: transposed [f][h] (CSR path: one contiguous row per
// non-zero) or zero-padded [h][ldp] with 16-byte rows (dense path). Tiled through shared memory, coalesced both ways.
__global__ void stage_w0_kernel(const float* __restrict__ w0, int64_t ldw, int h, int f, float* __restrict__ out, int64_t ldp, int transpose,
                                int* __restrict__ counters, int num_counters) {
  __shared__ float tile[32][33];
  if (blockIdx.x == 0 && blockIdx.y == 0)
    for (int k = threadIdx.x; k < num_counters; k += blockDim.x) counters[k] = 0;      // re-arm the stream-K panel counters
  const int f0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 256 threads
  for (int r = ty; r < 32; r += 8) {
    const int o = o0 + r, ff = f0 + tx;
    tile[r][tx] = (o < h && ff < f) ? w0[(int64_t)o * ldw + ff] : 0.f;
  }
  __syncthreads();
  if (transpose) {
    for (int r = ty; r < 32; r += 8) {
      const int ff = f0 + r, o = o0 + tx;
      if (ff < f && o < h) out[(int64_t)ff * h + o] = tile[tx][r];
    }
  } else {
    for (int r = ty; r < 32; r += 8) {
      const int o = o0 + r, ff = f0 + tx;
      if (o < h && ff < (int)ldp) out[(int64_t)o * ldp + ff] = tile[r][tx];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// P1 = dropout(X) W0^T + b0, operand (r*P1)^T.  One warp = 4 rows; lanes stride over feature quads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FEAT_THREADS)
feat_linear_kernel(const float* __restrict__ x, int64_t ldx, int n, int f, const float* __restrict__ w0, int64_t ldw, const float* __restrict__ b0, int h,
                   DropCfg dc, const float* __restrict__ rs, float* __restrict__ p1, int64_t ldr,
                   __nv_bfloat16* __restrict__ bt_hi, __nv_bfloat16* __restrict__ bt_lo, int64_t ldb, int hp,
                   int row0, float* __restrict__ opnd, int64_t ld_opnd) {
  pdl_prologue();
  __shared__ float tile[EPI_MAXW][EPI_ROWS + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * EPI_ROWS;
  const int f4 = (f + 3) >> 2;
  const int iters = (f4 + 31) >> 5;
  for (int idx = threadIdx.x; idx < EPI_MAXW * (EPI_ROWS + 1); idx += FEAT_THREADS) (&tile[0][0])[idx] = 0.f;
  __syncthreads();

  for (int oc = 0; oc < h; oc += 16) {
    float acc[4][16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[r][o] = 0.f;
    for (int it = 0; it < iters; ++it) {
      const int q = it * 32 + lane;
      const bool valid = q < f4;
      float4 xv[4];
      bool nz = false;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + 4 * warp + r;
        xv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && i < n) {
          xv[r] = *reinterpret_cast<const float4*>(x + (int64_t)i * ldx + 4 * q);
          if (dc.p > 0.f) {
            float k[4];
            if (dc.explicit_keep) {
#pragma unroll
              for (int e = 0; e < 4; ++e) k[e] = (4 * q + e < f && dc.explicit_keep[(int64_t)i * f + 4 * q + e]) ? dc.scale : 0.f;
            } else {
              uint32_t w[4];
              philox4x32_10((uint32_t)q, (uint32_t)(row0 + i), dc.key, w);
#pragma unroll
              for (int e = 0; e < 4; ++e) k[e] = (philox_to_uniform(w[e]) < dc.keep_thresh) ? dc.scale : 0.f;
            }
            xv[r].x *= k[0]; xv[r].y *= k[1]; xv[r].z *= k[2]; xv[r].w *= k[3];
          }
          nz = nz || (xv[r].x != 0.f) || (xv[r].y != 0.f) || (xv[r].z != 0.f) || (xv[r].w != 0.f);
        }
      }
      if (!__any_sync(0xffffffffu, nz)) continue;          // bag-of-words features are mostly zero
      if (valid) {
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          if (oc + o < h) {
            const float4 wv = *reinterpret_cast<const float4*>(w0 + (int64_t)(oc + o) * ldw + 4 * q);
#pragma unroll
            for (int r = 0; r < 4; ++r)
              acc[r][o] = fmaf(xv[r].x, wv.x, fmaf(xv[r].y, wv.y, fmaf(xv[r].z, wv.z, fmaf(xv[r].w, wv.w, acc[r][o]))));
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const float v = warp_sum(acc[r][o]);
        if (lane == 0 && oc + o < h) tile[oc + o][4 * warp + r] = v + b0[oc + o];
      }
  }
  __syncthreads();
  // P1 (transposed [h][ldr]: lanes = consecutive rows), then the scaled transposed operand
  if (i0 + lane < n) for (int c = warp; c < h; c += FEAT_THREADS / 32) p1[(int64_t)c * ldr + i0 + lane] = tile[c][lane];
  __syncthreads();
  {
    const int rr = threadIdx.x & 31;                         // scale column rr of the tile by r_i
    const int i = i0 + rr;
    const float ri = (i < n) ? rs[i] : 0.f;
    for (int c = threadIdx.x >> 5; c < hp; c += FEAT_THREADS / 32) tile[c][rr] *= ri;
  }
  __syncthreads();
  if (opnd) {                                                 // sharded: fp32 operand rows for the all-gather
    for (int r = 0; r < 4; ++r) {
      const int rr = 4 * warp + r, i = i0 + rr;
      if (i < n) for (int c = lane; c < h; c += 32) opnd[(int64_t)i * ld_opnd + c] = tile[c][rr];
    }
    return;
  }
  store_operand_tile(tile, hp, i0, ldb, bt_hi, bt_lo);
}

// ------------------------------------------------------------------------------------------------
// Sparse variant of the feature GEMM: x given as CSR (bag-of-words features are ~1% dense). One warp per row:
// lanes load up to 32 non-zeros (+ their dropout draw, keyed on (row, col) exactly like the dense kernel), then
// every non-zero is broadcast and the lanes FMA one contiguous row of w0t = W0^T each. Same result as the dense
// kernel up to fp32 summation order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EPI_THREADS)
feat_sparse_kernel(const int32_t* __restrict__ crow, const int32_t* __restrict__ xcol, const float* __restrict__ xval, int n, int f,
                   const float* __restrict__ w0t, const float* __restrict__ b0, int h,
                   DropCfg dc, const float* __restrict__ rs, float* __restrict__ p1, int64_t ldr,
                   __nv_bfloat16* __restrict__ bt_hi, __nv_bfloat16* __restrict__ bt_lo, int64_t ldb, int hp,
                   int row0, float* __restrict__ opnd, int64_t ld_opnd) {
  pdl_prologue();
  __shared__ float tile[EPI_MAXW][EPI_ROWS + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * EPI_ROWS;
  for (int idx = threadIdx.x; idx < EPI_MAXW * (EPI_ROWS + 1); idx += EPI_THREADS) (&tile[0][0])[idx] = 0.f;
  __syncthreads();
  for (int rr = warp; rr < EPI_ROWS; rr += EPI_THREADS / 32) {
    const int i = i0 + rr;
    if (i >= n) continue;
    const int beg = crow[i], end = crow[i + 1];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int base = beg; base < end; base += 32) {
      const int idx = base + lane;
      int col = 0; float v = 0.f;
      if (idx < end) {
        col = xcol[idx]; v = xval[idx];
        if (dc.p > 0.f) v = drop_keep(dc, i, row0 + i, col, f) ? v * dc.scale : 0.f;
      }
      const int cnt = min(32, end - base);
      for (int j = 0; j < cnt; j += 8) {                        // 8 non-zeros per trip: their w0t rows are loaded together
        float vj[8], wv[8][4];
        int cj[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {                           // lanes past `cnt` hold v = 0, col = 0: harmless
          vj[u] = __shfl_sync(0xffffffffu, v, (j + u) & 31);
          cj[u] = __shfl_sync(0xffffffffu, col, (j + u) & 31);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float* wrow = w0t + (int64_t)cj[u] * h;
#pragma unroll
          for (int t = 0; t < 4; ++t) { const int o = lane + 32 * t; wv[u][t] = (o < h && j + u < cnt) ? wrow[o] : 0.f; }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[t] = fmaf(vj[u], wv[u][t], acc[t]);
      }
    }
    const float ri = rs[i];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int o = lane + 32 * t;
      if (o < h) {
        const float pv = acc[t] + b0[o];
        tile[o][rr] = pv;
        if (opnd) opnd[(int64_t)i * ld_opnd + o] = ri * pv;
      }
    }
  }
  __syncthreads();
  // P1 (transposed [h][ldr]: lanes = consecutive rows); then scale the tile by r_i for the operand
  const float rl = (i0 + lane < n) ? rs[i0 + lane] : 0.f;
  for (int c = warp; c < h; c += EPI_THREADS / 32) {
    const float pv = tile[c][lane];
    if (i0 + lane < n) p1[(int64_t)c * ldr + i0 + lane] = pv;
    tile[c][lane] = rl * pv;
  }
  if (opnd) return;
  __syncthreads();
  store_operand_tile(tile, hp, i0, ldb, bt_hi, bt_lo);
}

// batched evaluation on the multi-launch plans: out[0:2] = mean over the graphs of tmp[0:2]
__global__ void eval_accumulate_kernel(float* out, const float* tmp, int s, int total, float tag) {
  if (threadIdx.x < 2) out[threadIdx.x] = (s ? out[threadIdx.x] : 0.f) + tmp[threadIdx.x] / (float)total;
  if (threadIdx.x == 0 && tag != 0.f) { __threadfence_system(); out[2] = tag; }
}

// forward-only steps: sum the per-panel partials in a fixed order (what the BWD2 launch does in a training step)
__global__ void finalize_scalars_kernel(const __grid_constant__ EpiArgs ea) {
  if (threadIdx.x == 0) finalize_scalars(ea);
}

}  // namespace lds

using namespace lds;

extern "C" int64_t lds_outer_step_workspace_bytes(int32_t n, int32_t f, int32_t h, int32_t c) {
  OuterLayout L;
  if (!make_layout(n, n, f, h, c, L)) return -1;
  return L.total;
}

extern "C" int64_t lds_outer_step_shard_workspace_bytes(int32_t n, int32_t rows, int32_t f, int32_t h, int32_t c) {
  OuterLayout L;
  if (!make_layout(n, rows, f, h, c, L)) return -1;
  return L.total;
}

extern "C" int64_t lds_outer_step_factor_ld(int32_t h, int32_t c) { return round_up(h + c, 4); }
// padded width of the operand that propagation phase `phase` (LDS_PHASE_LAYER1 .. LDS_PHASE_BWD1) multiplies
extern "C" int32_t lds_outer_step_operand_hp(int32_t h, int32_t c, uint32_t phase) {
  return k2_padded_width((phase == LDS_PHASE_LAYER1 || phase == LDS_PHASE_BWD1) ? h : c);
}
extern "C" int64_t lds_outer_step_packed_k(int32_t h, int32_t c) { return k3_packed_k(h, c); }
extern "C" int64_t lds_outer_step_state_ld(int32_t rows) { return round_up(rows, 32); }

extern "C" void* lds_outer_step_shard_buffer(void* workspace, int32_t n, int32_t rows, int32_t f, int32_t h, int32_t c, int32_t which) {
  OuterLayout L;
  if (!workspace || !make_layout(n, rows, f, h, c, L) || which < 0 || which > 16) return nullptr;
  const int b = (which == 14) ? B_OPND : (which == 15) ? B_F : (which == 16) ? B_BITS : which;
  return reinterpret_cast<uint8_t*>(workspace) + L.off[b];
}

extern "C" void* lds_outer_step_buffer(void* workspace, int32_t n, int32_t f, int32_t h, int32_t c, int32_t which) {
  return lds_outer_step_shard_buffer(workspace, n, n, f, h, c, which);
}

// Which launch plan lds_outer_step takes for these arguments: bit 0 = the fused small-graph kernel is eligible (the device may
// still refuse the cooperative launch), bit 1 = bit-packed A_tilde (else the bf16 plan).
extern "C" int32_t lds_outer_step_plan(const lds_outer_step_args* args) {
  if (!args || args->struct_bytes != sizeof(lds_outer_step_args)) return -1;
  const lds_outer_step_args& A = *args;
  const bool sharded = A.rows > 0 && A.rows < A.n;
  OuterLayout L;
  if (!make_layout(A.n, sharded ? A.rows : A.n, A.f, A.h, A.c, L)) return -1;
  int plan = 0;
  if (!sharded && A.x_crow != nullptr && !(A.k2_flags & (LDS_K2_NO_FUSE | LDS_K2_FORCE_STREAMK | LDS_K2_SIMT)) && L.fused) plan |= 1;
  if (!(A.k2_flags & (LDS_K2_FORCE_STREAMK | LDS_K2_SIMT | LDS_K2_BF16_ADJ)) && ((sharded ? A.row0 : 0) % 64 == 0)) plan |= 2;
  return plan;
}

extern "C" int32_t lds_outer_step(const lds_outer_step_args* args, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(args != nullptr, "lds_outer_step: null args");
  LDS_CHECK_ARG(args->struct_bytes == sizeof(lds_outer_step_args), "lds_outer_step: struct_bytes %u != %zu (header mismatch)", args->struct_bytes, sizeof(lds_outer_step_args));
  const lds_outer_step_args& A = *args;
  const bool sharded = A.rows > 0 && A.rows < A.n;
  const int rows = sharded ? A.rows : A.n;
  const int row0 = sharded ? A.row0 : 0;
  const uint32_t phases = sharded ? A.phases : LDS_PHASE_ALL;
  OuterLayout L;
  if (!make_layout(A.n, rows, A.f, A.h, A.c, L, (A.k2_flags & LDS_K2_FORCE_STREAMK) != 0)) { set_error("lds_outer_step: unsupported shape n=%d rows=%d f=%d h=%d c=%d (h, c must be in [1,128])", A.n, rows, A.f, A.h, A.c); return LDS_ERR_UNSUPPORTED; }
  LDS_CHECK_ARG(A.theta_full && A.w0 && A.b0 && A.w1 && A.b1 && A.y && A.mask && A.out_scalars, "lds_outer_step: null pointer");
  LDS_CHECK_ARG(A.ld_w0 >= A.f, "lds_outer_step: ld_w0 must be >= f");
  LDS_CHECK_ARG(A.ld_theta >= A.n && A.ld_theta % 4 == 0, "lds_outer_step: ld_theta must be >= n and a multiple of 4");
  if (sharded) {
    LDS_CHECK_ARG(row0 >= 0 && row0 + rows <= A.n && (row0 & 1) == 0, "lds_outer_step: shard rows [%d, %d) invalid (row0 must be even)", row0, row0 + rows);
    LDS_CHECK_ARG(phases != 0 && (phases & ~LDS_PHASE_ALL) == 0, "lds_outer_step: sharded calls need a phase mask");
    if (phases & (LDS_PHASE_LAYER1 | LDS_PHASE_LAYER2 | LDS_PHASE_BWD2 | LDS_PHASE_BWD1))
      LDS_CHECK_ARG(A.opnd_full && (phases & (phases - 1)) == 0, "lds_outer_step: a sharded propagation phase runs alone and needs opnd_full");
    if (phases & LDS_PHASE_UPDATE) {
      const bool tc = A.opt_kind == LDS_OPT_SGD && !(A.k3_flags & LDS_K3_SIMT);
      LDS_CHECK_ARG(A.c_full && (tc ? A.f_full != nullptr : (A.fa_full && A.fb_full)), "lds_outer_step: PHASE_UPDATE needs the gathered factors (f_full for the tensor-core SGD update, fa_full/fb_full otherwise) and c_full");
    }
  }
  const bool sparse_x = A.x_crow != nullptr;
  if (sparse_x) {
    LDS_CHECK_ARG(A.x_col && A.x_val, "lds_outer_step: the CSR feature path needs x_col and x_val");
  } else {
    LDS_CHECK_ARG(A.x, "lds_outer_step: null pointer (x)");
    LDS_CHECK_ARG(A.ld_x == round_up(A.f, 4), "lds_outer_step: dense x must be zero padded to ld_x == round_up(f, 4)");
    LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(A.x) & 15) == 0, "lds_outer_step: x must be 16-byte aligned");
  }
  LDS_CHECK_ARG(A.mask_count > 0, "lds_outer_step: mask_count must be positive");
  LDS_CHECK_ARG(A.dropout_p >= 0.f && A.dropout_p < 1.f, "lds_outer_step: dropout_p must be in [0, 1)");
  if (!A.workspace || A.workspace_bytes < L.total) { set_error("lds_outer_step: workspace too small (%lld < %lld)", (long long)A.workspace_bytes, (long long)L.total); return LDS_ERR_WORKSPACE; }
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(A.workspace) & 1023) == 0, "lds_outer_step: workspace must be 1024-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(A.workspace);
  auto buf = [&](int b) { return ws + L.off[b]; };
  auto fbuf = [&](int b) { return reinterpret_cast<float*>(ws + L.off[b]); };
  // sharded: the next operand's rows leave either as fp32 rows (re-laid out by a prep kernel after the all-gather) or,
  // with opnd_send, directly as the K-major bf16 hi/lo block the propagation's 3-D tensor map reads after the gather
  const bool packed_xchg = sharded && A.opnd_send != nullptr;
  if (packed_xchg) LDS_CHECK_ARG(A.opnd_rank_rows >= rows && A.opnd_rank_rows % 64 == 0 && (reinterpret_cast<uintptr_t>(A.opnd_send) & 1023) == 0,
                                 "lds_outer_step: opnd_send needs opnd_rank_rows >= rows, a multiple of 64, and a 1024-byte aligned buffer");
  float* opnd = (sharded && !packed_xchg) ? fbuf(B_OPND) : nullptr;
  const int64_t ld_opnd = A.h > A.c ? A.h : A.c;
  int* counters = reinterpret_cast<int*>(buf(B_CNT));
  auto* bt_hi = reinterpret_cast<__nv_bfloat16*>(buf(B_BTHI));
  auto* bt_lo = reinterpret_cast<__nv_bfloat16*>(buf(B_BTLO));
  const int64_t per = A.opnd_rank_rows;
  auto* snd = reinterpret_cast<__nv_bfloat16*>(A.opnd_send);
  // where an epilogue / the feature kernel writes the next operand (padded width hp_next), and with which row stride
  // Operand ping-pong: a propagation reads operand buffer `cur` while its epilogues write the NEXT operand into the other
  // one. With a single buffer, a panel that completes early overwrote operand rows that other CTAs were still streaming
  // (found by the full-size reproducibility test at N = 20 000; at Cora / Citeseer size every MMA retires before the
  // first epilogue starts, which hid it).
  __nv_bfloat16* pp_hi[2] = {bt_hi, reinterpret_cast<__nv_bfloat16*>(buf(B_BT2HI))};
  __nv_bfloat16* pp_lo[2] = {bt_lo, reinterpret_cast<__nv_bfloat16*>(buf(B_BT2LO))};
  // which buffer holds the operand of each propagation: SAMPLE -> 0, LAYER1 writes 1, LAYER2 writes 0, BWD2 writes 1
  auto out_hi = [&](int hp_next, int which) { (void)hp_next; return packed_xchg ? snd : pp_hi[which]; };
  auto out_lo = [&](int hp_next, int which) { return packed_xchg ? snd + (int64_t)hp_next * per : pp_lo[which]; };
  const int64_t out_ld = packed_xchg ? per : L.ldb;
  const bool use_lo = !(A.k2_flags & LDS_K2_SINGLE_BF16);

  DropCfg dx, dh;
  dx.p = dh.p = A.dropout_p;
  dx.keep_thresh = dh.keep_thresh = (float)(1.0 - (double)A.dropout_p);
  dx.scale = dh.scale = 1.0f / dx.keep_thresh;
  dx.explicit_keep = A.keep_x; dh.explicit_keep = A.keep_h;
  // Batched evaluation (LDS_K2_FORWARD_ONLY with num_samples = S > 1): S graphs drawn at Philox steps step .. step + S - 1
  // (sample 0), log-probabilities to out_logp[s][n][c], out_scalars = mean (loss, acc) over the graphs.
  const bool eval_batch = (A.k2_flags & LDS_K2_FORWARD_ONLY) && A.num_samples > 1;
  if (eval_batch) LDS_CHECK_ARG(!sharded && A.dropout_p == 0.f && A.sample_index == 0, "lds_outer_step: a batched evaluation is unsharded, without dropout, sample_index 0");
  const int S = (A.num_samples > 1 && !eval_batch) ? A.num_samples : 1;
  const uint32_t smp = S > 1 ? (uint32_t)A.sample_index : (uint32_t)(A.sample_index > 0 ? A.sample_index : 0);
  if (S > 1) {
    LDS_CHECK_ARG(A.sample_index >= 0 && A.sample_index < S, "lds_outer_step: sample_index %d outside [0, %d)", A.sample_index, S);
    LDS_CHECK_ARG(A.fpack_multi != nullptr && (reinterpret_cast<uintptr_t>(A.fpack_multi) & 15) == 0, "lds_outer_step: num_samples > 1 needs a 16-byte aligned fpack_multi");
    if (sharded || A.opt_kind != LDS_OPT_SGD || (A.k3_flags & LDS_K3_SIMT)) { set_error("lds_outer_step: the multi-sample estimator runs unsharded with the tensor-core SGD update"); return LDS_ERR_UNSUPPORTED; }
  }
  dx.key = philox_key(A.seed, A.step, LDS_STREAM_DROP_X, smp);
  dh.key = philox_key(A.seed, A.step, LDS_STREAM_DROP_H, smp);

  int32_t rc;
  profile_mark(stream, -1);

  // ---- small graphs: K1 + feature GEMM + the four propagations in one cooperative kernel (lds_fused_small.cu) ----
  bool fused_done = false;
  const bool tc_update = A.opt_kind == LDS_OPT_SGD && !(A.k3_flags & LDS_K3_SIMT);
  if (!sharded && sparse_x && !(A.k2_flags & (LDS_K2_NO_FUSE | LDS_K2_FORCE_STREAMK | LDS_K2_SIMT)) &&
      L.fused && A.ld_theta >= round_up(A.n, 64)) {
    FusedSmallArgs F;
    memset(&F, 0, sizeof(F));
    F.n = A.n;
    F.k1.theta = A.theta_full; F.k1.ldt = A.ld_theta; F.k1.n = A.n; F.k1.row0 = 0; F.k1.rows = A.n;
    F.k1.nt = (int)ceil_div(A.n, 64); F.k1.lo_t = 0; F.k1.my_tiles = F.k1.nt;
    F.k1.u = A.u_explicit; F.k1.ldu = A.ld_u;
    F.k1.bits = reinterpret_cast<uint32_t*>(buf(B_BITS)); F.k1.kblocks = pk_kblocks(A.n);
    F.k1.cnt = reinterpret_cast<int*>(buf(B_ROWCNT)); F.k1.ticket = nullptr;
    static const int k1dbg = getenv("LDS_K1P_DEBUG") ? atoi(getenv("LDS_K1P_DEBUG")) : 0;   // measurement switches of lds_k1_tile.cuh
    F.k1.chunk = 1; F.k1.dbg = k1dbg; F.k1.prefetch = 1;
    F.rounds = philox_rounds(philox_key(A.seed, A.step, LDS_STREAM_EDGES, smp));
    F.a_dump = (A.k2_flags & LDS_K2_DUMP_ADJ) ? reinterpret_cast<__nv_bfloat16*>(buf(B_A)) : nullptr; F.lda = L.lda;
    F.deg = fbuf(B_DEG); F.rs = fbuf(B_RS);
    F.crow = A.x_crow; F.xcol = A.x_col; F.xval = A.x_val; F.f = A.f;
    F.w0 = A.w0; F.ldw = A.ld_w0; F.w0t = fbuf(B_W0S); F.b0 = A.b0; F.dx = dx;
    F.num_phases = (A.k2_flags & LDS_K2_FORWARD_ONLY) ? 2 : 4;
    F.eval_samples = eval_batch ? A.num_samples : 1; F.step0 = A.step;
    F.rows_per_cta = (int)ceil_div(A.n, L.sf.grid);
    F.s = L.sf; F.kb_real = L.kb_real; F.partial = fbuf(B_PARTIAL); F.counters = counters; F.use_lo = use_lo ? 1 : 0;
    F.gridbar = reinterpret_cast<unsigned*>(buf(B_GBAR));
    F.timeline = A.k2_timeline;
    EpiArgs& E = F.ea;
    E.n = rows; E.h = A.h; E.c = A.c; E.hp1 = L.hp1; E.hp2 = L.hp2; E.row0 = 0;
    E.deg = fbuf(B_DEG); E.rs = fbuf(B_RS);
    E.p1 = fbuf(B_P1); E.z1 = fbuf(B_Z1); E.p2 = fbuf(B_P2); E.z2 = fbuf(B_Z2); E.dz2 = fbuf(B_DZ2); E.dp2 = fbuf(B_DP2);
    E.dz1 = fbuf(B_DZ1); E.dp1 = fbuf(B_DP1); E.ldr = L.ldr; E.cvec = fbuf(B_C);
    E.ldf = L.ldf; E.kf = L.kf;
    if (tc_update) { E.fpack = reinterpret_cast<__nv_bfloat16*>(buf(B_F)); E.ld_fpack = L.kf; }
    else { E.fa = fbuf(B_FA); E.fb = fbuf(B_FB); }
    if (S > 1) { E.fpack = reinterpret_cast<__nv_bfloat16*>(A.fpack_multi) + (int64_t)smp * L.kf; E.ld_fpack = (int64_t)S * L.kf; }
    E.c_accumulate = (S > 1 && smp > 0) ? 1 : 0; E.scal_scale = 1.0f / (float)(eval_batch ? A.num_samples : S); E.scal_accumulate = E.c_accumulate;
    E.scal_tag = ((int)smp == S - 1) ? A.scalars_tag : 0.f;
    E.w1 = A.w1; E.b1 = A.b1; E.y = A.y; E.mask = A.mask; E.inv_m = 1.0f / (float)A.mask_count;
    E.drop_h = dh; E.bt_hi = bt_hi; E.bt_lo = bt_lo; E.ldb = L.ldb;
    E.bt_hi_alt = reinterpret_cast<__nv_bfloat16*>(buf(B_BT2HI)); E.bt_lo_alt = reinterpret_cast<__nv_bfloat16*>(buf(B_BT2LO));
    E.loss_part = fbuf(B_LOSSP); E.corr_part = fbuf(B_CORRP); E.nblk = L.panels;
    E.out_scalars = A.out_scalars; E.out_logp = A.out_logp;
    rc = fused_small_launch(F, stream);
    if (rc == LDS_OK) { fused_done = true; profile_mark(stream, 10); }
    else if (rc != LDS_ERR_UNSUPPORTED) return rc;
  }
  if (eval_batch && !fused_done) {
    // any other launch plan: one forward-only pass per graph, the mean of the scalars accumulated on the device
    float* tmp = fbuf(B_EVALTMP);
    for (int s = 0; s < A.num_samples; ++s) {
      lds_outer_step_args one = A;
      one.num_samples = 1; one.step = A.step + (uint64_t)s; one.out_scalars = tmp; one.scalars_tag = 0.f;
      if (A.out_logp) one.out_logp = A.out_logp + (int64_t)s * A.n * A.c;
      if ((rc = lds_outer_step(&one, stream_)) != LDS_OK) return rc;
      eval_accumulate_kernel<<<1, 32, 0, stream>>>(A.out_scalars, tmp, s, A.num_samples, s == A.num_samples - 1 ? A.scalars_tag : 0.f);
      LDS_CHECK_LAUNCH("eval_accumulate_kernel");
    }
    return LDS_OK;
  }

  // Launch plan of everything that is not the fused small-graph kernel: the bit-packed A_tilde (tile-symmetric sampling,
  // on-chip expansion in the propagation) unless a flag asks for the bf16 plan (stream-K / validation switches).
  const bool packed = !(A.k2_flags & (LDS_K2_FORCE_STREAMK | LDS_K2_SIMT | LDS_K2_BF16_ADJ)) && (row0 % 64 == 0);
  if (!fused_done && !packed && !L.has_bf16_adj) {
    set_error("lds_outer_step: the bf16 A_tilde plan is provisioned up to n = %d (n = %d runs the bit-packed plan; row0 must be a multiple of 64)", kBf16AdjMaxN, A.n);
    return LDS_ERR_UNSUPPORTED;
  }
  if (!fused_done && (phases & LDS_PHASE_SAMPLE)) {
    {   // stage the layer_in weight (tiny; independent of K1) and re-arm the stream-K counters
      const int64_t ldp = round_up(A.f, 4);
      dim3 sgrid((unsigned)ceil_div(ldp, 32), (unsigned)ceil_div(A.h, 32));
      stage_w0_kernel<<<sgrid, 256, 0, stream>>>(A.w0, A.ld_w0, A.h, A.f, fbuf(B_W0S), ldp, sparse_x ? 1 : 0, counters, L.panels);
      LDS_CHECK_LAUNCH("stage_w0_kernel");
      profile_mark(stream, 8);
    }
    if (packed)
      rc = k1p_launch(A.theta_full, A.ld_theta, A.n, row0, rows, A.seed, A.step, smp, A.u_explicit, A.ld_u,
                      reinterpret_cast<uint32_t*>(buf(B_BITS)), reinterpret_cast<int*>(buf(B_ROWCNT)), fbuf(B_DEG), fbuf(B_RS), stream);
    else
      rc = lds_k1_sample_normalize(A.theta_full, A.ld_theta, A.n, row0, rows, A.seed, A.step, smp, A.u_explicit, A.ld_u,
                                   buf(B_A), L.lda, nullptr, 0, fbuf(B_DEG), fbuf(B_RS), A.u_explicit ? LDS_K1_EXPLICIT_U : 0u, stream_);
    if (rc != LDS_OK) return rc;
    profile_mark(stream, 0);
    const dim3 egrid((unsigned)L.nblk);
    if (sparse_x) {
      LDS_CHECK_CUDA(launch_dependent(feat_sparse_kernel, egrid, dim3(EPI_THREADS), 0, stream, A.x_crow, A.x_col, A.x_val, rows, A.f, fbuf(B_W0S), A.b0, A.h, dx,
                                      fbuf(B_RS), fbuf(B_P1), L.ldr, out_hi(L.hp1, 0), out_lo(L.hp1, 0), out_ld, L.hp1, row0, opnd, ld_opnd));
    } else {
      LDS_CHECK_CUDA(launch_dependent(feat_linear_kernel, egrid, dim3(FEAT_THREADS), 0, stream, A.x, A.ld_x, rows, A.f, fbuf(B_W0S), round_up(A.f, 4), A.b0, A.h, dx,
                                      fbuf(B_RS), fbuf(B_P1), L.ldr, out_hi(L.hp1, 0), out_lo(L.hp1, 0), out_ld, L.hp1, row0, opnd, ld_opnd));
    }
    profile_mark(stream, 1);
  }

  EpiArgs E;
  memset(&E, 0, sizeof(E));
  E.n = rows; E.h = A.h; E.c = A.c; E.hp1 = L.hp1; E.hp2 = L.hp2; E.row0 = row0;
  E.opnd = opnd; E.ld_opnd = ld_opnd;
  E.deg = fbuf(B_DEG); E.rs = fbuf(B_RS);
  E.p1 = fbuf(B_P1); E.z1 = fbuf(B_Z1); E.p2 = fbuf(B_P2); E.z2 = fbuf(B_Z2); E.dz2 = fbuf(B_DZ2); E.dp2 = fbuf(B_DP2);
  E.dz1 = fbuf(B_DZ1); E.dp1 = fbuf(B_DP1); E.ldr = L.ldr; E.cvec = fbuf(B_C);
  // factor rows: packed bf16 for the tensor-core update (own rows inside the n-row F buffer), row-major fp32 for the CUDA-core one
  E.ldf = L.ldf; E.kf = L.kf;
  if (tc_update) { E.fpack = reinterpret_cast<__nv_bfloat16*>(buf(B_F)) + (int64_t)row0 * L.kf; E.ld_fpack = L.kf; }
  else { E.fa = fbuf(B_FA); E.fb = fbuf(B_FB); }
  if (S > 1) { E.fpack = reinterpret_cast<__nv_bfloat16*>(A.fpack_multi) + (int64_t)smp * L.kf; E.ld_fpack = (int64_t)S * L.kf; }
  E.c_accumulate = (S > 1 && smp > 0) ? 1 : 0; E.scal_scale = 1.0f / (float)S; E.scal_accumulate = E.c_accumulate;
  E.scal_tag = (!sharded && (int)smp == S - 1) ? A.scalars_tag : 0.f;
  E.w1 = A.w1; E.b1 = A.b1; E.y = A.y; E.mask = A.mask; E.inv_m = 1.0f / (float)A.mask_count;
  E.drop_h = dh; E.bt_hi = bt_hi; E.bt_lo = bt_lo; E.ldb = L.ldb;
  E.loss_part = fbuf(B_LOSSP); E.corr_part = fbuf(B_CORRP); E.nblk = L.panels;
  E.out_scalars = A.out_scalars; E.out_logp = A.out_logp;

  if ((A.k2_flags & LDS_K2_SIMT) && (phases & (LDS_PHASE_LAYER1 | LDS_PHASE_LAYER2 | LDS_PHASE_BWD2 | LDS_PHASE_BWD1))) {
    set_error("lds_outer_step: LDS_K2_SIMT is only available through lds_k2_propagate"); return LDS_ERR_ARG;
  }
  auto propagate = [&](uint32_t phase, const K2Sched& s, const K2PSched& sp, int width, int epi, int mark, int hp_next, int cur) -> int32_t {
    if (fused_done || !(phases & phase)) return LDS_OK;
    const __nv_bfloat16* in_hi = pp_hi[cur];
    const __nv_bfloat16* in_lo = pp_lo[cur];
    int rank_rows = 0;
    if (packed_xchg) {                                        // gathered [rank][hi, lo][hp][per] bf16
      in_hi = reinterpret_cast<const __nv_bfloat16*>(A.opnd_full);
      in_lo = in_hi + (int64_t)s.hp * per;
      rank_rows = (int)per;
    } else if (sharded) {   // the gathered operand [n][width] fp32 -> K-major bf16 hi/lo terms
      const int32_t r0 = k2_launch_prep(A.opnd_full, ld_opnd, A.n, width, s.hp, nullptr, pp_hi[cur], pp_lo[cur], L.ldb, counters, 0, stream);
      if (r0 != LDS_OK) return r0;
    }
    E.bt_hi = out_hi(hp_next, cur ^ 1); E.bt_lo = out_lo(hp_next, cur ^ 1); E.ldb = out_ld;
    E.timeline = A.k2_timeline ? A.k2_timeline + (size_t)(mark - 3) * 512 * 8 : nullptr;
    const int32_t r = packed
        ? k2p_launch(buf(B_BITS), A.n, rows, in_hi, in_lo, L.ldb, fbuf(B_PARTIAL), sp, use_lo, epi, E, false, stream, rank_rows)
        : k2_launch_mma(buf(B_A), L.lda, A.n, rows, in_hi, in_lo, L.ldb, fbuf(B_PARTIAL), counters, s, use_lo, epi, E, stream, rank_rows);
    profile_mark(stream, mark);
    return r;
  };
  if ((rc = propagate(LDS_PHASE_LAYER1, L.s1, L.p1, A.h, K2_EPI_LAYER1, 3, L.hp2, 0)) != LDS_OK) return rc;   // Z1, H1, P2, operand (r P2)^T
  if ((rc = propagate(LDS_PHASE_LAYER2, L.s2, L.p2, A.c, K2_EPI_LAYER2, 4, L.hp2, 1)) != LDS_OK) return rc;   // Z2, log-softmax, loss, dZ2, operand (r dZ2)^T
  const bool fwd_only = (A.k2_flags & LDS_K2_FORWARD_ONLY) != 0;
  if (fwd_only && sharded) { set_error("lds_outer_step: LDS_K2_FORWARD_ONLY is not available for row-block shards"); return LDS_ERR_UNSUPPORTED; }
  if (fwd_only) {
    if (!fused_done && !sharded) {                            // (loss, acc) are normally finalised by the BWD2 launch
      finalize_scalars_kernel<<<1, 32, 0, stream>>>(E);
      LDS_CHECK_LAUNCH("finalize_scalars_kernel");
    }
    return LDS_OK;
  }
  if ((rc = propagate(LDS_PHASE_BWD2, L.s2, L.p2, A.c, K2_EPI_BWD2, 5, L.hp1, 0)) != LDS_OK) return rc;       // dP2, dZ1, operand (r dZ1)^T; loss/acc finalised
  if ((rc = propagate(LDS_PHASE_BWD1, L.s1, L.p1, A.h, K2_EPI_BWD1, 6, L.hp1, 1)) != LDS_OK) return rc;       // dP1, c, factor matrices

  if ((phases & LDS_PHASE_UPDATE) && A.update && (S == 1 || (int)smp == S - 1)) {
    const float* cv = sharded ? A.c_full : fbuf(B_C);
    if (S > 1) {                                              // mean of the S single-sample gradients: one pass, K = S * kf
      rc = k3_launch_tc(A.theta_full, A.ld_theta, A.n, 0, rows, A.fpack_multi, S * L.kf, cv, A.lr / (float)S, stream, (fused_done || packed) && !profile_active());
    } else if (tc_update) {
      const void* f = sharded ? A.f_full : buf(B_F);
      LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(cv) & 15) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0, "lds_outer_step: c_full / f_full must be 16-byte aligned");
      rc = k3_launch_tc(A.theta_full, A.ld_theta, A.n, row0, rows, f, L.kf, cv, A.lr, stream, (fused_done || packed) && !profile_active());
    } else {
      const float* fa = sharded ? A.fa_full : fbuf(B_FA);
      const float* fb = sharded ? A.fb_full : fbuf(B_FB);
      rc = lds_k3k4_theta_update(A.theta_full, A.ld_theta, A.n, row0, rows, fa, fb, L.ldf, A.h + A.c, cv,
                                 A.lr, A.opt_kind, A.adam_m, A.adam_v, A.beta1, A.beta2, A.eps, A.adam_t, nullptr, 0, 0u, stream_);
    }
    if (rc != LDS_OK) return rc;
    profile_mark(stream, 7);
  }
  return LDS_OK;
}
