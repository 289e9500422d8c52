// lds_epilogue.cuh — row epilogues of the four propagations, fused into the tail of the K2 kernel.
//
// Every epilogue of the GCN forward/backward chain is row-local: once a 128-row panel of
// A_tilde (r P) is complete, thread t of the epilogue warps owns row t of the panel with the whole operand
// width in registers and finishes the layer for that row — scaling by r_i, relu, dropout, the tiny second
// linear, log-softmax + NLL, the backward chain — and emits the NEXT propagation's operand (r * .)^T as bf16
// hi/lo terms (bt[c][i]: consecutive lanes = consecutive i, coalesced). Reference semantics per function below.
//
// Layout of the row-local state (P1, Z1, dZ1, dP1 [h][ldr]; P2, Z2, dZ2, dP2 [c][ldr]): TRANSPOSED, column-major in
// the row index, so that a warp (lane = row) touches one 128-byte line per access. With row-major [n][h] arrays every
// load/store instruction of the thread-per-row epilogue hit 32 different lines; the LSU serialises those, which made
// the epilogue of one 128-row panel cost 4-8 us at h = 16 and > 20 us at h = 64 (the tail of every K2 launch).
#pragma once
#include "lds_k2.cuh"
#include "lds_k3.cuh"
#include "lds_philox.cuh"

namespace lds {

enum K2Epi { K2_EPI_PLAIN = 0, K2_EPI_LAYER1 = 1, K2_EPI_LAYER2 = 2, K2_EPI_BWD2 = 3, K2_EPI_BWD1 = 4 };

struct DropCfg {           // dropout of one stream
  float p, keep_thresh, scale;
  const uint8_t* explicit_keep;   // [rows][cols] or nullptr
  PhiloxKey key;
};

__device__ __forceinline__ bool drop_keep(const DropCfg& dc, int row, int key_row, int col, int ncols) {
  if (dc.p <= 0.f) return true;
  if (dc.explicit_keep) return dc.explicit_keep[(int64_t)row * ncols + col] != 0;
  uint32_t w[4];
  philox4x32_10((uint32_t)(col >> 2), (uint32_t)key_row, dc.key, w);
  return philox_to_uniform(w[col & 3]) < dc.keep_thresh;
}

// keep flags (as multipliers: scale or 0) of columns 4q .. 4q+3 of `row`: one Philox call
// (`row` indexes the explicit mask — local rows of a shard; `key_row` = global row keys the Philox draw)
__device__ __forceinline__ void drop_quad(const DropCfg& dc, int row, int key_row, int q, int ncols, float (&k)[4]) {
  if (dc.p <= 0.f) { k[0] = k[1] = k[2] = k[3] = 1.f; return; }
  if (dc.explicit_keep) {
#pragma unroll
    for (int e = 0; e < 4; ++e) k[e] = (4 * q + e < ncols && dc.explicit_keep[(int64_t)row * ncols + 4 * q + e]) ? dc.scale : 0.f;
    return;
  }
  uint32_t w[4];
  philox4x32_10((uint32_t)q, (uint32_t)key_row, dc.key, w);
#pragma unroll
  for (int e = 0; e < 4; ++e) k[e] = (philox_to_uniform(w[e]) < dc.keep_thresh) ? dc.scale : 0.f;
}

struct EpiArgs {
  int n, h, c, hp1, hp2;       // n = number of (local) rows the epilogue owns
  int row0;                    // global index of local row 0 (row-block shard), keys the dropout draws
  float* opnd; int64_t ld_opnd; // sharded: write the next operand's rows as fp32 [n][ld_opnd] (all-gathered by the caller)
  const float* deg; const float* rs;
  float* p1; float* z1; float* p2; float* z2; float* dz2; float* dp2; float* dz1; float* dp1;   // transposed: [w][ldr]
  int64_t ldr;                 // row stride of the transposed state arrays (>= n)
  float* fa; float* fb; int64_t ldf;   // optional row-major fp32 factor rows [n][ldf] (CUDA-core / Adam update), else NULL
  __nv_bfloat16* fpack; int kf;        // optional packed bf16 factor rows [n][kf] of the tensor-core update (lds_k3.cuh), else NULL
  float* cvec;
  const float* w1; const float* b1;
  const int64_t* y; const uint8_t* mask; float inv_m;
  DropCfg drop_h;
  __nv_bfloat16* bt_hi; __nv_bfloat16* bt_lo; int64_t ldb;
  float* loss_part; float* corr_part; int nblk;       // one partial per 128-row panel
  float* out_scalars; float* out_logp;
  unsigned long long* timeline;   // optional debug: [grid][8] globaltimer stamps per CTA (scripts/k2_timeline.py), else NULL
  // K2_EPI_PLAIN (standalone lds_k2_propagate)
  float* z_out; int64_t ld_z; const float* scale_out; int rows; int width;
};

__device__ __forceinline__ void store_operand(const EpiArgs& a, int c, int i, float v) {
  if (a.opnd) { a.opnd[(int64_t)i * a.ld_opnd + c] = v; return; }
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  a.bt_hi[(int64_t)c * a.ldb + i] = hi;
  a.bt_lo[(int64_t)c * a.ldb + i] = lo;
}

// ---- plain: z = scale_out * sum -------------------------------------------------------------------------
template <int HP>
__device__ __forceinline__ void epi_plain(const EpiArgs& a, int i, const float (&v)[HP]) {
  if (i >= a.rows) return;
  const float so = a.scale_out ? a.scale_out[i] : 1.f;
#pragma unroll
  for (int c = 0; c < HP; ++c) if (c < a.width) a.z_out[(int64_t)i * a.ld_z + c] = v[c] * so;
}

// ---- layer 1: Z1 = r * sum, H1 = relu, dropout, P2 = H1' W1^T + b1, operand (r * P2)^T      (gcn.py:28-30, layers.py:43)
template <int HP>
__device__ __forceinline__ void epi_layer1(const EpiArgs& a, int i, float (&v)[HP]) {
  if (i >= a.n) return;                 // TMA never reads operand columns >= N
  const float ri = a.rs[i];
#pragma unroll
  for (int q = 0; q < HP / 4; ++q) {
    if (4 * q < a.h) {
      float k[4];
      drop_quad(a.drop_h, i, a.row0 + i, q, a.h, k);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = 4 * q + e;
        if (c < a.h) {
          const float z = ri * v[c];
          a.z1[(int64_t)c * a.ldr + i] = z;
          v[c] = fmaxf(z, 0.f) * k[e];
        } else v[c] = 0.f;
      }
    }
  }
  for (int o = 0; o < a.c; ++o) {
    float acc = a.b1[o];
    const float* wrow = a.w1 + (int64_t)o * a.h;
#pragma unroll
    for (int c = 0; c < HP; ++c) if (c < a.h) acc = fmaf(v[c], wrow[c], acc);
    a.p2[(int64_t)o * a.ldr + i] = acc;
    store_operand(a, o, i, ri * acc);
  }
  // operand rows >= C are left as they are: column c of the product depends on operand row c only, and no epilogue
  // reads columns >= its width
}

// ---- layer 2: Z2 = r * sum, log_softmax, masked NLL + accuracy, dZ2, operand (r * dZ2)^T     (gcn.py:34, outer.py:65-67)
// Returns this row's (loss, correct) contribution; the caller reduces over the panel.
template <int HP>
__device__ __forceinline__ void epi_layer2(const EpiArgs& a, int i, float (&v)[HP], float& loss_i, float& corr_i) {
  loss_i = 0.f; corr_i = 0.f;
  if (i >= a.n) return;
  const float ri = a.rs[i];
  float mx = -3.4e38f; int best = 0;
#pragma unroll
  for (int o = 0; o < HP; ++o) if (o < a.c) {
    v[o] = ri * v[o];
    a.z2[(int64_t)o * a.ldr + i] = v[o];
    if (v[o] > mx) { mx = v[o]; best = o; }                    // ascending o: first maximum wins (torch.argmax)
  }
  float se = 0.f;
#pragma unroll
  for (int o = 0; o < HP; ++o) if (o < a.c) se += expf(v[o] - mx);
  const float lse = mx + logf(se);
  const int yi = (int)a.y[i];
  const bool mk = a.mask[i] != 0;
#pragma unroll
  for (int o = 0; o < HP; ++o) if (o < a.c) {
    const float lp = v[o] - lse;
    if (a.out_logp) a.out_logp[(int64_t)i * a.c + o] = lp;
    if (mk && o == yi) loss_i = -lp;
    const float dz = mk ? (expf(lp) - (o == yi ? 1.f : 0.f)) * a.inv_m : 0.f;
    a.dz2[(int64_t)o * a.ldr + i] = dz;
    store_operand(a, o, i, ri * dz);
  }
  corr_i = (mk && best == yi) ? 1.f : 0.f;
}

// ---- backward 2: dP2 = r * sum, dH1' = dP2 W1, dZ1 = dropout' relu', operand (r * dZ1)^T
template <int HP>
__device__ __forceinline__ void epi_bwd2(const EpiArgs& a, int i, float (&v)[HP]) {
  if (i >= a.n) return;
  const float ri = a.rs[i];
#pragma unroll
  for (int o = 0; o < HP; ++o) {
    if (o < a.c) { v[o] = ri * v[o]; a.dp2[(int64_t)o * a.ldr + i] = v[o]; } else v[o] = 0.f;
  }
  for (int q = 0; 4 * q < a.h; ++q) {
    float k[4];
    drop_quad(a.drop_h, i, a.row0 + i, q, a.h, k);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      if (c >= a.h) break;
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < HP; ++o) if (o < a.c) acc = fmaf(v[o], a.w1[(int64_t)o * a.h + c], acc);
      const float dz = (a.z1[(int64_t)c * a.ldr + i] > 0.f) ? acc * k[e] : 0.f;
      a.dz1[(int64_t)c * a.ldr + i] = dz;
      store_operand(a, c, i, ri * dz);
    }
  }
}

// ---- backward 1: dP1 = r * sum, rho, kappa, c = -(rho+kappa)/(2 deg), factor rows fa = r(dZ1|dZ2), fb = r(P1|P2)
// The factor rows leave as the packed bf16 operand row of the tensor-core update (16 columns per step: a_hi, a_lo,
// b_hi, b_lo; the h-part padded to a multiple of 16 so that a step never straddles dZ1|dZ2 — zero columns add
// nothing to fa.fb) and / or as row-major fp32 rows for the CUDA-core update.
__device__ __forceinline__ void pack_step16(const float (&av)[16], const float (&bv)[16], __nv_bfloat16* dst) {
  // dst: 64 bf16 = [a_hi(16) | a_lo(16) | b_hi(16) | b_lo(16)], 16-byte aligned
  uint32_t w[32];
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(av[k], h0, l0); split_bf16(av[k + 1], h1, l1);
    w[k >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    w[8 + (k >> 1)] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    split_bf16(bv[k], h0, l0); split_bf16(bv[k + 1], h1, l1);
    w[16 + (k >> 1)] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    w[24 + (k >> 1)] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  uint4* out = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int u = 0; u < 8; ++u) out[u] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
}

template <int HP>
__device__ __forceinline__ void epi_bwd1(const EpiArgs& a, int i, float (&v)[HP]) {
  if (i >= a.n) return;
  const float ri = a.rs[i];
  const float di = a.deg[i];
  const int h16 = (a.h + 15) & ~15;
  float* fa = a.fa ? a.fa + (int64_t)i * a.ldf : nullptr;
  float* fb = a.fb ? a.fb + (int64_t)i * a.ldf : nullptr;
  __nv_bfloat16* fp = a.fpack ? a.fpack + (int64_t)i * a.kf : nullptr;
  float rho = 0.f, kappa = 0.f;
#pragma unroll
  for (int g = 0; g < HP / 16; ++g) {                          // hidden part, one packed step (16 columns) per trip
    if (16 * g < a.h) {
      float dz1[16], p1[16], z1[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {                           // coalesced: lane = row, one line per column
        const int c = 16 * g + k;
        const bool in = c < a.h;
        dz1[k] = in ? a.dz1[(int64_t)c * a.ldr + i] : 0.f;
        p1[k] = in ? a.p1[(int64_t)c * a.ldr + i] : 0.f;
        z1[k] = in ? a.z1[(int64_t)c * a.ldr + i] : 0.f;
      }
      float av[16], bv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int c = 16 * g + k;
        const float dp1 = ri * v[c];
        rho = fmaf(dz1[k], z1[k], rho);
        kappa = fmaf(p1[k], dp1, kappa);
        av[k] = ri * dz1[k]; bv[k] = ri * p1[k];
        if (c < a.h) {
          a.dp1[(int64_t)c * a.ldr + i] = dp1;
          if (fa) { fa[c] = av[k]; fb[c] = bv[k]; }
        }
      }
      if (fp) pack_step16(av, bv, fp + 64 * g);
    }
  }
  for (int o0 = 0; o0 < a.c; o0 += 16) {                       // class part
    float av[16], bv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int o = o0 + k;
      const bool in = o < a.c;
      const int64_t idx = (int64_t)o * a.ldr + i;
      const float dz2 = in ? a.dz2[idx] : 0.f, p2 = in ? a.p2[idx] : 0.f, z2 = in ? a.z2[idx] : 0.f, dp2 = in ? a.dp2[idx] : 0.f;
      rho = fmaf(dz2, z2, rho);
      kappa = fmaf(p2, dp2, kappa);
      av[k] = ri * dz2; bv[k] = ri * p2;
      if (in && fa) { fa[a.h + o] = av[k]; fb[a.h + o] = bv[k]; }
    }
    if (fp) pack_step16(av, bv, fp + 4 * h16 + 4 * o0);
  }
  if (fa) for (int k = a.h + a.c; k < (int)a.ldf; ++k) { fa[k] = 0.f; fb[k] = 0.f; }
  a.cvec[i] = -(rho + kappa) / (2.f * di);                     // both D^-1/2 factors depend on the row sum
}

}  // namespace lds
