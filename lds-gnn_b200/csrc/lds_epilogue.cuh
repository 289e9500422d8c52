// lds_epilogue.cuh — row epilogues of the four propagations, fused into the tail of the K2 kernel.
//
// Every epilogue of the GCN forward/backward chain is row-local: once a 128-row panel of
// A_tilde (r P) is complete, thread t of the epilogue warps owns row t of the panel with the whole operand
// width in registers and finishes the layer for that row — scaling by r_i, relu, dropout, the tiny second
// linear, log-softmax + NLL, the backward chain — and emits the NEXT propagation's operand (r * .)^T as bf16
// hi/lo terms (bt[c][i]: consecutive lanes = consecutive i, coalesced). Reference semantics per function below.
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
  float* p1; float* z1; float* p2; float* z2; float* dz2; float* dp2; float* dz1; float* dp1;
  float* fa; float* fb; int64_t ldf; float* cvec;
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
          a.z1[(int64_t)i * a.h + c] = z;
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
    a.p2[(int64_t)i * a.c + o] = acc;
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
    a.z2[(int64_t)i * a.c + o] = v[o];
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
    a.dz2[(int64_t)i * a.c + o] = dz;
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
    if (o < a.c) { v[o] = ri * v[o]; a.dp2[(int64_t)i * a.c + o] = v[o]; } else v[o] = 0.f;
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
      const float dz = (a.z1[(int64_t)i * a.h + c] > 0.f) ? acc * k[e] : 0.f;
      a.dz1[(int64_t)i * a.h + c] = dz;
      store_operand(a, c, i, ri * dz);
    }
  }
}

// ---- backward 1: dP1 = r * sum, rho, kappa, c = -(rho+kappa)/(2 deg), factor rows fa = r(dZ1|dZ2), fb = r(P1|P2)
template <int HP>
__device__ __forceinline__ void epi_bwd1(const EpiArgs& a, int i, float (&v)[HP]) {
  if (i >= a.n) return;
  const float ri = a.rs[i];
  const float di = a.deg[i];
  const int d = a.h + a.c;
  float* fa = a.fa + (int64_t)i * a.ldf;
  float* fb = a.fb + (int64_t)i * a.ldf;
  float dz1[HP], p1[HP], z1[HP];
#pragma unroll
  for (int c = 0; c < HP; ++c) {                               // every load of the row in flight before the first use
    const bool in = c < a.h;
    dz1[c] = in ? a.dz1[(int64_t)i * a.h + c] : 0.f;
    p1[c] = in ? a.p1[(int64_t)i * a.h + c] : 0.f;
    z1[c] = in ? a.z1[(int64_t)i * a.h + c] : 0.f;
  }
  float rho = 0.f, kappa = 0.f;
  for (int o0 = 0; o0 < a.c; o0 += 8) {
    float dz2[8], p2[8], z2[8], dp2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool in = o0 + u < a.c;
      const int64_t idx = (int64_t)i * a.c + o0 + u;
      dz2[u] = in ? a.dz2[idx] : 0.f; p2[u] = in ? a.p2[idx] : 0.f; z2[u] = in ? a.z2[idx] : 0.f; dp2[u] = in ? a.dp2[idx] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) if (o0 + u < a.c) {
      rho = fmaf(dz2[u], z2[u], rho);
      kappa = fmaf(p2[u], dp2[u], kappa);
      fa[a.h + o0 + u] = ri * dz2[u]; fb[a.h + o0 + u] = ri * p2[u];
    }
  }
#pragma unroll
  for (int c = 0; c < HP; ++c) if (c < a.h) {
    const float dp1 = ri * v[c];
    rho = fmaf(dz1[c], z1[c], rho);
    kappa = fmaf(p1[c], dp1, kappa);
    a.dp1[(int64_t)i * a.h + c] = dp1;
    fa[c] = ri * dz1[c]; fb[c] = ri * p1[c];
  }
  for (int k = d; k < (int)a.ldf; ++k) { fa[k] = 0.f; fb[k] = 0.f; }
  a.cvec[i] = -(rho + kappa) / (2.f * di);                     // both D^-1/2 factors depend on the row sum
}

}  // namespace lds
