// lds_epilogue.cuh — row epilogues of the four propagations, fused into the tail of the K2 kernel.
//
// Every epilogue of the GCN forward/backward chain is row-local: once a 128-row panel of
// A_tilde (r P) is complete, the CTA that completed it finishes the layer for those rows — scaling by r_i, relu,
// dropout, the tiny second linear, log-softmax + NLL, the backward chain — and emits the NEXT propagation's
// operand (r * .)^T as bf16 hi/lo terms. Reference semantics per function below.
//
// Work split: FOUR threads per row (a "quad" of adjacent lanes), each owning a contiguous quarter of the HP padded
// columns, all 16 epilogue warps of the CTA at once (512 threads = 128 rows x 4). The thread-per-row form this
// replaces ran ~1000-3000 dependent instructions on ONE warp per scheduler: 4-8 us per panel at h = 16, > 20 us at
// h = 64, all of it in the tail of the launch. Row-wide quantities (second linear, softmax, rho/kappa) are combined
// with two xor-shuffles in a fixed order, so results stay bitwise reproducible.
//
// Layout of the row-local state (P1, Z1, dZ1, dP1 [h][ldr]; P2, Z2, dZ2, dP2 [c][ldr]): TRANSPOSED, column-major in
// the row index: for a fixed column the 8 rows of a warp are one 32-byte sector.
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
  __nv_bfloat16* fpack; int kf;        // optional packed bf16 factor rows of the tensor-core update (lds_k3.cuh), else NULL:
  int64_t ld_fpack;                    // row i at fpack + i * ld_fpack, kf columns (multi-sample: column block s of [n][S * kf])
  int c_accumulate;                    // multi-sample: cvec += instead of =
  float scal_scale; int scal_accumulate;   // (loss, acc) *= scal_scale, added to out_scalars instead of stored (multi-sample mean)
  float scal_tag;                          // != 0: written to out_scalars[2] AFTER (loss, acc) with a system-scope fence, so a host that
                                           // polls pinned out_scalars sees the step's result as soon as it exists (lds_outer_step_args)
  float* cvec;
  const float* w1; const float* b1;
  const int64_t* y; const uint8_t* mask; float inv_m;
  DropCfg drop_h;
  __nv_bfloat16* bt_hi; __nv_bfloat16* bt_lo; int64_t ldb;       // where the NEXT propagation's operand is written: never the
  __nv_bfloat16* bt_hi_alt; __nv_bfloat16* bt_lo_alt;            // buffer the current one is still reading (ping-pong; `alt` selects)
  float* loss_part; float* corr_part; int nblk;       // one partial per 128-row panel
  float* out_scalars; float* out_logp;
  unsigned long long* timeline;   // optional debug: [grid][8] globaltimer stamps per CTA (scripts/k2_timeline.py), else NULL
  // K2_EPI_PLAIN (standalone lds_k2_propagate)
  float* z_out; int64_t ld_z; const float* scale_out; int rows; int width;
};

// (loss, acc) of the step from the per-panel partials, fixed order. One thread.
__device__ __forceinline__ void finalize_scalars(const EpiArgs& ea) {
  float l = 0.f, c = 0.f;
  for (int k = 0; k < ea.nblk; ++k) { l += ea.loss_part[k]; c += ea.corr_part[k]; }
  const float ls = l * ea.inv_m * ea.scal_scale, cs = c * ea.inv_m * ea.scal_scale;
  ea.out_scalars[0] = ea.scal_accumulate ? ea.out_scalars[0] + ls : ls;
  ea.out_scalars[1] = ea.scal_accumulate ? ea.out_scalars[1] + cs : cs;
  if (ea.scal_tag != 0.f) { __threadfence_system(); ea.out_scalars[2] = ea.scal_tag; }
}

__device__ __forceinline__ float quad_sum(float v) {           // fixed order: (x0 + x1) + (x2 + x3) seen from every lane
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

__device__ __forceinline__ void store_operand(const EpiArgs& a, bool alt, int c, int i, float v) {
  if (a.opnd) { a.opnd[(int64_t)i * a.ld_opnd + c] = v; return; }
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  (alt ? a.bt_hi_alt : a.bt_hi)[(int64_t)c * a.ldb + i] = hi;
  (alt ? a.bt_lo_alt : a.bt_lo)[(int64_t)c * a.ldb + i] = lo;
}

// In every function: i = row (may be >= a.n in the last panel: `live` guards the memory traffic, the shuffles are
// executed by all lanes), g = lane & 3 = which quarter of the columns, v = this thread's Q = HP/4 columns [g Q, g Q + Q).

// ---- plain: z = scale_out * sum -------------------------------------------------------------------------
template <int HP>
__device__ __forceinline__ void epi_plain(const EpiArgs& a, int i, int g, const float (&v)[HP / 4]) {
  constexpr int Q = HP / 4;
  if (i >= a.rows) return;
  const float so = a.scale_out ? a.scale_out[i] : 1.f;
#pragma unroll
  for (int k = 0; k < Q; ++k) { const int c = g * Q + k; if (c < a.width) a.z_out[(int64_t)i * a.ld_z + c] = v[k] * so; }
}

// ---- layer 1: Z1 = r * sum, H1 = relu, dropout, P2 = H1' W1^T + b1, operand (r * P2)^T      (gcn.py:28-30, layers.py:43)
template <int HP>
__device__ __forceinline__ void epi_layer1(const EpiArgs& a, int i, int g, float (&v)[HP / 4], bool alt = false) {
  constexpr int Q = HP / 4;
  const bool live = i < a.n;
  const int il = live ? i : a.n - 1;
  const float ri = a.rs[il];
#pragma unroll
  for (int kq = 0; kq < Q / 4; ++kq) {
    const int q = g * (Q / 4) + kq;                            // dropout quad = columns 4q .. 4q+3
    float k[4] = {0.f, 0.f, 0.f, 0.f};
    if (4 * q < a.h) drop_quad(a.drop_h, il, a.row0 + il, q, a.h, k);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      const float z = ri * v[4 * kq + e];
      if (live && c < a.h) a.z1[(int64_t)c * a.ldr + i] = z;
      v[4 * kq + e] = (c < a.h) ? fmaxf(z, 0.f) * k[e] : 0.f;
    }
  }
  for (int o0 = 0; o0 < a.c; o0 += 4) {                        // four outputs per trip; lane g keeps output o0 + g
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = o0 + u;
      float part = 0.f;
      if (o < a.c) {
        const float* wrow = a.w1 + (int64_t)o * a.h + g * Q;
#pragma unroll
        for (int k = 0; k < Q; ++k) if (g * Q + k < a.h) part = fmaf(v[k], wrow[k], part);
      }
      s[u] = quad_sum(part);
    }
    const float mine = (g == 0) ? s[0] : (g == 1) ? s[1] : (g == 2) ? s[2] : s[3];
    const int o = o0 + g;
    if (live && o < a.c) {
      const float acc = mine + a.b1[o];
      a.p2[(int64_t)o * a.ldr + i] = acc;
      store_operand(a, alt, o, i, ri * acc);
    }
  }
  // operand rows >= C are left as they are: column c of the product depends on operand row c only, and no epilogue
  // reads columns >= its width
}

// ---- layer 2: Z2 = r * sum, log_softmax, masked NLL + accuracy, dZ2, operand (r * dZ2)^T     (gcn.py:34, outer.py:65-67)
// Returns this thread's (loss, correct) contribution (non-zero on one lane of the quad); the caller reduces over the panel.
template <int HP>
__device__ __forceinline__ void epi_layer2(const EpiArgs& a, int i, int g, float (&v)[HP / 4], float& loss_i, float& corr_i, bool alt = false,
                                           int64_t logp_offset = 0) {
  constexpr int Q = HP / 4;
  loss_i = 0.f; corr_i = 0.f;
  const bool live = i < a.n;
  const int il = live ? i : a.n - 1;
  const float ri = a.rs[il];
  float mx = -3.4e38f; int best = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const int o = g * Q + k;
    if (o < a.c) {
      v[k] = ri * v[k];
      if (live) a.z2[(int64_t)o * a.ldr + i] = v[k];
      if (v[k] > mx) { mx = v[k]; best = o; }                  // ascending o: first maximum wins (torch.argmax)
    }
  }
#pragma unroll
  for (int sh = 1; sh <= 2; sh <<= 1) {                        // quad arg-max: larger value wins, ties go to the smaller index
    const float omx = __shfl_xor_sync(0xffffffffu, mx, sh);
    const int ob = __shfl_xor_sync(0xffffffffu, best, sh);
    if (omx > mx || (omx == mx && ob < best)) { mx = omx; best = ob; }
  }
  float se = 0.f;
#pragma unroll
  for (int k = 0; k < Q; ++k) if (g * Q + k < a.c) se += expf(v[k] - mx);
  se = quad_sum(se);
  const float lse = mx + logf(se);
  const int yi = (int)a.y[il];
  const bool mk = a.mask[il] != 0;
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const int o = g * Q + k;
    if (live && o < a.c) {
      const float lp = v[k] - lse;
      if (a.out_logp) a.out_logp[logp_offset + (int64_t)i * a.c + o] = lp;
      if (mk && o == yi) loss_i = -lp;
      const float dz = mk ? (expf(lp) - (o == yi ? 1.f : 0.f)) * a.inv_m : 0.f;
      a.dz2[(int64_t)o * a.ldr + i] = dz;
      store_operand(a, alt, o, i, ri * dz);
    }
  }
  corr_i = (live && g == 0 && mk && best == yi) ? 1.f : 0.f;
}

// ---- backward 2: dP2 = r * sum, dH1' = dP2 W1, dZ1 = dropout' relu', operand (r * dZ1)^T
// The quad first shares the whole dP2 row (C values), then lane g produces the hidden quads q = g, g + 4, ...
template <int HP>
__device__ __forceinline__ void epi_bwd2(const EpiArgs& a, int i, int g, float (&v)[HP / 4], bool alt = false) {
  constexpr int Q = HP / 4;
  const bool live = i < a.n;
  const int il = live ? i : a.n - 1;
  const float ri = a.rs[il];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const int o = g * Q + k;
    v[k] = (o < a.c) ? ri * v[k] : 0.f;
    if (live && o < a.c) a.dp2[(int64_t)o * a.ldr + i] = v[k];
  }
  float dp[HP];                                                // the row's dP2, every lane of the quad
#pragma unroll
  for (int o = 0; o < HP; ++o) dp[o] = (o < a.c) ? __shfl_sync(0xffffffffu, v[o % Q], (lane & ~3) + o / Q) : 0.f;
  for (int q = g; 4 * q < a.h; q += 4) {
    float k[4], z1[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
    drop_quad(a.drop_h, il, a.row0 + il, q, a.h, k);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const int c = 4 * q + e; z1[e] = (c < a.h) ? a.z1[(int64_t)c * a.ldr + il] : 0.f; }
#pragma unroll
    for (int o = 0; o < HP; ++o) {
      if (o < a.c) {
        const float* wrow = a.w1 + (int64_t)o * a.h + 4 * q;
#pragma unroll
        for (int e = 0; e < 4; ++e) if (4 * q + e < a.h) acc[e] = fmaf(dp[o], wrow[e], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      if (live && c < a.h) {
        const float dz = (z1[e] > 0.f) ? acc[e] * k[e] : 0.f;
        a.dz1[(int64_t)c * a.ldr + i] = dz;
        store_operand(a, alt, c, i, ri * dz);
      }
    }
  }
}

// ---- backward 1: dP1 = r * sum, rho, kappa, c = -(rho+kappa)/(2 deg), factor rows fa = r(dZ1|dZ2), fb = r(P1|P2)
// The factor rows leave as the packed bf16 operand row of the tensor-core update (lds_k3.cuh: 16 columns per step as
// a_hi, a_lo, b_hi, b_lo; the hidden part padded to a multiple of 16 so that a step never straddles dZ1|dZ2) and / or
// as row-major fp32 rows for the CUDA-core update.
// NC consecutive factor columns starting at padded column e0 (NC in {4, 8, 16}, e0 % NC == 0) of a packed row
template <int NC>
__device__ __forceinline__ void pack_cols(const float (&av)[NC], const float (&bv)[NC], __nv_bfloat16* row, int e0) {
  uint32_t w[4][NC / 2];
#pragma unroll
  for (int k = 0; k < NC; k += 2) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(av[k], h0, l0); split_bf16(av[k + 1], h1, l1);
    w[0][k >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    w[1][k >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    split_bf16(bv[k], h0, l0); split_bf16(bv[k + 1], h1, l1);
    w[2][k >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    w[3][k >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  __nv_bfloat16* base = row + 64 * (e0 >> 4) + (e0 & 15);      // step, then offset inside each 16-column block
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (NC == 4) *reinterpret_cast<uint2*>(base + 16 * t) = make_uint2(w[t][0], w[t][1]);
    else {
#pragma unroll
      for (int u = 0; u < NC / 8; ++u)
        *reinterpret_cast<uint4*>(base + 16 * t + 8 * u) = make_uint4(w[t][4 * u], w[t][4 * u + 1], w[t][4 * u + 2], w[t][4 * u + 3]);
    }
  }
}

template <int HP>
__device__ __forceinline__ void epi_bwd1(const EpiArgs& a, int i, int g, float (&v)[HP / 4]) {
  constexpr int Q = HP / 4;
  constexpr int NC = (Q < 16) ? Q : 16;                        // columns per packed group
  const bool live = i < a.n;
  const int il = live ? i : a.n - 1;
  const float ri = a.rs[il];
  const float di = a.deg[il];
  const int h16 = (a.h + 15) & ~15;
  float* fa = (a.fa && live) ? a.fa + (int64_t)i * a.ldf : nullptr;
  float* fb = (a.fb && live) ? a.fb + (int64_t)i * a.ldf : nullptr;
  __nv_bfloat16* fp = (a.fpack && live) ? a.fpack + (int64_t)i * a.ld_fpack : nullptr;
  float rho = 0.f, kappa = 0.f;
#pragma unroll
  for (int u = 0; u < Q / NC; ++u) {                           // hidden part: this thread's columns, NC at a time
    const int c0 = g * Q + u * NC;
    if (c0 < h16) {
      float dz1[NC], p1[NC], z1[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = c0 + k;
        const bool in = c < a.h;
        dz1[k] = in ? a.dz1[(int64_t)c * a.ldr + il] : 0.f;
        p1[k] = in ? a.p1[(int64_t)c * a.ldr + il] : 0.f;
        z1[k] = in ? a.z1[(int64_t)c * a.ldr + il] : 0.f;
      }
      float av[NC], bv[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = c0 + k;
        const float dp1 = ri * v[u * NC + k];
        rho = fmaf(dz1[k], z1[k], rho);
        kappa = fmaf(p1[k], dp1, kappa);
        av[k] = ri * dz1[k]; bv[k] = ri * p1[k];
        if (live && c < a.h) {
          a.dp1[(int64_t)c * a.ldr + i] = dp1;
          if (fa) { fa[c] = av[k]; fb[c] = bv[k]; }
        }
      }
      if (fp) pack_cols<NC>(av, bv, fp, c0);
    }
  }
  for (int o0 = 0; o0 < a.c; o0 += 16) {                       // class part: lane g takes columns o0 + 4g .. + 3 of each step
    float av[4], bv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int o = o0 + 4 * g + k;
      const bool in = o < a.c;
      const int64_t idx = (int64_t)(in ? o : 0) * a.ldr + il;
      const float dz2 = in ? a.dz2[idx] : 0.f, p2 = in ? a.p2[idx] : 0.f, z2 = in ? a.z2[idx] : 0.f, dp2 = in ? a.dp2[idx] : 0.f;
      rho = fmaf(dz2, z2, rho);
      kappa = fmaf(p2, dp2, kappa);
      av[k] = ri * dz2; bv[k] = ri * p2;
      if (in && fa) { fa[a.h + o] = av[k]; fb[a.h + o] = bv[k]; }
    }
    if (fp) pack_cols<4>(av, bv, fp, h16 + o0 + 4 * g);
  }
  rho = quad_sum(rho); kappa = quad_sum(kappa);
  if (g == 0 && live) {
    if (fa) for (int k = a.h + a.c; k < (int)a.ldf; ++k) { fa[k] = 0.f; fb[k] = 0.f; }
    const float cv = -(rho + kappa) / (2.f * di);              // both D^-1/2 factors depend on the row sum
    a.cvec[i] = a.c_accumulate ? a.cvec[i] + cv : cv;
  }
}

}  // namespace lds
