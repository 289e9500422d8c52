// lds_k1_tile.cuh — device code of the tile-symmetric, bit-packed sampling pass (see lds_k1_packed.cu for the design),
// shared by the stand-alone kernel (lds_k1_packed.cu) and the fused small-graph kernel (lds_fused_small.cu).
#pragma once
#include "lds_packed.cuh"
#include "lds_philox.cuh"

namespace lds {

constexpr int K1P_THREADS = 256;      // 8 warps = 8 chunks of one row-tile
constexpr int K1P_CHUNK = 4;          // consecutive column tiles per warp: row sums stay in registers along the strip
constexpr int K1P_GROUP = 4;          // row-pair steps whose theta loads are issued together

struct K1PArgs {
  const float* theta; int64_t ldt;
  int n, row0, rows;
  int nt;                              // column tiles = ceil(n / 64)
  int lo_t, my_tiles;                  // first global row tile of this shard, number of row tiles
  const float* u; int64_t ldu;         // explicit uniforms (parity mode) or NULL
  uint32_t* bits; int kblocks;
  int* cnt;                            // [rows] integer row sums (zero on entry)
  unsigned* ticket;                    // chunk counter (zero on entry), lives right behind cnt; NULL = static assignment (see k1p_warp_loop)
  int chunk;                           // tiles per ticket (K1P_CHUNK for the stand-alone kernel; 1 inside the fused small-graph kernel)
  int dbg;                             // measurement switches (LDS_K1P_DEBUG): 1 no theta loads, 2 no Philox rounds
  int prefetch;                        // != 0: pull a tile's theta rows towards L2 before its first load (fused small-graph kernel: a warp
                                       // samples about one tile, so nothing else hides the DRAM latency of its eight load groups)
};

__device__ __forceinline__ uint4 ld_cg_u4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// Interior tile (no diagonal block, all 64 columns < n, all 64 rows inside the shard): no per-cell masks, the orientation
// of the Philox blocks is uniform over the tile (LOWER: the tile lies below the diagonal, the two off-diagonal words swap),
// the packed row words leave as they are formed (the lane whose index equals the step stores the four ballots: rows 2t,
// 2t+1 are 16 contiguous bytes) and the column words shift in from the top: 2 instructions per cell. ~75 instructions per
// step of four cells, of which 40 are the ten Philox rounds; the generic path below costs ~290.
template <bool LOWER, bool MIRROR, int GROUP = K1P_GROUP>
__device__ __forceinline__ void k1p_tile_fast(const float* __restrict__ tptr, int64_t ldt, uint32_t p0, uint32_t q, const PhiloxRounds& R,
                                              uint32_t* __restrict__ row_dst, int lane, uint32_t (&c)[4], uint32_t c2, uint32_t c3, int dbg) {
  c[0] = c[1] = c[2] = c[3] = 0;
  for (int t0 = 0; t0 < 32; t0 += GROUP) {
    float2 th[GROUP][2];
#pragma unroll
    for (int u = 0; u < GROUP; ++u) {
      if (dbg & 1) { th[u][0] = make_float2(0.3f, 0.6f); th[u][1] = make_float2(0.2f, 0.9f); continue; }
      th[u][0] = __ldcs(reinterpret_cast<const float2*>(tptr + (int64_t)(2 * (t0 + u)) * ldt));
      th[u][1] = __ldcs(reinterpret_cast<const float2*>(tptr + (int64_t)(2 * (t0 + u) + 1) * ldt));
    }
#pragma unroll
    for (int u = 0; u < GROUP; ++u) {
      const int t = t0 + u;
      const uint32_t p = p0 + (uint32_t)t;
      uint32_t w[4];
      if (dbg & 2) { w[0] = p * 0x9E3779B9u; w[1] = q * 0x85EBCA6Bu; w[2] = w[0] ^ q; w[3] = w[1] ^ p; }
      else if (LOWER) philox4x32_10_rk(p, q, R, c2, c3, w); else philox4x32_10_rk(q, p, R, c2, c3, w);   // counter = (max block, min block)
      const uint32_t w01 = LOWER ? w[2] : w[1], w10 = LOWER ? w[1] : w[2];
      const bool s00 = (w[0] >> 8) < __float2uint_ru(th[u][0].x * 16777216.f);
      const bool s01 = (w01 >> 8)  < __float2uint_ru(th[u][0].y * 16777216.f);
      const bool s10 = (w10 >> 8)  < __float2uint_ru(th[u][1].x * 16777216.f);
      const bool s11 = (w[3] >> 8) < __float2uint_ru(th[u][1].y * 16777216.f);
      const uint32_t b00 = __ballot_sync(0xffffffffu, s00), b01 = __ballot_sync(0xffffffffu, s01);
      const uint32_t b10 = __ballot_sync(0xffffffffu, s10), b11 = __ballot_sync(0xffffffffu, s11);
      if (lane == t) *reinterpret_cast<uint4*>(row_dst + 4 * t) = make_uint4(b00, b01, b10, b11);
      if (MIRROR) {                                          // after 32 steps bit t of a word is the cell of step t
        c[0] >>= 1; if (s00) c[0] |= 0x80000000u;            // column 2l  : even rows -> word 0,
        c[1] >>= 1; if (s10) c[1] |= 0x80000000u;            //              odd rows  -> word 1
        c[2] >>= 1; if (s01) c[2] |= 0x80000000u;            // column 2l+1
        c[3] >>= 1; if (s11) c[3] |= 0x80000000u;
      }
    }
  }
}

// The ticket loop of one warp: takes chunks of K1P_CHUNK consecutive tiles until the trapezoid is exhausted.
// (c2, c3) = the step / stream / sample words of the Philox counter (a caller may advance the step between launches or,
// in a batched evaluation, between the graphs of ONE launch).
template <bool EXPLICIT_U, int GROUP = K1P_GROUP>
__device__ __forceinline__ void k1p_warp_loop(const K1PArgs& a, const PhiloxRounds& R, const uint32_t c2, const uint32_t c3, const int lane,
                                              long long chunk_first = 0, const long long chunk_stride = 0) {
  // Work = the trapezoid of tiles {(il, k): il < my_tiles, k < nt - il} (row tile il samples nt - il column tiles: all of
  // them minus the il it receives from mirrored tiles), linearised row by row; every warp takes K1P_CHUNK consecutive tiles.
  // (A 2-D grid with one CTA per (row tile, 8 chunks) left a quarter of the warp slots idle: warps past the end of a short
  // row exit at once but their CTA keeps its registers until the longest warp finishes.)
  const long long total = (long long)a.my_tiles * a.nt - (long long)a.my_tiles * (a.my_tiles - 1) / 2;
  const int hi_t = a.lo_t + a.my_tiles;
  const int n = a.n;
  const double bb = 2.0 * a.nt + 1.0;
  // Chunks are handed out dynamically (one atomic per chunk, re-armed by the finalize kernel): mirrored, plain and ragged
  // tiles cost different amounts, and a static split of 768 CTAs over 592 resident slots ran as two full waves.
  // a.ticket == NULL: static assignment instead — this warp takes chunks chunk_first, chunk_first + chunk_stride, ... (the fused
  // small-graph kernel has more warps than tiles: a few thousand same-address atomics in a burst cost more than the sampling)
  for (;;) {
  long long u = 0;
  if (a.ticket != nullptr) {
    if (lane == 0) u = (long long)atomicAdd(a.ticket, 1u) * a.chunk;
    u = __shfl_sync(0xffffffffu, u, 0);
  } else {
    u = chunk_first * a.chunk;
    chunk_first += chunk_stride;
  }
  if (u >= total) return;
  const long long u_end = min(total, u + a.chunk);
  // row tile of linear index u: largest il with il * nt - il (il - 1) / 2 <= u
  int il = (int)((bb - sqrt(bb * bb - 8.0 * (double)u)) * 0.5);
  il = max(0, min(il, a.my_tiles - 1));
  while (il > 0 && (long long)il * a.nt - (long long)il * (il - 1) / 2 > u) --il;
  while (il + 1 < a.my_tiles && (long long)(il + 1) * a.nt - (long long)(il + 1) * il / 2 <= u) ++il;
  int k = (int)(u - ((long long)il * a.nt - (long long)il * (il - 1) / 2));
  while (u < u_end) {
  const int ig = a.lo_t + il;                                 // global row tile
  const int count = a.nt - il;                                // column tiles this row tile samples: [0, lo_t) and [ig, nt)
  const int k_begin = k;
  const int k_end = (int)min((long long)count, (long long)k + (u_end - u));
  const int r_base = il * 64;                                 // local row of the tile's first row
  const float* trow = a.theta + (int64_t)r_base * a.ldt + 2 * lane;
  const int r_last = a.rows - 1 - r_base;                     // last valid row of this tile, relative to its first (>= 0)
  const bool rows_full = r_last >= 63;
  int sum0 = 0, sum1 = 0;                                     // row sums of local rows r_base + 2 lane, + 1 over the strip

  for (int kk = k_begin; kk < k_end; ++kk) {
    const int jg = (kk < a.lo_t) ? kk : ig + (kk - a.lo_t);
    const bool mirror = (jg > ig) && (jg < hi_t);             // the transposed tile belongs to this shard too: store it from here
    const int gj0 = jg * 64 + 2 * lane, gj1 = gj0 + 1;
    const uint32_t q = (uint32_t)(jg * 32 + lane);
    uint32_t* row_dst = a.bits + pk_word(r_base, jg, a.kblocks, 0);       // the tile's 64 rows x 2 words are contiguous
    uint32_t c[4] = {0u, 0u, 0u, 0u};                         // packed words of this lane's two COLUMNS (rows of the mirror tile)
    const bool fast = !EXPLICIT_U && rows_full && jg != ig && (jg + 1) * 64 <= n;     // warp-uniform
    if (a.prefetch) {                                         // rows 2 lane, 2 lane + 1 of the tile: two 128-byte lines each
      const float* pr = a.theta + (int64_t)(r_base + min(2 * lane, r_last)) * a.ldt + jg * 64;
      const float* ps = a.theta + (int64_t)(r_base + min(2 * lane + 1, r_last)) * a.ldt + jg * 64;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pr)); asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + 32));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ps)); asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + 32));
    }
    if ((a.dbg & 4) && !fast) continue;                     // measurement only: skip the generic tiles / the fast tiles
    if ((a.dbg & 8) && fast) continue;
    if (fast) {
      const float* tptr = trow + jg * 64;
      if (jg < ig) k1p_tile_fast<true, false, GROUP>(tptr, a.ldt, (uint32_t)(ig * 32), q, R, row_dst, lane, c, c2, c3, a.dbg);
      else if (mirror) k1p_tile_fast<false, true, GROUP>(tptr, a.ldt, (uint32_t)(ig * 32), q, R, row_dst, lane, c, c2, c3, a.dbg);
      else k1p_tile_fast<false, false, GROUP>(tptr, a.ldt, (uint32_t)(ig * 32), q, R, row_dst, lane, c, c2, c3, a.dbg);
    } else {
      // generic tile: diagonal blocks (self loops, both orientations), ragged edges, explicit uniforms
      for (int t0 = 0; t0 < 32; t0 += K1P_GROUP) {
        float2 th[K1P_GROUP][2];
#pragma unroll
        for (int u = 0; u < K1P_GROUP; ++u) {
          // unpredicated: rows past the shard are clamped to its last row (their cells are masked below)
          const int l0 = min(2 * (t0 + u), r_last), l1 = min(2 * (t0 + u) + 1, r_last);
          th[u][0] = __ldcs(reinterpret_cast<const float2*>(trow + (int64_t)l0 * a.ldt + jg * 64));
          th[u][1] = __ldcs(reinterpret_cast<const float2*>(trow + (int64_t)l1 * a.ldt + jg * 64));
        }
#pragma unroll 2
        for (int u = 0; u < K1P_GROUP; ++u) {
          const int t = t0 + u;
          const int r0 = r_base + 2 * t, r1 = r0 + 1;
          const int gi0 = a.row0 + r0, gi1 = gi0 + 1;
          const float2 th0 = th[u][0], th1 = th[u][1];
          bool s00, s01, s10, s11;
          const uint32_t p = (uint32_t)(ig * 32 + t);
          if (EXPLICIT_U) {
            // parity mode: element (i, j) uses U[min][max] (src/models/sampling.py:76), u < clamp(theta, 0, 1) in fp32
            auto draw = [&](int gi, int gj, float thv) {
              if (gi >= n || gj >= n) return false;
              const float uu = (gi <= gj) ? a.u[(int64_t)gi * a.ldu + gj] : a.u[(int64_t)gj * a.ldu + gi];
              return uu < fminf(fmaxf(thv, 0.f), 1.f);
            };
            s00 = draw(gi0, gj0, th0.x); s01 = draw(gi0, gj1, th0.y);
            s10 = draw(gi1, gj0, th1.x); s11 = draw(gi1, gj1, th1.y);
          } else {
            uint32_t w[4];
            philox4x32_10_rk(max(p, q), min(p, q), R, c2, c3, w);
            // word = 2 (a % 2) + (b % 2) of the canonical pair (a, b) = (min, max): below the diagonal the two off-diagonal cells swap
            const uint32_t w01 = (p <= q) ? w[1] : w[2], w10 = (p < q) ? w[2] : w[1];
            s00 = (w[0] >> 8) < __float2uint_ru(th0.x * 16777216.f);
            s01 = (w01 >> 8)  < __float2uint_ru(th0.y * 16777216.f);
            s10 = (w10 >> 8)  < __float2uint_ru(th1.x * 16777216.f);
            s11 = (w[3] >> 8) < __float2uint_ru(th1.y * 16777216.f);
          }
          if (p == q) { s00 = true; s11 = true; }             // self loops: diag := 1 (src/utils/graph.py:131-132)
          if (gj0 >= n) s00 = s10 = false;                    // (theta's padding columns are zero; do not depend on it)
          if (gj1 >= n) s01 = s11 = false;
          if (r0 >= a.rows) s00 = s01 = false;
          if (r1 >= a.rows) s10 = s11 = false;
          const uint32_t b00 = __ballot_sync(0xffffffffu, s00), b01 = __ballot_sync(0xffffffffu, s01);
          const uint32_t b10 = __ballot_sync(0xffffffffu, s10), b11 = __ballot_sync(0xffffffffu, s11);
          if (lane == t) *reinterpret_cast<uint4*>(row_dst + 4 * t) = make_uint4(b00, b01, b10, b11);
          c[0] |= (uint32_t)s00 << t; c[1] |= (uint32_t)s10 << t;   // column 2l  : even rows -> word 0, odd rows -> word 1
          c[2] |= (uint32_t)s01 << t; c[3] |= (uint32_t)s11 << t;   // column 2l+1
        }
      }
    }
    // row sums: every lane re-reads the two rows it accounts for (written by its own warp a moment ago: L2, not L1)
    __syncwarp();
    const uint4 mine = ld_cg_u4(row_dst + 4 * lane);
    sum0 += __popc(mine.x) + __popc(mine.y);
    sum1 += __popc(mine.z) + __popc(mine.w);
    if (mirror) {                                             // transposed tile: rows (jg - lo_t) * 64 + 2 lane, + 1 of k-block ig
      const int mr = (jg - a.lo_t) * 64 + 2 * lane;
      *reinterpret_cast<uint4*>(a.bits + pk_word(mr, ig, a.kblocks, 0)) = make_uint4(c[0], c[1], c[2], c[3]);
      const int m0 = __popc(c[0]) + __popc(c[1]), m1 = __popc(c[2]) + __popc(c[3]);
      if (mr < a.rows && m0) atomicAdd(a.cnt + mr, m0);
      if (mr + 1 < a.rows && m1) atomicAdd(a.cnt + mr + 1, m1);
    }
  }
  const int r = r_base + 2 * lane;
  if (r < a.rows && sum0) atomicAdd(a.cnt + r, sum0);
  if (r + 1 < a.rows && sum1) atomicAdd(a.cnt + r + 1, sum1);
  u += k_end - k_begin;                                       // next strip: the following row tile starts at its first column tile
  ++il; k = 0;
  }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Quarter-tile units for the fused small-graph kernel (lds_fused_small.cu). There a CTA has 18 warps and the whole graph is
// ~10 tiles per SM: with one tile per warp the phase lasts as long as ONE warp needs for 32 dependent steps (~17 us: a lone
// warp issues ~0.1 instructions per cycle on the Philox chains) while a third of the warps idle. The unit of work is
// therefore (tile, quarter): the 8 row-pair steps t0 .. t0 + 7 of the 64 x 64 tile (ig, jg), jg >= ig, of an UNSHARDED
// problem (row0 = 0, rows = n) — same draws, compares and bit layout as k1p_warp_loop. Rows leave as they are formed; the
// unit's share of the transposed tile is bits t0 .. t0 + 7 of every mirrored word, i.e. ONE BYTE per word (byte stores: no
// read-modify-write between the four warps that share a word); row sums come from the ballots (no re-read).
// EDGE: the tile touches the diagonal or the ragged end of the matrix (per-cell masks, both Philox orientations).
template <bool EDGE, bool EXPLICIT_U>
__device__ __forceinline__ void k1q_unit(const K1PArgs& a, const PhiloxRounds& R, const uint32_t c2, const uint32_t c3,
                                         const int ig, const int jg, const int t0, const int lane) {
  const int n = a.n;
  const int r_base = ig * 64;
  const int r_last = n - 1 - r_base;                          // last valid row of the tile, relative to its first (>= 0)
  const int gj0 = jg * 64 + 2 * lane, gj1 = gj0 + 1;
  const uint32_t q = (uint32_t)(jg * 32 + lane);
  const float* tp = a.theta + (int64_t)r_base * a.ldt + jg * 64 + 2 * lane;
  float2 th[8][2];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int l0 = EDGE ? min(2 * (t0 + u), r_last) : 2 * (t0 + u), l1 = EDGE ? min(2 * (t0 + u) + 1, r_last) : 2 * (t0 + u) + 1;
    if (a.dbg & 1) { th[u][0] = make_float2(0.3f, 0.6f); th[u][1] = make_float2(0.2f, 0.9f); continue; }
    // plain loads (not streaming): the update kernel that follows re-reads exactly these tiles, and at this size they fit in L2
    th[u][0] = __ldg(reinterpret_cast<const float2*>(tp + (int64_t)l0 * a.ldt));
    th[u][1] = __ldg(reinterpret_cast<const float2*>(tp + (int64_t)l1 * a.ldt));
  }
  uint32_t* row_dst = a.bits + pk_word(r_base, jg, a.kblocks, 0);         // the tile's 64 rows x 2 words are contiguous
  uint32_t c[4] = {0u, 0u, 0u, 0u};
  int sum0 = 0, sum1 = 0;                                     // row sums of rows 2 lane, 2 lane + 1 (lanes t0 .. t0 + 7)
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int t = t0 + u;
    const uint32_t p = (uint32_t)(ig * 32 + t);
    bool s00, s01, s10, s11;
    if (EXPLICIT_U) {
      // parity mode: element (i, j) uses U[min][max] (src/models/sampling.py:76), u < clamp(theta, 0, 1) in fp32
      const int gi0 = r_base + 2 * t, gi1 = gi0 + 1;
      auto draw = [&](int gi, int gj, float thv) {
        if (gi >= n || gj >= n) return false;
        const float uu = (gi <= gj) ? a.u[(int64_t)gi * a.ldu + gj] : a.u[(int64_t)gj * a.ldu + gi];
        return uu < fminf(fmaxf(thv, 0.f), 1.f);
      };
      s00 = draw(gi0, gj0, th[u][0].x); s01 = draw(gi0, gj1, th[u][0].y);
      s10 = draw(gi1, gj0, th[u][1].x); s11 = draw(gi1, gj1, th[u][1].y);
    } else {
      uint32_t w[4];
      if (a.dbg & 2) { w[0] = p * 0x9E3779B9u; w[1] = q * 0x85EBCA6Bu; w[2] = w[0] ^ q; w[3] = w[1] ^ p; }
      else if (EDGE) philox4x32_10_rk(max(p, q), min(p, q), R, c2, c3, w);      // counter = (max block, min block)
      else philox4x32_10_rk(q, p, R, c2, c3, w);                                // strictly above the diagonal: p < q
      // word = 2 (a % 2) + (b % 2) of the canonical pair (a, b) = (min, max): below the diagonal the two off-diagonal cells swap
      const uint32_t w01 = (!EDGE || p <= q) ? w[1] : w[2], w10 = (!EDGE || p < q) ? w[2] : w[1];
      s00 = (w[0] >> 8) < __float2uint_ru(th[u][0].x * 16777216.f);
      s01 = (w01 >> 8)  < __float2uint_ru(th[u][0].y * 16777216.f);
      s10 = (w10 >> 8)  < __float2uint_ru(th[u][1].x * 16777216.f);
      s11 = (w[3] >> 8) < __float2uint_ru(th[u][1].y * 16777216.f);
    }
    if (EDGE) {
      if (p == q) { s00 = true; s11 = true; }                 // self loops: diag := 1 (src/utils/graph.py:131-132)
      if (gj0 >= n) s00 = s10 = false;                        // (theta's padding columns are zero; do not depend on it)
      if (gj1 >= n) s01 = s11 = false;
      if (2 * t > r_last) s00 = s01 = false;
      if (2 * t + 1 > r_last) s10 = s11 = false;
    }
    const uint32_t b00 = __ballot_sync(0xffffffffu, s00), b01 = __ballot_sync(0xffffffffu, s01);
    const uint32_t b10 = __ballot_sync(0xffffffffu, s10), b11 = __ballot_sync(0xffffffffu, s11);
    if (lane == t) {
      *reinterpret_cast<uint4*>(row_dst + 4 * t) = make_uint4(b00, b01, b10, b11);
      sum0 = __popc(b00) + __popc(b01); sum1 = __popc(b10) + __popc(b11);
    }
    c[0] |= (uint32_t)s00 << u; c[1] |= (uint32_t)s10 << u;   // column 2l  : even rows -> word 0, odd rows -> word 1
    c[2] |= (uint32_t)s01 << u; c[3] |= (uint32_t)s11 << u;   // column 2l+1
  }
  const int r = r_base + 2 * lane;
  if (r < n && sum0) atomicAdd(a.cnt + r, sum0);
  if (r + 1 < n && sum1) atomicAdd(a.cnt + r + 1, sum1);
  if (jg > ig) {                                              // transposed tile: rows jg * 64 + 2 lane, + 1 of k-block ig
    const int mr = jg * 64 + 2 * lane;
    uint8_t* dst = reinterpret_cast<uint8_t*>(a.bits + pk_word(mr, ig, a.kblocks, 0)) + (t0 >> 3);
    dst[0] = (uint8_t)c[0]; dst[4] = (uint8_t)c[1]; dst[8] = (uint8_t)c[2]; dst[12] = (uint8_t)c[3];
    const int m0 = __popc(c[0]) + __popc(c[1]), m1 = __popc(c[2]) + __popc(c[3]);
    if (mr < n && m0) atomicAdd(a.cnt + mr, m0);
    if (mr + 1 < n && m1) atomicAdd(a.cnt + mr + 1, m1);
  }
}

// Units v = first, first + stride, ... of the 4 * nt (nt + 1) / 2 quarter-tile units (tile-major, tiles row by row).
template <bool EXPLICIT_U>
__device__ __forceinline__ void k1q_warp_loop(const K1PArgs& a, const PhiloxRounds& R, const uint32_t c2, const uint32_t c3, const int lane,
                                              const int first, const int stride) {
  const int nt = a.nt;
  const int units = 2 * nt * (nt + 1);
  const float bb = 2.0f * nt + 1.0f;
  auto tile_of = [&](int u, int& ig, int& jg) {
    // row tile of linear tile index u: largest ig with ig * nt - ig (ig - 1) / 2 <= u (float estimate, then exact correction)
    ig = (int)((bb - sqrtf(bb * bb - 8.0f * (float)u)) * 0.5f);
    ig = max(0, min(ig, nt - 1));
    while (ig > 0 && ig * nt - ig * (ig - 1) / 2 > u) --ig;
    while (ig + 1 < nt && (ig + 1) * nt - (ig + 1) * ig / 2 <= u) ++ig;
    jg = ig + (u - (ig * nt - ig * (ig - 1) / 2));
  };
  if (a.prefetch) {                                           // theta of ALL this warp's units towards L2 first: 16 rows x two 128-byte lines each
    for (int v = first; v < units; v += stride) {
      int ig, jg;
      tile_of(v >> 2, ig, jg);
      const int row = min(ig * 64 + 2 * (v & 3) * 8 + (lane >> 1), a.n - 1);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.theta + (int64_t)row * a.ldt + jg * 64 + 32 * (lane & 1)));
    }
  }
  for (int v = first; v < units; v += stride) {
    const int u = v >> 2, t0 = (v & 3) * 8;
    int ig, jg;
    tile_of(u, ig, jg);
    const bool edge = EXPLICIT_U || jg == ig || (jg + 1) * 64 > a.n;      // warp-uniform (the last row tile is a diagonal tile)
    if ((a.dbg & 4) && edge) continue;                        // measurement only
    if ((a.dbg & 8) && !edge) continue;
    if (edge) k1q_unit<true, EXPLICIT_U>(a, R, c2, c3, ig, jg, t0, lane);
    else k1q_unit<false, false>(a, R, c2, c3, ig, jg, t0, lane);
  }
}

}  // namespace lds
