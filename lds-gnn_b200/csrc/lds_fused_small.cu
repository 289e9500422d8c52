// lds_fused_small.cu — small-graph path of the outer step: K1 + feature GEMM + the four propagations as ONE
// persistent cooperative kernel with A_tilde resident in shared memory.
//
// At Cora / Citeseer shape (N = 2708 / 3327) the sampled adjacency (2 N^2 bytes: 14.7 / 22.1 MB) fits the combined
// shared memory of the 148 SMs (33 MB). The multi-kernel path writes A_tilde to HBM once and streams it back four
// times (10 N^2 of the step's 22 N^2 bytes) through five launches whose fixed costs (launch, TMEM / barrier setup,
// pipeline fill, tail) dominate at this size. Here every CTA owns the same contiguous range of (row panel, k-block)
// tiles for the whole kernel (the stream-K schedule of lds_k2.cuh, at most FS_MAX_TILES tiles of 128 x 64):
//
//   P0  sample its tiles straight into shared memory in the UMMA K-major SWIZZLE_128B layout (Philox keyed on the
//       canonical (min,max) 2x2 block, integer compare, self loops — bit-identical to k1_sample_kernel), write the
//       per-tile row sums; transpose its slice of layer_in.weight                                   — grid barrier
//   P1  deg = sum of the row sums, r = deg^-1/2; P1 = dropout(X) W0^T + b0 for its rows (CSR X, one warp per
//       row), operand (r P1)^T as bf16 hi/lo                                                         — grid barrier
//   P2  four times: TMA-load the operand k-blocks of its range, tcgen05.mma against the RESIDENT A tiles, drain /
//       count in / last-arriver reduction + row epilogue exactly as the K2 kernel (lds_k2_device.cuh) — grid barrier
//
// The theta update (K3+K4) stays its own launch. Reference semantics: see lds_k1_sample.cu, lds_outer_step.cu and
// lds_epilogue.cuh (src/models/sampling.py:47-85, src/utils/graph.py:123-153, src/models/gcn.py:23-34).
#include <stdlib.h>
#include <string.h>
#include "lds_fused_small.cuh"
#include "lds_k2_device.cuh"

namespace lds {

constexpr int FS_A_BYTES = K2_BLOCK_M * K2_BLOCK_K * 2;      // 16 KB: one resident A tile
constexpr int FS_B_BYTES = FS_HP * K2_BLOCK_K * 2;           // 2 KB: one bf16 term of one operand k-block
constexpr int FS_SMEM = FS_MAX_TILES * (FS_A_BYTES + 2 * FS_B_BYTES) + 1024 + 256;

__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t fs_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier over ONE monotonic 64-bit arrival counter (all CTAs are co-resident: cooperative launch). Every CTA of every
// launch on this workspace passes the same number of barriers, so the counter is a multiple of the grid size between launches
// and `base` = the value at kernel start rounded down to one (a CTA that starts late sees at most grid - 1 arrivals of
// barrier 0 on top of it). Barrier k is complete when the counter reaches base + (k + 1) * grid: one release-add, then
// acquire-polls of the same word — a separate generation word written by the last arriver costs one more L2 round trip per
// barrier, six times per step. Nothing to re-arm; 64 bits never wrap. Bounded spin: a protocol bug traps.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long base, unsigned& k, unsigned nctas) {
  __syncthreads();                                           // the CTA's writes happen-before thread 0's cumulative release below
  if (threadIdx.x == 0) {
    const unsigned long long want = base + (unsigned long long)(k + 1) * nctas;
    asm volatile("fence.proxy.async;" ::: "memory");         // generic-proxy writes -> TMA (async proxy) reads of other CTAs
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
    long long t0 = 0;
    for (unsigned spins = 0; ld_acquire_u64(bar) < want; ++spins) {
      if (spins == 64) t0 = clock64();
      if (spins > 64 && (spins & 255) == 0 && clock64() - t0 > 4000000000ll) {
        printf("liblds_b200: grid barrier timed out (block %d, barrier %u)\n", blockIdx.x, k);
        __trap();
      }
    }
  }
  __syncthreads();
  ++k;
}

// The same barrier in two halves: everything the CTA must publish is written, it arrives, does work that only touches its own
// shared memory, then waits. Hides the barrier's two L2 round trips (and the other CTAs' skew) behind that work.
__device__ __forceinline__ void grid_barrier_arrive(unsigned long long* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
  }
}
__device__ __forceinline__ void grid_barrier_wait(unsigned long long* bar, unsigned long long base, unsigned& k, unsigned nctas) {
  if (threadIdx.x == 0) {
    const unsigned long long want = base + (unsigned long long)(k + 1) * nctas;
    long long t0 = 0;
    for (unsigned spins = 0; ld_acquire_u64(bar) < want; ++spins) {
      if (spins == 64) t0 = clock64();
      if (spins > 64 && (spins & 255) == 0 && clock64() - t0 > 4000000000ll) {
        printf("liblds_b200: grid barrier timed out (block %d, barrier %u)\n", blockIdx.x, k);
        __trap();
      }
    }
  }
  __syncthreads();
  ++k;
}

// ---- thread-block cluster helpers (CLUSTER variant: the CTAs of one 128-row panel form a cluster) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {          // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(cta_rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}

// epi_bwd2 (lds_epilogue.cuh) for the CLUSTER variant at HP = 16: same arithmetic in the same order, but the row's dP2 goes
// to the other lanes of its quad through shared memory (`rowbuf`: the row's 16 floats of the partial-tile buffer, which this
// quad has just summed and nobody else reads) and the product with W1 is a rolled loop over the C classes: ~150 instructions
// instead of ~970 (16 shuffles and 16 predicated 4-wide steps, unrolled for the padded width).
__device__ __forceinline__ void fs_bwd2(const EpiArgs& a, int i, int g, float (&v)[4], bool alt, float* rowbuf) {
  const bool live = i < a.n;
  const int il = live ? i : a.n - 1;
  const float ri = a.rs[il];
  float k[4] = {0.f, 0.f, 0.f, 0.f}, z1[4] = {0.f, 0.f, 0.f, 0.f};
  const bool mine = 4 * g < a.h;
  if (mine) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { const int c = 4 * g + e; z1[e] = (c < a.h) ? a.z1[(int64_t)c * a.ldr + il] : 0.f; }
    drop_quad(a.drop_h, il, a.row0 + il, g, a.h, k);
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int o = g * 4 + kk;
    v[kk] = (o < a.c) ? ri * v[kk] : 0.f;
    if (live && o < a.c) a.dp2[(int64_t)o * a.ldr + i] = v[kk];
  }
  *reinterpret_cast<float4*>(rowbuf + 4 * g) = make_float4(v[0], v[1], v[2], v[3]);
  __syncwarp();
  if (mine) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int o = 0; o < a.c; ++o) {
      const float dpo = rowbuf[o];
      const float* wrow = a.w1 + o * a.h + 4 * g;
#pragma unroll
      for (int e = 0; e < 4; ++e) if (4 * g + e < a.h) acc[e] = fmaf(dpo, wrow[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * g + e;
      if (live && c < a.h) {
        const float dz = (z1[e] > 0.f) ? acc[e] * k[e] : 0.f;
        a.dz1[(int64_t)c * a.ldr + i] = dz;
        store_operand(a, alt, c, i, ri * dz);
      }
    }
  }
}

// CLUSTER = true: launched with cluster dimension = CTAs per panel. The split-K partial tiles of a panel stay in the
// shared memory of the CTAs that produced them (in the operand buffer, idle once the MMAs have retired); after a
// cluster barrier the cluster's rank-0 CTA sums them over DISTRIBUTED shared memory in rank order (deterministic) and
// runs the row epilogue. No partial tile, fence or arrival counter goes through L2: that chain was ~40 % of a phase.
template <bool CLUSTER>
__global__ void __launch_bounds__(K2_THREADS, 1)
fused_small_kernel(const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                   const __grid_constant__ CUtensorMap tm_bhi_alt, const __grid_constant__ CUtensorMap tm_blo_alt,
                   const __grid_constant__ FusedSmallArgs fa) {
  __shared__ K2EpiShared sh_epi;
  __shared__ unsigned long long sh_base;
  __shared__ int sh_rowpair;                                 // ticket of the feature-row pairs of this CTA
  __shared__ float sh_w1[FS_HP * FS_HP], sh_b1[FS_HP];       // layer_out's weights for the leader's row epilogues (CLUSTER)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                                        // resident A tiles
  uint8_t* sb = smem + FS_MAX_TILES * FS_A_BYTES;            // operand k-blocks of the current propagation: [tile][hi, lo]
  uint64_t* bfull = reinterpret_cast<uint64_t*>(sb + FS_MAX_TILES * 2 * FS_B_BYTES);
  uint64_t* tfull_bar = bfull + 1;                           // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const K2Sched& s = fa.s;
  const EpiArgs& ea = fa.ea;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int lo = cta * s.per_cta;                              // panel-aligned schedule: [lo, hi) lies inside ONE panel of the
  const int hi = lo + s.per_cta;                               // virtual (padded) k-range; k-blocks >= kb_real do not exist
  const int n = fa.n;
  const int my_p = lo / s.kblocks, my_kb0 = lo - my_p * s.kblocks;
  const int nt = min(s.per_cta, fa.kb_real - my_kb0);         // real tiles of this CTA (>= 1)

  unsigned long long* const gbar = reinterpret_cast<unsigned long long*>(fa.gridbar);
  if (tid == 0) { const unsigned long long cur = ld_acquire_u64(gbar); sh_base = cur - cur % gridDim.x; sh_rowpair = 0; }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_bhi); tma_prefetch_desc(&tm_blo); tma_prefetch_desc(&tm_bhi_alt); tma_prefetch_desc(&tm_blo_alt); }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(bfull, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 64);                                // two accumulators of [hi | lo] products: 2 x 32 columns
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const unsigned long long gbase = sh_base;
  unsigned gk = 0;
  auto stamp = [&](int k) {
    if (fa.timeline != nullptr && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      fa.timeline[(size_t)cta * 32 + k] = t;
    }
  };
  auto stamp_any = [&](int k) {                               // the calling thread stamps (finer phase breakdown)
    if (fa.timeline != nullptr) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      fa.timeline[(size_t)cta * 32 + k] = t;
    }
  };
  stamp(0);

  int acc_m = 0; uint32_t acc_phase_m = 0;                   // accumulator ring state of the MMA issuer
  int acc_e = 0; uint32_t acc_phase_e = 0;                   // ... and of the epilogue warps
  uint32_t bfull_uses = 0;                                   // completed uses of the operand barrier (its phase parity)
  // Batched evaluation (src/utils/evaluation.py:51-84): S graphs in ONE launch. The feature rows X W0^T + b0 are sample-
  // invariant in eval mode (no dropout), so they are computed for the first graph only; theta stays in L2 across samples.
  const int n_samples = fa.eval_samples > 1 ? fa.eval_samples : 1;
  for (int smp = 0; smp < n_samples; ++smp) {
  uint32_t c2s = fa.rounds.c2, c3s = fa.rounds.c3;
  if (n_samples > 1) {                                       // Philox step of this graph: step0 + smp, sample 0
    const unsigned long long st = fa.step0 + (unsigned long long)smp;
    c2s = (uint32_t)(st & 0xffffffffull);
    c3s = (c3s & 0xffffu) | (uint32_t)(((st >> 32) & 0xffffull) << 16);
  }
  // ================= P0: weight staging, tile-symmetric sampling to bits, feature rows =================
  const int r_lo = cta * fa.rows_per_cta, r_hi = min(n, r_lo + fa.rows_per_cta);
  if (smp == 0 && warp >= 16) {
    // Two warps stage this CTA's slice of layer_in.weight (transposed: one contiguous row per non-zero of X) before they start
    // sampling; the other sixteen start at once. Split barrier: arrive here, wait right before the feature rows, when every
    // CTA has long arrived. The count is re-armed by CTA 0 behind the next grid barrier.
    const int st = tid - 512;
#pragma unroll 4
    for (int idx = cta * 64 + st; idx < fa.f * ea.h; idx += gridDim.x * 64) {
      const int ff = idx / ea.h, o = idx - ff * ea.h;
      fa.w0t[idx] = fa.w0[(int64_t)o * fa.ldw + ff];
    }
    if (cta == 0) for (int k = st; k < s.panels; k += 64) fa.counters[k] = 0;
    named_bar_sync(4, 64);
    if (st == 0) { __threadfence(); atomicAdd(&fa.gridbar[2], 1u); }
    // Cold start (between two outer steps the inner problem trains: this step's inputs are not in L2): pull what the feature
    // rows of this CTA and the row epilogues will read towards L2 now, under the sampling.
    const int beg = fa.crow[r_lo], end = fa.crow[r_hi];
    for (int k = beg + st * 32; k < end; k += 64 * 32) { prefetch_l2(fa.xcol + k); prefetch_l2(fa.xval + k); }
    if (st < 4) prefetch_l2(ea.y + min(r_lo + st * 16, r_hi - 1));
    if (st == 4) { prefetch_l2(ea.mask + r_lo); prefetch_l2(ea.mask + r_hi - 1); }
    if (st >= 8 && st < 8 + 8 && (st - 8) * 32 < ea.c * ea.h) prefetch_l2(ea.w1 + (st - 8) * 32);
    if (st == 16) { prefetch_l2(ea.b1); prefetch_l2(fa.b0); }
  }
  // Every warp takes quarter tiles of the upper triangle (lds_k1_tile.cuh: the same draws and integer compare as the stand-
  // alone packed kernel): one Philox call and one theta read per UNORDERED 2 x 2 block, the tile and its transpose leave as
  // bits (L2), the row sums as integer atomics. Static assignment, interleaved over the CTAs: a global ticket would be a few
  // thousand same-address atomics in one burst, which alone took longer than the sampling.
  {
    const int first = warp * (int)gridDim.x + cta, stride = (K2_THREADS / 32) * (int)gridDim.x;
    if (fa.k1.u != nullptr) k1q_warp_loop<true>(fa.k1, fa.rounds, c2s, c3s, lane, first, stride);
    else k1q_warp_loop<false>(fa.k1, fa.rounds, c2s, c3s, lane, first, stride);
  }
  stamp(13);
  if (smp == 0) {
    // P1 = dropout(X) W0^T + b0 for this CTA's rows: independent of the sample, so the warps that are done sampling (a
    // warp gets at most about one tile) run it under the tail of the slow (diagonal / ragged) tiles. Two rows per warp:
    // h <= 16, so a half-warp owns a row (lane16 = output column; 16 non-zeros per trip); row pairs from a CTA-local ticket.
    if (lane == 0) while (ld_acquire_u32(&fa.gridbar[2]) < gridDim.x) { }
    __syncwarp();
    const int lane16 = lane & 15, half = lane >> 4;
    for (;;) {
      int pr = 0;
      if (lane == 0) pr = atomicAdd(&sh_rowpair, 1);
      pr = __shfl_sync(0xffffffffu, pr, 0);
      const int ib = r_lo + 2 * pr;
      if (ib >= r_hi) break;
      const int i = ib + half;
      const bool live = i < r_hi;
      const int il = live ? i : r_hi - 1;
      float acc = 0.f;
      const int beg = fa.crow[il], end = live ? fa.crow[il + 1] : beg;
      const int trips = (max(__shfl_sync(0xffffffffu, end - beg, 0), __shfl_sync(0xffffffffu, end - beg, 16)) + 15) >> 4;
      for (int tr = 0; tr < trips; ++tr) {                     // both halves run the same number of trips (shuffles are warp-wide)
        const int base = beg + 16 * tr;
        const int idx = base + lane16;
        int col = 0; float v = 0.f;
        if (idx < end) {
          col = fa.xcol[idx]; v = fa.xval[idx];
          if (fa.dx.p > 0.f) v = drop_keep(fa.dx, il, il, col, fa.f) ? v * fa.dx.scale : 0.f;
        }
        const int cnt = max(0, min(16, end - base));
#pragma unroll
        for (int j = 0; j < 16; j += 8) {                      // 8 non-zeros per trip: their w0t rows are loaded together
          float vj[8], wv[8];
          int cj[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { vj[u] = __shfl_sync(0xffffffffu, v, j + u, 16); cj[u] = __shfl_sync(0xffffffffu, col, j + u, 16); }
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = (lane16 < ea.h && j + u < cnt) ? fa.w0t[(int64_t)cj[u] * ea.h + lane16] : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) if (j + u < cnt) acc = fmaf(vj[u], wv[u], acc);
        }
      }
      if (live && lane16 < ea.h) ea.p1[(int64_t)lane16 * ea.ldr + i] = acc + fa.b0[lane16];
    }
  }
  stamp(1);
  grid_barrier(gbar, gbase, gk, gridDim.x);
  stamp(2);

  // ================= P1: expand this CTA's tiles into shared memory; degrees, r, first operand =================
  if (cta == 0 && tid == 0) fa.gridbar[2] = 0u;              // every CTA is past the split barrier: re-arm it for the next launch
  if (CLUSTER && smp == 0 && tid >= 512) {                   // warps 16-17 do not expand tiles: layer_out's weights -> shared memory
    for (int k = tid - 512; k < ea.c * ea.h; k += 64) sh_w1[k] = ea.w1[k];
    if (tid - 512 < ea.c) sh_b1[tid - 512] = ea.b1[tid - 512];
  }
  {
    // deg = integer row count (exact), r = deg^-1/2 with IEEE sqrt / divide like the reference; operand (r P1)^T as bf16 hi/lo
    const int nr = r_hi - r_lo;
    for (int idx = tid; idx < nr * FS_HP; idx += K2_THREADS) {
      const int c = idx / nr, i = r_lo + (idx - c * nr);
      const float d = (float)__ldcg(fa.k1.cnt + i);
      const float ri = 1.0f / sqrtf(d);
      if (c == 0) { fa.deg[i] = d; fa.rs[i] = ri; }
      if (c < ea.h) {
        __nv_bfloat16 bh, bl;
        split_bf16(ri * ea.p1[(int64_t)c * ea.ldr + i], bh, bl);
        ea.bt_hi[(int64_t)c * ea.ldb + i] = bh;
        ea.bt_lo[(int64_t)c * ea.ldb + i] = bl;
      }
    }
    // Everything other CTAs wait for (r, deg, the first operand) is written: arrive at the grid barrier now and expand the
    // tiles — work on this CTA's own shared memory — while the arrivals and the barrier's L2 round trips are in flight.
    grid_barrier_arrive(gbar);
    for (int idx = tid; idx < nr; idx += K2_THREADS) fa.k1.cnt[r_lo + idx] = 0;      // re-arm the row counters (this CTA is their only reader)
    if (tid == 0) sh_rowpair = 0;
  }
  if (tid < 512) {
    // row r of tile j: 64 bits (even / odd column words, lds_packed.cuh) -> 64 bf16 {0, 1} in the UMMA K-major SWIZZLE_128B
    // layout (16-byte chunk c of row r sits at chunk c ^ (r & 7)). The bits were written by other SMs: read them through L2.
    const int r = tid & 127, gi = my_p * K2_BLOCK_M + r;
    for (int j = tid >> 7; j < nt; j += 4) {
      const int kb = my_kb0 + j;
      uint2 w = make_uint2(0u, 0u);
      if (gi < n) w = __ldcg(reinterpret_cast<const uint2*>(fa.k1.bits + pk_word(gi, kb, fa.k1.kblocks, 0)));
      uint32_t pw[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {                            // bit deposit as in lds_k2_packed.cu (expand_row), then 0x4000 -> 0x3F80 = bf16 1.0
        const uint32_t a = ((q < 7) ? (w.x << (6 - q)) : (w.x >> 1)) & 0x40404040u;
        const uint32_t b = ((q < 7) ? (w.y << (6 - q)) : (w.y >> 1)) & 0x40404040u;
        const uint32_t x0 = fs_prmt(a, b, 0x4808u), x1 = fs_prmt(a, b, 0x5818u), x2 = fs_prmt(a, b, 0x6828u), x3 = fs_prmt(a, b, 0x7838u);
        pw[q] = x0 - (x0 >> 7); pw[q + 8] = x1 - (x1 >> 7); pw[q + 16] = x2 - (x2 >> 7); pw[q + 24] = x3 - (x3 >> 7);
      }
      uint8_t* row = sa + j * FS_A_BYTES + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = make_uint4(pw[4 * c], pw[4 * c + 1], pw[4 * c + 2], pw[4 * c + 3]);
      if (fa.a_dump != nullptr && gi < n) {                    // tests: the sampled A_tilde as the multi-kernel bf16 plan stores it
        uint4* d = reinterpret_cast<uint4*>(fa.a_dump + (int64_t)gi * fa.lda + kb * K2_BLOCK_K);
#pragma unroll
        for (int c = 0; c < 8; ++c) d[c] = make_uint4(pw[4 * c], pw[4 * c + 1], pw[4 * c + 2], pw[4 * c + 3]);
      }
    }
  }
  fence_proxy_async_smem();                                  // generic-proxy smem writes -> visible to tcgen05.mma
  stamp(3);
  grid_barrier_wait(gbar, gbase, gk, gridDim.x);
  stamp(4);

  // ================= P2: the four propagations from the resident tiles =================
  for (int ph = 0; ph < fa.num_phases; ++ph, ++bfull_uses) {
    // last phase of the launch: the update kernel that follows on the stream may be scheduled now (programmatic dependent
    // launch): its prologue and first instruction fetches overlap this phase; it waits for this grid's completion before it reads
    if (ph + 1 == fa.num_phases && smp + 1 == n_samples && tid == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 0) {
      if (lane == 0) {                                       // the operand k-blocks of this CTA's range (all fit: no ring)
        asm volatile("fence.proxy.async;" ::: "memory");     // the operand was written with generic stores by other CTAs
        mbar_expect_tx(bfull, (uint32_t)(nt * FS_B_BYTES * (fa.use_lo ? 2 : 1)));
        // Operand ping-pong: phase ph reads buffer ph & 1 and its epilogues write the other one — a leader that finishes
        // early must never overwrite operand rows another CTA has not loaded yet.
        const CUtensorMap* mh = (ph & 1) ? &tm_bhi_alt : &tm_bhi;
        const CUtensorMap* ml = (ph & 1) ? &tm_blo_alt : &tm_blo;
        for (int j = 0; j < nt; ++j) {                       // merged: hi and lo rows are one 32-row box (the two terms are adjacent)
          tma_load_2d(sb + j * 2 * FS_B_BYTES, mh, bfull, (my_kb0 + j) * K2_BLOCK_K, 0);
          if (fa.use_lo && !fa.merged_lo) tma_load_2d(sb + j * 2 * FS_B_BYTES + FS_B_BYTES, ml, bfull, (my_kb0 + j) * K2_BLOCK_K, 0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // one MMA of width 32 against the stacked operand [hi(16 rows); lo(16 rows)] (contiguous in smem): the hi and lo
        // products land in columns 0-15 / 16-31 of the accumulator and are added by the drain
        const uint32_t idesc = fa.use_lo ? umma_idesc_bf16(K2_BLOCK_M, 2 * FS_HP) : umma_idesc_bf16(K2_BLOCK_M, FS_HP);
        mbar_wait(bfull, bfull_uses & 1u);
        stamp_any(16 + 4 * ph);                              // operand k-blocks landed
        mbar_wait(&tempty_bar[acc_m], acc_phase_m ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc_m * 2 * FS_HP);
        for (int j = 0; j < nt; ++j) {
          const uint64_t adesc = umma_desc_k_sw128(smem_u32(sa + j * FS_A_BYTES));
          const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sb + j * 2 * FS_B_BYTES));
#pragma unroll
          for (int k = 0; k < K2_BLOCK_K / 16; ++k) tc_mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (j | k) != 0);
        }
        tc_commit(&tfull_bar[acc_m]);
        if (++acc_m == 2) { acc_m = 0; acc_phase_m ^= 1; }
      }
    } else if (!CLUSTER) {
      if (ph == 0) k2_epilogue_loop<FS_HP, K2_EPI_LAYER1, true>(s, ea, fa.partial, fa.counters, cta, lo, hi, tmem_base, tfull_bar, tempty_bar, acc_e, acc_phase_e, sh_epi, fa.use_lo != 0, ((ph + 1) & 1) != 0);
      else if (ph == 1) k2_epilogue_loop<FS_HP, K2_EPI_LAYER2, true>(s, ea, fa.partial, fa.counters, cta, lo, hi, tmem_base, tfull_bar, tempty_bar, acc_e, acc_phase_e, sh_epi, fa.use_lo != 0, ((ph + 1) & 1) != 0);
      else if (ph == 2) k2_epilogue_loop<FS_HP, K2_EPI_BWD2, true>(s, ea, fa.partial, fa.counters, cta, lo, hi, tmem_base, tfull_bar, tempty_bar, acc_e, acc_phase_e, sh_epi, fa.use_lo != 0, ((ph + 1) & 1) != 0);
      else k2_epilogue_loop<FS_HP, K2_EPI_BWD1, true>(s, ea, fa.partial, fa.counters, cta, lo, hi, tmem_base, tfull_bar, tempty_bar, acc_e, acc_phase_e, sh_epi, fa.use_lo != 0, ((ph + 1) & 1) != 0);
    } else if (warp < 6) {
      // CLUSTER: drain this CTA's accumulator into its own shared memory (the operand buffer is idle: the MMAs have retired)
      mbar_wait(&tfull_bar[acc_e], acc_phase_e);
      tc_fence_after();
      if (tid == 64) stamp_any(17 + 4 * ph);                 // accumulator complete
      const int quarter = warp & 3, row = quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc_e * 2 * FS_HP);
      uint32_t t16[16], u16[16];
      tc_ld16(taddr, t16);
      tc_ld16(taddr + FS_HP, u16);
      tc_wait_ld();
      if (fa.use_lo) {
#pragma unroll
        for (int q = 0; q < 16; ++q) t16[q] = __float_as_uint(__uint_as_float(t16[q]) + __uint_as_float(u16[q]));
      }
      uint4* dst = reinterpret_cast<uint4*>(sb + row * (FS_HP * 4));
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = make_uint4(t16[4 * q], t16[4 * q + 1], t16[4 * q + 2], t16[4 * q + 3]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc_e]);
      if (++acc_e == 2) { acc_e = 0; acc_phase_e ^= 1; }
    }
    if (CLUSTER) {
      cluster_sync_all();                                      // every CTA's partial tile is in its shared memory
      if (tid == 64) stamp_any(18 + 4 * ph);
      if (cluster_ctarank() == 0 && warp >= 2) {
        const int etid = tid - 64, row = etid >> 2, g = etid & 3;    // four threads per row, a quarter of the columns each
        const uint32_t mine = smem_u32(sb) + (uint32_t)(row * FS_HP + g * 4) * 4u;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        float4 t[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < fa.parts) t[r] = ld_dsmem_f4(mine, (uint32_t)r);       // all in flight, then a fixed-order sum
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < fa.parts) { v[0] += t[r].x; v[1] += t[r].y; v[2] += t[r].z; v[3] += t[r].w; }
        const int i = my_p * K2_BLOCK_M + row;
        const bool alt = ((ph + 1) & 1) != 0;
        if (tid == 64) stamp_any(19 + 4 * ph);               // partial tiles summed over distributed shared memory
        // (loss, acc): the layer-2 phase of EVERY panel is complete here — cluster 0 covers all k-blocks, so its CTAs have
        // consumed every panel's phase-1 flag, and each leader wrote its loss partial before its flag
        if (ph == 2 && my_p == 0 && etid == 0) finalize_scalars(ea);
        // W1 / b1 from shared memory: the cluster barrier's acquire has just invalidated L1, and data-dependent %globaltimer
        // stamps put 2.0 of the 3.1 us of the layer-1 epilogue on its two trips through rows of W1 in global memory
        EpiArgs es = ea;
        es.w1 = sh_w1; es.b1 = sh_b1;
        if (ph == 0) epi_layer1<FS_HP>(es, i, g, v, alt);
        else if (ph == 2) fs_bwd2(es, i, g, v, alt, reinterpret_cast<float*>(sb) + row * FS_HP);
        else if (ph == 3) epi_bwd1<FS_HP>(ea, i, g, v);
        else {
          float li, ci;
          epi_layer2<FS_HP>(ea, i, g, v, li, ci, alt, (int64_t)smp * n * ea.c);
          li = warp_sum(li); ci = warp_sum(ci);
          if (lane == 0) { sh_epi.red[warp - 2][0] = li; sh_epi.red[warp - 2][1] = ci; }
          named_bar_sync(2, 512);
          if (etid == 0) {
            float l = 0.f, c = 0.f;
            for (int w = 0; w < 16; ++w) { l += sh_epi.red[w][0]; c += sh_epi.red[w][1]; }   // fixed order
            ea.loss_part[my_p] = (smp ? ea.loss_part[my_p] : 0.f) + l;      // batched evaluation: sum over the graphs (this CTA is
            ea.corr_part[my_p] = (smp ? ea.corr_part[my_p] : 0.f) + c;      // the only writer of its panel's partial: fixed order)
          }
        }
      }
    }
    if (tid == 64) {                                           // first epilogue thread: its loop is done
      unsigned long long t;
      if (fa.timeline != nullptr) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); fa.timeline[(size_t)cta * 32 + 5 + 2 * ph] = t; }
    }
    // (Tried: per-panel ready flags instead of this barrier for phases 1-3 — a CTA only needs the operand rows of <= 7 panels — with
    // acquire-polling lanes in warp 0 and a release store by each panel's leader: 64.8 vs 61.7 us per launch, i.e. SLOWER than
    // the one-word barrier; reverted.)
    if (ph + 1 < fa.num_phases || smp + 1 < n_samples) grid_barrier(gbar, gbase, gk, gridDim.x);
    stamp(6 + 2 * ph);
  }
  }   // graphs of a batched evaluation

  if (CLUSTER) cluster_sync_all();                           // nobody exits while the cluster leader still reads its shared memory
  if (fa.num_phases == 2) {                                  // forward only: nobody runs the BWD2 prologue that finalises (loss, acc)
    grid_barrier(gbar, gbase, gk, gridDim.x);
    if (cta == 0 && tid == 0) finalize_scalars(ea);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

bool fused_small_schedule(int n, int hp1, int hp2, K2Sched& s, int& kb_real) {
  if (hp1 != FS_HP || hp2 != FS_HP || n < 2) return false;
  const int panels = (int)ceil_div(n, K2_BLOCK_M);
  kb_real = (int)ceil_div(n, K2_BLOCK_K);
  if (panels > kNumSMsB200) return false;
  int parts = kNumSMsB200 / panels;
  if (parts > kb_real) parts = kb_real;
  const int per = (int)ceil_div(kb_real, parts);
  if (per > FS_MAX_TILES) return false;
  parts = (int)ceil_div(kb_real, per);                       // drop parts that would own no real k-block
  s.hp = FS_HP; s.panels = panels; s.kblocks = parts * per; s.total = panels * s.kblocks;
  s.per_cta = per; s.grid = panels * parts; s.max_seg = 1;
  return true;
}

// How many clusters of `cs` CTAs of the CLUSTER kernel can be resident at once (cached per cluster size; 0 = cannot).
static int max_active_clusters(int cs) {
  static int cache_all[kMaxDevices][9];
  static PerDeviceOnce once;
  int* cache = cache_all[current_device()];
  if (first_use(once)) for (int k = 0; k < 9; ++k) cache[k] = -1;
  if (cs < 1 || cs > 8) return 0;
  if (cache[cs] >= 0) return cache[cs];
  int n = 0;
  if (cudaFuncSetAttribute(fused_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * 32)); cfg.blockDim = dim3(K2_THREADS); cfg.dynamicSmemBytes = FS_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, fused_small_kernel<true>, &cfg) != cudaSuccess) n = 0;
  }
  (void)cudaGetLastError();
  cache[cs] = n;
  return n;
}

int32_t fused_small_launch(const FusedSmallArgs& fa_in, cudaStream_t stream, bool allow_cluster) {
  static int coop_all[kMaxDevices];
  static PerDeviceOnce once;
  int& coop = coop_all[current_device()];
  if (first_use(once)) coop = -1;
  if (coop < 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) v = 0;
    if (v && num_sms() < fa_in.s.grid) v = 0;
    if (v) {
      if (cudaFuncSetAttribute(fused_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM) != cudaSuccess) v = 0;
    }
    (void)cudaGetLastError();
    coop = v;
  }
  if (!coop || num_sms() < fa_in.s.grid) return LDS_ERR_UNSUPPORTED;   // the caller falls back to the multi-kernel path
  FusedSmallArgs fa = fa_in;
  fa.parts = fa.s.grid / fa.s.panels;
  CUtensorMap tbh, tbl, tbh2, tbl2;
  int32_t rc;
  // the hi and lo terms of an operand buffer are adjacent ([16][ldb] each): one 32-row box fetches both
  fa.merged_lo = (fa.use_lo && fa.ea.bt_lo == fa.ea.bt_hi + (int64_t)FS_HP * fa.ea.ldb && fa.ea.bt_lo_alt == fa.ea.bt_hi_alt + (int64_t)FS_HP * fa.ea.ldb) ? 1 : 0;
  const int brows = fa.merged_lo ? 2 * FS_HP : FS_HP;
  if ((rc = make_tmap_2d(&tbh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, fa.ea.bt_hi, fa.n, brows, fa.ea.ldb, K2_BLOCK_K, brows)) != LDS_OK) return rc;
  if ((rc = make_tmap_2d(&tbl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, fa.ea.bt_lo, fa.n, FS_HP, fa.ea.ldb, K2_BLOCK_K, FS_HP)) != LDS_OK) return rc;
  if ((rc = make_tmap_2d(&tbh2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, fa.ea.bt_hi_alt, fa.n, brows, fa.ea.ldb, K2_BLOCK_K, brows)) != LDS_OK) return rc;
  if ((rc = make_tmap_2d(&tbl2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, fa.ea.bt_lo_alt, fa.n, FS_HP, fa.ea.ldb, K2_BLOCK_K, FS_HP)) != LDS_OK) return rc;
  void* params[] = {(void*)&tbh, (void*)&tbl, (void*)&tbh2, (void*)&tbl2, (void*)&fa};
  // LDS_FUSED_NO_CLUSTER: A/B switch for measurements. Nsight Compute (2025.2) dies on a launch that is both cooperative and
  // clustered (it reports grid (0,0,0) and exits with code 9), so under its injection library the L2 variant runs instead.
  static const bool env_no_cluster = getenv("LDS_FUSED_NO_CLUSTER") != nullptr || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr;
  static bool cluster_broken_all[kMaxDevices] = {};
  bool& cluster_broken = cluster_broken_all[current_device()];
  const bool cluster = allow_cluster && !env_no_cluster && !cluster_broken && fa.parts >= 2 && fa.parts <= 8 && max_active_clusters(fa.parts) >= fa.s.panels;
  if (cluster) {                                             // one cluster per panel: reduction over distributed shared memory
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)fa.s.grid); cfg.blockDim = dim3(K2_THREADS); cfg.dynamicSmemBytes = FS_SMEM; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)fa.parts; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    if (cudaLaunchKernelExC(&cfg, (const void*)fused_small_kernel<true>, params) == cudaSuccess) return LDS_OK;
    (void)cudaGetLastError();                                // e.g. the device cannot co-schedule the clusters after all:
    cluster_broken = true;                                   // use the variant without clusters from now on
  }
  if (fa.eval_samples > 1) return LDS_ERR_UNSUPPORTED;       // batched evaluation exists in the cluster variant only: the caller loops
  LDS_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)fused_small_kernel<false>, dim3((unsigned)fa.s.grid), dim3(K2_THREADS), params, FS_SMEM, stream));
  return LDS_OK;
}

}  // namespace lds
