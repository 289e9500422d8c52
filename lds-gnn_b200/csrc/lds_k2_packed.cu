// lds_k2_packed.cu — K2 on the bit-packed A_tilde: Z[rows x w] = A_tilde[rows x n] @ B[n x w], A_tilde read as N^2 / 8 bytes.
//
// Same product as lds_k2_propagate.cu (the dense `torch.mm(dense_adj, embeddings)` of MetaDenseGraphConvolution.forward,
// reference src/models/layers.py:44, and its transposes in backward) with the same operands — exact bf16 {0,1} tiles
// against the bf16 hi/lo split of the skinny matrix, fp32 accumulation in TMEM — but the adjacency arrives as bits
// (layout: lds_packed.cuh) and is expanded ON CHIP. With 2 bytes per element the propagation was an HBM stream of A_tilde
// (intensity = w flop/B, ceiling 25 % of the tensor peak at w = 64); with 1 bit per element the bytes per
// product drop 16x, A_tilde is L2-resident up to N ~ 30 000, and the kernel is bound by the tensor pipe (w = 64) or by the
// shared-memory writes of the expansion (w <= 32).
//
// Work unit = (super-panel of 256 rows, k-block of 64 columns). Operand roles are SWAPPED with respect to the bf16 kernel:
// the MMA computes Z^T = P^T A_tilde^T, i.e. the skinny operand is the M side (its hi and lo terms stacked: rows 0..HP-1 and
// 64..64+HP-1 of one 128-row tile, so ONE MMA forms both products) and the 256 adjacency rows are the N side. Measured on
// B200: a 128 x HP x 16 MMA with the big operand on the M side costs ~115 cycles whatever HP is (the A-operand read from
// shared memory is exposed at small N), which capped the bf16 kernel AND a first packed kernel at the same ~170 us per
// propagation at N = 20 000; with N = 256 a k-step is one 128-cycle MMA for 256 rows.
// Persistent stream-K over the linearised (super-panel, k-block) space, one CTA per SM, 15 warps:
//   warp 0 / 14  producers: TMA of the two operand terms (HP x 64, K-major) / one bulk copy of the 2 KB bit block per unit
//   warps 6-13   expanders: one row per thread — 64 bits -> 64 bf16 (a bit deposit: ~1.5 ALU instr per bf16 pair),
//                eight 16-byte stores into the K-major SWIZZLE_128B tile, fence.proxy.async, arrive on `a_ready`
//   warp 1       tcgen05.mma issuer: per unit 4 k-steps of one 128 x 256 x 16 MMA (two for HP = 128: hi, lo) into a
//                128-lane x 256-column fp32 accumulator (lane = feature, column = adjacency row), double-buffered in TMEM
//   warps 2-5    drain: tcgen05.ld -> fp32 partial tiles, transposed [plane][feature][256 rows] (L2)
// The split-K reduction (fixed order => bitwise reproducible; it also adds the hi and lo planes) and the row epilogue of
// lds_epilogue.cuh run as a second, tiny kernel over the 128-row panels: at the sizes this path serves the launch gap is
// < 2 % of the propagation, and keeping the epilogue out of the MMA kernel leaves its warps to the expansion.
#include <stdlib.h>
#include <string.h>
#include "lds_k2_packed.cuh"
#include "lds_tc.cuh"

namespace lds {

constexpr int K2P_THREADS = 480;                 // 15 warps
constexpr int K2P_EXP_WARPS = 8;
constexpr int K2P_TILE_BYTES = 128 * 64 * 2;     // one 128 x 64 bf16 tile
constexpr int K2P_NA = 3;                        // ring of expanded adjacency blocks (32 KB each)
constexpr int K2P_NB = 8;                        // ring of bit blocks (2 KB each): the expanders never wait for a load
constexpr float K2P_A_VALUE = 2.0f;              // the expanded tiles hold {0, 2.0}: bf16 2.0 = 0x4000 is a single bit (see expand_row)

template <int HP> struct K2PCfg {
  static constexpr bool STACKED = HP <= 64;                         // hi rows 0..HP-1 and lo rows 64..64+HP-1 of ONE 128-row tile
  static constexpr int NP = STACKED ? 6 : 3;                        // ring of skinny-operand tiles: TMA latency (~1.3 us) x MMA rate needs >= 5 in flight
  static constexpr int A_BYTES = 2 * K2P_TILE_BYTES;                // one expanded 256 x 64 adjacency block (N side of the MMA)
  static constexpr int B_BYTES = HP * 64 * 2;                       // one bf16 term of the skinny operand as TMA delivers it
  static constexpr int P_BYTES = STACKED ? K2P_TILE_BYTES : 2 * K2P_TILE_BYTES;   // the 128-row M-side tile(s)
  static constexpr int LO_OFFSET = STACKED ? 64 * 128 : K2P_TILE_BYTES;           // where the lo term lands inside P
  static constexpr int A_OFF = 0, P_OFF = K2P_NA * A_BYTES, BITS_OFF = P_OFF + NP * P_BYTES, BAR_OFF = BITS_OFF + K2P_NB * PK_UNIT_BYTES;
  static constexpr int TMEM_COLS = 512;                             // [2 buffers][256 columns = adjacency rows], lanes = features
  static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 512;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// One row of a unit: the 64 bits (w0 = even columns, w1 = odd columns) -> 64 bf16 {0, 2.0} written as eight 16-byte chunks
// of row `rr` of a K-major SWIZZLE_128B tile (chunk c of row rr sits at chunk position c ^ (rr & 7)).
// bf16 2.0 is the single bit 0x4000, so a cell is a bit deposit: shifting a word by 6 - q puts bits q, q+8, q+16, q+24 at
// bit 6 of its four bytes; one AND with 0x40404040 isolates them; prmt then builds the bf16 pair (2P, 2P+1), P = q + 8 j,
// from byte j of both words (-> bytes 1 and 3) and two zero bytes (the sign-replicate selector on a byte whose bit 7 is
// clear). 6 ALU-pipe instructions per four pairs (the shifts issue as IMAD.SHL on the FMA pipe); a first version that
// produced 1.0 = 0x3F80 needed a second AND per pair and kept the ALU pipe, which tops out at one warp instruction per two
// cycles per scheduler, ~90 % busy at the MMA rate. The factor 2 is removed by the epilogue kernel (x 0.5, exact).
__device__ __forceinline__ void expand_row(uint32_t w0, uint32_t w1, uint8_t* tile, int rr) {
  uint32_t pw[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t a = ((q < 7) ? (w0 << (6 - q)) : (w0 >> 1)) & 0x40404040u;
    const uint32_t b = ((q < 7) ? (w1 << (6 - q)) : (w1 >> 1)) & 0x40404040u;
    pw[q]      = prmt(a, b, 0x4808u);
    pw[q + 8]  = prmt(a, b, 0x5818u);
    pw[q + 16] = prmt(a, b, 0x6828u);
    pw[q + 24] = prmt(a, b, 0x7838u);
  }
  uint8_t* row = tile + rr * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(row + ((c ^ (rr & 7)) << 4)) = make_uint4(pw[4 * c], pw[4 * c + 1], pw[4 * c + 2], pw[4 * c + 3]);
}

// Three decoupled rings: bit blocks (2 KB, 8 deep: bulk copies run far ahead), skinny-operand tiles (TMA), expanded
// adjacency blocks (written by the expanders, read by the MMAs). A first version with ONE ring of (bits, operand, expanded
// block) stages ran the tensor pipe at 50 %: every unit paid the load latency, the expansion and the MMAs back to back in
// the same stage, four stages deep.
template <int HP>
__global__ void __launch_bounds__(K2P_THREADS, 1)
k2p_mma_kernel(const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
               const uint32_t* __restrict__ bits, float* __restrict__ partial, const K2PSched s, const int use_lo, const int b_rank_rows,
               const int dbg) {
  using Cfg = K2PCfg<HP>;
  constexpr int NP = Cfg::NP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* bits_full = bars;                    // [NB] bit block landed (bulk copy)
  uint64_t* bits_empty = bits_full + K2P_NB;     // [NB] expanders have read it (8 warps)
  uint64_t* p_full = bits_empty + K2P_NB;        // [NP] operand terms landed (TMA)
  uint64_t* p_empty = p_full + NP;               // [NP] MMAs that read it retired
  uint64_t* a_ready = p_empty + NP;              // [NA] adjacency block expanded (8 warps)
  uint64_t* a_empty = a_ready + K2P_NA;          // [NA] MMAs that read it retired
  uint64_t* tfull_bar = a_empty + K2P_NA;        // [2] accumulator complete
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int lo = cta * s.per_cta;
  const int hi = min(lo + s.per_cta, s.total);

  if (threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the epilogue kernel may be scheduled (see pdl_prologue)
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_bhi); tma_prefetch_desc(&tm_blo); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < K2P_NB; ++i) { mbar_init(&bits_full[i], 1); mbar_init(&bits_empty[i], K2P_EXP_WARPS); }
      for (int i = 0; i < NP; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], 1); }
      for (int i = 0; i < K2P_NA; ++i) { mbar_init(&a_ready[i], K2P_EXP_WARPS); mbar_init(&a_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");         // everything above overlaps the previous kernel of the step (dependent launch)

  if (warp == 0) {
    // ===== producer of the skinny-operand tiles =====
    if (lane == 0) {
      const uint32_t tx_bytes = Cfg::B_BYTES * (use_lo ? 2 : 1);
      int st = 0; uint32_t phase = 0;
      for (int pos = lo; pos < hi && !(dbg & 8); ++pos) {
        const int sp = pos / s.kblocks, kb = pos - sp * s.kblocks;
        mbar_wait(&p_empty[st], phase ^ 1);
        uint8_t* dst = smem + Cfg::P_OFF + st * Cfg::P_BYTES;
        mbar_expect_tx(&p_full[st], tx_bytes);
        if (b_rank_rows > 0) {                               // operand as gathered: [rank][hi, lo][HP][rows of that rank] (sharded step)
          const int rk = (kb * 64) / b_rank_rows, i0 = kb * 64 - rk * b_rank_rows;
          tma_load_3d(dst, &tm_bhi, &p_full[st], i0, 0, rk);
          if (use_lo) tma_load_3d(dst + Cfg::LO_OFFSET, &tm_blo, &p_full[st], i0, 0, rk);
        } else {
          tma_load_2d(dst, &tm_bhi, &p_full[st], kb * 64, 0);
          if (use_lo) tma_load_2d(dst + Cfg::LO_OFFSET, &tm_blo, &p_full[st], kb * 64, 0);
        }
        if (++st == NP) { st = 0; phase ^= 1; }
      }
    }
  } else if (warp == 14) {
    // ===== producer of the bit blocks =====
    if (lane == 0) {
      int st = 0; uint32_t phase = 0;
      for (int pos = lo; pos < hi && !(dbg & 4); ++pos) {
        mbar_wait(&bits_empty[st], phase ^ 1);
        mbar_expect_tx(&bits_full[st], PK_UNIT_BYTES);
        bulk_load(smem + Cfg::BITS_OFF + st * PK_UNIT_BYTES, bits + (int64_t)pos * PK_UNIT_WORDS, PK_UNIT_BYTES, &bits_full[st]);
        if (++st == K2P_NB) { st = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
      int ps = 0; uint32_t pphase = 0;
      int as = 0; uint32_t aphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int pos = lo; pos < hi;) {
        const int sp = pos / s.kblocks, kb0 = pos - sp * s.kblocks;
        const int cnt = min(s.kblocks - kb0, hi - pos);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);          // the drain warps have emptied this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * 256);
        for (int it = 0; it < cnt; ++it) {
          if (!(dbg & 8)) mbar_wait(&p_full[ps], pphase);    // operand terms landed (async proxy)
          mbar_wait(&a_ready[as], aphase);                   // adjacency block expanded (generic proxy + fence.proxy.async)
          tc_fence_after();
          const uint64_t adj = umma_desc_k_sw128(smem_u32(smem + Cfg::A_OFF + as * Cfg::A_BYTES));      // N side: 256 rows x 64 k
          const uint32_t p_addr = smem_u32(smem + Cfg::P_OFF + ps * Cfg::P_BYTES);
          const uint64_t ph = umma_desc_k_sw128(p_addr);                                                  // M side: [hi; lo] (or hi alone)
          if (!(dbg & 2)) {
#pragma unroll
          for (int k = 0; k < 4; ++k)                        // +32 B per UMMA_K inside the swizzle atom = +2 in the address field
            tc_mma_bf16(d, ph + 2 * k, adj + 2 * k, idesc, (it | k) != 0);
          }
          if (!Cfg::STACKED && use_lo) {
            const uint64_t pl = umma_desc_k_sw128(p_addr + Cfg::LO_OFFSET);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_bf16(d, pl + 2 * k, adj + 2 * k, idesc, 1u);
          }
          tc_commit(&p_empty[ps]);                           // both tiles reusable once these MMAs retire
          tc_commit(&a_empty[as]);
          if (++ps == NP) { ps = 0; pphase ^= 1; }
          if (++as == K2P_NA) { as = 0; aphase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        pos += cnt;
      }
    }
  } else if (warp < 6) {
    // ===== drain: TMEM -> fp32 partial tiles, transposed: [cta][seg][plane][feature][256 rows] =====
    // lane = feature: lanes [0, HP) hold the hi products, lanes [64, 64 + HP) the lo products (stacked operand); HP = 128:
    // one plane, hi and lo already accumulated together.
    const int quarter = warp & 3;                            // TMEM lane quarter this warp may access (warp id mod 4)
    const int m = quarter * 32 + lane;                       // accumulator row
    const int plane = Cfg::STACKED ? (m >> 6) : 0, feat = Cfg::STACKED ? (m & 63) : m;
    const bool useful = feat < HP && (plane == 0 || use_lo);
    constexpr int PLANES = Cfg::STACKED ? 2 : 1;
    int acc = 0; uint32_t acc_phase = 0;
    int seg = 0;
    for (int pos = lo; pos < hi; ++seg) {
      const int sp = pos / s.kblocks, kb0 = pos - sp * s.kblocks;
      pos += min(s.kblocks - kb0, hi - pos);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if ((Cfg::STACKED ? ((quarter & 1) * 32 < HP) : true)) {                       // warp-uniform: this lane quarter holds features
        float* dst = partial + (((int64_t)(cta * s.max_seg + seg) * PLANES + plane) * HP + (useful ? feat : 0)) * 256;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll 4
        for (int c0 = 0; c0 < 256; c0 += 16) {
          uint32_t t16[16];
          tc_ld16(taddr + c0, t16);
          tc_wait_ld();
          if (useful && !(dbg & 16)) {
#pragma unroll
            for (int q = 0; q < 16; q += 4)
              *reinterpret_cast<uint4*>(dst + c0 + q) = make_uint4(t16[q], t16[q + 1], t16[q + 2], t16[q + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 6 + K2P_EXP_WARPS) {
    // ===== expanders: warps 6..13, one row of the unit per thread =====
    const int r = (warp - 6) * 32 + lane;                    // row of the 256-row unit
    int bs = 0; uint32_t bphase = 0;
    int as = 0; uint32_t aphase = 0;
    for (int pos = lo; pos < hi; ++pos) {
      if (!(dbg & 4)) mbar_wait(&bits_full[bs], bphase);     // the bit block has landed
      const uint2 w = *reinterpret_cast<const uint2*>(smem + Cfg::BITS_OFF + bs * PK_UNIT_BYTES + r * 8);
      mbar_wait(&a_empty[as], aphase ^ 1);                   // the MMAs that read this block's previous contents retired
      if (!(dbg & 1)) expand_row(w.x, w.y, smem + Cfg::A_OFF + as * Cfg::A_BYTES, r);
      fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) { mbar_arrive(&a_ready[as]); if (!(dbg & 4)) mbar_arrive(&bits_empty[bs]); }      // (the words have been consumed)
      if (++bs == K2P_NB) { bs = 0; bphase ^= 1; }
      if (++as == K2P_NA) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)Cfg::TMEM_COLS); }
}

// Split-K reduction (fixed order) + row epilogue of one 128-row panel: 512 threads, four per row (lds_epilogue.cuh).
template <int HP, int EPI>
__global__ void __launch_bounds__(512)
k2p_epilogue_kernel(const float* __restrict__ partial, const K2PSched s, const __grid_constant__ EpiArgs ea, const int alt, const int use_lo) {
  __shared__ float red[16][2];
  pdl_prologue();
  constexpr int Q = HP / 4;
  const int etid = threadIdx.x, warp = etid >> 5, lane = etid & 31;
  const int p = blockIdx.x, sp = p >> 1, tile = p & 1;
  if (EPI == K2_EPI_BWD2 && p == 0 && etid == 0) finalize_scalars(ea);     // the layer-2 launch has completed: (loss, acc)
  const int c_first = (sp * s.kblocks) / s.per_cta;
  const int c_last = ((sp + 1) * s.kblocks - 1) / s.per_cta;
  const int row = etid >> 2, g = etid & 3;
  float v[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) v[k] = 0.f;
  // partial tiles are transposed: [cta][seg][plane][feature][256 rows]; for a fixed feature the 8 rows of a warp are one sector
  constexpr bool STACKED = HP <= 64;
  constexpr int PLANES = STACKED ? 2 : 1;
  const int col = tile * 128 + row;
  for (int c = c_first; c <= c_last; ++c) {                  // fixed order: CTA by CTA, hi plane then lo plane
    const int sg = sp - (c * s.per_cta) / s.kblocks;
    const float* base = partial + ((int64_t)(c * s.max_seg + sg) * PLANES * HP + g * Q) * 256 + col;
    float th[Q], tl[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      th[k] = __ldcg(base + (int64_t)k * 256);
      tl[k] = (STACKED && use_lo) ? __ldcg(base + (int64_t)(HP + k) * 256) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) { v[k] += th[k]; if (STACKED && use_lo) v[k] += tl[k]; }
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) v[k] *= 1.0f / K2P_A_VALUE;     // the expanded tiles hold {0, 2.0} (expand_row): exact
  const int i = p * 128 + row;
  if (EPI == K2_EPI_PLAIN) epi_plain<HP>(ea, i, g, v);
  else if (EPI == K2_EPI_LAYER1) epi_layer1<HP>(ea, i, g, v, alt != 0);
  else if (EPI == K2_EPI_BWD2) epi_bwd2<HP>(ea, i, g, v, alt != 0);
  else if (EPI == K2_EPI_BWD1) epi_bwd1<HP>(ea, i, g, v);
  else {
    float li, ci;
    epi_layer2<HP>(ea, i, g, v, li, ci, alt != 0);
    li = warp_sum(li); ci = warp_sum(ci);
    if (lane == 0) { red[warp][0] = li; red[warp][1] = ci; }
    __syncthreads();
    if (etid == 0) {
      float l = 0.f, c = 0.f;
      for (int w = 0; w < 16; ++w) { l += red[w][0]; c += red[w][1]; }       // fixed order
      ea.loss_part[p] = l;
      ea.corr_part[p] = c;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
K2PSched k2p_make_schedule(int n, int rows, int hp) {
  K2PSched s;
  s.hp = hp;
  s.superpanels = pk_superpanels(rows);
  s.kblocks = pk_kblocks(n);
  s.total = s.superpanels * s.kblocks;
  int grid = kNumSMsB200;                                    // a pure function of the shape: workspace sizing needs no device query
  if (grid > s.total) grid = s.total;
  s.per_cta = (int)ceil_div(s.total, grid);
  s.grid = (int)ceil_div(s.total, s.per_cta);
  s.max_seg = (s.per_cta - 1 + s.kblocks - 1) / s.kblocks + 1;
  return s;
}

template <int HP>
static int32_t launch_mma(const CUtensorMap& tbh, const CUtensorMap& tbl, const uint32_t* bits, float* partial, const K2PSched& s,
                          bool use_lo, int b_rank_rows, cudaStream_t stream) {
  static PerDeviceOnce once;
  if (first_use(once)) LDS_CHECK_CUDA(cudaFuncSetAttribute(k2p_mma_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2PCfg<HP>::SMEM_BYTES));
  static const int dbg = getenv("LDS_K2P_DEBUG") ? atoi(getenv("LDS_K2P_DEBUG")) : 0;      // measurement switches: 1 no expansion, 2 no MMAs, 4 no bit loads, 8 no operand loads, 16 no drain stores
  LDS_CHECK_CUDA(launch_dependent(k2p_mma_kernel<HP>, dim3((unsigned)s.grid), dim3(K2P_THREADS), (size_t)K2PCfg<HP>::SMEM_BYTES, stream,
                                  tbh, tbl, bits, partial, s, use_lo ? 1 : 0, b_rank_rows, dbg));
  return LDS_OK;
}

template <int HP>
static int32_t launch_epi(int epi, const float* partial, const K2PSched& s, const EpiArgs& ea, int alt, int use_lo, int panels, cudaStream_t stream) {
  cudaError_t err = cudaSuccess;
  switch (epi) {
    case K2_EPI_PLAIN:  err = launch_dependent(k2p_epilogue_kernel<HP, K2_EPI_PLAIN>, dim3((unsigned)panels), dim3(512), 0, stream, partial, s, ea, alt, use_lo); break;
    case K2_EPI_LAYER1: err = launch_dependent(k2p_epilogue_kernel<HP, K2_EPI_LAYER1>, dim3((unsigned)panels), dim3(512), 0, stream, partial, s, ea, alt, use_lo); break;
    case K2_EPI_LAYER2: err = launch_dependent(k2p_epilogue_kernel<HP, K2_EPI_LAYER2>, dim3((unsigned)panels), dim3(512), 0, stream, partial, s, ea, alt, use_lo); break;
    case K2_EPI_BWD2:   err = launch_dependent(k2p_epilogue_kernel<HP, K2_EPI_BWD2>, dim3((unsigned)panels), dim3(512), 0, stream, partial, s, ea, alt, use_lo); break;
    case K2_EPI_BWD1:   err = launch_dependent(k2p_epilogue_kernel<HP, K2_EPI_BWD1>, dim3((unsigned)panels), dim3(512), 0, stream, partial, s, ea, alt, use_lo); break;
    default: set_error("k2 (packed): unknown epilogue %d", epi); return LDS_ERR_ARG;
  }
  if (err != cudaSuccess) return cuda_fail(err, "k2p_epilogue_kernel");
  return LDS_OK;
}

int32_t k2p_launch(const void* bits, int n, int rows, const void* bt_hi, const void* bt_lo, int64_t ldb, float* partial,
                   const K2PSched& s, bool use_lo, int epi, const EpiArgs& ea, bool alt, cudaStream_t stream, int b_rank_rows) {
  CUtensorMap tbh, tbl;
  int32_t rc;
  if (b_rank_rows > 0) {      // rank-blocked gathered operand: block r = [hi: hp x b_rank_rows][lo: hp x b_rank_rows]
    const int64_t blocks = ceil_div(n, b_rank_rows), bstride = 2 * (int64_t)s.hp * b_rank_rows;
    if ((rc = make_tmap_3d_bf16(&tbh, bt_hi, b_rank_rows, s.hp, blocks, b_rank_rows, bstride, s.hp)) != LDS_OK) return rc;
    if ((rc = make_tmap_3d_bf16(&tbl, bt_lo, b_rank_rows, s.hp, blocks, b_rank_rows, bstride, s.hp)) != LDS_OK) return rc;
  } else {
    if ((rc = make_tmap_2d(&tbh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, bt_hi, n, s.hp, ldb, 64, s.hp)) != LDS_OK) return rc;
    if ((rc = make_tmap_2d(&tbl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, bt_lo, n, s.hp, ldb, 64, s.hp)) != LDS_OK) return rc;
  }
  const uint32_t* b = reinterpret_cast<const uint32_t*>(bits);
  switch (s.hp) {
    case 16:  rc = launch_mma<16>(tbh, tbl, b, partial, s, use_lo, b_rank_rows, stream); break;
    case 32:  rc = launch_mma<32>(tbh, tbl, b, partial, s, use_lo, b_rank_rows, stream); break;
    case 64:  rc = launch_mma<64>(tbh, tbl, b, partial, s, use_lo, b_rank_rows, stream); break;
    case 128: rc = launch_mma<128>(tbh, tbl, b, partial, s, use_lo, b_rank_rows, stream); break;
    default: set_error("k2 (packed): unsupported padded width %d", s.hp); return LDS_ERR_UNSUPPORTED;
  }
  if (rc != LDS_OK) return rc;
  const int panels = (int)ceil_div(rows, 128);
  switch (s.hp) {
    case 16:  return launch_epi<16>(epi, partial, s, ea, alt ? 1 : 0, use_lo ? 1 : 0, panels, stream);
    case 32:  return launch_epi<32>(epi, partial, s, ea, alt ? 1 : 0, use_lo ? 1 : 0, panels, stream);
    case 64:  return launch_epi<64>(epi, partial, s, ea, alt ? 1 : 0, use_lo ? 1 : 0, panels, stream);
    default:  return launch_epi<128>(epi, partial, s, ea, alt ? 1 : 0, use_lo ? 1 : 0, panels, stream);
  }
}

}  // namespace lds

using namespace lds;

// Standalone entry (tests, composable callers): z_out = scale_out * (A_tilde_bits @ (scale_in * p)).
extern "C" int64_t lds_k2_packed_workspace_bytes(int32_t n, int32_t rows, int32_t width) {
  const int hp = k2_padded_width(width);
  if (hp < 0 || n <= 0 || rows <= 0) return -1;
  const K2PSched s = k2p_make_schedule(n, rows, hp);
  return round_up(2 * k2_operand_bytes(n, hp), 1024) + round_up(k2p_partial_bytes(s), 1024) + 1024;
}

extern "C" int32_t lds_k2_propagate_packed(const void* bits, int32_t n, int32_t rows, const float* p, int64_t ld_p, int32_t width,
                                           const float* scale_in, const float* scale_out, float* z_out, int64_t ld_z,
                                           void* workspace, int64_t workspace_bytes, uint32_t flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(bits && p && z_out, "lds_k2_propagate_packed: null pointer");
  LDS_CHECK_ARG(n > 0 && rows > 0 && rows <= n, "lds_k2_propagate_packed: need 0 < rows <= n");
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(bits) & 15) == 0, "lds_k2_propagate_packed: bits must be 16-byte aligned");
  LDS_CHECK_ARG(ld_p >= width && ld_z >= width, "lds_k2_propagate_packed: ld_p / ld_z smaller than width");
  const int hp = k2_padded_width(width);
  if (hp < 0) { set_error("lds_k2_propagate_packed: width %d outside [1, 128]", width); return LDS_ERR_UNSUPPORTED; }
  const int64_t need = lds_k2_packed_workspace_bytes(n, rows, width);
  if (!workspace || workspace_bytes < need) { set_error("lds_k2_propagate_packed: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need); return LDS_ERR_WORKSPACE; }
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "lds_k2_propagate_packed: workspace must be 1024-byte aligned");
  const K2PSched s = k2p_make_schedule(n, rows, hp);
  const int64_t ldb = k2_operand_ld(n);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* bt_hi = ws;
  void* bt_lo = ws + k2_operand_bytes(n, hp);
  float* partial = reinterpret_cast<float*>(ws + round_up(2 * k2_operand_bytes(n, hp), 1024));
  int32_t rc;
  if ((rc = k2_launch_prep(p, ld_p, n, width, hp, scale_in, bt_hi, bt_lo, ldb, nullptr, 0, stream)) != LDS_OK) return rc;
  EpiArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.z_out = z_out; ea.ld_z = ld_z; ea.scale_out = scale_out; ea.rows = rows; ea.width = width;
  return k2p_launch(bits, n, rows, bt_hi, bt_lo, ldb, partial, s, !(flags & LDS_K2_SINGLE_BF16), K2_EPI_PLAIN, ea, false, stream, 0);
}
