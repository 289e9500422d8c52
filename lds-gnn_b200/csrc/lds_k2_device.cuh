// lds_k2_device.cuh — device code shared by the K2 propagate kernel (lds_k2_propagate.cu) and the fused small-graph
// kernel (lds_fused_small.cu): the epilogue-warp loop of one propagation (drain TMEM -> partial tiles, per-panel
// arrival counter, last-arriver reduction + row epilogue).
#pragma once
#include "lds_epilogue.cuh"
#include "lds_tc.cuh"

namespace lds {

struct K2EpiShared {           // static shared memory of the epilogue loop
  int last[2];
  float red[16][2];
};

// Runs on the 16 epilogue warps (threadIdx.x in [64, 576)) of a CTA that owns the linearised (panel, k-block) range
// [lo, hi) of schedule `s`. `acc` / `acc_phase` carry the TMEM accumulator ring state across calls.
// ===== epilogue warps 2..17 (512 threads). Warps 2-5 ("drain", one TMEM lane quarter each) move every finished
// accumulator segment to an fp32 partial tile in global memory (L2) and count the CTA in on the panel; the LAST
// CTA to arrive sums the panel's tiles in a fixed order (deterministic) and runs the row epilogue with all 16
// warps, four threads per row (lds_epilogue.cuh). A panel covered by one CTA takes the same path. =====
// STACKED: the accumulator holds the hi-term and the lo-term products side by side (one MMA of width 2 HP against the
// stacked operand [hi; lo]); the drain adds the two halves (`add_lo`) while it moves them out.
template <int HP, int EPI, bool STACKED = false>
__device__ __forceinline__ void k2_epilogue_loop(const K2Sched& s, const EpiArgs& ea, float* __restrict__ partial, int* __restrict__ counters,
                                                 int cta, int lo, int hi, uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                 int& acc, uint32_t& acc_phase, K2EpiShared& sh, bool add_lo = false, bool alt = false) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    constexpr int Q = HP / 4;
    const int etid = (int)threadIdx.x - 64;                  // 0..511
    const bool drain = warp < 6;
    const int quarter = warp & 3;                            // TMEM lane quarter a drain warp may access
    auto stamp = [&](int k) {
      if (ea.timeline != nullptr && etid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        ea.timeline[(size_t)cta * 8 + k] = t;
      }
    };
    stamp(0);                                                // epilogue warps past the setup barrier
    if (EPI == K2_EPI_BWD2 && cta == 0 && etid == 0) finalize_scalars(ea);   // the layer-2 launch / phase is complete
    int seg = 0;
    for (int pos = lo; pos < hi; ++seg) {
      const int p = pos / s.kblocks, kb0 = pos - p * s.kblocks;
      const int cnt = min(s.kblocks - kb0, hi - pos);
      const int c_first = (p * s.kblocks) / s.per_cta;
      const int c_last = ((p + 1) * s.kblocks - 1) / s.per_cta;
      pos += cnt;
      if (drain) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        stamp(1);                                            // accumulator of this segment complete (all MMAs retired)
        const int row = quarter * 32 + lane;
        float* dst = partial + ((int64_t)(cta * s.max_seg + seg) * K2_BLOCK_M + row) * HP;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * (STACKED ? 2 * HP : HP));
#pragma unroll
        for (int c0 = 0; c0 < HP; c0 += 16) {
          uint32_t t16[16];
          tc_ld16(taddr + c0, t16);
          if (STACKED) {
            uint32_t u16[16];
            tc_ld16(taddr + HP + c0, u16);
            tc_wait_ld();
            if (add_lo) {
#pragma unroll
              for (int q = 0; q < 16; ++q) t16[q] = __float_as_uint(__uint_as_float(t16[q]) + __uint_as_float(u16[q]));
            }
          } else {
            tc_wait_ld();
          }
#pragma unroll
          for (int q = 0; q < 16; q += 4)
            *reinterpret_cast<uint4*>(dst + c0 + q) = make_uint4(t16[q], t16[q + 1], t16[q + 2], t16[q + 3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);        // accumulator drained: the next segment's MMAs may start
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        stamp(2);                                            // partial tile written
        __threadfence();                                     // every writer publishes its part of the tile before the CTA is counted in
        named_bar_sync(1, 128);
        if (etid == 0) {
          const int old = atomicAdd(&counters[p], 1);
          const int last = (old == c_last - c_first) ? 1 : 0;
          if (last) counters[p] = 0;                         // everyone has arrived: re-arm for the next launch
          sh.last[seg & 1] = last;
        }
      }
      named_bar_sync(2, 512);                                // the drain warps' verdict reaches all 16 warps
      stamp(3);                                              // fence + counter round trip done
      if (sh.last[seg & 1] != 0) {                           // uniform over the 512 epilogue threads
        __threadfence();
        const int row = etid >> 2, g = etid & 3;             // four threads per row, a quarter of the columns each
        float v[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) v[k] = 0.f;
        // fixed order => bitwise reproducible; RB contributors' tiles are in flight before any is added
        constexpr int RB = (Q <= 4) ? 8 : (Q <= 8) ? 4 : (Q <= 16) ? 2 : 1;
        const int64_t tile_elems = (int64_t)K2_BLOCK_M * HP;
        const float* rowbase = partial + (int64_t)row * HP + g * Q;
        for (int c = c_first; c <= c_last; c += RB) {
          float4 t[RB][Q / 4];
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int cc = min(c + u, c_last);
            const int sg = p - (cc * s.per_cta) / s.kblocks;
            const float4* src = reinterpret_cast<const float4*>(rowbase + (int64_t)(cc * s.max_seg + sg) * tile_elems);
#pragma unroll
            for (int k = 0; k < Q / 4; ++k) t[u][k] = __ldcg(src + k);
          }
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            if (c + u <= c_last) {
#pragma unroll
              for (int k = 0; k < Q / 4; ++k) {
                v[4 * k] += t[u][k].x; v[4 * k + 1] += t[u][k].y; v[4 * k + 2] += t[u][k].z; v[4 * k + 3] += t[u][k].w;
              }
            }
          }
        }
        stamp(4);                                            // partial tiles reduced
        const int i = p * K2_BLOCK_M + row;
        if (EPI == K2_EPI_PLAIN) epi_plain<HP>(ea, i, g, v);
        else if (EPI == K2_EPI_LAYER1) epi_layer1<HP>(ea, i, g, v, alt);
        else if (EPI == K2_EPI_BWD2) epi_bwd2<HP>(ea, i, g, v, alt);
        else if (EPI == K2_EPI_BWD1) epi_bwd1<HP>(ea, i, g, v);
        else if (EPI == K2_EPI_LAYER2) {
          float li, ci;
          epi_layer2<HP>(ea, i, g, v, li, ci, alt);
          li = warp_sum(li); ci = warp_sum(ci);
          if (lane == 0) { sh.red[warp - 2][0] = li; sh.red[warp - 2][1] = ci; }
          named_bar_sync(2, 512);
          if (etid == 0) {
            float l = 0.f, c = 0.f;
            for (int w = 0; w < 16; ++w) { l += sh.red[w][0]; c += sh.red[w][1]; }   // fixed order
            ea.loss_part[p] = l;
            ea.corr_part[p] = c;
          }
        }
        stamp(5);                                            // row epilogue done
      }
    }
    stamp(6);
  }
}

}  // namespace lds
