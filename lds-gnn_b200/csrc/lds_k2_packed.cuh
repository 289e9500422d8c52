// lds_k2_packed.cuh — internal interface of the packed propagation (lds_k2_packed.cu), used by lds_outer_step.cu.
#pragma once
#include "lds_epilogue.cuh"
#include "lds_packed.cuh"

namespace lds {

// Stream-K over the linearised (super-panel, k-block) space, super-panel-major; CTA c owns units [c per_cta, (c+1) per_cta).
// A CTA writes one transposed fp32 partial block ([hi, lo planes][HP features][256 rows]; one plane for HP = 128) per
// super-panel its range touches into slot (cta * max_seg + segment); the epilogue kernel sums a panel's slots in CTA order.
struct K2PSched {
  int hp;            // padded operand width (16/32/64/128)
  int superpanels;   // ceil(rows / 256)
  int kblocks;       // ceil(n / 64)
  int total;         // superpanels * kblocks
  int per_cta;       // units per CTA
  int grid;          // CTAs launched
  int max_seg;       // partial-tile pairs a CTA can write
};

K2PSched k2p_make_schedule(int n, int rows, int hp);
static inline int64_t k2p_partial_bytes(const K2PSched& s) { return (int64_t)s.grid * s.max_seg * (s.hp <= 64 ? 2 : 1) * s.hp * 256 * 4; }

// bits: packed A_tilde of `rows` local rows (lds_packed.cuh); bt_hi / bt_lo: K-major operand terms [hp][ldb] (or, with
// b_rank_rows > 0, the gathered rank-blocked operand of the sharded step). Enqueues the MMA kernel and the epilogue kernel.
int32_t k2p_launch(const void* bits, int n, int rows, const void* bt_hi, const void* bt_lo, int64_t ldb, float* partial,
                   const K2PSched& s, bool use_lo, int epi, const EpiArgs& ea, bool alt, cudaStream_t stream, int b_rank_rows = 0);

}  // namespace lds
