// lds_packed.cuh — the bit-packed A_tilde: layout shared by the packed sampling kernel (lds_k1_packed.cu), the packed
// propagation (lds_k2_packed.cu) and lds_outer_step.cu.
//
// A_tilde is a {0,1} matrix. As bf16 it costs 2 N^2 bytes to write and 2 N^2 bytes per propagation to read — 10 N^2 of the
// 22 N^2 bytes of an outer step. Packed it is N^2 / 8 bytes, L2-resident up to N ~ 30 000, and the propagation becomes
// tensor-bound instead of HBM-bound: its producer warps expand 256 x 64-bit blocks into the K-major SWIZZLE_128B bf16 tiles
// tcgen05.mma reads (the arithmetic is unchanged: the same exact bf16 {0,1} operand).
//
// Layout (local rows of a row-block shard, r = 0 .. rows-1; kb = column / 64):
//   unit (sp, kb) = rows [256 sp, 256 sp + 256) x columns [64 kb, 64 kb + 64)  -> 2 KB contiguous: [256 rows][2 words]
//   word 0 of a row = the 32 EVEN columns of the block (bit q <-> column 64 kb + 2 q), word 1 = the ODD columns.
//   units are stored [sp][kb] (a CTA's k-range of one super-panel is contiguous: one bulk copy per unit).
// The even/odd split is what a Philox 2 x 2 block produces naturally with one lane per column pair (two ballots per row),
// and the expander undoes it for free (it picks bit q of both words for the bf16 pair (2q, 2q + 1)).
// Rows >= n of the last super-panel and columns >= n hold zero bits.
#pragma once
#include "lds_common.cuh"

namespace lds {

constexpr int PK_SP_ROWS = 256;                     // rows of a super-panel (two 128-row MMA tiles share one operand stage)
constexpr int PK_KB_COLS = 64;                      // columns of a k-block
constexpr int PK_UNIT_WORDS = PK_SP_ROWS * 2;       // 512 words = 2 KB
constexpr int PK_UNIT_BYTES = PK_UNIT_WORDS * 4;

static inline int pk_superpanels(int rows) { return (int)ceil_div(rows, PK_SP_ROWS); }
static inline int pk_kblocks(int n) { return (int)ceil_div(n, PK_KB_COLS); }
static inline int64_t pk_bytes(int n, int rows) { return (int64_t)pk_superpanels(rows) * pk_kblocks(n) * PK_UNIT_BYTES; }

// word index of (local row r, k-block kb, word w)
__host__ __device__ __forceinline__ int64_t pk_word(int r, int kb, int kblocks, int w) {
  return (((int64_t)(r >> 8) * kblocks + kb) * PK_SP_ROWS + (r & 255)) * 2 + w;
}

// Packed sampling pass (lds_k1_packed.cu): bits, integer row counts -> deg, r = deg^-1/2. `cnt` (int32 [rows + 1]) must be
// zero on entry and is zero again on return (the finalize kernel re-arms it).
int32_t k1p_launch(const float* theta, int64_t ldt, int n, int row0, int rows, uint64_t seed, uint64_t step, uint32_t sample,
                   const float* u_explicit, int64_t ldu, uint32_t* bits, int* cnt, float* deg, float* rs, cudaStream_t stream);

}  // namespace lds
