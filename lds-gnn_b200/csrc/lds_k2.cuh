// lds_k2.cuh — internal interface of the K2 propagation (shared by lds_k2_propagate.cu and lds_outer_step.cu).
#pragma once
#include "lds_common.cuh"

namespace lds {

constexpr int K2_BLOCK_M = 128;      // rows of A_tilde per output tile (= UMMA M, cta_group::1)
constexpr int K2_BLOCK_K = 64;       // bf16 elements per k-block = one 128-byte swizzle atom
constexpr int K2_THREADS = 576;      // warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-17 epilogue (2-5 drain TMEM)

// Stream-K schedule: the (panel, k-block) space is linearised (panel-major) and cut into equal contiguous
// ranges, one per CTA. A CTA writes one fp32 partial tile (128 x HP) per panel its range touches into
// slot (cta * max_seg + segment); the last CTA to finish a panel sums its slots in a fixed order (deterministic).
struct K2Sched {
  int hp;          // padded operand width (16/32/64/128)
  int panels;      // ceil(rows / 128)
  int kblocks;     // ceil(n / 64)
  int total;       // panels * kblocks
  int per_cta;     // k-blocks per CTA
  int grid;        // CTAs launched
  int max_seg;     // partial tiles a CTA can write
};

int k2_padded_width(int width);                        // 16/32/64/128, or -1
K2Sched k2_make_schedule(int n, int rows, int hp, bool force_streamk = false);
static inline int64_t k2_operand_ld(int n) { return round_up(n, K2_BLOCK_K); }
static inline int64_t k2_operand_bytes(int n, int hp) { return (int64_t)hp * k2_operand_ld(n) * 2; }            // one bf16 term
static inline int64_t k2_partial_bytes(const K2Sched& s) { return (int64_t)s.grid * s.max_seg * K2_BLOCK_M * s.hp * 4; }

struct EpiArgs;   // lds_epilogue.cuh

// Enqueue the tcgen05 kernel: A[rows][n] (bf16, ld_a) x Bt (bf16 hi/lo terms, [hp][ldb], K-major), then the row
// epilogue `epi` (K2Epi) on every completed 128-row panel. `counters`: one zero-initialised int per panel.
int32_t k2_launch_mma(const void* a, int64_t ld_a, int n, int rows, const void* bt_hi, const void* bt_lo, int64_t ldb,
                      float* partial, int* counters, const K2Sched& s, bool use_lo, int epi, const EpiArgs& ea, cudaStream_t stream,
                      int b_rank_rows = 0, bool dependent = false);   // > 0: the operand is the gathered rank-blocked array [rank][hi, lo][hp][b_rank_rows] (sharded step)
// Operand preparation: bt_hi/lo[c][i] = bf16 split of scale_in[i] * p[i][c] (transposed, K-major), c < hp; zeroes the counters.
int32_t k2_launch_prep(const float* p, int64_t ld_p, int n, int width, int hp, const float* scale_in,
                       void* bt_hi, void* bt_lo, int64_t ldb, int* counters, int num_counters, cudaStream_t stream);
// CUDA-core validation kernel producing the same partial layout (slot 0 of each panel's first CTA only is NOT used:
// it writes complete sums into z directly). Tests only.
int32_t k2_launch_simt(const void* a, int64_t ld_a, int n, int rows, const float* p, int64_t ld_p, int width,
                       const float* scale_in, const float* scale_out, float* z, int64_t ld_z, cudaStream_t stream);

}  // namespace lds
