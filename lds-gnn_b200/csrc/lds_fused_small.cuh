// lds_fused_small.cuh — interface of the small-graph fused kernel (lds_fused_small.cu), used by lds_outer_step.cu.
#pragma once
#include "lds_epilogue.cuh"
#include "lds_k1_tile.cuh"

namespace lds {

constexpr int FS_MAX_TILES = 11;     // resident 128 x 64 A tiles per CTA (176 KB of shared memory)
constexpr int FS_HP = 16;            // padded operand width of all four propagations (h <= 16 and C <= 16)

struct FusedSmallArgs {
  // sampling (K1): tile-symmetric, to bits (lds_k1_tile.cuh) — theta, explicit uniforms (parity mode), bits, row counters
  // (int32 [n + 1]: zero on entry, zero again on return; the ticket lives behind them), chunk = 1
  K1PArgs k1; int n;
  PhiloxRounds rounds;                         // edge stream of this (seed, step)
  __nv_bfloat16* a_dump; int64_t lda;          // optional copy of A_tilde in global memory (tests), else NULL
  float* deg; float* rs;
  // feature GEMM
  const int32_t* crow; const int32_t* xcol; const float* xval; int f;
  const float* w0; int64_t ldw; float* w0t; const float* b0; DropCfg dx;
  int rows_per_cta;
  // propagations
  int num_phases;                              // 4: forward + backward; 2: forward only (layer 1, layer 2 + loss)
  int eval_samples;                            // forward only: graphs evaluated in THIS launch (Philox steps step0 .. step0 + S - 1, sample 0);
  unsigned long long step0;                    // the feature rows are computed once, log-probs go to out_logp[s][n][c] (cluster variant only)
  K2Sched s;                                   // panel-aligned: s.kblocks is the PADDED k-range (parts * per_cta), max_seg = 1
  int kb_real;                                 // ceil(n / 64): k-blocks that exist
  int parts;                                   // CTAs per panel (= cluster size of the CLUSTER variant); set by the launcher
  float* partial; int* counters; int use_lo;
  int merged_lo;                               // set by the launcher: the lo term follows the hi term in memory, one TMA box fetches both
  unsigned* gridbar;                           // 16 bytes: {monotonic 64-bit arrival counter of the grid barrier, arrivals of the staged-weight split barrier (zero between launches), spare}
  unsigned long long* timeline;                // optional debug: [grid][16] %globaltimer stamps per CTA, else NULL
  EpiArgs ea;
};

// Panel-aligned schedule of the fused kernel: every CTA owns `per_cta` consecutive k-blocks of ONE 128-row panel.
// Returns false when the graph does not fit (more than FS_MAX_TILES tiles per CTA on 148 SMs) or h, c > 16.
bool fused_small_schedule(int n, int hp1, int hp2, K2Sched& s, int& kb_real);
// LDS_ERR_UNSUPPORTED when the device cannot run the cooperative launch: the caller uses the multi-kernel path.
// allow_cluster: use the thread-block-cluster variant (split-K reduction over distributed shared memory) when the device can
// keep one cluster per panel resident.
int32_t fused_small_launch(const FusedSmallArgs& fa, cudaStream_t stream, bool allow_cluster = true);

}  // namespace lds
