// lds_k2_propagate.cu — K2: Z[rows x w] = A_tilde[rows x n] (bf16 {0,1}) @ B[n x w] on the 5th-gen tensor cores.
//
// Replaces the dense `torch.mm(dense_adj, embeddings)` of MetaDenseGraphConvolution.forward
// (reference src/models/layers.py:44) and its transposed products in autograd's backward. The
// normalisation D^-1/2 A D^-1/2 (src/utils/graph.py:150-152, two N^3 SGEMMs in the reference) never
// becomes a GEMM here: r = deg^-1/2 is folded into the skinny operand before and after the product.
//
// Shape of the problem: M = rows (<= N), K = N, "N" = w in {7..128}: arithmetic intensity = w flop/byte,
// far below the B200 ridge (~214 flop/B) => the kernel is an HBM stream of A_tilde; the tensor core is
// used because it is the only unit that consumes 16 KB tiles at stream rate without spending issue slots.
//
// Kernel: persistent stream-K, one CTA per SM slot.
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of A (128 x 64, SWIZZLE_128B) and of the two
//               bf16 terms of the operand (HP x 64, K-major) into a STAGES-deep mbarrier ring
//   warp 1      tcgen05.mma issuer (one lane): D[tmem 128 x HP fp32] += A[smem] x Bhi[smem] (+ Blo), UMMA 128xHPx16,
//               tcgen05.commit -> frees the smem stage / publishes the accumulator; owns the TMEM allocation
//   warps 2-5   drain: tcgen05.ld 32x32b -> fp32 partial tile in global memory (L2), per-panel arrival counter
//   warps 2-17  the last CTA to arrive on a panel reduces its partial tiles (fixed order) and runs the row epilogue,
//               four threads per row (lds_epilogue.cuh)
//   accumulators are double-buffered in TMEM so the next segment's MMAs overlap the epilogue.
#include <stdlib.h>
#include <string.h>
#include "lds_k2_device.cuh"

namespace lds {

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int HP> struct K2Cfg {
  static constexpr int STAGES = (HP == 16) ? 8 : (HP == 32) ? 7 : (HP == 64) ? 6 : 4;
  static constexpr int A_BYTES = K2_BLOCK_M * K2_BLOCK_K * 2;        // 16 KB
  static constexpr int B_BYTES = HP * K2_BLOCK_K * 2;                // one bf16 term
  static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  static constexpr int TMEM_COLS = (2 * HP < 32) ? 32 : 2 * HP;      // two accumulators, power of two >= 32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers + tmem slot*/;
  static constexpr int CTAS_PER_SM = (SMEM_BYTES <= 113 * 1024) ? 2 : 1;
};

template <int HP, int EPI>
__global__ void __launch_bounds__(K2_THREADS, 1)
k2_mma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bhi,
              const __grid_constant__ CUtensorMap tm_blo, float* __restrict__ partial, int* __restrict__ counters,
              const K2Sched s, const int use_lo, const __grid_constant__ EpiArgs ea, const int b_rank_rows) {
  __shared__ K2EpiShared sh_epi;
  using Cfg = K2Cfg<HP>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int lo = cta * s.per_cta;
  const int hi = min(lo + s.per_cta, s.total);

  if (EPI == K2_EPI_PLAIN && threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // stand-alone propagate: see pdl_prologue
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_bhi); tma_prefetch_desc(&tm_blo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
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
  // programmatic dependent launch (lds_k2_propagate: this kernel follows the operand pack): the setup above overlaps the pack;
  // nothing below may start before the pack's writes (operand, re-armed counters) are visible. A plain launch returns at once.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t tx_bytes = Cfg::A_BYTES + Cfg::B_BYTES * (use_lo ? 2 : 1);
      int stage = 0; uint32_t phase = 0;
      for (int pos = lo; pos < hi;) {
        const int p = pos / s.kblocks, kb0 = pos - p * s.kblocks;
        const int cnt = min(s.kblocks - kb0, hi - pos);
        for (int kb = kb0; kb < kb0 + cnt; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_2d(st, &tm_a, &full_bar[stage], kb * K2_BLOCK_K, p * K2_BLOCK_M);
          if (b_rank_rows > 0) {                             // operand as gathered: [rank][HP][rows of that rank] (sharded step)
            const int rk = (kb * K2_BLOCK_K) / b_rank_rows, i0 = kb * K2_BLOCK_K - rk * b_rank_rows;
            tma_load_3d(st + Cfg::A_BYTES, &tm_bhi, &full_bar[stage], i0, 0, rk);
            if (use_lo) tma_load_3d(st + Cfg::A_BYTES + Cfg::B_BYTES, &tm_blo, &full_bar[stage], i0, 0, rk);
          } else {
            tma_load_2d(st + Cfg::A_BYTES, &tm_bhi, &full_bar[stage], kb * K2_BLOCK_K, 0);
            if (use_lo) tma_load_2d(st + Cfg::A_BYTES + Cfg::B_BYTES, &tm_blo, &full_bar[stage], kb * K2_BLOCK_K, 0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        pos += cnt;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(K2_BLOCK_M, HP);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int pos = lo; pos < hi;) {
        const int p = pos / s.kblocks, kb0 = pos - p * s.kblocks;
        const int cnt = min(s.kblocks - kb0, hi - pos);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * HP);
        for (int it = 0; it < cnt; ++it) {
          mbar_wait(&full_bar[stage], phase);                // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(a_addr);
          const uint64_t bhdesc = umma_desc_k_sw128(a_addr + Cfg::A_BYTES);
          const uint64_t bldesc = umma_desc_k_sw128(a_addr + Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < K2_BLOCK_K / 16; ++k)          // +32 B per UMMA_K inside the swizzle atom = +2 in the address field
            tc_mma_bf16(d_tmem, adesc + 2 * k, bhdesc + 2 * k, idesc, (it | k) != 0);
          if (use_lo) {
#pragma unroll
            for (int k = 0; k < K2_BLOCK_K / 16; ++k)
              tc_mma_bf16(d_tmem, adesc + 2 * k, bldesc + 2 * k, idesc, 1u);
          }
          tc_commit(&empty_bar[stage]);                      // smem stage reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);                          // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        pos += cnt;
        (void)p;
      }
    }
  } else {
    // ===== epilogue warps 2..17: drain, per-panel arrival, last-arriver reduction + row epilogue (lds_k2_device.cuh) =====
    int acc = 0; uint32_t acc_phase = 0;
    k2_epilogue_loop<HP, EPI>(s, ea, partial, counters, cta, lo, hi, tmem_base, tfull_bar, tempty_bar, acc, acc_phase, sh_epi);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// operand preparation, partial reduction, validation kernel
// ------------------------------------------------------------------------------------------------
// Bt_hi/lo[c][i] = bf16 split of scale_in[i] * p[i][c]  (transposed: K-major operand), rows c >= width zero.
__global__ void k2_prep_kernel(const float* __restrict__ p, int64_t ld_p, int n, int width, int hp, const float* __restrict__ scale_in,
                               __nv_bfloat16* __restrict__ bt_hi, __nv_bfloat16* __restrict__ bt_lo, int64_t ldb,
                               int* __restrict__ counters, int num_counters) {
  __shared__ float tile[32][33];
  pdl_prologue();                                              // the propagation that follows may set itself up now; wait for our own producer
  if (blockIdx.x == 0 && blockIdx.y == 0)
    for (int k = threadIdx.x; k < num_counters; k += blockDim.x) counters[k] = 0;
  const int i0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 256 threads: ty in [0, 8)
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, c = c0 + tx;
    float v = 0.f;
    if (i < n && c < width) v = p[(int64_t)i * ld_p + c] * (scale_in ? scale_in[i] : 1.f);
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, i = i0 + tx;
    if (c < hp && i < (int)ldb) {
      __nv_bfloat16 h, l;
      split_bf16(tile[tx][r], h, l);
      bt_hi[(int64_t)c * ldb + i] = h;
      bt_lo[(int64_t)c * ldb + i] = l;
    }
  }
}

// CUDA-core validation kernel (tests only): one warp per output row, fp32 FMA over bf16 A.
__global__ void k2_simt_kernel(const __nv_bfloat16* __restrict__ a, int64_t ld_a, int n, int rows,
                               const float* __restrict__ p, int64_t ld_p, int width,
                               const float* __restrict__ scale_in, const float* __restrict__ scale_out,
                               float* __restrict__ z, int64_t ld_z) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  for (int c0 = 0; c0 < width; c0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < n; j += 32) {
      const float av = __bfloat162float(a[(int64_t)row * ld_a + j]);
      if (av != 0.f) {
        const float sj = scale_in ? scale_in[j] : 1.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) if (c0 + e < width) acc[e] = fmaf(av, sj * p[(int64_t)j * ld_p + c0 + e], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = warp_sum(acc[e]);
      if (lane == 0 && c0 + e < width) z[(int64_t)row * ld_z + c0 + e] = v * (scale_out ? scale_out[row] : 1.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int k2_padded_width(int width) {
  if (width <= 0) return -1;
  if (width <= 16) return 16;
  if (width <= 32) return 32;
  if (width <= 64) return 64;
  if (width <= 128) return 128;
  return -1;
}

static int k2_ctas_per_sm(int hp) {
  switch (hp) {
    case 16: return K2Cfg<16>::CTAS_PER_SM;
    case 32: return K2Cfg<32>::CTAS_PER_SM;
    case 64: return K2Cfg<64>::CTAS_PER_SM;
    default: return K2Cfg<128>::CTAS_PER_SM;
  }
}

K2Sched k2_make_schedule(int n, int rows, int hp, bool force_streamk) {
  K2Sched s;
  s.hp = hp;
  s.panels = (int)ceil_div(rows, K2_BLOCK_M);
  s.kblocks = (int)ceil_div(n, K2_BLOCK_K);
  s.total = s.panels * s.kblocks;
  int grid_max = kNumSMsB200 * k2_ctas_per_sm(hp);           // a pure function of the shape: workspace sizing needs no device query
  if (const char* e = getenv("LDS_K2_GRID_MAX")) { const int v = atoi(e); if (v > 0 && v <= grid_max) grid_max = v; }   // tuning knob
  // Stream-K over the linearised (panel, k-block) space. Measured on B200 (Citeseer shape, 26 panels): one CTA per
  // panel streams at only ~35 GB/s per SM (33 us per propagate) — every SM has to pull on the TMA path, so panels are
  // split even when they could each own a CTA (20 us). `force_streamk` is kept for tests of larger splits.
  int per = (int)ceil_div(s.total, grid_max);
  if (per < (force_streamk ? 2 : 4)) per = force_streamk ? 2 : 4;
  if (per > s.total) per = s.total;
  s.per_cta = per;
  s.grid = (int)ceil_div(s.total, per);
  s.max_seg = (per - 1 + s.kblocks - 1) / s.kblocks + 1;
  return s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 2-D row-major tensor map [rows][cols] (row stride ld elements), box = box_rows x box_cols, 128-byte swizzle.
int32_t make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, int elem_bytes, const void* base, int64_t cols, int64_t rows,
                     int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LDS_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dtype, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%lld rows=%lld ld=%lld)", (int)r, (long long)cols, (long long)rows, (long long)ld); return LDS_ERR_CUDA; }
  return LDS_OK;
}

int32_t make_tmap_3d_bf16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t blocks, int64_t row_stride_elems,
                          int64_t block_stride_elems, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LDS_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)blocks};
  cuuint64_t gstr[2] = {(cuuint64_t)row_stride_elems * 2, (cuuint64_t)block_stride_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)K2_BLOCK_K, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (cols=%lld rows=%lld blocks=%lld)", (int)r, (long long)cols, (long long)rows, (long long)blocks); return LDS_ERR_CUDA; }
  return LDS_OK;
}

static int32_t make_tmap_bf16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_rows) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, cols, rows, ld, K2_BLOCK_K, box_rows);
}

template <int HP, int EPI>
static int32_t launch_mma_t(const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tbl, float* partial, int* counters,
                            const K2Sched& s, bool use_lo, const EpiArgs& ea, int b_rank_rows, cudaStream_t stream, bool dependent = false) {
  static PerDeviceOnce once;
  if (first_use(once)) LDS_CHECK_CUDA(cudaFuncSetAttribute(k2_mma_kernel<HP, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2Cfg<HP>::SMEM_BYTES));
  static const bool no_pdl = getenv("LDS_NO_PDL") != nullptr;
  if (dependent && !no_pdl) {                                  // the previous kernel on the stream (operand pack) triggers its dependents
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)s.grid); cfg.blockDim = dim3(K2_THREADS); cfg.dynamicSmemBytes = K2Cfg<HP>::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int ul = use_lo ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, k2_mma_kernel<HP, EPI>, ta, tbh, tbl, partial, counters, s, ul, ea, b_rank_rows) == cudaSuccess) return LDS_OK;
    (void)cudaGetLastError();                                  // fall through to the plain launch
  }
  k2_mma_kernel<HP, EPI><<<s.grid, K2_THREADS, K2Cfg<HP>::SMEM_BYTES, stream>>>(ta, tbh, tbl, partial, counters, s, use_lo ? 1 : 0, ea, b_rank_rows);
  LDS_CHECK_LAUNCH("k2_mma_kernel");
  return LDS_OK;
}

template <int HP>
static int32_t launch_mma_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tbl, float* partial, int* counters,
                              const K2Sched& s, bool use_lo, const EpiArgs& ea, int b_rank_rows, cudaStream_t stream, bool dependent) {
  switch (epi) {
    case K2_EPI_PLAIN:  return launch_mma_t<HP, K2_EPI_PLAIN>(ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream, dependent);
    case K2_EPI_LAYER1: return launch_mma_t<HP, K2_EPI_LAYER1>(ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream);
    case K2_EPI_LAYER2: return launch_mma_t<HP, K2_EPI_LAYER2>(ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream);
    case K2_EPI_BWD2:   return launch_mma_t<HP, K2_EPI_BWD2>(ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream);
    case K2_EPI_BWD1:   return launch_mma_t<HP, K2_EPI_BWD1>(ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream);
  }
  set_error("k2: unknown epilogue %d", epi);
  return LDS_ERR_ARG;
}

int32_t k2_launch_mma(const void* a, int64_t ld_a, int n, int rows, const void* bt_hi, const void* bt_lo, int64_t ldb,
                      float* partial, int* counters, const K2Sched& s, bool use_lo, int epi, const EpiArgs& ea, cudaStream_t stream,
                      int b_rank_rows, bool dependent) {
  CUtensorMap ta, tbh, tbl;
  int32_t rc;
  if ((rc = make_tmap_bf16(&ta, a, n, rows, ld_a, K2_BLOCK_M)) != LDS_OK) return rc;
  if (b_rank_rows > 0) {      // rank-blocked gathered operand: block r = [hi: hp x b_rank_rows][lo: hp x b_rank_rows], bt_hi / bt_lo point into block 0
    const int64_t blocks = ceil_div(n, b_rank_rows), bstride = 2 * (int64_t)s.hp * b_rank_rows;
    if ((rc = make_tmap_3d_bf16(&tbh, bt_hi, b_rank_rows, s.hp, blocks, b_rank_rows, bstride, s.hp)) != LDS_OK) return rc;
    if ((rc = make_tmap_3d_bf16(&tbl, bt_lo, b_rank_rows, s.hp, blocks, b_rank_rows, bstride, s.hp)) != LDS_OK) return rc;
  } else {
  if ((rc = make_tmap_bf16(&tbh, bt_hi, n, s.hp, ldb, s.hp)) != LDS_OK) return rc;
  if ((rc = make_tmap_bf16(&tbl, bt_lo, n, s.hp, ldb, s.hp)) != LDS_OK) return rc;
  }
  switch (s.hp) {
    case 16: return launch_mma_epi<16>(epi, ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream, dependent);
    case 32: return launch_mma_epi<32>(epi, ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream, dependent);
    case 64: return launch_mma_epi<64>(epi, ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream, dependent);
    case 128: return launch_mma_epi<128>(epi, ta, tbh, tbl, partial, counters, s, use_lo, ea, b_rank_rows, stream, dependent);
  }
  set_error("k2: unsupported padded width %d", s.hp);
  return LDS_ERR_UNSUPPORTED;
}

int32_t k2_launch_simt(const void* a, int64_t ld_a, int n, int rows, const float* p, int64_t ld_p, int width,
                       const float* scale_in, const float* scale_out, float* z, int64_t ld_z, cudaStream_t stream) {
  const int warps = 8;
  k2_simt_kernel<<<(int)ceil_div(rows, warps), warps * 32, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a), ld_a, n, rows, p, ld_p, width, scale_in, scale_out, z, ld_z);
  LDS_CHECK_LAUNCH("k2_simt_kernel");
  return LDS_OK;
}

int32_t k2_launch_prep(const float* p, int64_t ld_p, int n, int width, int hp, const float* scale_in,
                       void* bt_hi, void* bt_lo, int64_t ldb, int* counters, int num_counters, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div(ldb, 32), (unsigned)ceil_div(hp, 32));
  LDS_CHECK_CUDA(launch_dependent(k2_prep_kernel, grid, dim3(256), 0, stream, p, ld_p, n, width, hp, scale_in, reinterpret_cast<__nv_bfloat16*>(bt_hi),
                                  reinterpret_cast<__nv_bfloat16*>(bt_lo), ldb, counters, num_counters));
  return LDS_OK;
}

}  // namespace lds

using namespace lds;

extern "C" int64_t lds_k2_workspace_bytes(int32_t n, int32_t rows, int32_t width) {
  const int hp = k2_padded_width(width);
  if (hp < 0 || n <= 0 || rows <= 0) return -1;
  const K2Sched s = k2_make_schedule(n, rows, hp, true);      // sized for the stream-K schedule (the larger one)
  return round_up(2 * k2_operand_bytes(n, hp), 1024) + round_up(k2_partial_bytes(s), 1024) + round_up((int64_t)s.panels * 4, 1024) + 1024;
}

extern "C" int32_t lds_k2_propagate(const void* a, int64_t ld_a, int32_t n, int32_t rows,
                                    const float* p, int64_t ld_p, int32_t width,
                                    const float* scale_in, const float* scale_out,
                                    float* z_out, int64_t ld_z,
                                    void* workspace, int64_t workspace_bytes, uint32_t flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LDS_CHECK_ARG(a && p && z_out, "lds_k2_propagate: null pointer");
  LDS_CHECK_ARG(n > 0 && rows > 0, "lds_k2_propagate: need n > 0 and rows > 0");
  LDS_CHECK_ARG(ld_a >= n && ld_a % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0, "lds_k2_propagate: A must be 16-byte aligned with ld_a >= n, ld_a %% 8 == 0");
  LDS_CHECK_ARG(ld_p >= width && ld_z >= width, "lds_k2_propagate: ld_p / ld_z smaller than width");
  const int hp = k2_padded_width(width);
  if (hp < 0) { set_error("lds_k2_propagate: width %d outside [1, 128]", width); return LDS_ERR_UNSUPPORTED; }
  if (flags & LDS_K2_SIMT) return k2_launch_simt(a, ld_a, n, rows, p, ld_p, width, scale_in, scale_out, z_out, ld_z, stream);
  const int64_t need = lds_k2_workspace_bytes(n, rows, width);
  if (!workspace || workspace_bytes < need) { set_error("lds_k2_propagate: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need); return LDS_ERR_WORKSPACE; }
  LDS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "lds_k2_propagate: workspace must be 1024-byte aligned");
  const K2Sched s_max = k2_make_schedule(n, rows, hp, true);
  const K2Sched s = k2_make_schedule(n, rows, hp, (flags & LDS_K2_FORCE_STREAMK) != 0);
  const int64_t ldb = k2_operand_ld(n);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* bt_hi = ws;
  void* bt_lo = ws + k2_operand_bytes(n, hp);
  float* partial = reinterpret_cast<float*>(ws + round_up(2 * k2_operand_bytes(n, hp), 1024));
  int* counters = reinterpret_cast<int*>(ws + round_up(2 * k2_operand_bytes(n, hp), 1024) + round_up(k2_partial_bytes(s_max), 1024));
  int32_t rc;
  if ((rc = k2_launch_prep(p, ld_p, n, width, hp, scale_in, bt_hi, bt_lo, ldb, counters, s.panels, stream)) != LDS_OK) return rc;
  EpiArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.z_out = z_out; ea.ld_z = ld_z; ea.scale_out = scale_out; ea.rows = rows; ea.width = width;
  return k2_launch_mma(a, ld_a, n, rows, bt_hi, bt_lo, ldb, partial, counters, s, !(flags & LDS_K2_SINGLE_BF16), K2_EPI_PLAIN, ea, stream, 0,
                       /*dependent=*/true);
}
