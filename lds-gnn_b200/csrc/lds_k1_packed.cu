// lds_k1_packed.cu — K1, tile-symmetric and bit-packed: Bernoulli sample / mirror / self loops / row-sum degrees /
// D^-1/2 for rows [row0, row0 + rows) of theta, writing A_tilde as BITS (layout: lds_packed.cuh).
//
// Same reference semantics as lds_k1_sample.cu (src/models/sampling.py:47-85, src/utils/graph.py:123-153) and the same
// draws: Philox4x32-10 keyed on the canonical (min, max) 2 x 2 block, (w >> 8) < ceil(theta 2^24) in integers. What changes
// is the work decomposition. One warp owns a 64 x 64 tile (I, J): lane l holds the column pair (2l, 2l+1), step t the row
// pair (2t, 2t+1), so one Philox call yields the lane's four cells. Per step, four ballots give the two packed words of both
// rows (even / odd columns); OR-ing the lane's own four results into bit t of four accumulators gives, after 32 steps, the
// two packed words of ITS two columns — i.e. of two rows of the transposed tile. For a tile pair whose rows both belong to
// this shard, tile (I, J) with J > I is therefore sampled once and stored twice: every Philox block and every theta element
// of the shard's diagonal block is touched once per unordered pair (theta is symmetric, so theta(J, I) is never read).
// Algorithmic bytes (unsharded): 2 N^2 (upper triangle of theta) + N^2 / 8 (bits) instead of 6 N^2.
// Tiles that pair with another rank's rows are sampled by both owners from the same counters: no exchange.
// Row sums are integer popcounts accumulated with integer atomics: exact and order-independent.
#include <stdlib.h>
#include "lds_k1_tile.cuh"

namespace lds {

template <bool EXPLICIT_U>
__global__ void __launch_bounds__(K1P_THREADS, 4)
k1p_sample_kernel(const __grid_constant__ K1PArgs a, const __grid_constant__ PhiloxRounds R) {
  if (threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the finalize kernel may be scheduled (pdl_prologue)
  k1p_warp_loop<EXPLICIT_U>(a, R, R.c2, R.c3, (int)(threadIdx.x & 31));
}

// deg = row sum (integer-valued, exact), r = deg^-1/2 with IEEE sqrt / divide like the reference; re-arms the counters.
__global__ void k1p_finalize_kernel(int* __restrict__ cnt, int rows, float* __restrict__ deg, float* __restrict__ rs) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) cnt[rows] = 0;                                  // the chunk ticket
  if (i >= rows) return;
  const float d = (float)cnt[i];
  cnt[i] = 0;
  deg[i] = d;
  rs[i] = 1.0f / sqrtf(d);
}

int32_t k1p_launch(const float* theta, int64_t ldt, int n, int row0, int rows, uint64_t seed, uint64_t step, uint32_t sample,
                   const float* u_explicit, int64_t ldu, uint32_t* bits, int* cnt, float* deg, float* rs, cudaStream_t stream) {
  LDS_CHECK_ARG(theta && bits && cnt && deg && rs, "k1 (packed): null pointer");
  LDS_CHECK_ARG(n > 0 && rows > 0 && row0 >= 0 && row0 + rows <= n, "k1 (packed): rows [%d, %d) outside [0, %d)", row0, row0 + rows, n);
  LDS_CHECK_ARG(row0 % 64 == 0, "k1 (packed): row0 must be a multiple of 64 (got %d)", row0);
  LDS_CHECK_ARG(ldt >= round_up(n, 64) && ldt % 2 == 0 && (reinterpret_cast<uintptr_t>(theta) & 7) == 0,
                "k1 (packed): theta needs ld >= round_up(n, 64) and 8-byte aligned rows");
  K1PArgs a;
  a.theta = theta; a.ldt = ldt; a.n = n; a.row0 = row0; a.rows = rows;
  a.nt = (int)ceil_div(n, 64); a.lo_t = row0 / 64; a.my_tiles = (int)ceil_div(rows, 64);
  a.u = u_explicit; a.ldu = ldu; a.bits = bits; a.kblocks = pk_kblocks(n); a.cnt = cnt;
  a.ticket = reinterpret_cast<unsigned*>(cnt + rows);
  static const int dbg = getenv("LDS_K1P_DEBUG") ? atoi(getenv("LDS_K1P_DEBUG")) : 0;
  a.dbg = dbg; a.chunk = K1P_CHUNK; a.prefetch = 0;
  const PhiloxRounds R = philox_rounds(philox_key(seed, step, LDS_STREAM_EDGES, sample));
  const int64_t total = (int64_t)a.my_tiles * a.nt - (int64_t)a.my_tiles * (a.my_tiles - 1) / 2;
  const int64_t want = ceil_div(total, (int64_t)K1P_CHUNK * (K1P_THREADS / 32));
  const int64_t resident = (int64_t)num_sms() * 4;            // persistent: 4 CTAs per SM (64 registers per thread)
  dim3 grid((unsigned)(want < resident ? want : resident));
  if (u_explicit) k1p_sample_kernel<true><<<grid, K1P_THREADS, 0, stream>>>(a, R);
  else k1p_sample_kernel<false><<<grid, K1P_THREADS, 0, stream>>>(a, R);
  LDS_CHECK_LAUNCH("k1p_sample_kernel");
  LDS_CHECK_CUDA(launch_dependent(k1p_finalize_kernel, dim3((unsigned)ceil_div(rows, 256)), dim3(256), 0, stream, cnt, rows, deg, rs));
  return LDS_OK;
}

}  // namespace lds

using namespace lds;

extern "C" int64_t lds_packed_adj_bytes(int32_t n, int32_t rows) {
  if (n <= 0 || rows <= 0 || rows > n) return -1;
  return pk_bytes(n, rows);
}

extern "C" int32_t lds_k1_sample_packed(const float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                        uint64_t seed, uint64_t step, uint32_t sample, const float* u_explicit, int64_t ld_u,
                                        void* bits_out, int32_t* count_scratch, float* deg_out, float* rsqrt_out, void* stream) {
  if (u_explicit) LDS_CHECK_ARG(ld_u >= n, "lds_k1_sample_packed: explicit uniforms need ld_u >= n");
  return k1p_launch(theta_full, ld_theta, n, row0, rows, seed, step, sample, u_explicit, ld_u, reinterpret_cast<uint32_t*>(bits_out),
                    count_scratch, deg_out, rsqrt_out, (cudaStream_t)stream);
}
