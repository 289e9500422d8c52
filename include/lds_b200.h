/* lds_b200.h — C ABI of liblds_b200.so: the B200 (sm_100a) hot path of LDS-GNN's outer step.
 *
 * The reference (andreas-grafberger/lds-gnn) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it offers for this path is its Python API. Every entry point below names
 * the reference function(s) it replaces (paths relative to the reference repo). INTEGRATION.md shows
 * the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the comment says "host". The library never allocates or
 *    frees caller-visible memory and keeps no pointer after a call returns. Scratch is passed in
 *    (`workspace`, `workspace_bytes`) and sized by the matching `*_workspace_bytes` query.
 *  - `stream` is a `cudaStream_t` passed as `void*`. Calls are asynchronous, never synchronise the
 *    host, never allocate, and are CUDA-graph capturable.
 *  - Return value: LDS_OK or an LDS_ERR_* code; `lds_last_error()` gives a thread-local message.
 *    Nothing throws or aborts across the ABI.
 *  - theta is stored as a full symmetric N x ld fp32 matrix (`ld = lds_padded_ld(N)`), unclamped like
 *    the reference's `probs` (src/models/graph.py:60-61); the (T,) upper-triangle vector of the
 *    reference is converted at the API edge (lds_theta_triu_to_full / lds_theta_full_to_triu).
 *  - A_tilde is the sampled adjacency WITH self loops as bf16 {0,1}, N x ld. The normalisation
 *    D^-1/2 A_tilde D^-1/2 (src/utils/graph.py:136-153) is carried as the fp32 vector r = deg^-1/2
 *    and folded into the skinny operands:  A_hat P = r * (A_tilde (r * P)).
 */
#ifndef LDS_B200_H
#define LDS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDS_OK               0
#define LDS_ERR_ARG          1   /* bad shape / alignment / null pointer (mirrors the reference's asserts) */
#define LDS_ERR_CUDA         2   /* a CUDA runtime/driver call failed; message holds the CUDA error string  */
#define LDS_ERR_UNSUPPORTED  3   /* device is not sm_100 or a size is outside the compiled range             */
#define LDS_ERR_WORKSPACE    4   /* workspace too small                                                       */

/* K1 flags */
#define LDS_K1_EXPLICIT_U    1u  /* read uniforms from `u_explicit` (parity mode) instead of Philox           */
/* K2 flags */
#define LDS_K2_SIMT          1u  /* CUDA-core validation kernel instead of the tcgen05 kernel (tests only)    */
#define LDS_K2_SINGLE_BF16   2u  /* operand as one bf16 term instead of the hi+lo split                       */
#define LDS_K2_FORCE_STREAMK 4u  /* split panels across CTAs (stream-K) even when every panel could own a CTA   */
#define LDS_K2_NO_FUSE       8u  /* lds_outer_step: never take the fused small-graph kernel (one launch per stage)   */
#define LDS_K2_DUMP_ADJ     16u  /* lds_outer_step: the fused small-graph kernel also writes A_tilde to the workspace */
#define LDS_K2_BF16_ADJ     64u  /* lds_outer_step: keep A_tilde as bf16 in HBM (the pre-packed launch plan; n <= 8192) instead of bits */
#define LDS_K2_FORWARD_ONLY 32u  /* lds_outer_step: sample + GCN forward + loss/accuracy (+ out_logp) only: no backward,
                                    no update — the evaluation pass of empirical_mean_loss (src/utils/evaluation.py:51-84).
                                    With num_samples = S > 1 (sample_index 0, dropout 0): S graphs in ONE call, drawn at Philox
                                    steps step .. step + S - 1; out_logp is [S][n][c], out_scalars the mean (loss, acc). Small
                                    graphs run all S in one launch and compute the sample-invariant X W0^T + b0 once.      */
/* K3 flags */
#define LDS_K3_DENSE_GRAD    1u  /* write dL/dA_tilde (dense, not symmetrised) instead of updating theta      */
#define LDS_K3_ACCUMULATE    2u  /* with DENSE_GRAD: add into grad_out instead of overwriting                  */
#define LDS_K3_SIMT          4u  /* lds_outer_step: use the CUDA-core update kernel instead of the tcgen05 one   */
/* optimiser kinds (src/models/factory.py:66-69 uses SGD; Adam is the north-star's option)                   */
#define LDS_OPT_SGD          0
#define LDS_OPT_ADAM         1
/* Philox streams */
#define LDS_STREAM_EDGES     0u
#define LDS_STREAM_DROP_X    1u
#define LDS_STREAM_DROP_H    2u

int32_t     lds_version(void);
const char* lds_last_error(void);                 /* host; thread-local, never NULL                        */
int32_t     lds_device_check(void);               /* LDS_OK iff the current device is compute capability 10.x */
int64_t     lds_padded_ld(int32_t n);             /* row stride (elements) used for theta / A_tilde: round_up(n, 64) */

/* Host-side restatement of the device draw, for the oracle: the uniform in [0,1) that element (i, j)
 * of `stream` receives at (seed, step, sample). Edges are keyed on (min(i,j), max(i,j)). */
float lds_philox_uniform(uint64_t seed, uint64_t step, uint32_t stream, uint32_t sample, uint32_t i, uint32_t j);

/* ---- a1/a2: theta layouts. Replaces get_triu_values (src/utils/graph.py:41-45) and the scatter+mirror of
 * triu_values_to_symmetric_matrix (src/utils/graph.py:166-181). `clamp01` applies that function's clamp. */
int32_t lds_theta_triu_to_full(const float* triu, float* theta_full, int64_t ld, int32_t n, int32_t clamp01, void* stream);
/* sym_sum = 0: out[idx(i,j)] = in[i][j];  sym_sum = 1: out[idx(i,j)] = in[i][j] + in[j][i] (i<j), in[i][i] (i=j)
 * — the backward of the mirror (autograd of src/utils/graph.py:35-37 + 176). */
int32_t lds_theta_full_to_triu(const float* full, int64_t ld, float* triu, int32_t n, int32_t sym_sum, void* stream);
/* project_parameters -> ParameterClamper (src/models/graph.py:16-20, 63-64) on the full matrix. */
int32_t lds_theta_clamp(float* theta_full, int64_t ld, int32_t n, void* stream);
/* statistics() (src/models/graph.py:69-78): out4 (device, double[4]) = { sum of clamp(theta_full) over all N*N,
 * sum of theta over the upper triangle incl. diagonal, min, max over that triangle }. */
int32_t lds_theta_stats(const float* theta_full, int64_t ld, int32_t n, double* out4, void* stream);

/* ---- K1 (a2,a3,a5,a6,a7): Bernoulli sample from the upper-triangle draw, mirror, self loops, row-sum
 * degrees, r = deg^-1/2 in one pass over rows [row0, row0+rows) of theta. Replaces
 * BernoulliGraphModel.forward + Sampler.sample/sample_graph (src/models/graph.py:29-32,66-67;
 * src/models/sampling.py:47-85) + add_self_loops/normalize_adjacency_matrix (src/utils/graph.py:123-153).
 * theta_full points at global row `row0` (row-block shard); row0 must be even.
 * a_out   bf16 [rows][ld_a]  A_tilde (diag = 1), columns >= n zeroed.           (may be NULL)
 * sample_out fp32 [rows][ld_s] the raw sample incl. the SAMPLED diagonal (what sample() returns). (may be NULL)
 * deg_out, rsqrt_out fp32 [rows].
 * u_explicit fp32 [n][ld_u] full matrix of uniforms (flag LDS_K1_EXPLICIT_U); element (i,j) uses
 * u[min][max], exactly the reference's "upper-triangle draw wins" (src/models/sampling.py:76). */
int32_t lds_k1_sample_normalize(const float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                uint64_t seed, uint64_t step, uint32_t sample,
                                const float* u_explicit, int64_t ld_u,
                                void* a_out, int64_t ld_a, float* sample_out, int64_t ld_s,
                                float* deg_out, float* rsqrt_out, uint32_t flags, void* stream);

/* Same pass with the Philox step taken from DEVICE memory: step = *step_base + step_offset, read by the kernel. A CUDA graph
 * captured over a whole bilevel block (tau inner steps + the hyper step, src/trainers/bilevel.py:53-73) then draws fresh
 * graphs on every replay; the block bumps the counter itself. Same draws as lds_k1_sample_normalize at that step. */
int32_t lds_k1_sample_normalize_dstep(const float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                      uint64_t seed, const uint64_t* step_base, uint64_t step_offset, uint32_t sample,
                                      void* a_out, int64_t ld_a, float* deg_out, float* rsqrt_out, void* stream);

/* ---- K1 on the bit-packed A_tilde (same reference rows and the same draws as lds_k1_sample_normalize): every Philox block and
 * every theta element of the shard's diagonal block is touched once per UNORDERED 64 x 64 tile pair — tile (I, J) is sampled
 * once and stored with its transpose — and A_tilde leaves as bits: 2 N^2 + N^2 / 8 bytes instead of 6 N^2 (unsharded).
 * Layout of bits_out (lds_packed_adj_bytes(n, rows) bytes, 16-byte aligned, ZERO-FILLED before its first use):
 *   unit (sp, kb) = local rows [256 sp, 256 sp + 256) x columns [64 kb, 64 kb + 64): 2 KB = [256 rows][2 x uint32],
 *   word 0 = the 32 even columns of the block (bit q <-> column 64 kb + 2 q), word 1 = the odd columns; units stored [sp][kb].
 * count_scratch: int32 [rows + 1], zero on entry, zero again on return. row0 must be a multiple of 64. */
int64_t lds_packed_adj_bytes(int32_t n, int32_t rows);
int32_t lds_k1_sample_packed(const float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                             uint64_t seed, uint64_t step, uint32_t sample, const float* u_explicit, int64_t ld_u,
                             void* bits_out, int32_t* count_scratch, float* deg_out, float* rsqrt_out, void* stream);
/* K2 on the packed A_tilde: z_out = scale_out * (A_tilde[rows][n] @ (scale_in * p[n][width])). The bits are expanded to the
 * bf16 {0,1} tiles tcgen05.mma reads inside the kernel, so the product is the one lds_k2_propagate computes while A_tilde
 * costs N^2 / 8 bytes per pass instead of 2 N^2 (tensor-bound at width 64, not HBM-bound). */
int64_t lds_k2_packed_workspace_bytes(int32_t n, int32_t rows, int32_t width);
int32_t lds_k2_propagate_packed(const void* bits, int32_t n, int32_t rows, const float* p, int64_t ld_p, int32_t width,
                                const float* scale_in, const float* scale_out, float* z_out, int64_t ld_z,
                                void* workspace, int64_t workspace_bytes, uint32_t flags, void* stream);

/* ---- K2 (a8's torch.mm(dense_adj, .), src/models/layers.py:44, and its transposes in backward):
 *   z_out[rows][width] = scale_out * ( A[rows][n] @ (scale_in * p[n][width]) )
 * A bf16 (TMA -> tcgen05.mma, fp32 accumulation in TMEM), p split into bf16 hi+lo terms.
 * scale_in [n] / scale_out [rows] may be NULL (= 1). width <= 128. */
int64_t lds_k2_workspace_bytes(int32_t n, int32_t rows, int32_t width);
int32_t lds_k2_propagate(const void* a, int64_t ld_a, int32_t n, int32_t rows,
                         const float* p, int64_t ld_p, int32_t width,
                         const float* scale_in, const float* scale_out,
                         float* z_out, int64_t ld_z,
                         void* workspace, int64_t workspace_bytes, uint32_t flags, void* stream);

/* ---- sparse feature products of the unrolled inner steps: MetaLinear's F.linear(dropout(X), W0) and its autograd transposes
 * (src/models/layers.py:43, src/models/gcn.py:27-28) for bag-of-words X kept in CSR.
 *   y[i][c] = sum_{k in [ptr[i], ptr[i+1])} val[perm ? perm[k] : k] * b[idx[k] * ldb_row + c * ldb_col],  i < rows, c < w
 * (ptr, idx) = CSR of S; for S = X^T pass the CSC arrays of X and `perm` = position of each entry in X's CSR value order, so
 * one (dropout-scaled) value array serves both directions. b may be a strided view (e.g. W0^T: ldb_row = 1, ldb_col = F). */
int32_t lds_spmm_csr(const int32_t* ptr, const int32_t* idx, const float* val, const int32_t* perm, int32_t rows,
                     const float* b, int64_t ldb_row, int64_t ldb_col, int32_t w,
                     float* y, int64_t ldy, void* stream);

/* ---- skinny dense products around the second GCN layer in the unrolled inner steps (MetaLinear of layer_out,
 * src/models/layers.py:43, src/models/gcn.py:29-30, and what autograd derives from it). Closed under differentiation:
 *   lds_row_linear : y[n][j] = sum_k x[n][k] * w[j*sw0 + k*sw1] (+ bias[j]),   k, m <= 128   (X W^T with a tiny, strided W)
 *   lds_gram_tn    : out[i][j] = sum_n a[n][i] * b[n][j],                       a, b <= 128   (weight / bias gradients)
 * lds_gram_tn reduces deterministically (fixed partial-tile order). Its workspace must be zero-filled before the FIRST use
 * (the kernel leaves its arrival counter re-armed) and be 256-byte aligned. */
int32_t lds_row_linear(const float* x, int64_t ldx, int32_t k, const float* w, int64_t sw0, int64_t sw1, int32_t m,
                       const float* bias, float* y, int64_t ldy, int64_t n_rows, void* stream);
int64_t lds_gram_tn_workspace_bytes(int32_t a, int32_t b);
int32_t lds_gram_tn(const float* a_mat, int64_t lda, int32_t a, const float* b_mat, int64_t ldb, int32_t b, int64_t n_rows,
                    float* out, int64_t ldo, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- row-local pieces of the unrolled inner steps (csrc/lds_rowops.cu).
 * Masked NLL of log-softmax ON THE LOGITS z [n][ld_z] (c classes): F.nll_loss(log_softmax(z)[mask], y[mask]) and the accuracy
 * (src/models/gcn.py:34, src/trainers/inner.py:63-66, src/trainers/outer.py:65-67) -> out_loss_acc[2]; its gradient w.r.t. z
 * (dense [n][c], zero outside the mask, scaled by the device scalar *grad_loss); and the backward of that gradient given u =
 * upstream of dz: out_z (may be NULL) and the per-row terms of the gradient w.r.t. *grad_loss (out_g_rows [n], may be NULL; the
 * caller sums them). rows [m] = global rows of the mask (int64), slot [n] = position among them or -1, y [n] labels (int64).
 * lds_row_dot2: out[i] = (<a1_i, b1_i> + <a2_i, b2_i>) / r[i] — the gradient of the normalised propagation w.r.t.
 * r = deg^-1/2 (src/utils/graph.py:148-152), all four [n][ld] with w used columns. */
int32_t lds_masked_nll_forward(const float* z, int64_t ld_z, int32_t c, const int64_t* rows, const int64_t* y, int32_t m,
                               float* out_loss_acc, void* stream);
int32_t lds_masked_nll_grad(const float* z, int64_t ld_z, int32_t c, const int32_t* slot, const int64_t* y, int32_t n, int32_t m,
                            const float* grad_loss, float* dz, int64_t ld_dz, void* stream);
int32_t lds_masked_nll_grad_grad(const float* u, int64_t ld_u, const float* z, int64_t ld_z, int32_t c, const int32_t* slot,
                                 const int64_t* y, int32_t n, int32_t m, const float* grad_loss,
                                 float* out_z, int64_t ld_out, float* out_g_rows, void* stream);
int32_t lds_row_dot2(const float* a1, const float* b1, const float* a2, const float* b2, int64_t ld, int32_t w,
                     const float* r, int32_t n, float* out, void* stream);

/* ---- differentiable Adam of the unrolled inner problem (higher.optim.DifferentiableAdam as the reference drives it,
 * src/trainers/inner.py:6, 42-50, 71; src/trainers/bilevel.py:53-73), over the flat parameter vector, ONE launch per step:
 *   g2 = g + wd p;  m' = m + (1 - b1)(g2 - m);  v' = b2 v + (1 - b2) g2^2;
 *   p' = p - step_size m' / (sqrt(max(v', 1e-30)) root_scale + eps)      step_size = lr / (1 - b1^t), root_scale = 1 / sqrt(1 - b2^t)
 * and the vector-Jacobian product of that elementwise map as one launch (what the hyper step's backward needs of it; the
 * second-order terms of the hypergradient live in g's own graph). step_size / root_scale are taken from device memory when
 * step_size_dev / root_scale_dev are non-NULL (a captured CUDA graph must not bake the step count in), else by value.
 * weight_decay_vec (per-element weight decay: the reference decays layer_in only, inner.py:42-46) overrides weight_decay when non-NULL.
 * Backward: grad_m_out / grad_v_out may be NULL (= 0); any of dp, dm, dv, dg may be NULL (not needed). No aliasing. */
int32_t lds_adam_step(const float* p, const float* m, const float* v, const float* g, int64_t n,
                      float weight_decay, const float* weight_decay_vec, float beta1, float beta2, float eps,
                      float step_size, float root_scale, const float* step_size_dev, const float* root_scale_dev,
                      float* p_out, float* m_out, float* v_out, void* stream);
int32_t lds_adam_step_backward(const float* grad_p_out, const float* grad_m_out, const float* grad_v_out,
                               const float* p, const float* m, const float* v, const float* g, int64_t n,
                               float weight_decay, const float* weight_decay_vec, float beta1, float beta2, float eps,
                               float step_size, float root_scale, const float* step_size_dev, const float* root_scale_dev,
                               float* dp, float* dm, float* dv, float* dg, void* stream);

/* ---- K3+K4 (a10's theta part, a11): closed-form straight-through gradient + optimiser step + projection
 * for rows [row0, row0+rows):   g_ij = fa_i.fb_j + fb_i.fa_j + c_i + c_j  (i != j),  0 on the diagonal,
 * masked where theta is outside [0,1] (clamp backward), then
 *   SGD : theta <- clamp(theta - lr g, 0, 1)                  (src/models/factory.py:66-69, outer.py:78-83)
 *   Adam: torch.optim.Adam defaults on (m, v), step count t (1-based), then the same clamp.
 * fa = r * (dZ1 | dZ2), fb = r * (P1 | P2), each [n][ld_f] fp32 with `d` used columns (d <= 512).
 * With LDS_K3_DENSE_GRAD nothing is updated: grad_out[rows][ld_g] (+)= fa_i.fb_j + c_i (0 on the diagonal),
 * i.e. dL/d(sample) of ONE propagate, for the composable autograd path. */
int32_t lds_k3k4_theta_update(float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                              const float* fa, const float* fb, int64_t ld_f, int32_t d, const float* cvec,
                              float lr, int32_t opt_kind, float* adam_m, float* adam_v,
                              float beta1, float beta2, float eps, int32_t t,
                              float* grad_out, int64_t ld_g, uint32_t flags, void* stream);

/* Tensor-core variant of the SGD update (the one lds_outer_step uses): the rank-2d term runs as a bf16 hi/lo-split
 * GEMM on tcgen05 (fp32 accumulation in TMEM) with theta streamed through shared memory by TMA; same result as
 * the CUDA-core kernel within ~1e-6 relative, and exactly symmetric. workspace: lds_k3_workspace_bytes(n, d). */
int64_t lds_k3_workspace_bytes(int32_t n, int32_t d);
int32_t lds_k3k4_theta_update_tc(float* theta_full, int64_t ld_theta, int32_t n, int32_t row0, int32_t rows,
                                 const float* fa, const float* fb, int64_t ld_f, int32_t d, const float* cvec,
                                 float lr, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fused direct outer step (a4..a11): one OuterProblemTrainer.train_step (src/trainers/outer.py:57-87)
 * with gcn_predict_fct = InnerProblemTrainer.model_forward (src/trainers/inner.py:76-78) at fixed weights,
 * regularize = False. Enqueues K1, the feature GEMM with fused dropout, 4 x K2 with their row epilogues
 * (scaling, relu, dropout, second linear, log-softmax, NLL/accuracy, backward chain), K3+K4.
 * Small graphs (h, c <= 16, CSR features, A_tilde <= ~23 MB, e.g. Cora / Citeseer shape) run everything up to the
 * update as ONE cooperative kernel that keeps A_tilde in shared memory (it never touches HBM), then K3+K4.
 * The workspace must be zero-filled before its FIRST use (the library leaves its counters re-armed after every call). */
typedef struct lds_outer_step_args {
  uint32_t struct_bytes;      /* sizeof(lds_outer_step_args), checked                                   */
  int32_t  n, f, h, c;        /* nodes, features, hidden, classes                                       */
  float*   theta_full;        /* [n][ld_theta] fp32, updated in place                                   */
  int64_t  ld_theta;
  const float*   x;           /* [n][ld_x] fp32, ld_x % 4 == 0 (dense features; may be NULL if x_crow given) */
  int64_t  ld_x;
  const int32_t* x_crow;      /* optional CSR copy of x: [n+1] row offsets, [nnz] columns, [nnz] values.      */
  const int32_t* x_col;       /* Bag-of-words features are ~1% dense; with CSR the feature GEMM gathers rows  */
  const float*   x_val;       /* of w0t instead of streaming N x F zeros. Same result for any x.              */
  const float*   reserved_ptr; /* must be NULL                                                                */
  const float*   w0;          /* [h][ld_w0] fp32 (layer_in.fc.weight), any ld_w0 >= f: staged (padded or      */
  int64_t  ld_w0;             /* transposed) into the workspace by the first kernel of the step               */
  const float*   b0;          /* [h]                                                                    */
  const float*   w1;          /* [c][h] contiguous (layer_out.fc.weight)                                */
  const float*   b1;          /* [c]                                                                    */
  const int64_t* y;           /* [n] labels                                                             */
  const uint8_t* mask;        /* [n] 0/1: rows in the outer objective (opt_mask)                        */
  int32_t  mask_count;        /* number of ones in mask                                                 */
  float    dropout_p;         /* 0 disables dropout (eval mode)                                         */
  uint64_t seed, step;        /* Philox key / step counter                                              */
  const float* u_explicit;    /* optional [n][ld_u] explicit edge uniforms (parity mode), else NULL     */
  int64_t  ld_u;
  const uint8_t* keep_x;      /* optional explicit dropout keep masks [n][f], [n][h] (parity mode)      */
  const uint8_t* keep_h;
  float    lr;                /* learning rate used by this step                                        */
  int32_t  opt_kind;          /* LDS_OPT_*                                                              */
  float*   adam_m; float* adam_v; float beta1, beta2, eps; int32_t adam_t;
  int32_t  update;            /* 1: apply K3+K4; 0: forward + backward factors only                     */
  float*   out_scalars;       /* [4] fp32: loss, accuracy, scalars_tag (see below), reserved             */
  float*   out_logp;          /* optional [n][c] log-probabilities, else NULL                           */
  void*    workspace; int64_t workspace_bytes;
  uint32_t k2_flags;          /* LDS_K2_* for the four propagations                                      */
  uint32_t k3_flags;          /* LDS_K3_SIMT or 0                                                        */
  /* ---- row-block sharding (multi-GPU). rows == 0: the whole matrix, all phases, fields below ignored.
   * With rows > 0 this rank owns global rows [row0, row0+rows): theta_full, x (or its CSR), y, mask, keep_x/keep_h and
   * out_logp are the LOCAL row blocks; every row-local buffer of the workspace holds `rows` rows. A step is run as
   * phases with one exchange between them (the caller all-gathers, e.g. with NCCL):
   *   PHASE_SAMPLE   K1 + feature GEMM            -> local operand rows  (buffer 14, [rows][h]  fp32 = r * P1)
   *   PHASE_LAYER1   needs opnd_full [n][h]       -> local operand rows  (buffer 14, [rows][c]  = r * P2)
   *   PHASE_LAYER2   needs opnd_full [n][c]       -> local operand rows  ([rows][c] = r * dZ2), out_scalars = LOCAL sums / mask_count
   *   PHASE_BWD2     needs opnd_full [n][c]       -> local operand rows  ([rows][h] = r * dZ1)
   *   PHASE_BWD1     needs opnd_full [n][h]       -> local factor rows: packed bf16 rows [row0, row0+rows) of buffer 15 and c
   *                                                  (13) for the tensor-core SGD update, else fp32 rows (buffers 11, 12, 13)
   *   PHASE_UPDATE   needs c_full [n] and f_full [n][lds_outer_step_packed_k(h, c)] bf16 (tensor-core SGD update) or
   *                  fa_full, fb_full [n][ld_f] fp32 (Adam / LDS_K3_SIMT)    -> theta rows updated in place
   * mask_count is the GLOBAL number of masked rows; seed/step must agree on all ranks. No theta / A_tilde traffic. */
  int32_t  row0, rows;
  uint32_t phases;            /* bitmask of LDS_PHASE_*                                                  */
  uint32_t reserved2;
  const float* opnd_full;
  const float* fa_full; const float* fb_full; const float* c_full;
  const void*  f_full;        /* gathered packed factor rows (bf16), see PHASE_UPDATE                    */
  unsigned long long* k2_timeline;   /* optional debug buffer [4][512][8] of %globaltimer stamps, else NULL            */
  /* ---- multi-sample estimator (BASELINE config 3: S Bernoulli samples per outer step). num_samples <= 1: one sample.
   * Otherwise the caller makes S calls with sample_index = 0 .. S-1 and the same (seed, step): each call draws its own graph
   * and dropout masks (Philox sample = sample_index), runs forward + backward and leaves its packed factor rows in column
   * block sample_index of fpack_multi [n][S * lds_outer_step_packed_k(h, c)] bf16 (caller-allocated, 16-byte aligned);
   * c and (loss, acc) are accumulated; the LAST call applies theta <- clamp(theta - lr * mean_s g_s) in ONE update pass
   * (the S rank-2d gradients are concatenated along K of the update GEMM). Tensor-core SGD update, unsharded only. */
  int32_t  num_samples, sample_index;
  void*    fpack_multi;
  /* ---- sharded step, operand exchange in its final layout (optional; opnd_send == NULL keeps the fp32 row exchange).
   * Each phase leaves the NEXT propagation's operand rows of this rank in opnd_send as bf16 [hi: hp x opnd_rank_rows]
   * [lo: hp x opnd_rank_rows] (K-major, hp = padded operand width of the next propagation: 16/32/64/128; rows beyond this
   * rank's row count stay as the caller zero-filled them). The caller all-gathers the first 2 * hp * opnd_rank_rows
   * elements of every rank into opnd_full = [rank][hi, lo][hp][opnd_rank_rows]; the propagation reads it through a 3-D
   * tensor map, so no re-layout kernel runs. opnd_rank_rows = rows per rank (a multiple of 64, the same on every rank). */
  void*    opnd_send;
  int32_t  opnd_rank_rows, reserved3;
  /* ---- early result. scalars_tag != 0 (unsharded calls): once (loss, acc) are final — after the second propagation, i.e.
   * before the backward half and the update have run — the kernel writes them, a system-scope fence, then the tag into
   * out_scalars[2]. With out_scalars in pinned host memory the host can poll for the tag and return the step's metrics while
   * the rest of the step is still executing; everything later stays ordered by the stream. */
  float    scalars_tag; float reserved4;
} lds_outer_step_args;

#define LDS_PHASE_SAMPLE   1u
#define LDS_PHASE_LAYER1   2u
#define LDS_PHASE_LAYER2   4u
#define LDS_PHASE_BWD2     8u
#define LDS_PHASE_BWD1    16u
#define LDS_PHASE_UPDATE  32u
#define LDS_PHASE_ALL     63u

int64_t lds_outer_step_workspace_bytes(int32_t n, int32_t f, int32_t h, int32_t c);
int64_t lds_outer_step_shard_workspace_bytes(int32_t n, int32_t rows, int32_t f, int32_t h, int32_t c);
int32_t lds_outer_step(const lds_outer_step_args* args, void* stream);
/* Launch plan lds_outer_step takes for these arguments: bit 0 = fused small-graph kernel eligible, bit 1 = bit-packed A_tilde
 * (buffer 16, layout above) instead of the bf16 copy (buffer 0). -1: invalid arguments. Host only. */
int32_t lds_outer_step_plan(const lds_outer_step_args* args);
/* Device pointers into a workspace laid out by lds_outer_step (for tests / the composable path / the sharded exchange):
 * which: 0 A_tilde(bf16) 1 deg 2 rsqrt 3 P1 4 Z1 5 P2 6 Z2 7 dZ2 8 dP2 9 dZ1 10 dP1 11 fa 12 fb 13 cvec 14 operand rows
 * 15 packed factor rows (bf16 [n][lds_outer_step_packed_k], this rank's rows at row0). 16 bit-packed A_tilde (lds_k1_sample_packed).
 * The row-local state 3..10 is stored TRANSPOSED, [width][lds_outer_step_state_ld(rows)] fp32 (coalesced for the
 * thread-per-row epilogues); 11/12 are written only when the CUDA-core update runs (Adam or LDS_K3_SIMT).
 * `rows` = n for the unsharded layout. */
void*   lds_outer_step_buffer(void* workspace, int32_t n, int32_t f, int32_t h, int32_t c, int32_t which);
void*   lds_outer_step_shard_buffer(void* workspace, int32_t n, int32_t rows, int32_t f, int32_t h, int32_t c, int32_t which);
int64_t lds_outer_step_factor_ld(int32_t h, int32_t c);
int32_t lds_outer_step_operand_hp(int32_t h, int32_t c, uint32_t phase);   /* padded operand width (16/32/64/128) of a propagation phase */
int64_t lds_outer_step_packed_k(int32_t h, int32_t c);   /* columns of a packed bf16 factor row            */
int64_t lds_outer_step_state_ld(int32_t rows);           /* row stride of the transposed row-local state    */

/* ---- sharded step over NVLink peer memory: all-gather by direct stores. Copies `bytes` from src into
 * dst_bases[p] + dst_offset_bytes for every peer p in [0, world) (device array of `world` P2P-mapped base pointers, this
 * rank included) with 128-bit stores in one kernel; the caller then runs a cross-GPU barrier. 16-byte alignment. */
int32_t lds_peer_push(const void* src, void* const* dst_bases, int32_t world, int64_t dst_offset_bytes, int64_t bytes, void* stream);

/* ---- theta_0 construction on the device (one-off per run in the reference, on the host).
 * lds_knn_graph: sklearn.neighbors.kneighbors_graph(x, k, mode="connectivity", metric, include_self=loop) as the reference
 * calls it (src/data/utils.py:165-183, KNNGraph src/data/transforms.py:15-28): adj_out[i][j] = 1 iff j is one of the k nearest
 * points of i. metric 0 = cosine, 1 = euclidean (= minkowski p = 2). Ties between equidistant neighbours go to the smaller
 * column index (sklearn leaves them unspecified). symmetrize != 0 additionally applies MakeUndirected on the 0/1 matrix:
 * max(A, A^T) (src/data/transforms.py:31-38). x fp32 [n][ld_x]; adj_out fp32 [n][ld_adj], ld_adj >= n (padding columns zeroed);
 * workspace: lds_knn_workspace_bytes(n). k <= 64. */
int64_t lds_knn_workspace_bytes(int32_t n);
int32_t lds_knn_graph(const float* x, int64_t ld_x, int32_t n, int32_t f, int32_t k, int32_t metric, int32_t loop,
                      int32_t symmetrize, float* adj_out, int64_t ld_adj, void* workspace, int64_t workspace_bytes, void* stream);
/* In place A <- max(A, A^T): to_undirected(adj) (src/utils/graph.py:27-34) / MakeUndirected on a dense matrix. */
int32_t lds_symmetrize_max(float* adj, int64_t ld, int32_t n, void* stream);
/* to_dense_adj (src/utils/graph.py:80-116, single graph): adj_out (zero-filled here) [src][dst] = 1 for every edge of
 * edge_index int64 [2][num_edges]; symmetric != 0 also sets [dst][src] (to_undirected on the edge list). *bad_count (device
 * int32) receives the number of edges with an endpoint outside [0, n) — the reference raises an IndexError for those. */
int32_t lds_edges_to_dense(const int64_t* edge_index, int64_t num_edges, int32_t n, int32_t symmetric,
                           float* adj_out, int64_t ld_adj, int32_t* bad_count, void* stream);
/* remove_edges_from_directed_graph / remove_edges_from_undirected_graph (src/data/utils.py:197-227) in two calls:
 * lds_edge_offsets: offsets[i] (device int64 [n+1]) = number of non-zeros in rows < i (of the upper triangle incl. diagonal
 * when triu != 0), offsets[n] = nnz — the position of an edge in Tensor.nonzero() order is offsets[row] + its rank in the row.
 * lds_remove_edges_apply: out (zero-filled here) keeps edge e iff e is among perm[0 .. num_keep) — `perm` is the SAME
 * torch.randperm(nnz) the reference draws, so the retained set is identical; with triu != 0 kept entries are mirrored
 * (to_undirected(from_triu_only=True)). keep_flags: nnz bytes of scratch. out must not alias adj. */
int32_t lds_edge_offsets(const float* adj, int64_t ld, int32_t n, int32_t triu, int64_t* offsets, void* stream);
int32_t lds_remove_edges_apply(const float* adj, int64_t ld, int32_t n, int32_t triu, const int64_t* offsets,
                               const int64_t* perm, int64_t nnz, int64_t num_keep,
                               float* out, int64_t ld_out, uint8_t* keep_flags, void* stream);

/* ---- empirical_mean_loss (src/utils/evaluation.py:51-84): from the [samples][n][c] log-probabilities of a batched
 * evaluation (lds_outer_step with LDS_K2_FORWARD_ONLY), out4 (device) = { mean NLL on mask_a, accuracy on mask_a, mean NLL on
 * mask_b, accuracy on mask_b } — each the mean over the samples of the per-sample masked mean (F.nll_loss / accuracy on
 * predictions[mask], evaluation.py:76-80), arg-max ties to the smaller class like torch.argmax. Deterministic (fixed-order sums).
 * count_a / count_b = number of ones in the masks. workspace: lds_eval_metrics_workspace_bytes(n, samples). */
int64_t lds_eval_metrics_workspace_bytes(int32_t n, int32_t samples);
int32_t lds_eval_metrics(const float* logp, int32_t samples, int32_t n, int32_t c, const int64_t* y,
                         const uint8_t* mask_a, int32_t count_a, const uint8_t* mask_b, int32_t count_b,
                         float* out4, void* workspace, int64_t workspace_bytes, void* stream);

/* Host -> device upload of a step's inputs on a copy stream owned by the library. The reference keeps its GCN weights on the
 * device (src/trainers/inner.py:42-50); a caller whose fast weights arrive from the host every step (bench.py's e2e arm) uploads
 * them with this call instead of a copy on its compute stream: the transfer overlaps the work still executing on `stream`, and
 * everything enqueued on `stream` afterwards is ordered behind it. `dst_device` must not be read or written by work already
 * enqueued (alternate two buffers); `src_pinned_host` must stay unchanged until the copy has run. */
int32_t lds_upload_async(void* dst_device, const void* src_pinned_host, int64_t bytes, void* stream);

/* ---- measurement hook for bench.py (not part of the reference-facing surface). Between begin and end,
 * lds_outer_step records a CUDA event on its stream after every kernel launch. lds_profile_end synchronises on
 * the last event and returns the number of intervals written: ms_out[i] = device time of the launch whose id is
 * ids_out[i] (host, both arrays of capacity `cap`). ids: 0 K1, 1 feature GEMM, 3/4/5/6 the four K2 propagations
 * (layer 1, layer 2, backward 2, backward 1, each with its fused row epilogue), 7 K3+K4, 8 weight staging, 10 the fused
 * small-graph kernel (replaces 8, 0, 1, 3-6). Not graph-capturable while active. */
int32_t lds_profile_begin(void);
int32_t lds_profile_end(float* ms_out, int32_t* ids_out, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* LDS_B200_H */
