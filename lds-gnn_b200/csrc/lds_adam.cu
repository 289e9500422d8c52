// lds_adam.cu — one step of the differentiable Adam of the unrolled inner problem as ONE launch, and its backward as one.
//
// Reference: the inner optimiser of LDS is `higher.optim.DifferentiableAdam` (src/trainers/inner.py:6, 42-50, 71), i.e.
// torch.optim.Adam's rule applied out of place so that the hyper step can differentiate through tau of them
// (src/trainers/bilevel.py:53-73). Over the flat parameter vector (W0 | b0 | W1 | b1, ~60 k floats at Citeseer shape):
//     g2 = g + wd p;   m' = m + (1 - b1)(g2 - m);   v' = b2 v + (1 - b2) g2^2         (wd per element: weight decay on layer_in only)
//     p' = p - step_size * m' / (sqrt(max(v', 1e-30)) * root_scale + eps)
// step_size = lr / (1 - b1^t), root_scale = 1 / sqrt(1 - b2^t) (the bias correction scales sqrt(v) BEFORE eps is added).
// As ATen ops this is 12 launches forward and ~22 in the backward of the hyper step, per inner step — a fifth of the ~870
// kernels of a captured bilevel block, each 1-2 us on 60 k elements. The map (p, m, v, g) -> (p', m', v') is elementwise, so
// its vector-Jacobian product is elementwise too (the second-order terms of the hypergradient live in g's own graph, not here):
//     dm't = gm' - gp' step / den;            dden = gp' step m' / den^2
//     dv't = gv' + dden root_scale / (2 root)       (0 where v' was floored)
//     dg2  = (1 - b1) dm't + 2 (1 - b2) g2 dv't;     dm = b1 dm't;   dv = b2 dv't;   dg = dg2;   dp = gp' + wd dg2
// step_size / root_scale come by value or from device memory (a captured block reads them from a table, trainers/diffopt.py).
#include "lds_common.cuh"

namespace lds {

struct AdamHyper { float wd, b1, b2, eps, step_size, root_scale; const float* step_size_dev; const float* root_scale_dev;
                   const float* wd_vec; };      // per-element weight decay (parameter groups), else NULL: `wd` for every element

__global__ void __launch_bounds__(256)
adam_step_kernel(const float* __restrict__ p, const float* __restrict__ m, const float* __restrict__ v, const float* __restrict__ g, int64_t n,
                 AdamHyper hp, float* __restrict__ p_out, float* __restrict__ m_out, float* __restrict__ v_out) {
  pdl_prologue();
  const float step = hp.step_size_dev ? *hp.step_size_dev : hp.step_size;
  const float rs = hp.root_scale_dev ? *hp.root_scale_dev : hp.root_scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = p[i], mi = m[i], vi = v[i];
    const float wd = hp.wd_vec ? hp.wd_vec[i] : hp.wd;
    const float g2 = (wd != 0.f) ? __fadd_rn(g[i], __fmul_rn(wd, pi)) : g[i];
    const float mn = __fadd_rn(mi, __fmul_rn(1.f - hp.b1, __fsub_rn(g2, mi)));
    const float vn = __fadd_rn(__fmul_rn(vi, hp.b2), __fmul_rn(__fmul_rn(1.f - hp.b2, g2), g2));
    const float root = sqrtf(fmaxf(vn, 1e-30f));
    const float den = __fadd_rn(__fmul_rn(root, rs), hp.eps);
    p_out[i] = __fsub_rn(pi, __fmul_rn(step, __fdiv_rn(mn, den)));
    m_out[i] = mn;
    v_out[i] = vn;
  }
}

__global__ void __launch_bounds__(256)
adam_step_backward_kernel(const float* __restrict__ gp, const float* __restrict__ gm, const float* __restrict__ gv,
                          const float* __restrict__ p, const float* __restrict__ m, const float* __restrict__ v, const float* __restrict__ g,
                          int64_t n, AdamHyper hp, float* __restrict__ dp, float* __restrict__ dm, float* __restrict__ dv, float* __restrict__ dg) {
  pdl_prologue();
  const float step = hp.step_size_dev ? *hp.step_size_dev : hp.step_size;
  const float rs = hp.root_scale_dev ? *hp.root_scale_dev : hp.root_scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = p[i], mi = m[i], vi = v[i];
    const float wd = hp.wd_vec ? hp.wd_vec[i] : hp.wd;
    const float g2 = (wd != 0.f) ? __fadd_rn(g[i], __fmul_rn(wd, pi)) : g[i];
    const float mn = __fadd_rn(mi, __fmul_rn(1.f - hp.b1, __fsub_rn(g2, mi)));
    const float vn = __fadd_rn(__fmul_rn(vi, hp.b2), __fmul_rn(__fmul_rn(1.f - hp.b2, g2), g2));
    const bool floored = !(vn > 1e-30f);
    const float root = sqrtf(fmaxf(vn, 1e-30f));
    const float den = __fadd_rn(__fmul_rn(root, rs), hp.eps);
    const float gpi = gp ? gp[i] : 0.f;
    const float sd = step / den;
    const float dmt = (gm ? gm[i] : 0.f) - gpi * sd;
    const float dden = gpi * sd * (mn / den);
    const float dvt = (gv ? gv[i] : 0.f) + (floored ? 0.f : dden * rs * (0.5f / root));
    const float dg2 = (1.f - hp.b1) * dmt + 2.f * (1.f - hp.b2) * g2 * dvt;
    if (dm) dm[i] = hp.b1 * dmt;
    if (dv) dv[i] = hp.b2 * dvt;
    if (dg) dg[i] = dg2;
    if (dp) dp[i] = gpi + wd * dg2;
  }
}

static int adam_grid(int64_t n) {
  const int64_t want = ceil_div(n, (int64_t)256);
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace lds

using namespace lds;

extern "C" int32_t lds_adam_step(const float* p, const float* m, const float* v, const float* g, int64_t n,
                                 float weight_decay, const float* weight_decay_vec, float beta1, float beta2, float eps,
                                 float step_size, float root_scale, const float* step_size_dev, const float* root_scale_dev,
                                 float* p_out, float* m_out, float* v_out, void* stream) {
  LDS_CHECK_ARG(p && m && v && g && p_out && m_out && v_out, "lds_adam_step: null pointer");
  LDS_CHECK_ARG(n > 0, "lds_adam_step: n must be positive");
  LDS_CHECK_ARG((step_size_dev == nullptr) == (root_scale_dev == nullptr), "lds_adam_step: step_size_dev and root_scale_dev come together");
  AdamHyper hp{weight_decay, beta1, beta2, eps, step_size, root_scale, step_size_dev, root_scale_dev, weight_decay_vec};
  LDS_CHECK_CUDA(launch_dependent(adam_step_kernel, dim3((unsigned)adam_grid(n)), dim3(256), 0, (cudaStream_t)stream, p, m, v, g, n, hp, p_out, m_out, v_out));
  return LDS_OK;
}

extern "C" int32_t lds_adam_step_backward(const float* grad_p_out, const float* grad_m_out, const float* grad_v_out,
                                          const float* p, const float* m, const float* v, const float* g, int64_t n,
                                          float weight_decay, const float* weight_decay_vec, float beta1, float beta2, float eps,
                                          float step_size, float root_scale, const float* step_size_dev, const float* root_scale_dev,
                                          float* dp, float* dm, float* dv, float* dg, void* stream) {
  LDS_CHECK_ARG(p && m && v && g, "lds_adam_step_backward: null pointer");
  LDS_CHECK_ARG(n > 0, "lds_adam_step_backward: n must be positive");
  LDS_CHECK_ARG((step_size_dev == nullptr) == (root_scale_dev == nullptr), "lds_adam_step_backward: step_size_dev and root_scale_dev come together");
  AdamHyper hp{weight_decay, beta1, beta2, eps, step_size, root_scale, step_size_dev, root_scale_dev, weight_decay_vec};
  LDS_CHECK_CUDA(launch_dependent(adam_step_backward_kernel, dim3((unsigned)adam_grid(n)), dim3(256), 0, (cudaStream_t)stream, grad_p_out, grad_m_out, grad_v_out, p, m, v, g, n, hp, dp, dm, dv, dg));
  return LDS_OK;
}
