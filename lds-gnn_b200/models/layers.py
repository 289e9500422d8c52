"""Dense graph-convolution layer with the reference's API (src/models/layers.py:29-44) and the small part
of torchmeta it relies on (MetaLinear / get_subdict: `F.linear` with an overridable weight dict).

`forward(node_features, dense_adj, params)` = A (X W^T + b): the bias is added BEFORE propagation
(layers.py:43-44). `dense_adj` is either a dense tensor (generic path, torch.mm like the reference) or a
`FactoredAdjacency` of a sampled graph, in which case the product runs on the tcgen05 kernel
(K2, csrc/lds_k2_propagate.cu) as r * (A_tilde (r * P)) with a closed-form backward.
"""
import re
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Parameter
from torch.nn.init import xavier_uniform_

from ..utils.graph import FactoredAdjacency
from .sampling import FactoredGraph, FactoredNormalizedAdjacency


def get_subdict(dictionary, key=None):
    """None -> None; otherwise the entries under prefix `key.` with the prefix removed (torchmeta semantics)."""
    if dictionary is None:
        return None
    if not key:
        return dictionary
    pattern = re.compile(r"^{0}\.(.+)".format(re.escape(key)))
    return OrderedDict((pattern.sub(r"\1", k), v) for k, v in dictionary.items() if pattern.match(k) is not None)


class MetaModule(nn.Module):
    """Marker base class: modules whose forward accepts `params=` overrides."""


class SparseFeatureMatrix:
    """CSR + CSC index structure of a static, mostly-zero feature matrix (bag of words, ~1 % dense), built once per tensor.
    The CSC side stores, for every entry, its position in the CSR value order (`perm`), so one per-step value array — the
    dropout-scaled values — serves X' B and X'^T B alike."""

    def __init__(self, x: torch.Tensor):
        csr = x.detach().to(torch.float32).to_sparse_csr()
        self.rows, self.cols = int(x.shape[0]), int(x.shape[1])
        crow, col = csr.crow_indices(), csr.col_indices()
        self.crow, self.col = crow.to(torch.int32).contiguous(), col.to(torch.int32).contiguous()
        self.val = csr.values().to(torch.float32).contiguous()
        row_of = torch.repeat_interleave(torch.arange(self.rows, device=x.device), crow[1:] - crow[:-1])
        order = torch.argsort(col.to(torch.int64) * self.rows + row_of)           # column-major order of the entries
        self.perm = order.to(torch.int32).contiguous()
        self.csc_row = row_of[order].to(torch.int32).contiguous()
        counts = torch.bincount(col.to(torch.int64), minlength=self.cols)
        self.csc_ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)]).to(torch.int32).contiguous()

    def with_values(self, values: torch.Tensor) -> "SparseFeatures":
        return SparseFeatures(self, values)


class SparseFeatures:
    """X' = the static structure + this step's (dropout-scaled) values."""

    def __init__(self, matrix: SparseFeatureMatrix, values: torch.Tensor):
        self.matrix, self.values = matrix, values.contiguous()

    def size(self, dim=None):
        shape = torch.Size((self.matrix.rows, self.matrix.cols))
        return shape if dim is None else shape[dim]

    def product(self, b: torch.Tensor, transposed: bool) -> torch.Tensor:
        from .. import kernels
        m = self.matrix
        if transposed:
            return kernels.spmm_csr(m.csc_ptr, m.csc_row, self.values, m.perm, m.cols, b)
        return kernels.spmm_csr(m.crow, m.col, self.values, None, m.rows, b)


class _SparseProduct(torch.autograd.Function):
    """Y = S B with S = X' (or X'^T) constant: linear in B, so the backward is the transposed product through this same
    Function — differentiable to any order, which the unrolled hypergradient needs (src/trainers/inner.py:71)."""

    @staticmethod
    def forward(ctx, b, feats, transposed):
        ctx.feats, ctx.transposed = feats, transposed
        return feats.product(b, transposed)

    @staticmethod
    def backward(ctx, dy):
        return _SparseProduct.apply(dy, ctx.feats, not ctx.transposed), None, None


class _RowLinear(torch.autograd.Function):
    """Y = X W^T (+ b) for a skinny W (hidden x classes) as one launch (lds_row_linear); closed under differentiation together
    with _GramTN: dX = dY W, dW = dY^T X, db = dY^T 1 — so a double backward (src/trainers/inner.py:71) stays on these two."""

    @staticmethod
    def forward(ctx, x, w, b):
        from .. import kernels
        ctx.save_for_backward(x, w)
        return kernels.row_linear(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = _RowLinear.apply(dy, w.t(), None) if ctx.needs_input_grad[0] else None
        dw = _GramTN.apply(dy, x) if ctx.needs_input_grad[1] else None
        db = _column_sum(dy) if ctx.needs_input_grad[2] else None
        return dx, dw, db


class _GramTN(torch.autograd.Function):
    """G = A^T B, a deterministic reduction over all rows to a few dozen numbers (lds_gram_tn). dA = B dG^T, dB = A dG."""

    @staticmethod
    def forward(ctx, a, b):
        from .. import kernels
        ctx.save_for_backward(a, b)
        return kernels.gram_tn(a, b)

    @staticmethod
    def backward(ctx, dg):
        a, b = ctx.saved_tensors
        da = _RowLinear.apply(b, dg, None) if ctx.needs_input_grad[0] else None
        db = _RowLinear.apply(a, dg.t(), None) if ctx.needs_input_grad[1] else None
        return da, db


_ONES = {}


def _column_sum(y: torch.Tensor) -> torch.Tensor:
    """y.sum(0) as 1^T y through _GramTN (torch's column reduction of an [N, 16] tensor is a 20 us single-block kernel)."""
    key = (y.shape[0], y.device)
    ones = _ONES.get(key)
    if ones is None:
        ones = _ONES[key] = torch.ones((y.shape[0], 1), dtype=torch.float32, device=y.device)
    return _GramTN.apply(ones, y).reshape(-1)


class _BiasAdd(torch.autograd.Function):
    """x + bias with the bias gradient as a _GramTN column sum (any-order differentiable: both directions are linear)."""

    @staticmethod
    def forward(ctx, x, bias):
        return x + bias

    @staticmethod
    def backward(ctx, dy):
        return (dy if ctx.needs_input_grad[0] else None), (_column_sum(dy) if ctx.needs_input_grad[1] else None)


def _skinny_ok(t: torch.Tensor, *widths) -> bool:
    return t.is_cuda and t.dim() == 2 and t.dtype == torch.float32 and all(0 < w <= 128 for w in widths)


SKINNY_LINEAR = [True]         # MetaLinear with in/out features <= 128 on CUDA runs on lds_row_linear / lds_gram_tn
SPARSE_DENSITY = 0.25          # same rule as the fused outer step (kernels.OuterStep.SPARSE_DENSITY)


def sparse_companion(x: torch.Tensor):
    """The SparseFeatureMatrix of a static CUDA feature tensor (cached on the tensor), or None when it is too dense."""
    cached = getattr(x, "_lds_sparse", None)
    if cached is None:
        cached = False
        if x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and not x.requires_grad:
            if int((x != 0).sum().item()) < SPARSE_DENSITY * x.numel():
                cached = SparseFeatureMatrix(x)
        x._lds_sparse = cached
    return cached or None


class MetaLinear(nn.Linear, MetaModule):
    def forward(self, input, params=None):
        if params is None:
            params = OrderedDict(self.named_parameters())
        if isinstance(input, SparseFeatures):
            out = _SparseProduct.apply(params["weight"].t(), input, False)
            bias = params.get("bias", None)
            if bias is None:
                return out
            return _BiasAdd.apply(out, bias) if _skinny_ok(out, out.shape[1]) else out + bias
        weight = params["weight"]
        if SKINNY_LINEAR[0] and _skinny_ok(input, weight.shape[0], weight.shape[1]) and weight.dtype == torch.float32:
            # a skinny layer (hidden x classes): as a matmul its weight gradient is a [C x N] x [N x h] product that cuBLAS
            # runs as a 30 us "large-K" SGEMM for a 6 x 16 result, 16 times per bilevel block; here every order of its backward
            # is one ~3 us launch of lds_row_linear / lds_gram_tn.
            return _RowLinear.apply(input, weight, params.get("bias", None))
        return F.linear(input, weight, params.get("bias", None))


class _Propagate(torch.autograd.Function):
    """Z = A_hat P for a sampled graph in factored form. Backward (A_hat symmetric):
    dP = A_hat dZ;  dL/d(sample)_ij = (r_i dZ_i).(r_j P_j) + c_i (i != j), c = -(rho + kappa) / (2 deg)."""

    @staticmethod
    def forward(ctx, p, graph, handle):
        from .. import kernels
        p = p.contiguous()
        z = kernels.k2_propagate(handle.adj, handle.n, p, handle.rsqrt, handle.rsqrt)
        ctx.save_for_backward(p, z)
        ctx.handle = handle
        return z

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dz):
        from .. import kernels
        p, z = ctx.saved_tensors
        h = ctx.handle
        dz = dz.contiguous()
        dp = kernels.k2_propagate(h.adj, h.n, dz, h.rsqrt, h.rsqrt)
        dgraph = None
        if ctx.needs_input_grad[1]:
            rho = (dz * z).sum(dim=1)
            kappa = (p * dp).sum(dim=1)
            cvec = -(rho + kappa) / (2.0 * h.deg)
            fa = (h.rsqrt[:, None] * dz).contiguous()
            fb = (h.rsqrt[:, None] * p).contiguous()
            dgraph = kernels.k3_dense_grad(h.n, fa, fb, cvec.contiguous())
        return dp, dgraph, None


def propagate(dense_adj, embeddings):
    if isinstance(dense_adj, FactoredNormalizedAdjacency):       # unrolled bilevel loop: differentiable to any order
        return dense_adj.propagate(embeddings)
    if isinstance(dense_adj, FactoredGraph):
        raise NotImplementedError("a FactoredGraph is propagated in normalised form (MetaDenseGCN(normalize_adj=True))")
    if isinstance(dense_adj, FactoredAdjacency):
        return _Propagate.apply(embeddings, dense_adj.graph, dense_adj.handle)
    return torch.mm(dense_adj, embeddings)


class MetaDenseGraphConvolution(MetaModule):
    """Graph convolution on a dense adjacency; weights can be overridden per call (fast weights)."""

    def __init__(self, in_features, out_features, use_bias=True):
        super().__init__()
        self.fc = MetaLinear(in_features, out_features, bias=use_bias)
        self.reset_weights()

    def reset_weights(self):
        """Xavier-uniform weight, zero bias, as fresh Parameters (layers.py:38-40)."""
        self.fc.weight = Parameter(xavier_uniform_(self.fc.weight.clone()))
        self.fc.bias = Parameter(self.fc.bias.clone().zero_())

    def forward(self, node_features, dense_adj, params=None):
        embeddings = self.fc.forward(node_features, params=get_subdict(params, "fc"))
        return propagate(dense_adj, embeddings)
