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


class MetaLinear(nn.Linear, MetaModule):
    def forward(self, input, params=None):
        if params is None:
            params = OrderedDict(self.named_parameters())
        return F.linear(input, params["weight"], params.get("bias", None))


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
