"""Dense graph tensor utilities with the reference's names and semantics (src/utils/graph.py).

Two kinds of functions live here:
  * on the hot path — `triu_values_to_symmetric_matrix` (a2) runs the CUDA layout kernel and
    `normalize_adjacency_matrix` (a7) keeps a sampled graph in factored form (bf16 A_tilde + r = deg^-1/2)
    so the propagation runs on the tcgen05 kernel; both need liblds_b200 and a B200;
  * generic dense helpers (`to_undirected`, `get_triu_values`, `add_self_loops`, `split_mask`, and
    normalisation of an arbitrary dense matrix) written with elementwise torch ops — O(N^2), never the
    reference's N x N x N diagonal matmuls (src/utils/graph.py:150-152) — and differentiable to any order.
"""
from math import sqrt
from typing import Tuple, Union

import numpy as np
import torch
from torch import Tensor


class DenseData:
    """Attribute bag standing in for the reference's `DenseData(torch_geometric.data.Data)`
    (src/utils/graph.py:15-24): x, y, dense_adj, edge_index, train/val/test masks, num_classes, name."""

    def __init__(self, **kwargs):
        self.dense_adj = None
        self.train_mask = None
        self.val_mask = None
        self.test_mask = None
        self.num_classes = -1
        self.name = ""
        for key, value in kwargs.items():
            setattr(self, key, value)

    def to(self, device):
        for key, value in list(vars(self).items()):
            if torch.is_tensor(value):
                setattr(self, key, value.to(device))
        return self

    @property
    def num_features(self):
        return self.x.size(1)

    @property
    def num_nodes(self):
        return self.x.size(0)


def is_square_matrix(tensor: Tensor) -> bool:
    return tensor.dim() == 2 and tensor.size(0) == tensor.size(1)


def num_nodes_from_triu_shape(n_triu_values: int) -> int:
    """N from T = N(N+1)/2, the reference's formula (src/utils/graph.py:184-192)."""
    return int(0.5 * sqrt((8 * n_triu_values + 1) - 1))


def to_undirected(adj: Tensor, from_triu_only: bool = False) -> Tensor:
    """max(A, A^T), or mirror the strict upper triangle and keep the diagonal (src/utils/graph.py:27-38)."""
    assert is_square_matrix(adj)
    if not from_triu_only:
        return torch.max(adj, adj.t())
    upper = adj.triu(1)
    return upper + upper.t() + torch.diag(adj.diag())


def get_triu_values(adj: Tensor) -> Tensor:
    """Row-major upper triangle incl. the diagonal (src/utils/graph.py:41-45)."""
    assert adj.size(0) == adj.size(1)
    rows, cols = torch.triu_indices(adj.size(0), adj.size(0), device=adj.device)
    return adj[rows, cols]


def split_mask(mask: Tensor, ratio: float = 0.5, shuffle: bool = True,
               device: Union[str, torch.device] = "cpu") -> Tuple[Tensor, Tensor]:
    """Split a boolean node mask in two (src/utils/graph.py:48-76); shuffling uses numpy's global RNG."""
    chosen = mask.nonzero().flatten()
    if shuffle:
        order = np.arange(chosen.numel())
        np.random.shuffle(order)
        chosen = chosen[torch.as_tensor(order, device=chosen.device)]
    cut = int(chosen.numel() * ratio)
    first = torch.zeros_like(mask, dtype=torch.bool, device=device)
    second = torch.zeros_like(mask, dtype=torch.bool, device=device)
    first[chosen[:cut]] = True
    second[chosen[cut:]] = True
    return first, second


def add_self_loops(adj: Tensor) -> Tensor:
    """Clone with the diagonal set to 1; the diagonal receives no gradient (src/utils/graph.py:123-133)."""
    assert is_square_matrix(adj)
    looped = adj.clone()
    looped.fill_diagonal_(1.0)
    return looped


# ------------------------------------------------------------------------------------------ hot path: a2
class _TriuToSymmetric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, triu_values):
        from .. import kernels
        full = kernels.theta_triu_to_full(triu_values.detach(), clamp=True)
        n = full.shape[0]
        ctx.save_for_backward(triu_values)
        ctx.n = n
        out = full[:, :n]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        from .. import kernels
        (triu_values,) = ctx.saved_tensors
        g = kernels.theta_full_to_triu(grad.contiguous(), ctx.n, sym_sum=True)      # mirror backward
        inside = (triu_values >= 0.0) & (triu_values <= 1.0)                          # clamp backward
        return g * inside


def triu_values_to_symmetric_matrix(triu_values: Tensor) -> Tensor:
    """(T,) -> symmetric (N, N) clamped to [0,1] (src/utils/graph.py:166-181), one CUDA pass, no index tensors."""
    assert triu_values.dim() == 1
    out = _TriuToSymmetric.apply(triu_values)
    out._lds_symmetric = True
    return out


# ------------------------------------------------------------------------------------------ hot path: a7
class FactoredAdjacency:
    """D^-1/2 (A+I) D^-1/2 of a sampled graph kept as (bf16 A_tilde, r): what the tcgen05 propagate consumes."""

    def __init__(self, graph: Tensor, handle):
        self.graph = graph          # the dense sample tensor (autograd link to theta)
        self.handle = handle        # models.sampling.SampleHandle


def normalize_adjacency_matrix(dense_adj: Tensor, materialize: bool = True):
    """GCN normalisation (src/utils/graph.py:136-153). For a graph produced by `Sampler.sample` and
    `materialize=False` the result stays factored; otherwise a dense tensor is returned, computed with
    O(N^2) elementwise ops (r_i * A_ij * r_j) instead of the reference's two N^3 products."""
    assert is_square_matrix(dense_adj)
    handle = getattr(dense_adj, "_lds_handle", None)
    if handle is not None and not materialize:
        return FactoredAdjacency(dense_adj, handle)
    looped = add_self_loops(dense_adj)
    inv_sqrt = 1.0 / looped.sum(dim=1).sqrt()
    return inv_sqrt[:, None] * looped * inv_sqrt[None, :]
