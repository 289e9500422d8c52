"""theta_0 construction with the reference's function names (src/data/utils.py:134-227, src/utils/graph.py:80-116) on the
device kernels of csrc/lds_theta0.cu: kNN graph, edge list <-> dense adjacency, random edge removal. Inputs must be CUDA
tensors; there is no CPU fallback (the reference's CPU/sklearn path is what these replace)."""
import ctypes
from typing import Optional

import torch

from .. import _lib, kernels

_METRICS = {"cosine": 0, "euclidean": 1, "minkowski": 1, "l2": 1}


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: theta_0 construction runs on the B200 kernels (no CPU fallback)")
    _lib.require_device()


def knn_graph_dense(x: torch.Tensor, k: int, loop: bool = True, metric: str = "cosine", symmetrize: bool = False) -> torch.Tensor:
    """Directed kNN connectivity matrix [N, N] (float32 0/1): row i has ones at its k nearest neighbours; `loop=True` counts
    the point itself as its first neighbour (sklearn's include_self=True). src/data/utils.py:165-175."""
    if metric not in _METRICS:
        raise ValueError(f"unsupported knn metric {metric!r} (cosine / minkowski (p=2) / euclidean)")
    _need_cuda(x, "x")
    n, f = x.shape
    if not 0 < k <= n - (0 if loop else 1):
        raise ValueError(f"k={k} out of range for {n} points")
    xf = x.detach().to(torch.float32)
    xf = xf if xf.stride(1) == 1 else xf.contiguous()
    ld = kernels.padded_ld(n)
    adj = torch.empty((n, ld), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    need = int(lib.lds_knn_workspace_bytes(n))
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    _lib.check(lib.lds_knn_graph(kernels._ptr(xf), xf.stride(0), n, f, int(k), _METRICS[metric], int(bool(loop)), int(bool(symmetrize)),
                                 kernels._ptr(adj), ld, kernels._ptr(ws), need, kernels._stream()), "lds_knn_graph")
    return adj[:, :n]


def knn_init_adjacency(x: torch.Tensor, k: int = 10, metric: str = "cosine", loop: bool = False) -> torch.Tensor:
    """Symmetrised kNN graph used as theta_0 (KNNGraph followed by MakeUndirected, src/data/transforms.py:15-38)."""
    return knn_graph_dense(x, k, loop=loop, metric=metric, symmetrize=True)


def dense_adj_to_edge_index(adj: torch.Tensor) -> torch.Tensor:
    """src/data/utils.py:134-135."""
    return adj.nonzero().t()


def knn_graph(x: torch.Tensor, k: int, loop: bool = True, metric: str = "cosine") -> torch.Tensor:
    """Edge list [2, E] of the directed kNN graph (src/data/utils.py:178-183)."""
    return dense_adj_to_edge_index(knn_graph_dense(x=x, k=k, loop=loop, metric=metric))


def to_dense_adj(edge_index: torch.Tensor, batch=None, edge_attr=None, num_max_nodes: Optional[int] = None, symmetric: bool = False) -> torch.Tensor:
    """Dense float adjacency of ONE graph from its edge list (src/utils/graph.py:80-116 with batch = None, edge_attr = None —
    the only way the LDS pipeline calls it: data/transforms.py:27,37, data/dataloader.py). `symmetric=True` folds
    `to_undirected(edge_index)` (MakeUndirected) into the same scatter."""
    if batch is not None or edge_attr is not None:
        raise NotImplementedError("to_dense_adj: batched graphs / edge attributes are outside the LDS path")
    _need_cuda(edge_index, "edge_index")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape [2, E]")
    ei = edge_index.to(torch.int64).contiguous()
    e = ei.size(1)
    n = int(num_max_nodes) if num_max_nodes else (int(ei.max().item()) + 1 if e else 0)
    if n <= 0:
        return torch.zeros((0, 0), dtype=torch.float32, device=edge_index.device)
    ld = kernels.padded_ld(n)
    adj = torch.empty((n, ld), dtype=torch.float32, device=edge_index.device)
    bad = torch.zeros(1, dtype=torch.int32, device=edge_index.device)
    _lib.check(_lib.load().lds_edges_to_dense(kernels._ptr(ei), e, n, int(bool(symmetric)), kernels._ptr(adj), ld, kernels._ptr(bad),
                                              kernels._stream()), "lds_edges_to_dense")
    if int(bad.item()):
        raise IndexError(f"to_dense_adj: {int(bad.item())} edge(s) reference a node outside [0, {n})")
    return adj[:, :n]


class PytorchSeedOverwrite:
    """src/data/utils.py:230-245: temporarily seed torch's global generator."""

    def __init__(self, seed: Optional[int] = None):
        self.seed = seed

    def __enter__(self):
        if self.seed is not None:
            self.state = torch.random.get_rng_state()
            torch.manual_seed(self.seed)

    def __exit__(self, *exc):
        if self.seed is not None:
            torch.random.set_rng_state(self.state)


def _remove(adj: torch.Tensor, remove_edges_percentage: float, seed, triu: bool) -> torch.Tensor:
    assert 0.0 <= remove_edges_percentage <= 1.0
    assert adj.dim() == 2 and adj.size(0) == adj.size(1)
    _need_cuda(adj, "adj")
    n = adj.size(0)
    a = adj.detach().to(torch.float32)
    a = a if a.stride(1) == 1 else a.contiguous()
    lib = _lib.load()
    off = torch.empty(n + 1, dtype=torch.int64, device=adj.device)
    _lib.check(lib.lds_edge_offsets(kernels._ptr(a), a.stride(0), n, int(triu), kernels._ptr(off), kernels._stream()), "lds_edge_offsets")
    nnz = int(off[n].item())
    num_keep = int(nnz * (1.0 - remove_edges_percentage))
    with PytorchSeedOverwrite(seed):
        perm = torch.randperm(nnz)                      # the host generator, exactly the reference's draw (src/data/utils.py:207-208)
    perm = perm.to(adj.device)
    ld = kernels.padded_ld(n)
    out = torch.empty((n, ld), dtype=torch.float32, device=adj.device)
    flags = torch.empty(max(nnz, 1), dtype=torch.uint8, device=adj.device)
    _lib.check(lib.lds_remove_edges_apply(kernels._ptr(a), a.stride(0), n, int(triu), kernels._ptr(off), kernels._ptr(perm), nnz, num_keep,
                                          kernels._ptr(out), ld, kernels._ptr(flags), kernels._stream()), "lds_remove_edges_apply")
    return out[:, :n].to(adj.dtype)


def remove_edges_from_directed_graph(adj: torch.Tensor, remove_edges_percentage: float, seed: int = None) -> torch.Tensor:
    """src/data/utils.py:197-214."""
    return _remove(adj, remove_edges_percentage, seed, triu=False)


def remove_edges_from_undirected_graph(adj: torch.Tensor, remove_edges_percentage: float, seed: int = None) -> torch.Tensor:
    """src/data/utils.py:217-227: removal on the upper triangle (diagonal included), then mirrored."""
    assert adj.t().equal(adj)
    return _remove(adj, remove_edges_percentage, seed, triu=True)


def remove_edges(dense_adj: torch.Tensor, is_directed: bool, remove_edges_percentage: float, seed: int = None) -> torch.Tensor:
    """src/data/utils.py:186-194."""
    if is_directed:
        return remove_edges_from_directed_graph(dense_adj, remove_edges_percentage, seed=seed)
    return remove_edges_from_undirected_graph(dense_adj, remove_edges_percentage, seed=seed)
