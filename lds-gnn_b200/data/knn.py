"""kNN initial graph on the device: the reference builds theta_0 for its kNN variants with
`sklearn.neighbors.kneighbors_graph(x, k, mode="connectivity", metric=..., include_self=loop)` on the CPU
(src/data/utils.py:165-175, src/data/transforms.py:15-37) and symmetrises it (`MakeUndirected`). Here the same graph
comes from one similarity GEMM + a row top-k on whatever device `x` lives on (library calls: this is one-off set-up,
not the hot path). Ties between equidistant neighbours are broken by column index (sklearn leaves them unspecified)."""
import torch


def knn_graph_dense(x: torch.Tensor, k: int, loop: bool = True, metric: str = "cosine", block_rows: int = 8192) -> torch.Tensor:
    """Directed kNN connectivity matrix [N, N] (float32 0/1): row i has ones at its k nearest neighbours.
    `loop=True` counts the point itself as its first neighbour (sklearn's include_self=True)."""
    if metric not in ("cosine", "minkowski", "euclidean"):
        raise ValueError(f"unsupported knn metric {metric!r} (cosine / minkowski (p=2) / euclidean)")
    n = x.shape[0]
    if not 0 < k <= n - (0 if loop else 1):
        raise ValueError(f"k={k} out of range for {n} points")
    xf = x.to(torch.float64)
    if metric == "cosine":
        xf = xf / xf.norm(dim=1, keepdim=True).clamp_min(1e-300)
    sq = (xf * xf).sum(1)
    adj = torch.zeros((n, n), dtype=torch.float32, device=x.device)
    for r0 in range(0, n, block_rows):
        blk = xf[r0:r0 + block_rows]
        if metric == "cosine":
            dist = 1.0 - blk @ xf.t()
        else:
            dist = (sq[r0:r0 + block_rows, None] + sq[None, :] - 2.0 * (blk @ xf.t())).clamp_min(0.0)
        rows = torch.arange(r0, r0 + blk.shape[0], device=x.device)
        dist[rows - r0, rows] = -1.0 if loop else float("inf")          # self: always first / never a neighbour
        idx = torch.topk(dist, k, dim=1, largest=False).indices
        adj[rows[:, None].expand_as(idx), idx] = 1.0
    return adj


def knn_init_adjacency(x: torch.Tensor, k: int = 10, metric: str = "cosine", loop: bool = False) -> torch.Tensor:
    """Symmetrised kNN graph used as theta_0 (KNNGraph followed by MakeUndirected): A_ij = max(kNN_ij, kNN_ji)."""
    a = knn_graph_dense(x, k, loop=loop, metric=metric)
    return torch.maximum(a, a.t())
