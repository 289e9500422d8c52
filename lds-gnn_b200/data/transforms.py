"""Graph transforms with the reference's class names (src/data/transforms.py:15-55) for the theta_0 pipeline:
`KNNGraph`, `MakeUndirected`, `RemoveEdges`. They act on a `DenseData`-like object (x, edge_index, dense_adj, num_nodes) whose
tensors live on the CUDA device and return a shallow copy with `edge_index` and `dense_adj` replaced."""
import copy as _copy
from typing import Optional

from ..utils.tracking import setup_basic_logger
from .utils import dense_adj_to_edge_index, knn_graph_dense, remove_edges, to_dense_adj

logger = setup_basic_logger()


class Transform:
    def __call__(self, data):
        raise NotImplementedError

    def __repr__(self):
        return f"{self.__class__.__name__}()"


def _clone(data):
    return data.clone() if hasattr(data, "clone") else _copy.copy(data)


class KNNGraph(Transform):
    def __init__(self, loop: bool, k: int, metric: str = "cosine"):
        self.k, self.loop, self.metric = k, loop, metric

    def __call__(self, data):
        logger.info(f"Constructing knn-graph with k={self.k}, self-loop={self.loop}")
        out = _clone(data)
        out.dense_adj = knn_graph_dense(data.x, self.k, loop=self.loop, metric=self.metric)
        out.edge_index = dense_adj_to_edge_index(out.dense_adj)
        return out


class MakeUndirected(Transform):
    def __call__(self, data):
        logger.info("Making graph undirected (if not already)")
        out = _clone(data)
        out.dense_adj = to_dense_adj(data.edge_index, num_max_nodes=data.num_nodes, symmetric=True)
        out.edge_index = dense_adj_to_edge_index(out.dense_adj)
        return out


class RemoveEdges(Transform):
    def __init__(self, remove_edges_percentage: float, seed: Optional[int] = None):
        assert 0.0 <= remove_edges_percentage <= 1.0
        self.remove_edges_percentage, self.seed = remove_edges_percentage, seed

    def __call__(self, data):
        assert getattr(data, "dense_adj", None) is not None
        logger.info(f"Using {(1.0 - self.remove_edges_percentage) * 100}% of original edges")
        out = _clone(data)
        directed = not bool(data.dense_adj.t().equal(data.dense_adj))
        out.dense_adj = remove_edges(data.dense_adj, directed, self.remove_edges_percentage, seed=self.seed)
        out.edge_index = dense_adj_to_edge_index(out.dense_adj)
        return out
