"""`DataFactory` with the reference's ingredient and config keys (src/data/dataloader.py:12-53) over the synthetic datasets of
this package (the Planetoid / UCI loaders are host-side preprocessing outside the hot path).

`DataFactory.load()` returns host data exactly like the reference, and the reference's scripts continue with
`.to(device)` (src/scripts/bilevel.py:74): the graph transforms the config asks for (KNNGraph, MakeUndirected, RemoveEdges —
theta_0 construction) run on the device kernels, so they are applied when the data arrives there."""
from typing import List

from ..config import Ingredient
from ..utils.graph import DenseData
from .synthetic import SHAPES, make_dataset
from .transforms import KNNGraph, MakeUndirected, RemoveEdges, Transform


class DeviceTransformedData(DenseData):
    """DenseData whose pending graph transforms run on the first `.to(<cuda device>)`."""

    def to(self, device):
        super().to(device)
        pending = self.__dict__.pop("_pending", [])
        import torch
        if pending and torch.device(device).type != "cuda":
            raise RuntimeError("the configured graph transforms (kNN graph / MakeUndirected / RemoveEdges) run on the B200 kernels: "
                               "move the data to a CUDA device (there is no CPU fallback)")
        data = self
        for transform in pending:
            data = transform(data)
        if data is not self:
            self.__dict__.update(vars(data))
        return self

    def clone(self):
        out = DeviceTransformedData()
        out.__dict__.update(vars(self))
        return out


def create_transformations(remove_edges_percentage: float, make_undirected: bool, nearest_neighbor_k: int, knn_metric: str,
                           seed: int = None) -> List[Transform]:
    """The graph part of the reference's chain, in its order (src/data/dataloader.py:100-117)."""
    transforms: List[Transform] = []
    if nearest_neighbor_k:
        transforms.append(KNNGraph(k=nearest_neighbor_k, loop=False, metric=knn_metric))
    if make_undirected:
        transforms.append(MakeUndirected())
    if remove_edges_percentage:
        transforms.append(RemoveEdges(remove_edges_percentage=remove_edges_percentage, seed=seed))
    return transforms


class DataFactory:
    _data_ingredient = Ingredient("data")
    INGREDIENTS = {"data": _data_ingredient}

    @staticmethod
    @_data_ingredient.config
    def _data_config():
        dataset: str = "cora"                    # noqa: F841  (a synthetic graph of that dataset's shape)
        remove_edges_percentage: float = 0.0     # noqa: F841
        normalize_features: bool = True          # noqa: F841  (the synthetic features are row-normalised already)
        shuffle_splits: bool = True              # noqa: F841  (the synthetic splits are random already)
        make_undirected: bool = True             # noqa: F841
        nearest_neighbor_k: int = None           # noqa: F841
        use_largest_subgraph: bool = False       # noqa: F841
        split_seed: int = None                   # noqa: F841
        knn_metric: str = "cosine"               # noqa: F841

    @staticmethod
    @_data_ingredient.capture
    def load(dataset: str, remove_edges_percentage: float, normalize_features: bool, shuffle_splits: bool, make_undirected: bool,
             nearest_neighbor_k: int, use_largest_subgraph: bool, knn_metric: str, split_seed: int) -> DenseData:
        if dataset not in SHAPES:
            raise NotImplementedError(f"dataset {dataset!r}: this package ships synthetic graphs of the shapes {sorted(SHAPES)}")
        if use_largest_subgraph:
            raise NotImplementedError("LargestSubgraph (src/data/transforms.py:72-83) is host-side preprocessing outside the LDS path")
        base = make_dataset(dataset, seed=0 if split_seed is None else int(split_seed))
        data = DeviceTransformedData()
        data.__dict__.update(vars(base))
        data._pending = create_transformations(remove_edges_percentage, make_undirected, nearest_neighbor_k, knn_metric, seed=split_seed)
        return data
