"""Stand-in for torch-scatter (src/utils/graph.py:8,98 — only `to_dense_adj`). TEST INFRASTRUCTURE ONLY."""
import torch


def scatter_add(src, index, dim=0, dim_size=None):
    size = list(src.size())
    size[dim] = int(dim_size if dim_size is not None else int(index.max()) + 1)
    out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.index_add_(dim, index, src)
