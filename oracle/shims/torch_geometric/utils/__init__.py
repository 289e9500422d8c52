def to_undirected(*a, **k):
    raise RuntimeError("torch_geometric shim: not available")


def to_scipy_sparse_matrix(*a, **k):
    raise RuntimeError("torch_geometric shim: not available")
