"""kNN initial graph — kept as an import path; the implementation lives in data/utils.py
(device kernels of csrc/lds_theta0.cu, reference src/data/utils.py:165-183)."""
from .utils import knn_graph, knn_graph_dense, knn_init_adjacency      # noqa: F401
