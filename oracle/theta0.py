"""oracle/theta0.py — TEST INFRASTRUCTURE. CPU restatement of the reference's theta_0 construction.

  knn_connectivity      sklearn.neighbors.kneighbors_graph exactly as src/data/utils.py:165-175 calls it (sklearn is the
                        reference's own dependency for this step and is present in the image)
  knn_distances         float64 distance matrix of the same metric: kNN graphs are compared "up to ties" against it
  to_dense_adj          src/utils/graph.py:80-116 for one graph without edge attributes
  remove_edges_*        src/data/utils.py:197-227 (torch CPU ops, same generator calls: manual_seed + randperm)
tests/test_oracle.py checks these against the live reference's functions whenever /root/reference exists.
"""
import numpy as np
import torch


def knn_connectivity(x, k, metric="cosine", loop=True):
    from sklearn.neighbors import kneighbors_graph
    return kneighbors_graph(np.asarray(x), n_neighbors=k, mode="connectivity", metric=metric, include_self=loop).toarray().astype(np.float32)


def knn_distances(x, metric="cosine"):
    x = np.asarray(x, dtype=np.float64)
    if metric == "cosine":
        norm = np.sqrt((x * x).sum(1, keepdims=True))
        xn = np.divide(x, norm, out=np.zeros_like(x), where=norm > 0)
        return 1.0 - xn @ xn.T
    sq = (x * x).sum(1)
    return np.maximum(sq[:, None] + sq[None, :] - 2.0 * (x @ x.T), 0.0)


def to_dense_adj(edge_index, num_nodes):
    adj = np.zeros((num_nodes, num_nodes), dtype=np.float32)
    adj[np.asarray(edge_index[0]), np.asarray(edge_index[1])] = 1.0
    return adj


def remove_edges_from_directed_graph(adj, remove_edges_percentage, seed=None):
    """src/data/utils.py:197-214."""
    adj = torch.as_tensor(adj)
    nonzero_indices = adj.nonzero()
    num_edges = nonzero_indices.size(0)
    num_keep = int(num_edges * (1.0 - remove_edges_percentage))
    state = torch.random.get_rng_state()
    if seed is not None:
        torch.manual_seed(seed)
    perm = torch.randperm(num_edges)
    if seed is not None:
        torch.random.set_rng_state(state)
    keep = nonzero_indices.t()[:, perm[:num_keep]]
    new_adj = torch.zeros_like(adj)
    new_adj[keep[0], keep[1]] = adj[keep[0], keep[1]]
    return new_adj


def remove_edges_from_undirected_graph(adj, remove_edges_percentage, seed=None):
    """src/data/utils.py:217-227 (+ to_undirected(from_triu_only=True), src/utils/graph.py:35-37)."""
    adj = torch.as_tensor(adj)
    removed = remove_edges_from_directed_graph(adj.clone().triu(), remove_edges_percentage, seed=seed)
    upper = removed.triu(1)
    return upper + upper.t() + torch.diag(removed.diag())
