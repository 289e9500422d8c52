"""-m gpu tests of theta_0 construction on the device (csrc/lds_theta0.cu behind lds_gnn_b200/data/utils.py and
data/transforms.py) against the CPU restatement of the reference (oracle/theta0.py: sklearn's kneighbors_graph exactly as
src/data/utils.py:165-175 calls it, to_dense_adj src/utils/graph.py:80-116, remove_edges* src/data/utils.py:186-227)."""
import numpy as np
import pytest
import torch

from oracle import theta0 as T0

pytestmark = pytest.mark.gpu
CUDA = "cuda"


def bag_of_words(rng, n, f, density):
    x = (rng.random((n, f)) < density).astype(np.float32)
    x[np.arange(n), rng.integers(0, f, n)] = 1.0
    return x / x.sum(1, keepdims=True)


@pytest.mark.parametrize("metric,loop,k", [("cosine", False, 10), ("cosine", True, 10), ("minkowski", False, 20), ("cosine", False, 1), ("minkowski", True, 33)])
def test_knn_graph_matches_sklearn_on_generic_data(metric, loop, k):
    """Continuous features: no ties, the neighbour sets must be identical to sklearn's."""
    from lds_gnn_b200.data.utils import knn_graph_dense, knn_init_adjacency
    rng = np.random.default_rng(k)
    x = rng.random((333, 37)).astype(np.float32)
    ref = T0.knn_connectivity(x, k, metric, loop)
    got = knn_graph_dense(torch.as_tensor(x).to(CUDA), k, loop=loop, metric=metric).cpu().numpy()
    assert np.array_equal(got, ref)
    sym = knn_init_adjacency(torch.as_tensor(x).to(CUDA), k, metric=metric, loop=loop).cpu().numpy()
    assert np.array_equal(sym, np.maximum(ref, ref.T))


@pytest.mark.parametrize("n,f,density,metric", [(2708, 1433, 0.0127, "cosine"), (1000, 300, 0.02, "minkowski"), (65, 3, 0.5, "cosine")])
def test_knn_graph_on_bag_of_words_is_a_knn_graph_up_to_ties(n, f, density, metric):
    """Binary bag-of-words rows produce many exactly equal distances: sklearn leaves the choice among tied candidates
    unspecified, so the check is the definition — every row has exactly k ones, none on the diagonal, and every chosen
    neighbour is at most as far as the k-th nearest (float64 distances, 1e-6 slack for the fp32 arithmetic)."""
    from lds_gnn_b200.data.utils import knn_graph_dense
    rng = np.random.default_rng(n)
    x = bag_of_words(rng, n, f, density)
    k = 10
    got = knn_graph_dense(torch.as_tensor(x).to(CUDA), k, loop=False, metric=metric).cpu().numpy()
    assert set(np.unique(got)) <= {0.0, 1.0} and np.all(got.sum(1) == k) and np.all(np.diag(got) == 0)
    d = T0.knn_distances(x, metric)
    np.fill_diagonal(d, np.inf)
    kth = np.sort(d, axis=1)[:, k - 1]
    worst_chosen = np.where(got > 0, d, -np.inf).max(1)
    assert np.all(worst_chosen <= kth + 1e-6 * np.maximum(1.0, np.abs(kth)))
    # rows whose k-th and (k+1)-th distances are well separated must agree with sklearn exactly
    ref = T0.knn_connectivity(x, k, metric, False)
    gap = np.sort(d, axis=1)[:, k] - kth
    clear = gap > 1e-5
    assert clear.sum() > 0 and np.array_equal(got[clear], ref[clear])


def test_to_dense_adj_and_make_undirected():
    from lds_gnn_b200.data.transforms import MakeUndirected
    from lds_gnn_b200.data.utils import dense_adj_to_edge_index, to_dense_adj
    from lds_gnn_b200.utils.graph import DenseData
    rng = np.random.default_rng(1)
    n = 301
    ei = np.stack([rng.integers(0, n, 900), rng.integers(0, n, 900)])
    ref = T0.to_dense_adj(ei, n)
    got = to_dense_adj(torch.as_tensor(ei).to(CUDA), num_max_nodes=n)
    assert got.shape == (n, n) and np.array_equal(got.cpu().numpy(), ref)
    assert np.array_equal(to_dense_adj(torch.as_tensor(ei).to(CUDA)).cpu().numpy(), ref[:ei.max() + 1, :ei.max() + 1])
    data = DenseData(x=torch.zeros((n, 2), device=CUDA), edge_index=torch.as_tensor(ei).to(CUDA))
    und = MakeUndirected()(data)
    assert np.array_equal(und.dense_adj.cpu().numpy(), np.maximum(ref, ref.T))
    assert torch.equal(und.edge_index, dense_adj_to_edge_index(und.dense_adj)) and data.dense_adj is None
    with pytest.raises(IndexError):
        to_dense_adj(torch.tensor([[0, 5], [1, 2]], device=CUDA), num_max_nodes=4)


@pytest.mark.parametrize("pct", [0.0, 0.25, 0.5, 0.9, 1.0])
@pytest.mark.parametrize("seed", [0, 11])
def test_remove_edges_keeps_exactly_the_reference_edge_set(pct, seed):
    from lds_gnn_b200.data.transforms import RemoveEdges
    from lds_gnn_b200.data.utils import remove_edges_from_directed_graph, remove_edges_from_undirected_graph
    from lds_gnn_b200.utils.graph import DenseData
    rng = np.random.default_rng(seed)
    n = 257
    a = (rng.random((n, n)) < 0.05).astype(np.float32) * (0.25 + rng.random((n, n)).astype(np.float32))
    sym = np.maximum(a, a.T)
    np.fill_diagonal(sym, (rng.random(n) < 0.3) * 0.5)
    ref = T0.remove_edges_from_undirected_graph(sym, pct, seed=seed).numpy()
    state = torch.random.get_rng_state()
    got = remove_edges_from_undirected_graph(torch.as_tensor(sym).to(CUDA), pct, seed=seed)
    assert torch.equal(torch.random.get_rng_state(), state)              # the seed override restores the generator
    assert np.array_equal(got.cpu().numpy(), ref) and np.array_equal(ref, ref.T)
    refd = T0.remove_edges_from_directed_graph(a, pct, seed=seed).numpy()
    assert np.array_equal(remove_edges_from_directed_graph(torch.as_tensor(a).to(CUDA), pct, seed=seed).cpu().numpy(), refd)
    data = DenseData(x=torch.zeros((n, 2), device=CUDA), dense_adj=torch.as_tensor(sym).to(CUDA))
    out = RemoveEdges(pct, seed=seed)(data)
    assert np.array_equal(out.dense_adj.cpu().numpy(), ref)
    kept = int((np.triu(ref) != 0).sum())
    assert kept == int(int((np.triu(sym) != 0).sum()) * (1.0 - pct))


def test_knn_theta0_feeds_the_graph_model():
    """The config-3 pipeline end to end: KNNGraph -> MakeUndirected -> BernoulliGraphModel(dense_adj) (models/factory.py:60-62)."""
    from lds_gnn_b200.data.transforms import KNNGraph, MakeUndirected
    from lds_gnn_b200.models.graph import BernoulliGraphModel
    from lds_gnn_b200.utils.graph import DenseData
    rng = np.random.default_rng(5)
    x = torch.as_tensor(rng.random((200, 16)).astype(np.float32)).to(CUDA)
    data = MakeUndirected()(KNNGraph(loop=False, k=10, metric="cosine")(DenseData(x=x)))
    ref = T0.knn_connectivity(x.cpu().numpy(), 10, "cosine", False)
    assert np.array_equal(data.dense_adj.cpu().numpy(), np.maximum(ref, ref.T))
    model = BernoulliGraphModel(data.dense_adj).to(CUDA)
    assert abs(model.statistics()["expected_num_edges"] - np.maximum(ref, ref.T).sum()) < 1e-3
