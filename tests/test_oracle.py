"""CPU tests of the oracle itself (no GPU): it is only trustworthy if it reproduces
 (1) every known-answer test the reference holds for this path, and
 (2) the golden vectors produced by the live reference (oracle/make_golden.py), and — in the build
     container, where /root/reference exists — the live reference directly, bit for bit where it must.
"""
import numpy as np
import pytest
import torch

from oracle import live_reference as L
from oracle import philox as PH
from oracle import reference_port as P
from oracle import restatement as R

needs_reference = pytest.mark.skipif(not L.available(), reason="/root/reference is only present in the build container")


def rel_inf(x, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(np.asarray(x, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


# ------------------------------------------------------------------ (1) the reference's known-answer tests
def test_known_answer_to_undirected():                   # tst/utils/test_graph.py:32-52
    adj = np.zeros((10, 10), np.float32)
    adj[1, :] = 1.0
    out = R.to_undirected(adj)
    exp = adj.copy(); exp[:, 1] = 1.0
    assert np.array_equal(out, exp)
    tri = R.to_undirected(adj, from_triu_only=True)
    exp2 = np.zeros((10, 10), np.float32); exp2[1, 1:] = 1.0; exp2[1:, 1] = 1.0
    assert np.array_equal(tri, exp2)


def test_known_answer_triu_to_symmetric():               # tst/utils/test_graph.py:213-221
    out = R.theta_full_from_triu(np.array([1, 2, 3, 4, 5, 6], np.float32) / 10)
    assert np.array_equal(out, np.array([[1, 2, 3], [2, 4, 5], [3, 5, 6]], np.float32) / 10)


@pytest.mark.parametrize("n", [10, 100, 1000, 2708, 3327, 50000, 500000])
def test_known_answer_num_nodes_from_triu_shape(n):      # tst/utils/test_graph.py:232-235
    assert R.num_nodes_from_triu_shape(n * (n + 1) // 2) == n


def test_known_answer_sampling_of_binary_theta():        # tst/models/test_sampling.py:149-160
    n = 10
    theta = np.triu(np.ones((n, n), np.float32), 1)
    u = np.random.default_rng(0).random((n, n)).astype(np.float32)
    assert np.array_equal(R.sample_graph(theta, u, undirected=False), theta)
    assert np.array_equal(R.sample_graph(theta, u, undirected=True), 1.0 - np.eye(n, dtype=np.float32))


def test_known_answer_bernoulli_forward():               # tst/models/test_bernoulli_model.py:56-64
    adj = np.eye(10, dtype=np.float32); adj[1, :] = 1.0
    adj = R.to_undirected(adj)
    assert np.array_equal(R.theta_full_from_triu(R.get_triu_values(adj)), np.clip(adj + adj.T, 0, 1))


def test_known_answer_accuracy():                        # tst/utils/test_evaluation.py:12-18
    logp = np.log(np.array([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7]]))
    _, acc = R.nll_and_accuracy(logp, np.array([1, 0, 0]), np.array([True, True, True]))
    assert abs(acc - 2 / 3) < 1e-7


def test_gradient_structure_facts():                     # tst/utils/test_graph.py:66-72, 169-178
    """Mirror backward touches only the upper triangle; self loops kill the diagonal gradient."""
    rng = np.random.default_rng(1)
    n, f, h, c = 12, 5, 4, 3
    theta = rng.random(n * (n + 1) // 2).astype(np.float32)
    o = R.outer_step(theta, rng.random((n, n)).astype(np.float32), rng.random((n, f)), rng.standard_normal((h, f)),
                     np.zeros(h), rng.standard_normal((c, h)), np.zeros(c), rng.integers(0, c, n), np.ones(n, bool), lr=0.1)
    assert np.all(np.diag(o["d_theta_full"]) == 0)
    iu = np.triu_indices(n)
    assert np.all(o["d_theta_triu"][iu[0] == iu[1]] == 0) and np.count_nonzero(o["d_theta_triu"]) == n * (n - 1) // 2


# ------------------------------------------------------------------ (2) golden vectors from the live reference
def test_restatement_matches_golden(golden):
    g = golden
    theta = g["theta_triu"].astype(np.float32)
    p = float(g["p"])
    for s in range(int(g["steps"])):
        kx = g[f"keep_x{s}"] if p > 0 else None
        kh = g[f"keep_h{s}"] if p > 0 else None
        o = R.outer_step(theta, g[f"U{s}"], g["x"], g["w0"], g["b0"], g["w1"], g["b1"], g["y"], g["mask"],
                         lr=float(g[f"lr_used{s}_f64"]), p=p, keep_x=kx, keep_h=kh)
        assert np.array_equal(o["sample"].astype(np.uint8), g[f"sample{s}"])                  # mask: bit exact
        tol = 1e-12 if s == 0 else 1e-6
        assert rel_inf(o["logp"], g[f"logp{s}_f64"]) < tol
        assert abs(o["loss"] - float(g[f"loss{s}_f64"])) < tol and abs(o["acc"] - float(g[f"acc{s}_f64"])) < 1e-7
        assert rel_inf(o["d_theta_triu"], g[f"grad_triu{s}_f64"]) < (1e-10 if s == 0 else 1e-5)   # closed form == autograd
        # later steps restart from the reference's fp32 state while its fp64 run carried fp64 state: rounding-level offset
        assert np.abs(o["theta_new"] - g[f"theta_new{s}_f64"]).max() < (1e-12 if s == 0 else 1e-7)
        assert rel_inf(o["logp"], g[f"logp{s}_f32"]) < 1e-4                                    # fp32 reference run, for scale
        theta = g[f"theta_new{s}_f32"].astype(np.float32)
    st = R.statistics(theta)
    assert abs(st["expected_num_edges"] - float(g["stat_expected_num_edges_f32"])) <= 1e-3 * max(1.0, st["expected_num_edges"])
    assert st["min_prob"] == float(g["stat_min_prob_f32"]) and st["max_prob"] == float(g["stat_max_prob_f32"])


def test_port_matches_golden_on_deterministic_case():
    """theta in {0,1}: the port's torch.bernoulli draw is deterministic, so it must reproduce the live reference."""
    from conftest import load_golden
    g = load_golden("n96_binary")
    t = lambda k: torch.as_tensor(g[k])
    port = P.ReferenceOuterStep(t("theta_triu"), t("x"), t("y"), t("mask"), t("w0"), t("b0"), t("w1"), t("b1"),
                                lr=float(g["lr"]), lr_decay=float(g["lr_decay"]), dropout=0.0)
    loss, acc = port.step(training=False)
    assert abs(loss - float(g["loss0_f32"])) < 1e-6 and abs(acc - float(g["acc0_f32"])) < 1e-7
    assert np.array_equal(port.last["graph"].numpy().astype(np.uint8), g["sample0"])
    assert np.abs(port.theta.detach().numpy() - g["theta_new0_f32"]).max() < 1e-7


# ------------------------------------------------------------------ Philox restatement
def test_philox_known_answer_vectors():
    """Random123 known-answer test for Philox4x32-10 (kat_vectors): counter/key all zeros and all ones."""
    out = PH.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    ff = 0xFFFFFFFF
    out = PH.philox4x32_10(ff, ff, ff, ff, ff, ff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_philox_uniform_mapping_properties():
    u = PH.edge_uniforms(33, seed=5, step=2)
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1 and np.array_equal(u, u.T)
    assert not np.array_equal(u, PH.edge_uniforms(33, seed=5, step=3))
    assert PH.edge_uniform(5, 2, 0, 7, 19) == u[7, 19] == u[19, 7]
    keep = PH.dropout_keep_mask(200, 64, 0.5, seed=5, step=2, stream=PH.STREAM_DROP_X)
    assert 0.45 < keep.mean() < 0.55
    assert not np.array_equal(keep[:, :16], PH.dropout_keep_mask(200, 16, 0.5, 5, 2, PH.STREAM_DROP_H))


def test_blockwise_philox_generators_equal_the_elementwise_definition():
    """edge_uniforms / dropout_uniforms draw one Philox call per 2x2 block / per 4-column group; the element-by-element
    statement of the mapping (one call per element, the header's definition) must give the same words."""
    for n in (1, 2, 5, 33, 130):
        for sample in (0, 3):
            assert np.array_equal(PH.edge_uniforms(n, 77, 5, sample=sample), PH.edge_uniforms_elementwise(n, 77, 5, sample=sample))
    for rows, cols in ((7, 1), (5, 4), (9, 13), (64, 50)):
        r = np.arange(rows, dtype=np.int64)[:, None]; c = np.arange(cols, dtype=np.int64)[None, :]
        c2, c3 = PH._c23(2, PH.STREAM_DROP_X, 2)
        w = PH.philox4x32_10(c // 4, r, c2, c3, 9, 0)
        assert np.array_equal(PH.dropout_uniforms(rows, cols, 9, 2, PH.STREAM_DROP_X, sample=2), PH.to_uniform(np.choose((c % 4) + 0 * r, w)))


# ------------------------------------------------------------------ live reference (build container only)
@needs_reference
def test_port_is_bit_identical_to_live_reference_under_shared_rng():
    rng = np.random.default_rng(7)
    n, f, h, c = 40, 16, 8, 4
    theta = rng.random(n * (n + 1) // 2).astype(np.float32)
    x = rng.random((n, f)).astype(np.float32)
    w0 = (rng.standard_normal((h, f)) * 0.3).astype(np.float32); b0 = np.zeros(h, np.float32)
    w1 = (rng.standard_normal((c, h)) * 0.3).astype(np.float32); b1 = np.zeros(c, np.float32)
    y = rng.integers(0, c, n); mask = rng.random(n) < 0.5
    model, gcn, trainer, data = L.build(theta, x, w0, b0, w1, b1, y, mask, lr=0.4, lr_decay=0.9, p=0.5)
    gcn.train()
    port = P.ReferenceOuterStep(torch.as_tensor(theta), torch.as_tensor(x), torch.as_tensor(y), torch.as_tensor(mask),
                                torch.as_tensor(w0), torch.as_tensor(b0), torch.as_tensor(w1), torch.as_tensor(b1),
                                lr=0.4, lr_decay=0.9, dropout=0.5)
    for step in range(3):
        torch.manual_seed(100 + step)
        m = trainer.train_step(lambda g_: gcn(data.x, g_, params=None), retain_graph=False)
        torch.manual_seed(100 + step)
        loss, acc = port.step(training=True)
        assert loss == m.loss and acc == m.acc
        assert torch.equal(port.theta.detach(), model.probs.detach())


@needs_reference
def test_restatement_matches_live_reference_fp64():
    rng = np.random.default_rng(11)
    n, f, h, c = 57, 20, 8, 5
    theta = rng.random(n * (n + 1) // 2).astype(np.float32)
    theta[rng.random(len(theta)) < 0.3] = 0.0
    x = rng.random((n, f)).astype(np.float32)
    w0 = (rng.standard_normal((h, f)) * 0.3).astype(np.float32); b0 = (rng.standard_normal(h) * 0.1).astype(np.float32)
    w1 = (rng.standard_normal((c, h)) * 0.3).astype(np.float32); b1 = (rng.standard_normal(c) * 0.1).astype(np.float32)
    y = rng.integers(0, c, n); mask = rng.random(n) < 0.5
    u = PH.edge_uniforms(n, 77, 0)
    kx = PH.dropout_keep_mask(n, f, 0.5, 77, 0, PH.STREAM_DROP_X); kh = PH.dropout_keep_mask(n, h, 0.5, 77, 0, PH.STREAM_DROP_H)
    res, _ = L.outer_steps(theta, [u], x, w0, b0, w1, b1, y, mask, lr=0.7, lr_decay=0.99, p=0.5, keep_masks=[kx, kh],
                           dtype=torch.float64)
    o = R.outer_step(theta, u, x, w0, b0, w1, b1, y, mask, lr=0.7, p=0.5, keep_x=kx, keep_h=kh)
    assert np.array_equal(res[0]["sample"], o["sample"])
    assert rel_inf(o["logp"], res[0]["logp"]) < 1e-12 and rel_inf(o["d_theta_triu"], res[0]["grad_triu"]) < 1e-10


# ------------------------------------------------------------------ theta_0 construction
@needs_reference
def test_theta0_restatement_matches_live_reference_functions():
    from oracle import theta0 as T0
    L.enable()
    from src.data import utils as RU
    from src.utils.graph import to_dense_adj as ref_to_dense
    rng = np.random.default_rng(0)
    n = 60
    a = (rng.random((n, n)) < 0.1).astype(np.float32) * rng.random((n, n)).astype(np.float32)
    sym = np.maximum(a, a.T)
    np.fill_diagonal(sym, (rng.random(n) < 0.3) * 0.5)
    for pct in (0.0, 0.25, 0.9, 1.0):
        for seed in (0, 7):
            ref = RU.remove_edges_from_undirected_graph(torch.as_tensor(sym), pct, seed=seed)
            assert torch.equal(T0.remove_edges_from_undirected_graph(sym, pct, seed=seed), ref)
            refd = RU.remove_edges_from_directed_graph(torch.as_tensor(a), pct, seed=seed)
            assert torch.equal(T0.remove_edges_from_directed_graph(a, pct, seed=seed), refd)
    x = rng.random((n, 9)).astype(np.float32)
    for metric, loop in (("cosine", True), ("cosine", False), ("minkowski", False)):
        assert np.array_equal(T0.knn_connectivity(x, 5, metric, loop), RU.knn_graph_dense(torch.as_tensor(x), 5, loop=loop, metric=metric).numpy())
    ei = torch.as_tensor(np.stack([rng.integers(0, n, 200), rng.integers(0, n, 200)]))
    assert np.array_equal(T0.to_dense_adj(ei.numpy(), n), ref_to_dense(ei, num_max_nodes=n).numpy())
