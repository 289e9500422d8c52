"""-m gpu parity tests of the fused direct outer step (lds_outer_step through the C ABI).

Checked against (1) the committed golden vectors produced by the LIVE reference
(`OuterProblemTrainer.train_step`, src/trainers/outer.py:57-87; oracle/make_golden.py) and
(2) the numpy restatement on fresh seeded inputs, including the Philox path where the oracle
regenerates the device's uniforms on the CPU. Tolerances are the north-star's: mask bit-exact,
logits and dL/dtheta rel <= 1e-3 (inf-norm relative), loss within 1e-4.
"""
import numpy as np
import pytest
import torch

from oracle import philox as PH
from oracle import restatement as R

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def rel_inf(x, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(np.asarray(x, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def grad_from_factors(eng, n, theta_triu):
    """Rebuild dL/dtheta_triu in fp64 from the device factor matrices (what K3 consumes)."""
    fa = eng.buffer("fa").double().cpu().numpy()
    fb = eng.buffer("fb").double().cpu().numpy()
    c = eng.buffer("cvec").double().cpu().numpy()
    g = fa @ fb.T + fb @ fa.T + c[:, None] + c[None, :]
    np.fill_diagonal(g, 0.0)
    gt = g[np.triu_indices(n)]
    return gt * ((theta_triu >= 0) & (theta_triu <= 1))


def make_engine(g, sparse):
    from lds_gnn_b200 import kernels as K
    n = int(g["n"])
    eng = K.OuterStep(n, dev(g["x"]), dev(g["y"]), dev(g["mask"]), hidden=int(g["h"]), classes=int(g["c"]), sparse_features=sparse)
    eng.set_weights(dev(g["w0"]), dev(g["b0"]), dev(g["w1"]), dev(g["b1"]))
    return K, eng, n


@pytest.mark.parametrize("sparse", [False, True], ids=["dense_x", "csr_x"])
def test_golden_outer_step(golden, sparse):
    g = golden
    K, eng, n = make_engine(g, sparse)
    p = float(g["p"])
    theta = g["theta_triu"].astype(np.float32)
    for s in range(int(g["steps"])):
        full = K.theta_triu_to_full(dev(theta))
        logp = torch.empty((n, int(g["c"])), dtype=torch.float32, device="cuda")
        kx = dev(g[f"keep_x{s}"].astype(np.uint8)) if p > 0 else None
        kh = dev(g[f"keep_h{s}"].astype(np.uint8)) if p > 0 else None
        lr = float(g[f"lr_used{s}_f64"])
        sc = eng.run(full, lr=lr, seed=0, step=s, dropout_p=p, update=True, u=dev(g[f"U{s}"]), keep_x=kx, keep_h=kh,
                     out_logp=logp)
        torch.cuda.synchronize()
        # sampled mask: bit exact (the reference's sample keeps the sampled diagonal; A_tilde sets it to 1)
        sample = g[f"sample{s}"].astype(np.float32)
        a_ref = sample.copy(); np.fill_diagonal(a_ref, 1.0)
        assert np.array_equal(eng.buffer("adj")[:, :n].float().cpu().numpy(), a_ref), "sampled mask differs"
        assert np.array_equal(eng.buffer("deg").cpu().numpy(), a_ref.sum(1))
        # forward
        assert rel_inf(logp.cpu().numpy(), g[f"logp{s}_f64"]) < 1e-3
        loss, acc = float(sc[0].item()), float(sc[1].item())
        assert abs(loss - float(g[f"loss{s}_f64"])) < 1e-4 * max(1.0, abs(float(g[f"loss{s}_f64"])))
        assert abs(acc - float(g[f"acc{s}_f64"])) < 1e-6
        # hypergradient (factor form) and projected update
        gt = grad_from_factors(eng, n, theta)
        ref_g = g[f"grad_triu{s}_f64"]
        assert rel_inf(gt, ref_g) < 1e-3, f"dL/dtheta rel err {rel_inf(gt, ref_g)}"
        new = K.theta_full_to_triu(full, n).cpu().numpy()
        ref_new = g[f"theta_new{s}_f64"]
        tol = 1e-3 * lr * max(np.abs(ref_g).max(), 1e-12) + 2e-7
        assert np.abs(new - ref_new).max() <= tol
        sym = full[:, :n].cpu().numpy()
        assert np.array_equal(sym, sym.T)
        theta = g[f"theta_new{s}_f32"].astype(np.float32)          # continue from the reference's own state
    # statistics() of the final state (src/models/graph.py:69-78)
    st = K.theta_stats(K.theta_triu_to_full(dev(theta)), n).cpu().numpy()
    assert abs(st[0] - float(g["stat_expected_num_edges_f64"])) <= 1e-4 * max(1.0, float(g["stat_expected_num_edges_f64"]))
    assert abs(st[1] / len(theta) - float(g["stat_mean_prob_f64"])) < 1e-5


@pytest.mark.parametrize("sparse", [False, True], ids=["dense_x", "csr_x"])
@pytest.mark.parametrize("n,f,h,c,p", [(50, 30, 16, 7, 0.0), (301, 120, 16, 7, 0.5), (700, 64, 64, 7, 0.5),
                                       (1200, 200, 32, 10, 0.0), (2708, 1433, 16, 7, 0.5), (3327, 3703, 16, 6, 0.5)],
                         ids=["n50", "n301", "n700_h64", "n1200_h32", "cora_shape", "citeseer_shape_as_benchmarked"])
def test_philox_outer_step_matches_oracle(n, f, h, c, p, sparse):
    """Philox mode end to end: the oracle regenerates the edge uniforms and both dropout masks on the CPU."""
    from oracle.make_golden import make_inputs
    from lds_gnn_b200 import kernels as K
    inp = make_inputs(seed=n, n=n, f=f, h=h, c=c, theta_kind="mixed" if n < 2000 else "sparse", p=p, mask_frac=0.2)
    seed, step, lr = 0xABCDEF0123, 11, 0.3
    eng = K.OuterStep(n, dev(inp["x"]), dev(inp["y"]), dev(inp["mask"]), hidden=h, classes=c, sparse_features=sparse)
    eng.set_weights(dev(inp["w0"]), dev(inp["b0"]), dev(inp["w1"]), dev(inp["b1"]))
    full = K.theta_triu_to_full(dev(inp["theta_triu"]))
    logp = torch.empty((n, c), dtype=torch.float32, device="cuda")
    sc = eng.run(full, lr=lr, seed=seed, step=step, dropout_p=p, update=True, out_logp=logp)
    torch.cuda.synchronize()
    u = PH.edge_uniforms(n, seed, step)
    kx = PH.dropout_keep_mask(n, f, p, seed, step, PH.STREAM_DROP_X) if p > 0 else None
    kh = PH.dropout_keep_mask(n, h, p, seed, step, PH.STREAM_DROP_H) if p > 0 else None
    o = R.outer_step(inp["theta_triu"], u, inp["x"], inp["w0"], inp["b0"], inp["w1"], inp["b1"], inp["y"], inp["mask"],
                     lr=lr, p=p, keep_x=kx, keep_h=kh)
    assert np.array_equal(eng.buffer("adj")[:, :n].float().cpu().numpy(), o["a_tilde"].astype(np.float32))
    assert rel_inf(eng.buffer("p1").cpu().numpy(), o["p1"]) < 1e-4
    assert rel_inf(eng.buffer("z1").cpu().numpy(), o["z1"]) < 1e-3
    assert rel_inf(logp.cpu().numpy(), o["logp"]) < 1e-3
    assert abs(float(sc[0].item()) - o["loss"]) < 1e-4 * max(1.0, abs(o["loss"]))
    assert abs(float(sc[1].item()) - o["acc"]) < 1e-6
    for name in ("dz2", "dp2", "dz1", "dp1"):
        assert rel_inf(eng.buffer(name).cpu().numpy(), o[name]) < 1e-3, name
    assert rel_inf(eng.buffer("cvec").cpu().numpy(), o["c"]) < 1e-3
    gt = grad_from_factors(eng, n, inp["theta_triu"])
    assert rel_inf(gt, o["d_theta_triu"]) < 1e-3
    new = K.theta_full_to_triu(full, n).cpu().numpy()
    assert np.abs(new - o["theta_new"]).max() <= 1e-3 * lr * np.abs(o["d_theta_triu"]).max() + 2e-7


@pytest.mark.parametrize("n,f,h,c", [(301, 120, 16, 7), (2708, 1433, 16, 7), (700, 64, 64, 7)])
def test_stream_k_schedule_gives_the_same_step(n, f, h, c):
    """Large graphs (> 148 row panels) split panels across CTAs; force that schedule at test sizes and compare."""
    from oracle.make_golden import make_inputs
    from lds_gnn_b200 import _lib, kernels as K
    inp = make_inputs(seed=n + 1, n=n, f=f, h=h, c=c, theta_kind="mixed", p=0.5, mask_frac=0.2)
    eng = K.OuterStep(n, dev(inp["x"]), dev(inp["y"]), dev(inp["mask"]), hidden=h, classes=c)
    eng.set_weights(dev(inp["w0"]), dev(inp["b0"]), dev(inp["w1"]), dev(inp["b1"]))
    full = K.theta_triu_to_full(dev(inp["theta_triu"]))
    outs = []
    for flags in (0, _lib.K2_FORCE_STREAMK, _lib.K2_FORCE_STREAMK):
        t = full.clone()
        lp = torch.empty((n, c), device="cuda")
        sc = eng.run(t, lr=0.3, seed=5, step=2, dropout_p=0.5, update=True, out_logp=lp, k2_flags=flags).clone()
        outs.append((t, lp, sc))
    assert torch.equal(outs[1][0], outs[2][0]) and torch.equal(outs[1][1], outs[2][1])          # stream-K is bitwise reproducible
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-5 * outs[0][1].abs().max().item()  # same logits up to summation order
    assert (outs[0][0] - outs[1][0]).abs().max().item() < 1e-6
    assert abs(outs[0][2][0].item() - outs[1][2][0].item()) < 1e-5


def test_outer_step_update_flag_and_determinism():
    from oracle.make_golden import make_inputs
    from lds_gnn_b200 import kernels as K
    n, f, h, c = 400, 60, 16, 7
    inp = make_inputs(seed=9, n=n, f=f, h=h, c=c, theta_kind="uniform", p=0.5)
    eng = K.OuterStep(n, dev(inp["x"]), dev(inp["y"]), dev(inp["mask"]), hidden=h, classes=c)
    eng.set_weights(dev(inp["w0"]), dev(inp["b0"]), dev(inp["w1"]), dev(inp["b1"]))
    full = K.theta_triu_to_full(dev(inp["theta_triu"]))
    before = full.clone()
    eng.run(full, lr=0.5, seed=3, step=0, dropout_p=0.5, update=False)
    assert torch.equal(full, before)                                     # update=0 leaves theta untouched
    outs = []
    for _ in range(2):
        t = before.clone()
        sc = eng.run(t, lr=0.5, seed=3, step=0, dropout_p=0.5, update=True).clone()
        outs.append((t.clone(), sc))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])   # bitwise reproducible
    t2 = before.clone()
    eng.run(t2, lr=0.5, seed=3, step=1, dropout_p=0.5, update=True)
    assert not torch.equal(t2, outs[0][0])                               # a new step draws a new graph


def test_error_paths_do_not_abort():
    from lds_gnn_b200 import _lib, kernels as K
    lib = _lib.load()
    full = torch.zeros((8, 64), device="cuda")
    with pytest.raises(RuntimeError, match="row0 must be even"):
        K.k1_sample_normalize(full, 8, 0, 0, row0=1, rows=2)
    with pytest.raises(RuntimeError, match="outside"):
        K.k1_sample_normalize(full, 8, 0, 0, row0=6, rows=4)
    assert lib.lds_k2_workspace_bytes(100, 100, 300) == -1
    args = _lib.OuterStepArgs()
    args.struct_bytes = 8
    import ctypes
    assert lib.lds_outer_step(ctypes.byref(args), None) != 0
    assert "struct_bytes" in _lib.last_error()


@pytest.mark.parametrize("n,f,h,c,p,world,packed", [(700, 64, 16, 7, 0.5, 3, True), (1500, 200, 32, 10, 0.0, 4, True), (2708, 1433, 16, 7, 0.5, 2, True),
                                                    (600, 40, 64, 7, 0.5, 5, True), (700, 64, 16, 7, 0.5, 3, False), (600, 40, 64, 7, 0.5, 5, False)])
def test_row_block_sharded_step_matches_single_device(n, f, h, c, p, world, packed):
    """the oracle of the sharded run is the single-device result. All `world` ranks are emulated on one GPU
    (same kernels, same phases, exchange = concatenation); the sampled mask must be identical bit for bit (no exchange
    of random bits), the rest equal up to fp32 summation order."""
    from oracle.make_golden import make_inputs
    from lds_gnn_b200 import kernels as K, sharded as S
    inp = make_inputs(seed=n + world, n=n, f=f, h=h, c=c, theta_kind="mixed", p=p, mask_frac=0.2)
    w = [dev(inp[k]) for k in ("w0", "b0", "w1", "b1")]
    x, y, mask = dev(inp["x"]), dev(inp["y"]), dev(inp["mask"])
    full = K.theta_triu_to_full(dev(inp["theta_triu"]))
    seed, step, lr = 4242, 7, 0.4
    # single device
    eng = K.OuterStep(n, x, y, mask, hidden=h, classes=c)
    eng.set_weights(*w)
    ref_theta = full.clone()
    ref_sc = eng.run(ref_theta, lr=lr, seed=seed, step=step, dropout_p=p, update=True).clone()
    # `world` row-block shards
    bounds = [S.shard_bounds(n, world, r) for r in range(world)]
    bounds = [b for b in bounds if b[1] > 0]
    assert sum(b[1] for b in bounds) == n
    shards, thetas = [], []
    for lo, cnt in bounds:
        sh = S.ShardedOuterStep(n, lo, cnt, x[lo:lo + cnt], y[lo:lo + cnt], mask[lo:lo + cnt], int(mask.sum().item()), h, c)
        sh.set_weights(*w)
        sh.packed = packed            # operand exchange in its final bf16 layout (3-D tensor map) / legacy fp32 rows + re-layout kernel
        shards.append(sh)
        thetas.append(full[lo:lo + cnt].clone())
    sc = S.run_local_group(shards, thetas, lr=lr, seed=seed, step=step, dropout_p=p, update=True)
    torch.cuda.synchronize()
    adj = torch.cat([s.eng.buffer("adj") for s in shards], dim=0)
    assert torch.equal(adj, eng.buffer("adj")), "sharded sampling must reproduce the same graph without any exchange"
    assert torch.equal(torch.cat([s.eng.buffer("deg") for s in shards]), eng.buffer("deg"))
    for name in ("z1", "z2", "dz1", "dp1", "cvec"):
        got = torch.cat([s.eng.buffer(name) for s in shards], dim=0)
        ref = eng.buffer(name)
        assert (got - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1e-30), name
    assert abs(sc[0].item() - ref_sc[0].item()) < 1e-5 and abs(sc[1].item() - ref_sc[1].item()) < 1e-6
    new = torch.cat(thetas, dim=0)
    assert (new - ref_theta).abs().max().item() < 5e-6       # fp32 summation order of the split propagations differs
    assert torch.equal(new[:, :n], new[:, :n].t())            # shards stay mutually consistent (exact symmetry)


# ------------------------------------------------------------------------------------------------
# Multi-sample straight-through estimator (BASELINE.json config 3: S Bernoulli samples per outer step)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,f,h,c,S,sparse,p", [(300, 40, 16, 7, 3, True, 0.5), (300, 40, 16, 7, 3, False, 0.0), (530, 64, 32, 5, 4, True, 0.5),
                                               (1100, 50, 16, 6, 16, True, 0.5)],
                         ids=["fused-kernel-S3", "dense-x-S3", "h32-S4", "S16"])
def test_multi_sample_step_is_the_mean_of_the_single_sample_gradients(n, f, h, c, S, sparse, p):
    """theta <- clamp(theta - lr * mean_s g_s): the S samples' rank-2d gradients are concatenated along K of ONE update pass.
    Expected value: the per-sample factor matrices of single-sample calls with the same (seed, step, sample) — themselves
    checked against the oracle above — combined in fp64."""
    from lds_gnn_b200 import kernels as K
    rng = np.random.default_rng(n + S)
    x = (rng.random((n, f)) < 0.1).astype(np.float32) * rng.random((n, f)).astype(np.float32)
    y = rng.integers(0, c, n)
    mask = rng.random(n) < 0.3
    w = [dev((rng.standard_normal(s_) * 0.3).astype(np.float32)) for s_ in ((h, f), (h,), (c, h), (c,))]
    a = (rng.random((n, n)) < 0.05)
    th = np.triu(a, 1); th = (th + th.T).astype(np.float32) * 0.9 + 0.05 * np.eye(n, dtype=np.float32)
    th[5, 7] = th[7, 5] = 1.3; th[9, 11] = th[11, 9] = -0.2                       # outside [0, 1]: clamp backward masks them
    theta0 = K.theta_triu_to_full(dev(th[np.triu_indices(n)]))
    eng = K.OuterStep(n, dev(x), dev(y), dev(mask), hidden=h, classes=c, sparse_features=sparse)
    eng.set_weights(*w)
    seed, step, lr = 99, 3, 0.6
    g = np.zeros((n, n)); losses, accs, adjs = [], [], []
    for s in range(S):
        t = theta0.clone()
        sc = eng.run(t, lr=lr, seed=seed, step=step, dropout_p=p, update=False, sample=s).clone()
        assert torch.equal(t, theta0)
        fa, fb, cv = (eng.buffer(k).double().cpu().numpy() for k in ("fa", "fb", "cvec"))
        gs = fa @ fb.T + fb @ fa.T + cv[:, None] + cv[None, :]
        np.fill_diagonal(gs, 0.0)
        g += gs
        losses.append(sc[0].item()); accs.append(sc[1].item())
        adjs.append(eng.buffer("adj")[:, :n].float().cpu().numpy().copy())
    assert any(not np.array_equal(adjs[0], adjs[s]) for s in range(1, S)), "samples must differ"
    th64 = theta0[:, :n].double().cpu().numpy()
    g *= ((th64 >= 0) & (th64 <= 1))
    ref = np.clip(th64 - lr * g / S, 0, 1)
    theta = theta0.clone()
    sc = eng.run_multi(theta, S, lr=lr, seed=seed, step=step, dropout_p=p, update=True)
    out = theta[:, :n].cpu().numpy()
    assert np.array_equal(out, out.T)
    assert np.abs(out - ref).max() < 3e-6 + 2e-5 * lr * np.abs(g).max() / S
    assert abs(sc[0].item() - np.mean(losses)) < 1e-5 and abs(sc[1].item() - np.mean(accs)) < 1e-6
    # S = 1 through run_multi is the plain step
    t1, t2 = theta0.clone(), theta0.clone()
    eng.run_multi(t1, 1, lr=lr, seed=seed, step=step, dropout_p=p, update=True)
    eng.run(t2, lr=lr, seed=seed, step=step, dropout_p=p, update=True)
    assert torch.equal(t1, t2)


def test_cora_shape_knn_theta_16_samples_matches_oracle():
    """BASELINE config 3 at its exact geometry: Cora shape (N = 2708, F = 1433, C = 7, h = 16), theta_0 from a cosine kNN graph
    (k = 10, max-symmetrised; src/data/utils.py:165-175, transforms.py:15-37), S = 16 Bernoulli samples per outer step with
    dropout 0.5. Oracle: the fp64 restatement run once per sample with that sample's regenerated uniforms and dropout masks;
    expected update = clamp(theta - lr * mean_s g_s)."""
    from sklearn.neighbors import kneighbors_graph
    from lds_gnn_b200 import kernels as K
    n, f, h, c, S, p = 2708, 1433, 16, 7, 16, 0.5
    rng = np.random.default_rng(3)
    x = (rng.random((n, f)) < 0.0127).astype(np.float32)
    x[np.arange(n), rng.integers(0, f, n)] = 1.0
    x /= x.sum(1, keepdims=True)
    knn = kneighbors_graph(x, 10, metric="cosine", mode="connectivity", include_self=False).toarray().astype(np.float32)
    adj = np.maximum(knn, knn.T)
    full0 = 0.7 * adj + 0.3 * adj * rng.random((n, n)).astype(np.float32)             # edges in (0.7, 1.0]: the 16 graphs differ
    full0 = np.triu(full0, 1); full0 = full0 + full0.T
    theta_triu = full0[np.triu_indices(n)].astype(np.float32)
    lim0, lim1 = np.sqrt(6.0 / (f + h)), np.sqrt(6.0 / (h + c))
    w = [rng.uniform(-lim0, lim0, (h, f)).astype(np.float32), (rng.standard_normal(h) * 0.05).astype(np.float32),
         rng.uniform(-lim1, lim1, (c, h)).astype(np.float32), (rng.standard_normal(c) * 0.05).astype(np.float32)]
    y = rng.integers(0, c, n).astype(np.int64)
    mask = np.zeros(n, dtype=bool); mask[rng.permutation(n)[:250]] = True
    seed, step, lr = 0xC0FFEE, 4, 1.0
    eng = K.OuterStep(n, dev(x), dev(y), dev(mask), hidden=h, classes=c)
    eng.set_weights(*(dev(a) for a in w))
    theta = K.theta_triu_to_full(dev(theta_triu))
    sc = eng.run_multi(theta, S, lr=lr, seed=seed, step=step, dropout_p=p, update=True)
    torch.cuda.synchronize()
    grad, losses, accs, degs = 0.0, [], [], []
    for s in range(S):
        u = PH.edge_uniforms(n, seed, step, sample=s)
        kx = PH.dropout_keep_mask(n, f, p, seed, step, PH.STREAM_DROP_X, sample=s)
        kh = PH.dropout_keep_mask(n, h, p, seed, step, PH.STREAM_DROP_H, sample=s)
        o = R.outer_step(theta_triu, u, x, *w, y, mask, lr=lr, p=p, keep_x=kx, keep_h=kh)
        grad = grad + o["d_theta_triu"] / S
        losses.append(o["loss"]); accs.append(o["acc"]); degs.append(o["deg"].sum())
    assert len(set(degs)) > 1, "the 16 sampled graphs must differ"
    assert abs(sc[0].item() - np.mean(losses)) < 1e-4 * max(1.0, np.mean(losses)) and abs(sc[1].item() - np.mean(accs)) < 1e-6
    ref_new = np.clip(theta_triu.astype(np.float64) - lr * grad, 0.0, 1.0)
    new = K.theta_full_to_triu(theta, n).cpu().numpy()
    tol = 1e-3 * lr * np.abs(grad).max() + 2e-7
    assert np.abs(new - ref_new).max() <= tol, (np.abs(new - ref_new).max(), tol)
    assert np.abs(ref_new - np.clip(theta_triu, 0, 1)).max() > 50 * tol                  # the step taken dwarfs the tolerance
    sym = theta[:, :n].cpu().numpy()
    assert np.array_equal(sym, sym.T)


@pytest.mark.parametrize("n,f,h,c,flags", [(301, 120, 16, 7, "fused"), (301, 120, 16, 7, "packed"), (301, 120, 16, 7, "bf16"),
                                           (1000, 64, 64, 7, "packed"), (2708, 300, 16, 7, "fused"), (130, 40, 16, 3, "fused")])
def test_outer_step_writes_stay_inside_their_buffers(n, f, h, c, flags):
    """Guard bands instead of compute-sanitizer (closed on this GPU pool): theta and the step's workspace sit inside larger
    allocations filled with a sentinel; after several steps on each launch plan the bands, theta's padding columns and the
    rows past n are untouched, and the step still agrees with an engine that ran on plain allocations."""
    from lds_gnn_b200 import _lib, kernels as K
    rng = np.random.default_rng(n + f)
    x = (rng.random((n, f)) < 0.05).astype(np.float32) * rng.random((n, f)).astype(np.float32)
    y = rng.integers(0, c, n)
    mask = np.zeros(n, dtype=bool); mask[rng.permutation(n)[: max(4, n // 5)]] = True
    w = [rng.standard_normal((h, f)).astype(np.float32) * 0.1, np.zeros(h, np.float32),
         rng.standard_normal((c, h)).astype(np.float32) * 0.1, np.zeros(c, np.float32)]
    sym = rng.random((n, n)).astype(np.float32); sym = (sym + sym.T) / 2
    k2 = {"fused": 0, "packed": _lib.K2_NO_FUSE, "bf16": _lib.K2_NO_FUSE | _lib.K2_BF16_ADJ}[flags]

    def engine():
        eng = K.OuterStep(n, dev(x), dev(y), dev(mask), hidden=h, classes=c, sparse_features=True)
        eng.set_weights(*(dev(a) for a in w))
        return eng

    ld = K.padded_ld(n)
    band = 64
    SENT = 7.25
    big = torch.full((n + 2 * band, ld), SENT, device="cuda")
    theta = big[band:band + n]
    theta.zero_(); theta[:, :n] = dev(sym)
    plain_theta = torch.zeros((n, ld), device="cuda"); plain_theta[:, :n] = dev(sym)

    eng = engine()
    nbytes = eng.ws_bytes
    guard = 1 << 16
    raw = torch.full((nbytes + 2 * guard + 1024,), 0xA5, dtype=torch.uint8, device="cuda")
    off = guard + ((-(raw.data_ptr() + guard)) % 1024)
    raw[off:off + nbytes] = 0
    eng._ws_raw, eng.ws = raw, raw[off:off + nbytes]
    eng._init_static_args()                                              # the argument block caches the workspace pointer
    ref = engine()
    for step in range(3):
        eng.run(theta, lr=0.5, seed=11, step=step, dropout_p=0.5, k2_flags=k2, want_adj=False)
        ref.run(plain_theta, lr=0.5, seed=11, step=step, dropout_p=0.5, k2_flags=k2, want_adj=False)
    torch.cuda.synchronize()
    assert bool((raw[:off] == 0xA5).all()) and bool((raw[off + nbytes:] == 0xA5).all()), "write outside the workspace"
    assert bool((big[:band] == SENT).all()) and bool((big[band + n:] == SENT).all()), "write outside theta's rows"
    assert bool((theta[:, n:] == 0).all()), "theta's padding columns were written"
    assert torch.equal(theta, plain_theta)
    assert torch.equal(eng.scalars[:2], ref.scalars[:2])
