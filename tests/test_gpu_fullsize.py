"""-m gpu tests at BASELINE.json's full single-GPU size (config 4: N = 20 000, F = 512, hidden 64, C = 7), where the oracle
cannot run (25 dense N x N fp32 temporaries, N^3 products). Size-independent properties of the path instead:

  * the sampled A_tilde is a symmetric {0,1} matrix with unit diagonal; deg = its row sums EXACTLY; r = deg^-1/2;
    its edge density matches theta's mean (5 sigma);
  * checksum through the tensor cores: A_tilde @ 1 reproduces deg bit for bit (sums of ones are exact in fp32);
  * linearity of the propagation;
  * a row-block shard regenerates exactly the rows of the full graph (no exchange of random bits);
  * the update keeps theta symmetric bit for bit, inside [0, 1], leaves the padding alone, is idempotent at lr = 0,
    and the whole step is bitwise reproducible.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
N = 20000


@pytest.fixture(scope="module")
def big():
    import bench
    from lds_gnn_b200 import kernels as K
    dev = torch.device("cuda")
    d = bench.make_large_rows("n20k", dev, 0, N, seed=3)
    rng = np.random.default_rng(0)
    f, h, c = d["f"], d["h"], d["c"]
    w = [torch.as_tensor((rng.standard_normal(s) * 0.2).astype(np.float32), device=dev) for s in ((h, f), (h,), (c, h), (c,))]
    return K, d, w


def test_full_size_sample_degree_and_tensor_core_checksum(big):
    K, d, _ = big
    theta = d["theta"]
    adj, _, deg, rs = K.k1_sample_normalize(theta, N, seed=11, step=5)
    a = adj[:, :N]
    assert torch.equal(a, a.t()), "A_tilde must be symmetric"
    assert bool(((a == 0) | (a == 1)).all()) and bool((torch.diagonal(a) == 1).all())
    assert bool((adj[:, N:] == 0).all())
    rowsum = a.float().sum(dim=1)
    assert torch.equal(rowsum, deg), "deg must be the exact row sums"
    assert torch.allclose(rs, deg.rsqrt(), rtol=2e-7, atol=0)
    # density: E[A_ij] = theta_ij (i != j); both triangles carry the same draws -> N(N-1)/2 independent Bernoullis
    p = theta[:, :N].clamp(0, 1)
    mean_theta = (p.sum().item() - torch.diagonal(p).sum().item()) / (N * (N - 1))
    mean_adj = (a.float().sum().item() - N) / (N * (N - 1))
    sigma = (0.25 / (N * (N - 1) / 2)) ** 0.5
    assert abs(mean_adj - mean_theta) < 5 * sigma
    # checksum through TMA + tcgen05: A_tilde @ ones == deg exactly, every column
    ones = torch.ones((N, 16), device="cuda")
    z = K.k2_propagate(adj, N, ones)
    assert torch.equal(z, deg[:, None].expand(-1, 16).contiguous())
    # a row-block shard regenerates the same rows without any exchange
    r0, rows = 8192, 4096
    sh_adj, _, sh_deg, _ = K.k1_sample_normalize(theta[r0:r0 + rows], N, seed=11, step=5, row0=r0, rows=rows)
    assert torch.equal(sh_adj, adj[r0:r0 + rows]) and torch.equal(sh_deg, deg[r0:r0 + rows])


def test_full_size_propagation_is_linear(big):
    K, d, _ = big
    adj = K.k1_sample_normalize(d["theta"], N, seed=2, step=1)[0]
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    p = torch.randn((N, 64), device="cuda", generator=g)
    q = torch.randn((N, 64), device="cuda", generator=g)
    zp, zq = K.k2_propagate(adj, N, p), K.k2_propagate(adj, N, q)
    zl = K.k2_propagate(adj, N, 0.5 * p - 2.0 * q)
    ref = 0.5 * zp - 2.0 * zq
    assert (zl - ref).abs().max().item() < 2e-5 * ref.abs().max().item()     # hi/lo split: ~2^-17 relative per term


def test_full_size_step_symmetry_projection_idempotence_determinism(big):
    K, d, w = big
    eng = K.OuterStep(N, d["x"], d["y"], d["mask"], hidden=d["h"], classes=d["c"])
    eng.set_weights(*w)
    theta0 = d["theta"].clone()
    theta0[3, 5] = theta0[5, 3] = 1.4                   # outside [0, 1]: projected by the step
    theta0[7, 9] = theta0[9, 7] = -0.3
    pad = theta0[:, N:].clone()
    # lr = 0: the step is the projection, and applying it twice changes nothing
    t = theta0.clone()
    eng.run(t, lr=0.0, seed=5, step=0, dropout_p=0.5, update=True)
    assert torch.equal(t[:, :N], theta0[:, :N].clamp(0, 1))
    t2 = t.clone()
    eng.run(t2, lr=0.0, seed=5, step=1, dropout_p=0.5, update=True)
    assert torch.equal(t2, t)
    # a real step: symmetric bit for bit, inside [0, 1], padding untouched, finite loss, and reproducible
    a, b = theta0.clone(), theta0.clone()
    sa = eng.run(a, lr=50.0, seed=5, step=2, dropout_p=0.5, update=True).clone()
    sb = eng.run(b, lr=50.0, seed=5, step=2, dropout_p=0.5, update=True).clone()
    assert torch.equal(a, b) and torch.equal(sa, sb), "the step must be bitwise reproducible"
    v = a[:, :N]
    assert torch.equal(v, v.t()) and float(v.min()) >= 0.0 and float(v.max()) <= 1.0
    assert torch.equal(a[:, N:], pad)
    assert not torch.equal(v, theta0[:, :N].clamp(0, 1)), "lr = 50 must move theta"
    assert np.isfinite(sa[0].item()) and 0.0 <= sa[1].item() <= 1.0
    # degrees of the step's graph are consistent with its own A_tilde
    assert torch.equal(eng.buffer("adj")[:, :N].float().sum(1), eng.buffer("deg"))


def test_full_size_propagations_and_update_against_plain_fp32_products(big):
    """Independent check at N = 20 000 (no oracle exists at this size): every propagation of the step against a plain
    torch fp32 product on the same sampled graph, and the theta update against the closed form in fp64 on random rows.
    Tolerance 1e-4 (inf-norm relative; bar of the north star: 1e-3)."""
    K, d, w = big
    torch.backends.cuda.matmul.allow_tf32 = False
    eng = K.OuterStep(N, d["x"], d["y"], d["mask"], hidden=d["h"], classes=d["c"])
    eng.set_weights(*w)
    theta0 = d["theta"].clone()
    t = theta0.clone()
    lr = 30.0
    eng.run(t, lr=lr, seed=21, step=4, dropout_p=0.5, update=True)
    a = eng.buffer("adj")[:, :N].float()
    rs = eng.buffer("rsqrt")[:, None]

    def rel(x, ref):
        return float((x - ref).abs().max() / ref.abs().max().clamp_min(1e-30))

    for src, dst in (("p1", "z1"), ("p2", "z2"), ("dz2", "dp2"), ("dz1", "dp1")):
        ref = rs * (a @ (rs * eng.buffer(src)))
        assert rel(eng.buffer(dst), ref) < 1e-4, (src, dst)
    del a
    # theta update on 96 random rows (all columns), closed form of SURVEY.md App. A.2 in fp64
    g = torch.Generator(device="cpu"); g.manual_seed(1)
    rows = torch.randperm(N, generator=g)[:96].cuda()
    fa, fb, cv = eng.buffer("fa").double(), eng.buffer("fb").double(), eng.buffer("cvec").double()
    grad = fa[rows] @ fb.t() + fb[rows] @ fa.t() + cv[rows][:, None] + cv[None, :]
    grad[torch.arange(96, device="cuda"), rows] = 0.0
    th = theta0[rows, :N].double()
    grad = grad * ((th >= 0) & (th <= 1))
    ref = (th - lr * grad).clamp(0, 1)
    assert float((t[rows, :N].double() - ref).abs().max()) < 2e-6 + 1e-5 * lr * float(grad.abs().max())


def test_full_size_row_block_shards_match_the_single_device_step(big):
    """SURVEY.md 8e at N = 20 000: three row-block shards (emulated on one GPU: same kernels, same phases, exchange =
    concatenation of the packed operand blocks) against the unsharded step — same graph bit for bit, theta equal up to
    fp32 summation order, shards mutually consistent (exact symmetry)."""
    K, d, w = big
    from lds_gnn_b200 import sharded as S
    eng = K.OuterStep(N, d["x"], d["y"], d["mask"], hidden=d["h"], classes=d["c"])
    eng.set_weights(*w)
    ref = d["theta"].clone()
    lr = 30.0
    sc_ref = eng.run(ref, lr=lr, seed=9, step=3, dropout_p=0.5, update=True).clone()
    deg_ref = eng.buffer("deg").clone()
    del eng
    world = 3
    bounds = [S.shard_bounds(N, world, r) for r in range(world)]
    shards, thetas = [], []
    for lo, cnt in bounds:
        sh = S.ShardedOuterStep(N, lo, cnt, d["x"][lo:lo + cnt], d["y"][lo:lo + cnt], d["mask"][lo:lo + cnt], int(d["mask"].sum().item()),
                                d["h"], d["c"])
        sh.set_weights(*w)
        shards.append(sh)
        thetas.append(d["theta"][lo:lo + cnt].clone())
    sc = S.run_local_group(shards, thetas, lr=lr, seed=9, step=3, dropout_p=0.5, update=True)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([s.eng.buffer("deg") for s in shards]), deg_ref), "shards must sample the same graph"
    new = torch.cat(thetas, dim=0)
    assert (new - ref).abs().max().item() < 2e-5
    assert torch.equal(new[:, :N], new[:, :N].t())
    assert abs(sc[0].item() - sc_ref[0].item()) < 1e-5 and abs(sc[1].item() - sc_ref[1].item()) < 1e-6


def test_full_size_unrolled_block_factored_route_matches_dense_autograd(big):
    """The unrolled bilevel block (two inner steps with the differentiable Adam, then the hyper step through them,
    src/trainers/bilevel.py:53-73) at N = 20 000: the factored route (K2 for every product with a graph, factor pairs of all
    three graphs folded into theta by one K3+K4 pass, d ~ 430 columns) against dense N x N autograd through the same unroll
    (the composable route, ~20 GB of temporaries) on the same Philox graphs. Also pins the hi/lo-split K3 at large d and N."""
    import lds_gnn_b200.models.gcn as gcn_mod
    from lds_gnn_b200.models.gcn import MetaDenseGCN
    from lds_gnn_b200.models.graph import BernoulliGraphModel
    from lds_gnn_b200.models.sampling import PHILOX
    from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
    from lds_gnn_b200.trainers.inner import InnerProblemTrainer
    from lds_gnn_b200.trainers.outer import OuterProblemTrainer
    from lds_gnn_b200.utils.graph import DenseData
    K, d, w = big
    dev = torch.device("cuda")
    train_mask = torch.zeros(N, dtype=torch.bool, device=dev)
    train_mask[torch.arange(0, N, 143, device=dev)] = True
    data = DenseData(x=d["x"], y=d["y"], train_mask=train_mask, val_mask=d["mask"], test_mask=d["mask"], num_classes=d["c"])
    theta0 = d["theta"][:, :N].contiguous()
    results = {}
    for route in ("factored", "composable"):
        torch.manual_seed(4)
        gcn = MetaDenseGCN(d["f"], d["h"], d["c"], dropout=0.0).to(dev)
        with torch.no_grad():
            for p, v in zip((gcn.layer_in.fc.weight, gcn.layer_in.fc.bias, gcn.layer_out.fc.weight, gcn.layer_out.fc.bias), w):
                p.copy_(v)
        inner = InnerProblemTrainer(gcn, data, lr=0.01, weight_decay=5e-4)
        # Adam's update lr m / (sqrt(v_hat) + eps) is sign-like where |g| ~ eps, and its derivative (which the hypergradient
        # flows through) is a difference of terms ~ 1 / |g|: with the default eps = 1e-8 and weight gradients of ~1e-6 at this
        # size, two equally valid fp32 evaluations of the same unroll differ by percents. A larger eps keeps the comparison of
        # the two ROUTES well conditioned (what is tested here is the factored machinery, not Adam's conditioning).
        for group in inner.optimizer.param_groups:
            group["eps"] = 1e-4
        inner.optimizer._vectors = None
        model = BernoulliGraphModel(theta0).to(dev)
        outer = OuterProblemTrainer(optimizer=torch.optim.SGD(model.parameters(), lr=2.0e4), data=data,   # lr: gradients are ~1e-7 at degree ~10 000
                                     opt_mask=d["mask"], model=model,
                                    smoothness_factor=0.0, disconnection_factor=0.0, sparsity_factor=0.0, regularize=False, lr_decay=None)
        outer.factored_enabled = route == "factored"
        runner = BilevelProblemRunner(inner, outer, data)
        PHILOX.manual_seed(31)
        for _ in range(2):
            runner.inner_opt_step()
        before = model.theta_full().clone()
        m = outer.train_step(inner.model_forward, retain_graph=False)
        assert outer.last_route == route
        inner.detach()
        after = model.theta_full()
        step = (after - before).abs().max().item()
        results[route] = (m, after[:, :N].clone(), step)
        del runner, outer, model, inner, gcn, before, after
        torch.cuda.empty_cache()
    (mf, tf, sf), (mc, tc, sc) = results["factored"], results["composable"]
    assert abs(mf.loss - mc.loss) < 1e-4 and abs(mf.acc - mc.acc) < 1e-6
    assert sc > 1e-4, "the hyper step must move theta for the comparison to mean anything"
    assert torch.equal(tf, tf.t()), "theta must stay exactly symmetric"
    assert (tf - tc).abs().max().item() <= 2e-3 * sc + 1e-6, ((tf - tc).abs().max().item(), sc)
