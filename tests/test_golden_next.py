"""Goldens of SURVEY.md §8 rows (f)1, (f)2 and BASELINE config 3, produced by the LIVE reference
(oracle/make_golden.py::next_rows): `empirical_mean_loss` (src/utils/evaluation.py:51-84), the S-sample
accumulated-gradient outer step (the reference's own objects composed as src/trainers/outer.py:57-87 does) and the
unrolled bilevel block (src/trainers/bilevel.py:103-113 + src/trainers/inner.py:55-74 over the `higher` stand-in).

CPU tests (-m "not gpu") pin the oracle restatement and the package's host logic to those goldens; the -m gpu tests
compare the CUDA path through the reference-facing API. Parity of `higher.DifferentiableAdam` ITSELF stays unpinned:
the package is absent and the goldens were produced over oracle/shims/higher/optim.py, which restates
higher master's published update rule.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import philox as PH
from oracle import restatement as R

EVAL = ["eval_n130_s4", "eval_n257_s16"]
MULTI = ["multi_n130_s4", "multi_n300_s16"]
BLOCK = ["blk_n130_tau3", "blk_n257_tau5", "blk_n96_tau1"]
WNAMES = ("w0", "b0", "w1", "b1")


def rel_inf(x, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(np.asarray(x, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


# ====================================================================================================== CPU: oracle + host logic
@pytest.mark.parametrize("name", EVAL)
def test_restatement_reproduces_the_reference_empirical_mean(name):
    g = load_golden(name)
    n, S, seed, step0 = int(g["n"]), int(g["S"]), int(g["philox_seed"]), int(g["step0"])
    rows = []
    for s in range(S):
        u = PH.edge_uniforms(n, seed, step0 + s)
        full = R.theta_full_from_triu(g["theta_triu"])
        a_hat, _, _, _ = R.normalize_adjacency_matrix(R.sample_graph(full, u))
        fwd = R.gcn_forward(a_hat, *(g[k].astype(np.float64) for k in ("x", "w0", "b0", "w1", "b1")))
        rows.append(R.nll_and_accuracy(fwd["logp"], g["y"], g["val_mask"]) + R.nll_and_accuracy(fwd["logp"], g["y"], g["test_mask"]))
    got = np.mean(np.asarray(rows, dtype=np.float64), axis=0)
    ref = g["metrics_f64"]
    assert abs(got[0] - ref[0]) < 1e-12 and abs(got[2] - ref[2]) < 1e-12
    assert abs(got[1] - ref[1]) < 1e-7 and abs(got[3] - ref[3]) < 1e-7          # accuracies pass through fp32 `.mean()`
    assert np.abs(g["metrics_f32"] - ref).max() < 1e-5


@pytest.mark.parametrize("name", MULTI)
def test_restatement_reproduces_the_reference_multi_sample_step(name):
    g = load_golden(name)
    n, f, h, S, p = int(g["n"]), int(g["f"]), int(g["h"]), int(g["S"]), float(g["p"])
    seed, step, lr = int(g["philox_seed"]), int(g["step"]), float(g["lr"])
    grad = 0.0
    losses, accs = [], []
    for s in range(S):
        u = PH.edge_uniforms(n, seed, step, sample=s)
        kx = PH.dropout_keep_mask(n, f, p, seed, step, PH.STREAM_DROP_X, sample=s) if p > 0 else None
        kh = PH.dropout_keep_mask(n, h, p, seed, step, PH.STREAM_DROP_H, sample=s) if p > 0 else None
        o = R.outer_step(g["theta_triu"], u, g["x"], g["w0"], g["b0"], g["w1"], g["b1"], g["y"], g["mask"], lr=lr, p=p, keep_x=kx, keep_h=kh)
        grad = grad + o["d_theta_triu"] / S
        losses.append(o["loss"]); accs.append(o["acc"])
    assert rel_inf(grad, g["grad_triu_f64"]) < 1e-10
    assert abs(np.mean(losses) - float(g["loss_f64"])) < 1e-12 and abs(np.mean(accs) - float(g["acc_f64"])) < 1e-7
    theta32 = g["theta_triu"].astype(np.float64)
    assert np.abs(np.clip(theta32 - lr * grad, 0, 1) - g["theta_new_f64"]).max() < 1e-12
    inside = (g["theta_triu"] >= 0) & (g["theta_triu"] <= 1)
    assert (~inside).sum() == 2 and np.all(g["grad_triu_f64"][~inside] == 0)     # clamp backward masks values outside [0, 1]


def _block_objects(g, device, dtype, dropout=0.0):
    from lds_gnn_b200.models.gcn import MetaDenseGCN
    from lds_gnn_b200.trainers.inner import InnerProblemTrainer
    from lds_gnn_b200.utils.graph import DenseData
    tt = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(a)).to(device=device, dtype=dt)
    n = int(g["n"])
    full = np.zeros((n, n), np.float32)
    full[np.triu_indices(n)] = g["theta_triu"]
    full = np.triu(full, 1) + np.triu(full, 1).T + np.diag(np.diag(full))
    data = DenseData(x=tt(g["x"], dtype), y=tt(g["y"]), train_mask=tt(g["train_mask"]), val_mask=tt(g["opt_mask"]),
                     test_mask=tt(g["opt_mask"]), dense_adj=tt(full, dtype), num_classes=int(g["c"]))
    gcn = MetaDenseGCN(int(g["f"]), int(g["h"]), int(g["c"]), dropout=dropout).to(device=device, dtype=dtype)
    with torch.no_grad():
        for p_, k in zip(gcn.parameters(), WNAMES):
            p_.copy_(tt(g[k], dtype))
    inner = InnerProblemTrainer(gcn, data, lr=float(g["inner_lr"]), weight_decay=float(g["weight_decay"]))
    return data, gcn, inner, full


@pytest.mark.parametrize("name", BLOCK)
def test_factored_unroll_host_logic_reproduces_the_reference_bilevel_block(name, monkeypatch):
    """The package's inner steps (differentiable Adam, factored graphs) and the factored hypergradient, with the one kernel
    they call replaced by a torch stand-in IN THIS TEST (fp64, CPU), against the live reference's block: fast weights after
    every inner step, losses, and probs.grad of the hyper step (second-order terms through the unroll included)."""
    import torch.nn.functional as F
    from lds_gnn_b200 import kernels
    from lds_gnn_b200.models.sampling import FactoredGraph, FactorSink, SampleHandle

    def k2_stub(adj, n, q, scale_in=None, scale_out=None, **kw):
        y = adj @ (q if scale_in is None else scale_in[:, None] * q)
        return y if scale_out is None else scale_out[:, None] * y
    monkeypatch.setattr(kernels, "k2_propagate", k2_stub)
    g = load_golden(name)
    dt = torch.float64
    n, tau, blocks = int(g["n"]), int(g["tau"]), int(g["blocks"])
    seed, step0 = int(g["philox_seed"]), int(g["step0"])
    data, gcn, inner, _ = _block_objects(g, "cpu", dt)
    theta = g["theta_triu"].astype(np.float64)
    lr = float(g["outer_lr"])
    draw = 0
    for b in range(blocks):
        link = torch.zeros(1, dtype=dt, requires_grad=True)
        sink = FactorSink()

        def graph():
            nonlocal draw
            u = PH.edge_uniforms(n, seed, step0 + draw)
            draw += 1
            a = R.add_self_loops(R.sample_graph(R.theta_full_from_triu(theta.astype(np.float32)), u))
            a = torch.as_tensor(a, dtype=dt)
            return FactoredGraph(SampleHandle(n, a, a.sum(1), None, 0, 0), link, sink)
        for k in range(tau):
            m = inner.train_step(graph())
            assert abs(m.loss - g[f"inner_loss{b}_f64"][k]) < 1e-10 and abs(m.acc - g[f"inner_acc{b}_f64"][k]) < 1e-6
            for w, wname in zip(inner.model_params.values(), WNAMES):
                assert rel_inf(w.detach().numpy(), g[f"{wname}_after{b}_{k}_f64"]) < 1e-9, (b, k, wname)
        assert sink.empty()
        pred = inner.model_forward(graph())
        loss = F.nll_loss(pred[data.val_mask], data.y[data.val_mask])
        assert abs(loss.item() - float(g[f"hyper_loss{b}_f64"])) < 1e-10
        loss.backward()
        fa, fb, cvec = sink.collect(n, "cpu")
        dense = fa @ fb.t() + cvec[:, None]
        dense = dense + dense.t()
        dense.fill_diagonal_(0.0)
        grad = dense.numpy()[np.triu_indices(n)] * ((theta >= 0) & (theta <= 1))
        assert rel_inf(grad, g[f"grad_triu{b}_f64"]) < 1e-8
        theta = np.clip(theta - lr * grad, 0.0, 1.0)
        assert np.abs(theta - g[f"theta_new{b}_f64"]).max() < 1e-10
        lr *= float(g["lr_decay"])
        assert abs(lr - float(g[f"lr_after{b}_f64"])) < 1e-12
        inner.detach()


# ====================================================================================================== GPU: CUDA path through the API
CUDA = "cuda"


@pytest.fixture()
def philox():
    from lds_gnn_b200.models.sampling import PHILOX, Sampler
    saved = dict(Sampler._ingredient.values)
    yield PHILOX
    Sampler._ingredient.values.update(saved)


def _outer_objects(g, data, lr, lr_decay, opt_mask):
    from lds_gnn_b200.models.graph import BernoulliGraphModel
    from lds_gnn_b200.trainers.outer import OuterProblemTrainer
    model = BernoulliGraphModel(data.dense_adj).to(CUDA)
    with torch.no_grad():                                   # values outside [0, 1] survive only through probs (the init clamps nothing, but be explicit)
        model.probs.copy_(torch.as_tensor(g["theta_triu"]).to(CUDA))
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    outer = OuterProblemTrainer(optimizer=opt, data=data, opt_mask=opt_mask, model=model, smoothness_factor=0.0,
                                disconnection_factor=0.0, sparsity_factor=0.0, regularize=False, lr_decay=lr_decay, pretrain=False)
    return model, outer


@pytest.mark.gpu
@pytest.mark.parametrize("name", EVAL)
def test_gpu_empirical_mean_loss_matches_live_reference(name, philox):
    from lds_gnn_b200.utils.evaluation import empirical_mean_loss
    g = dict(load_golden(name))
    g.update(train_mask=g["val_mask"], opt_mask=g["val_mask"], inner_lr=0.01, weight_decay=0.0)
    data, gcn, inner, _ = _block_objects(g, CUDA, torch.float32, dropout=0.5)        # eval mode must switch dropout off
    data.val_mask = torch.as_tensor(g["val_mask"]).to(CUDA)
    data.test_mask = torch.as_tensor(g["test_mask"]).to(CUDA)
    model, _ = _outer_objects(g, data, 1.0, None, data.val_mask)
    philox.seed, philox.step = int(g["philox_seed"]), int(g["step0"])
    val, test = empirical_mean_loss(gcn, model, n_samples=int(g["S"]), data=data, model_parameters=None)
    ref = g["metrics_f64"]
    assert philox.step == int(g["step0"]) + int(g["S"])
    assert abs(val.loss - ref[0]) < 1e-4 * max(1.0, abs(ref[0])) and abs(test.loss - ref[2]) < 1e-4 * max(1.0, abs(ref[2]))
    assert abs(val.acc - ref[1]) < 1e-6 and abs(test.acc - ref[3]) < 1e-6
    # explicit fast weights (what the bilevel runner passes) give the same numbers
    philox.seed, philox.step = int(g["philox_seed"]), int(g["step0"])
    val2, _ = empirical_mean_loss(gcn, model, n_samples=int(g["S"]), data=data, model_parameters=inner.model_params)
    assert val2 == val


@pytest.mark.gpu
@pytest.mark.parametrize("name", MULTI)
def test_gpu_multi_sample_outer_step_matches_live_reference(name, philox):
    """BASELINE config 3 through OuterProblemTrainer.train_step with n_samples = S."""
    g = dict(load_golden(name))
    g.update(train_mask=g["mask"], opt_mask=g["mask"], inner_lr=0.01, weight_decay=0.0)
    data, gcn, inner, _ = _block_objects(g, CUDA, torch.float32, dropout=float(g["p"]))
    model, outer = _outer_objects(g, data, float(g["lr"]), None, data.val_mask)
    outer.n_samples = int(g["S"])
    philox.seed, philox.step = int(g["philox_seed"]), int(g["step"])
    m = outer.train_step(inner.model_forward)
    assert outer.last_route == "fused" and philox.step == int(g["step"]) + 1
    assert abs(m.loss - float(g["loss_f64"])) < 1e-4 * max(1.0, float(g["loss_f64"])) and abs(m.acc - float(g["acc_f64"])) < 1e-6
    ref_g = g["grad_triu_f64"]
    new = model.probs.detach().cpu().numpy()
    assert np.abs(new - g["theta_new_f64"]).max() <= 1e-3 * float(g["lr"]) * np.abs(ref_g).max() + 2e-7
    moved = np.abs(g["theta_new_f64"] - np.clip(g["theta_triu"], 0, 1)).max()
    assert moved > 100 * (1e-3 * float(g["lr"]) * np.abs(ref_g).max() + 2e-7)          # the bound is tight relative to the step taken


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["step-by-step", "captured-graph"])
@pytest.mark.parametrize("name", BLOCK)
def test_gpu_bilevel_block_matches_live_reference(name, mode, philox):
    """tau x inner_opt_step + hyper_opt_step (src/trainers/bilevel.py:103-113) on the factored route, eagerly and replayed
    from the captured CUDA graph, against the live reference's block (dense autograd through the whole unroll)."""
    from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
    from lds_gnn_b200.trainers.graph_block import CapturedBilevelBlock
    g = load_golden(name)
    tau, blocks = int(g["tau"]), int(g["blocks"])
    data, gcn, inner, _ = _block_objects(g, CUDA, torch.float32)
    model, outer = _outer_objects(g, data, float(g["outer_lr"]), float(g["lr_decay"]), data.val_mask)
    runner = BilevelProblemRunner(inner, outer, data)
    philox.seed, philox.step = int(g["philox_seed"]), int(g["step0"])
    block = CapturedBilevelBlock(runner, tau) if mode == "captured-graph" else None
    for b in range(blocks):
        if block is not None:
            metrics = block.replay()
            inner_metrics, hyper = metrics[:tau], metrics[tau]
            weights = [list(block.params_after(k).values()) for k in range(tau)]
            assert outer.last_route == "factored-graph"
        else:
            inner_metrics, weights = [], []
            for _ in range(tau):
                inner_metrics.append(runner.inner_opt_step())
                weights.append([v.detach().clone() for v in inner.model_params.values()])
            hyper = outer.train_step(inner.model_forward)
            assert outer.last_route == "factored"
            inner.detach(); outer.detach()
        for k in range(tau):
            assert abs(inner_metrics[k].loss - g[f"inner_loss{b}_f64"][k]) < 1e-4 and abs(inner_metrics[k].acc - g[f"inner_acc{b}_f64"][k]) < 1e-6
            for w, wname in zip(weights[k], WNAMES):
                ref = g[f"{wname}_after{b}_{k}_f64"]
                # Adam's first steps are sign-like: update = lr g / (|g| + eps), so an element whose gradient is tiny (1e-7) but
                # far above eps = 1e-8 turns an ABSOLUTE gradient error of 1e-8 (the hi/lo-split operands carry 2^-17 relative)
                # into a few percent of lr. The bulk of the weights must agree tightly, the worst element within 10 % of the
                # steps taken; the reference's own fp32 run deviates from its fp64 run by up to 1e-3 lr on this case.
                err = np.abs(w.detach().cpu().numpy() - ref)
                lr_in = float(g["inner_lr"])
                assert np.median(err) < 1e-3 * lr_in * (k + 1) and err.max() < 0.1 * lr_in * (k + 1), (b, k, wname, np.median(err), err.max())
        assert abs(hyper.loss - float(g[f"hyper_loss{b}_f64"])) < 1e-4 and abs(hyper.acc - float(g[f"hyper_acc{b}_f64"])) < 1e-6
        ref_g = g[f"grad_triu{b}_f64"]
        lr_b = float(g["outer_lr"]) * float(g["lr_decay"]) ** b
        new = model.probs.detach().cpu().numpy()
        # The hypergradient flows through the derivative of Adam's update, lr eps / (|g| + eps)^2 per weight with eps = 1e-8: a
        # difference of terms ~ 1 / |g| that amplifies the 2^-17 relative error of the hi/lo-split propagation operands exactly
        # on the weights whose inner gradients nearly cancel. One unrolled step (tau = 1) and three (tau = 3, h = 16) meet the
        # 1e-3 bar of the direct step; five steps at h = 64 — the hardest case — land within 2 % (99th percentile) / 5 % (worst
        # element) of the largest step. For scale: the reference's own fp32 run is 1e-3 of that step away from its fp64 run,
        # and the same host logic run in fp64 on the CPU reproduces the reference to 1e-8 (test above).
        err = np.abs(new - g[f"theta_new{b}_f64"])
        step_max = lr_b * np.abs(ref_g).max()
        hard = tau >= 5
        assert np.percentile(err, 99) <= (2e-2 if hard else 1e-3) * step_max + 2e-7, (b, np.percentile(err, 99), step_max)
        assert err.max() <= (5e-2 if hard else 2e-3) * step_max + 1e-6, (b, err.max(), step_max)
        assert outer.get_learning_rates()[0] == pytest.approx(float(g[f"lr_after{b}_f64"]))
    assert philox.step == int(g["step0"]) + blocks * (tau + 1)
