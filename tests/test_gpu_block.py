"""-m gpu tests of the captured bilevel block (lds_gnn_b200/trainers/graph_block.py): tau inner steps + the hyper step
(src/trainers/bilevel.py:53-73) replayed from one CUDA graph must reproduce the step-by-step loop."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_api import _setup, t  # noqa: F401

pytestmark = pytest.mark.gpu
CUDA = "cuda"


@pytest.fixture(autouse=True)
def _fresh_config():
    from lds_gnn_b200.models.sampling import PHILOX, Sampler
    saved = dict(Sampler._ingredient.values)
    PHILOX.manual_seed(1234)
    yield
    Sampler._ingredient.values.update(saved)


@pytest.mark.parametrize("n,step0", [(130, 5), (257, (1 << 32) - 2), (64, (7 << 32) + 11)])
def test_k1_with_device_step_counter_draws_the_same_graph(n, step0):
    """lds_k1_sample_normalize_dstep: step = *step_base + offset read on the device == the by-value step (bit-exact mask)."""
    from lds_gnn_b200 import kernels
    torch.manual_seed(n)
    ld = kernels.padded_ld(n)
    theta = torch.zeros((n, ld), device=CUDA)
    sym = torch.rand(n, n, device=CUDA)
    theta[:, :n] = (sym + sym.t()) / 2
    base = torch.tensor([step0], dtype=torch.int64, device=CUDA)
    for offset in (0, 1, 3):
        a_ref, _, deg_ref, rs_ref = kernels.k1_sample_normalize(theta, n, 99, step0 + offset)
        a_dev, _, deg_dev, rs_dev = kernels.k1_sample_normalize(theta, n, 99, offset, step_base=base)
        assert torch.equal(a_ref, a_dev) and torch.equal(deg_ref, deg_dev) and torch.equal(rs_ref, rs_dev)
    other, _, _, _ = kernels.k1_sample_normalize(theta, n, 99, 0, step_base=base + 1)
    assert not torch.equal(other, a_ref[..., :])                              # the counter is really read


def _runner(g, dropout=0.0, lr=0.5, lr_decay=0.9):
    from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
    data, gcn, inner, model, outer = _setup(g, lr=lr, lr_decay=lr_decay, dropout=dropout)
    with torch.no_grad():
        model.probs.mul_(0.6).add_(0.2)
    runner = BilevelProblemRunner(inner, outer, data, n_samples_empirical_mean=2)
    return runner, inner, outer, model


@pytest.mark.parametrize("name,tau", [("n130_sparse", 3), ("n257_h64", 5), ("n33_twostep", 1)])
def test_captured_block_replays_match_the_eager_loop(name, tau):
    from lds_gnn_b200.models.sampling import PHILOX
    from lds_gnn_b200.trainers.graph_block import CapturedBilevelBlock
    g = load_golden(name)
    out = {}
    for mode in ("graph", "eager"):
        runner, inner, outer, model = _runner(g)
        PHILOX.manual_seed(77)
        assert CapturedBilevelBlock.eligible(runner)
        metrics = []
        if mode == "graph":
            block = CapturedBilevelBlock(runner, tau)
            for _ in range(3):
                metrics += block.replay()
            assert outer.last_route == "factored-graph"
            block.store_state(tau)
        else:
            for _ in range(3):
                for _ in range(tau):
                    metrics.append(runner.inner_opt_step())
                metrics.append(outer.train_step(inner.model_forward))
                assert outer.last_route == "factored"
                inner.detach()
        out[mode] = (metrics, model.probs.detach().clone(), [p.detach().clone() for p in inner.model_params.values()],
                     outer.get_learning_rates(), PHILOX.step, inner.optimizer.state["step"])
    mg, pg, wg, lg, sg, tg = out["graph"]
    me, pe, we, le, se, te = out["eager"]
    assert (sg, tg) == (se, te) == (3 * (tau + 1), 3 * tau) and lg == pytest.approx(le)
    # the replayed block executes the same kernels on the same numbers as the eager loop (learning rate and Adam step size
    # are rounded the same way in both), so the comparison is tight: a looser bound would hide a flipped Bernoulli draw
    for a, b in zip(mg, me):
        assert abs(a.loss - b.loss) < 1e-6 * max(1.0, abs(b.loss)) and abs(a.acc - b.acc) < 1e-6
    for a, b in zip(wg, we):
        assert (a - b).abs().max().item() < 1e-6
    assert (pg - pe).abs().max().item() < 1e-6
    assert (pg - (0.6 * t(g["theta_triu"]) + 0.2)).abs().max().item() > 1e-5          # theta really moved


def test_captured_block_draws_fresh_dropout_masks_and_graphs_per_replay():
    from lds_gnn_b200.trainers.graph_block import CapturedBilevelBlock
    runner, inner, outer, model = _runner(load_golden("n130_sparse"), dropout=0.5)
    block = CapturedBilevelBlock(runner, 2)
    first = block.replay()
    losses = [first[0].loss]
    for _ in range(3):
        inner.reset_weights(); inner.reset_optimizer()                          # same weights-independent randomness check:
        block.resident = False                                                   # every replay starts from fresh weights
        losses.append(block.replay()[0].loss)
    assert len({round(v, 6) for v in losses[1:]}) > 1                            # different graphs / masks every replay
    assert all(np.isfinite(v) for v in losses)


def test_runner_with_captured_blocks_follows_the_step_by_step_loop_incl_early_stop_rollback():
    """Whole `BilevelProblemRunner.train` (src/trainers/bilevel.py:34-101) with dropout 0 (all randomness from the Philox
    counter): blocks + post-hoc early stopping + undo of a hyper step the reference would not have taken == eager loop.
    inner_loop_max_epochs = 7 with tau = 3 stops at step 7, the first step of a block."""
    from lds_gnn_b200.models.sampling import PHILOX
    g = load_golden("n130_sparse")
    out = {}
    for mode in (True, False):
        runner, inner, outer, model = _runner(g)
        runner.graph_blocks = mode
        PHILOX.manual_seed(5)
        np.random.seed(0)
        torch.manual_seed(9)                                                     # reset_weights draws from torch's generator
        runner.train(patience=2, hyper_gradient_interval=3, inner_loop_max_epochs=7, outer_loop_max_epochs=1)
        res = runner.evaluate()
        out[mode] = (model.probs.detach().clone(), outer.get_learning_rates(), PHILOX.step, res,
                     [v.detach().clone() for v in runner.gcn_params.values()])
        if mode:
            assert runner._blocks and outer.last_route == "factored-graph"
    (pg, lg, sg, rg, wg), (pe, le, se, re_, we) = out[True], out[False]
    assert sg == se and lg == pytest.approx(le)
    assert (pg - pe).abs().max().item() < 1e-5
    for a, b in zip(wg, we):
        assert (a - b).abs().max().item() < 1e-5
    for key in rg:
        assert abs(rg[key] - re_[key]) < 1e-4


def test_capture_failure_falls_back_to_the_step_by_step_loop_with_state_intact(monkeypatch):
    """If the block cannot be captured (here: its body raises while the stream is capturing), the runner must carry on step by
    step from exactly the state at block entry — the eager warm-up pass and the aborted capture leave no trace in theta, the fast
    weights, the optimiser, the Philox counter or the LR schedule."""
    import lds_gnn_b200.trainers.graph_block as GB
    from lds_gnn_b200.models.sampling import PHILOX
    g = load_golden("n130_sparse")
    real_body = GB.CapturedBilevelBlock._body

    def failing_body(self):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("simulated capture failure")
        return real_body(self)

    out = {}
    for mode in ("broken-capture", "eager"):
        runner, inner, outer, model = _runner(g)
        runner.graph_blocks = mode == "broken-capture"
        if mode == "broken-capture":
            monkeypatch.setattr(GB.CapturedBilevelBlock, "_body", failing_body)
        else:
            monkeypatch.setattr(GB.CapturedBilevelBlock, "_body", real_body)
        PHILOX.manual_seed(5)
        np.random.seed(0)
        torch.manual_seed(9)
        runner.train(patience=2, hyper_gradient_interval=3, inner_loop_max_epochs=7, outer_loop_max_epochs=1)
        res = runner.evaluate()
        out[mode] = (model.probs.detach().clone(), outer.get_learning_rates(), PHILOX.step, res)
        if mode == "broken-capture":
            assert runner.graph_blocks is False and outer.last_route == "factored"      # it did fall back
    (pb, lb, sb, rb), (pe, le, se, re_) = out["broken-capture"], out["eager"]
    assert sb == se and lb == pytest.approx(le)
    assert (pb - pe).abs().max().item() < 1e-6
    for key in rb:
        assert abs(rb[key] - re_[key]) < 1e-5


@pytest.mark.parametrize("wd,on_device", [(5e-4, False), (0.0, True), ("vector", True)])
def test_fused_adam_step_matches_the_elementwise_form_and_its_autograd(wd, on_device):
    """lds_adam_step / lds_adam_step_backward (one launch each) against the same update written as differentiable fp64 torch
    ops — values and the vector-Jacobian product the hyper step takes through it (src/trainers/inner.py:42-50 over higher's
    Adam), incl. entries whose second moment is floored, hyper-parameters by value and from device memory."""
    from lds_gnn_b200.trainers.diffopt import _FusedAdamStep
    torch.manual_seed(7)
    n = 70_001
    b1, b2, eps, step_size, root_scale = 0.9, 0.999, 1e-8, 0.0123, 3.7
    p = torch.randn(n, device=CUDA)
    m = torch.randn(n, device=CUDA) * 0.1
    v = torch.rand(n, device=CUDA) * 1e-2 + 1e-3                              # (a vanishing second moment makes the map ill-conditioned in fp32)
    g = torch.randn(n, device=CUDA) * 0.3
    if isinstance(wd, float) and wd == 0.0:
        v[:64] = 0.0; g[:64] = 0.0                                            # v' == 0: the floored branch (no gradient through sqrt)
    ups = [torch.randn(n, device=CUDA) for _ in range(3)]
    if wd == "vector":                                                        # weight decay on the first "layer" only (parameter groups)
        wd = torch.zeros(n, device=CUDA); wd[: n // 2] = 5e-3
    wd64 = wd.double() if isinstance(wd, torch.Tensor) else wd

    def reference(p, m, v, g):
        g2 = g + wd64 * p
        mn = m + (1 - b1) * (g2 - m)
        vn = b2 * v + (1 - b2) * g2 * g2
        root = vn.clamp_min(1e-30).sqrt()
        return p - step_size * mn / (root * root_scale + eps), mn, vn

    ins64 = [x.double().requires_grad_(True) for x in (p, m, v, g)]
    outs64 = reference(*ins64)
    grads64 = torch.autograd.grad(outs64, ins64, [u.double() for u in ups])
    ins = [x.clone().requires_grad_(True) for x in (p, m, v, g)]
    ss = torch.tensor([step_size], device=CUDA) if on_device else step_size
    rs = torch.tensor([root_scale], device=CUDA) if on_device else root_scale
    outs = _FusedAdamStep.apply(*ins, wd, b1, b2, eps, ss, rs)
    for a, b in zip(outs, outs64):
        assert (a.double() - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item())
    grads = torch.autograd.grad(outs, ins, ups)
    for a, b in zip(grads, grads64):
        assert torch.isfinite(a).all()
        # fp32 against fp64: relative to the entry (tiny second moments give gradients ~ 1e4) plus an absolute term for entries whose
        # two contributions (through m' and through v') cancel
        assert ((a.double() - b).abs() <= 3e-5 * b.abs() + 1e-4 * b.abs().mean()).all()
    # only p' used downstream (the last step of an unroll): the state gradients arrive as None
    (gp_only,) = torch.autograd.grad(_FusedAdamStep.apply(*ins, wd, b1, b2, eps, ss, rs)[0], ins[3], ups[0])
    (ref_only,) = torch.autograd.grad(reference(*ins64)[0], ins64[3], ups[0].double())
    assert ((gp_only.double() - ref_only).abs() <= 3e-5 * ref_only.abs() + 1e-4 * ref_only.abs().mean()).all()


@pytest.mark.parametrize("n,c,m", [(300, 6, 40), (1000, 7, 1000), (257, 16, 1), (3327, 6, 120)])
def test_fused_masked_nll_matches_torch_to_second_order(n, c, m):
    """lds_masked_nll_{forward,grad,grad_grad} against F.nll_loss(F.log_softmax(z)[mask], y[mask]) in fp64: loss, accuracy, the
    gradient w.r.t. the logits, and the gradient of a functional of that gradient (the double backward of the hyper step,
    src/trainers/inner.py:63-71)."""
    import torch.nn.functional as F
    from lds_gnn_b200.utils.rowops import MaskInfo, masked_nll
    torch.manual_seed(n + c)
    z = (torch.randn(n, c, device=CUDA) * 2).requires_grad_(True)
    y = torch.randint(0, c, (n,), device=CUDA)
    mask = torch.zeros(n, dtype=torch.bool, device=CUDA)
    mask[torch.randperm(n, device=CUDA)[:m]] = True
    info = MaskInfo(mask, y)
    probe = torch.randn(n, c, device=CUDA)

    loss, acc = masked_nll(z, info)
    (dz,) = torch.autograd.grad(loss, z, create_graph=True)
    (ddz,) = torch.autograd.grad((dz * probe).sum(), z)

    z64 = z.detach().double().requires_grad_(True)
    logp = F.log_softmax(z64, dim=1)
    loss64 = F.nll_loss(logp[mask], y[mask])
    acc64 = (logp[mask].argmax(1) == y[mask]).double().mean()
    (dz64,) = torch.autograd.grad(loss64, z64, create_graph=True)
    (ddz64,) = torch.autograd.grad((dz64 * probe.double()).sum(), z64)
    assert abs(loss.item() - loss64.item()) <= 1e-6 * max(1.0, abs(loss64.item()))
    assert abs(acc.item() - acc64.item()) <= 1e-6
    assert (dz.double() - dz64).abs().max().item() <= 1e-6 * max(1e-3, dz64.abs().max().item())
    assert (ddz.double() - ddz64).abs().max().item() <= 2e-6 * max(1e-3, ddz64.abs().max().item())
    assert dz[~mask].abs().max().item() == 0.0 if m < n else True


def test_row_dot2_matches_torch():
    from lds_gnn_b200 import kernels
    torch.manual_seed(1)
    n, w = 3327, 16
    a1, b1, a2, b2 = (torch.randn(n, w, device=CUDA) for _ in range(4))
    r = torch.rand(n, device=CUDA) + 0.1
    out = kernels.row_dot2(a1, b1, a2, b2, r)
    ref = ((a1.double() * b1.double()).sum(1) + (a2.double() * b2.double()).sum(1)) / r.double()
    assert (out.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
