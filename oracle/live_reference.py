"""oracle/live_reference.py — TEST INFRASTRUCTURE. Runs the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference): used by `make_golden.py`
to produce `tests/golden/*.npz` and by `tests/test_oracle_vs_reference.py` (skipped when the
reference is absent). The reference has no hook for explicit uniforms, so `explicit_draws` patches
`torch.bernoulli` (reached by `Bernoulli(probs).sample()`, src/models/sampling.py:68) to return
`(U < p)` — torch's own CPU semantics — and `torch.nn.functional.dropout` (src/models/gcn.py:27,29)
to apply supplied keep-masks.
"""
import contextlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("LDS_REFERENCE_ROOT", "/root/reference")
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def enable():
    """Put the shims and the reference on sys.path (idempotent)."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [SHIMS, REFERENCE_ROOT]


@contextlib.contextmanager
def explicit_draws(uniforms=(), keep_masks=()):
    """uniforms: iterable of (N,N) fp32 arrays consumed by successive torch.bernoulli calls;
    keep_masks: iterable of bool arrays consumed by successive F.dropout(training=True, p>0) calls."""
    import torch.nn.functional as F
    u_iter = iter([torch.as_tensor(np.asarray(u, dtype=np.float32)) for u in uniforms])
    k_iter = iter([torch.as_tensor(np.asarray(k)) for k in keep_masks])
    real_bernoulli, real_dropout = torch.bernoulli, F.dropout

    def fake_bernoulli(p, *a, **k):
        u = next(u_iter)
        assert u.shape == p.shape, (u.shape, p.shape)
        return (u < p).to(p.dtype)

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        keep = next(k_iter).to(x.dtype)
        assert keep.shape == x.shape, (keep.shape, x.shape)
        return x * (keep / (1.0 - p))

    torch.bernoulli, F.dropout = fake_bernoulli, fake_dropout
    try:
        yield
    finally:
        torch.bernoulli, F.dropout = real_bernoulli, real_dropout


def symmetric_from_triu(theta_triu):
    t = len(theta_triu)
    n = int(0.5 * ((8 * t + 1) ** 0.5 - 1) + 0.5)
    full = np.zeros((n, n), dtype=np.float32)
    full[np.triu_indices(n)] = theta_triu
    return np.triu(full, 1) + np.triu(full, 1).T + np.diag(np.diag(full))


def build(theta_triu, x, w0, b0, w1, b1, y, mask, lr, lr_decay=None, p=0.0, dtype=torch.float32):
    """Construct the reference's model / GCN / trainer objects around the given tensors."""
    enable()
    from src.models.graph import BernoulliGraphModel
    from src.models.gcn import MetaDenseGCN
    from src.trainers.outer import OuterProblemTrainer

    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    model = BernoulliGraphModel(t(symmetric_from_triu(np.asarray(theta_triu, dtype=np.float32))))
    # the constructor reads the upper triangle incl. diagonal; overwrite to keep out-of-range values exact
    with torch.no_grad():
        model.probs.copy_(t(theta_triu))
    if dtype != torch.float32:
        model.probs.data = model.probs.data.to(dtype)
    f, h, c = x.shape[1], w0.shape[0], w1.shape[0]
    gcn = MetaDenseGCN(f, h, c, dropout=p)
    with torch.no_grad():
        gcn.layer_in.fc.weight.copy_(t(w0)); gcn.layer_in.fc.bias.copy_(t(b0))
        gcn.layer_out.fc.weight.copy_(t(w1)); gcn.layer_out.fc.bias.copy_(t(b1))
    gcn = gcn.to(dtype)
    data = SimpleNamespace(x=t(x), y=torch.as_tensor(np.asarray(y)).long())
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    trainer = OuterProblemTrainer(optimizer=opt, data=data, opt_mask=torch.as_tensor(np.asarray(mask)).bool(),
                                  model=model, smoothness_factor=0.0, disconnection_factor=0.0,
                                  sparsity_factor=0.0, regularize=False, lr_decay=lr_decay, pretrain=False)
    return model, gcn, trainer, data


def outer_steps(theta_triu, uniforms, x, w0, b0, w1, b1, y, mask, lr, lr_decay=None, p=0.0,
                keep_masks=(), dtype=torch.float32):
    """Run len(uniforms) reference `OuterProblemTrainer.train_step`s (src/trainers/outer.py:57-87)
    with explicit draws; returns a list of per-step dicts of numpy arrays."""
    old_default = torch.get_default_dtype()
    torch.set_default_dtype(dtype)      # the reference allocates with torch.zeros(...) (utils/graph.py:175)
    try:
        return _outer_steps(theta_triu, uniforms, x, w0, b0, w1, b1, y, mask, lr, lr_decay, p, keep_masks, dtype)
    finally:
        torch.set_default_dtype(old_default)


def _outer_steps(theta_triu, uniforms, x, w0, b0, w1, b1, y, mask, lr, lr_decay, p, keep_masks, dtype):
    model, gcn, trainer, data = build(theta_triu, x, w0, b0, w1, b1, y, mask, lr, lr_decay, p, dtype)
    gcn.train(p > 0.0)
    results = []
    captured = {}

    def predict(graph):
        captured["graph"] = graph.detach().clone()
        gcn.train(p > 0.0)
        pred = gcn(data.x, graph, params=None)
        captured["logp"] = pred.detach().clone()
        return pred

    with explicit_draws(uniforms, keep_masks):
        for _ in range(len(uniforms)):
            lr_used = trainer.get_learning_rates()[0]
            metrics = trainer.train_step(predict, retain_graph=False)
            results.append(dict(
                sample=captured["graph"].numpy().copy(),
                logp=captured["logp"].numpy().copy(),
                loss=np.float64(metrics.loss), acc=np.float64(metrics.acc),
                grad_triu=model.probs.grad.detach().numpy().copy(),
                theta_new=model.probs.detach().numpy().copy(),
                lr_used=np.float64(lr_used),
            ))
    stats = model.statistics()
    return results, stats


# ------------------------------------------------------------------------------------------------
# Rows (f)1, (f)2 and BASELINE config 3 of SURVEY.md §8: evaluation, unrolled bilevel block, S-sample step
# ------------------------------------------------------------------------------------------------
def _with_dtype(dtype, fn):
    old_default = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        return fn()
    finally:
        torch.set_default_dtype(old_default)


def empirical_mean(theta_triu, uniforms, x, w0, b0, w1, b1, y, val_mask, test_mask, dtype=torch.float32):
    """The reference's `empirical_mean_loss` (src/utils/evaluation.py:51-84) with one explicit uniform matrix per sample.
    Returns (val loss, val acc, test loss, test acc)."""
    def run():
        model, gcn, _, data = build(theta_triu, x, w0, b0, w1, b1, y, val_mask, lr=1.0, p=0.5, dtype=dtype)
        from src.utils.evaluation import empirical_mean_loss
        data.val_mask = torch.as_tensor(np.asarray(val_mask)).bool()
        data.test_mask = torch.as_tensor(np.asarray(test_mask)).bool()
        with explicit_draws(uniforms, ()):
            val, test = empirical_mean_loss(gcn, graph_model=model, n_samples=len(uniforms), data=data,
                                            model_parameters=None)
        return np.array([val.loss, val.acc, test.loss, test.acc], dtype=np.float64)
    return _with_dtype(dtype, run)


def multi_sample_step(theta_triu, uniforms, x, w0, b0, w1, b1, y, mask, lr, p=0.0, keep_masks=(), dtype=torch.float32):
    """BASELINE config 3: S Bernoulli samples per outer step. The reference has no such trainer method; this composes ITS
    objects the way its train_step does (src/trainers/outer.py:57-87): zero_grad; S x {model.sample(); gcn forward; nll / S;
    backward -> probs.grad accumulates}; optimizer.step(); project_parameters()."""
    def run():
        import torch.nn.functional as F
        model, gcn, trainer, data = build(theta_triu, x, w0, b0, w1, b1, y, mask, lr=lr, p=p, dtype=dtype)
        from src.utils.evaluation import accuracy
        s_total = len(uniforms)
        m = trainer.opt_mask
        losses, accs = [], []
        with explicit_draws(uniforms, keep_masks):
            trainer.model.train()
            trainer.optimizer.zero_grad()
            for _ in range(s_total):
                graph = trainer.model.sample()
                gcn.train(p > 0.0)
                pred = gcn(data.x, graph, params=None)
                loss = F.nll_loss(pred[m], data.y[m])
                losses.append(loss.item()); accs.append(accuracy(pred[m], data.y[m]))
                (loss / s_total).backward()
            grad = model.probs.grad.detach().numpy().copy()
            trainer.optimizer.step()
            trainer.model.project_parameters()
        return dict(loss=np.float64(np.mean(losses)), acc=np.float64(np.mean(accs)), grad_triu=grad,
                    theta_new=model.probs.detach().numpy().copy())
    return _with_dtype(dtype, run)


def bilevel_blocks(theta_triu, uniforms, x, w0, b0, w1, b1, y, train_mask, opt_mask, tau, blocks, outer_lr, lr_decay,
                   inner_lr, weight_decay, dtype=torch.float32):
    """`blocks` x {tau x BilevelProblemRunner.inner_opt_step; hyper_opt_step} of the unmodified reference
    (src/trainers/bilevel.py:103-113, src/trainers/inner.py:55-74, src/trainers/outer.py:57-87) over the `higher` stand-in
    (oracle/shims/higher/optim.py), dropout 0, one explicit uniform matrix per `sample()` call (tau + 1 per block).
    Returns per block: inner (loss, acc) per step, the fast weights after every inner step, the hyper step's (loss, acc),
    probs.grad and the updated probs."""
    def run():
        model, gcn, outer, data = build(theta_triu, x, w0, b0, w1, b1, y, opt_mask, lr=outer_lr, lr_decay=lr_decay, p=0.0, dtype=dtype)
        from src.trainers.bilevel import BilevelProblemRunner
        from src.trainers.inner import InnerProblemTrainer
        data.train_mask = torch.as_tensor(np.asarray(train_mask)).bool()
        data.val_mask = torch.as_tensor(np.asarray(opt_mask)).bool()
        data.test_mask = data.val_mask
        inner = InnerProblemTrainer(gcn, data, lr=inner_lr, weight_decay=weight_decay)
        runner = BilevelProblemRunner(inner, outer, data)
        out = []
        with explicit_draws(uniforms, ()):
            for _ in range(blocks):
                rec = dict(inner_loss=[], inner_acc=[], weights=[])
                for _ in range(tau):
                    m = runner.inner_opt_step()
                    rec["inner_loss"].append(m.loss); rec["inner_acc"].append(m.acc)
                    rec["weights"].append([v.detach().numpy().copy() for v in inner.model_params.values()])
                captured = {}
                real = outer.train_step

                def spy(fct, *a, **k):
                    captured["m"] = real(fct, *a, **k)
                    captured["grad"] = model.probs.grad.detach().numpy().copy()
                    return captured["m"]
                outer.train_step = spy
                runner.hyper_opt_step(0)
                outer.train_step = real
                rec.update(hyper_loss=captured["m"].loss, hyper_acc=captured["m"].acc, grad_triu=captured["grad"],
                           theta_new=model.probs.detach().numpy().copy(), lr_after=outer.get_learning_rates()[0])
                out.append(rec)
        return out
    return _with_dtype(dtype, run)
