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
