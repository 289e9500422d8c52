"""oracle/reference_port.py — TEST INFRASTRUCTURE. torch/CPU port of the reference's outer step.

This is the timed CPU baseline (`bench.py --impl reference`, `cpu_baseline.kind == "port"`): the
Python reference cannot travel to the GPU box, so this file restates, op for op, the tensor program
the reference's own CPU run executes for one `OuterProblemTrainer.train_step`
(src/trainers/outer.py:57-87) — including its per-call `triu_indices` rebuild
(src/utils/graph.py:174-176), the full N x N Bernoulli draw (src/models/sampling.py:68), the two
dense N x N x N products of the normalisation (src/utils/graph.py:150-152) and autograd's backward
through all of it. `tests/test_oracle_vs_reference.py` checks it bit-for-bit against the live
reference (same torch RNG state => same draws => identical tensors) whenever /root/reference exists.
"""
from math import sqrt

import torch
import torch.nn.functional as F


def theta_matrix(theta_triu):
    """src/utils/graph.py:166-181 (+ :27-38, :184-192)."""
    n = int(0.5 * sqrt((8 * theta_triu.size(0) + 1) - 1))
    rows, cols = torch.triu_indices(n, n, device=theta_triu.device)
    dense = torch.zeros((n, n), device=theta_triu.device)
    dense[rows, cols] = theta_triu
    upper = dense.triu(1)
    dense = (upper + upper.t()) + torch.diag(dense.diag())
    return dense.clamp(0.0, 1.0)


def draw_graph(theta):
    """src/models/sampling.py:47-85 with undirected=True, sparsification NONE, dense=False."""
    with torch.no_grad():
        drawn = torch.bernoulli(theta)
    upper = drawn.triu(1)
    drawn = (upper + upper.t()) + torch.diag(drawn.diag())
    return (drawn - theta).detach() + theta            # straight-through estimator


def normalise(adj):
    """src/utils/graph.py:123-153 — self loops, row-sum degrees, two dense matmuls with diag matrices."""
    looped = adj.clone()
    looped.fill_diagonal_(1.0)
    inv_sqrt = torch.diag(1.0 / looped.sum(dim=1).sqrt())
    return inv_sqrt @ looped @ inv_sqrt


def gcn_log_probs(x, adj, w0, b0, w1, b1, dropout, training):
    """src/models/gcn.py:23-34 + src/models/layers.py:42-44 (bias before propagation)."""
    a_hat = normalise(adj)
    hid = F.dropout(x, dropout, training=training)
    hid = F.relu(torch.mm(a_hat, F.linear(hid, w0, b0)))
    hid = F.dropout(hid, dropout, training=training)
    return F.log_softmax(torch.mm(a_hat, F.linear(hid, w1, b1)), dim=1)


class ReferenceOuterStep:
    """One object holding theta (T,) + SGD + StepLR, stepping like outer.py:57-87."""

    def __init__(self, theta_triu, x, y, mask, w0, b0, w1, b1, lr, lr_decay=None, dropout=0.0):
        self.theta = torch.nn.Parameter(theta_triu.clone())
        self.x, self.y, self.mask = x, y, mask
        self.weights = [t.clone().requires_grad_(True) for t in (w0, b0, w1, b1)]
        self.dropout = dropout
        self.opt = torch.optim.SGD([self.theta], lr=lr)
        self.sched = (torch.optim.lr_scheduler.StepLR(self.opt, step_size=1, gamma=lr_decay)
                      if lr_decay is not None else None)

    def step(self, training=True):
        self.opt.zero_grad()
        graph = draw_graph(theta_matrix(self.theta))
        logp = gcn_log_probs(self.x, graph, *self.weights, self.dropout, training)
        loss = F.nll_loss(logp[self.mask], self.y[self.mask])
        acc = (torch.argmax(logp[self.mask], dim=-1) == self.y[self.mask]).float().mean().item()
        loss.backward()
        self.opt.step()
        if self.sched is not None:
            self.sched.step()
        self.theta.data.clamp_(0.0, 1.0)
        self.last = dict(graph=graph.detach(), logp=logp.detach())
        return loss.item(), acc
