"""Graph sampling with the reference's API (src/models/sampling.py): `sample_graph`, `Sampler.sample`,
`straight_through_estimator`, `SPARSIFICATION`.

The LDS path (undirected=True, sparsification NONE, dense=False — the only combination any LDS config
uses, configs/sacred/lds/config.json, configs/seml/final/lds.yaml) runs K1: one fused CUDA pass that draws
counter-based Philox uniforms for the upper triangle only, mirrors, and at the same time produces the
bf16 self-looped adjacency, the degrees and deg^-1/2 that the GCN needs next. The returned tensor is the
dense {0,1} sample (sampled diagonal kept, like the reference) with a straight-through gradient to the
probabilities; the by-products ride along as `graph._lds_handle`.

KNN / EPS sparsification and `dense=True` belong to the GAE/GRCN model families (out of scope, SURVEY.md
§2 row 2) and raise NotImplementedError.
"""
from enum import Enum
from typing import Optional

import torch
from torch import Tensor

from ..config import Ingredient
from ..utils.graph import is_square_matrix, to_undirected


class SPARSIFICATION(Enum):
    NONE = 1
    KNN = 2
    EPS = 3


class PhiloxState:
    """Seed + step counter of the device RNG. The seed is drawn from torch's global generator on first use,
    so `torch.manual_seed` (what sacred's seeding does) makes runs reproducible.

    While a bilevel block is being captured into a CUDA graph (`begin_capture`), `next_step` hands out OFFSETS relative to
    a device-resident counter instead of absolute steps: the sampling kernel reads the counter itself, so every replay of
    the graph draws fresh graphs (the host sets the counter to `step` before a replay and advances `step` after it)."""

    def __init__(self):
        self.seed = None
        self.step = 0
        self.capture_base = None         # int64 device tensor [1] while capturing, else None
        self.capture_draws = 0
        self._derived_from = None        # torch.initial_seed() the current seed was derived from (None: set explicitly)
        self._torch_reseeded = False     # torch.manual_seed ran since the seed was derived (see _watch_torch_seed below)

    def manual_seed(self, seed: int):
        self.seed = int(seed) & ((1 << 64) - 1)
        self.step = 0
        self._derived_from = None

    def _ensure_seed(self):
        """A seed derived from torch's generator follows it: after a later `torch.manual_seed(s)` (sacred seeds every run
        that way) the next draw re-derives the seed and restarts the step counter, so a second run in the same process is
        reproducible too. A seed set with `manual_seed` (or assigned) stays."""
        if self.seed is not None and self._derived_from is not None and self.capture_base is None \
                and (self._torch_reseeded or torch.initial_seed() != self._derived_from):
            self.seed = None
        if self.seed is None:
            self._derived_from = torch.initial_seed()
            self._torch_reseeded = False
            self.seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            self.step = 0

    def next_step(self):
        self._ensure_seed()
        if self.capture_base is not None:
            offset = self.capture_draws
            self.capture_draws += 1
            return self.seed, offset
        step = self.step
        self.step += 1
        return self.seed, step

    def begin_capture(self, base: Tensor):
        self._ensure_seed()
        self.capture_base, self.capture_draws = base, 0

    def end_capture(self) -> int:
        draws = self.capture_draws
        self.capture_base, self.capture_draws = None, 0
        return draws


PHILOX = PhiloxState()


def _watch_torch_seed():
    """`torch.manual_seed(s)` with the SAME s as before leaves `torch.initial_seed()` unchanged, so a second sacred run with the
    same seed in one process could not be told from a continuation. The global seeding entry points are wrapped once to leave
    a note for PHILOX (nothing else changes: the wrappers call straight through)."""
    if getattr(torch.manual_seed, "_lds_watched", False):
        return
    original = torch.manual_seed

    def manual_seed(seed):
        PHILOX._torch_reseeded = True
        return original(seed)
    manual_seed._lds_watched = True
    manual_seed.__doc__ = original.__doc__
    torch.manual_seed = manual_seed
    if getattr(torch.random, "manual_seed", None) is original:
        torch.random.manual_seed = manual_seed


_watch_torch_seed()


class SampleHandle:
    """By-products of K1 for one sampled graph."""

    def __init__(self, n, adj, deg, rsqrt, seed, step):
        self.n, self.adj, self.deg, self.rsqrt, self.seed, self.step = n, adj, deg, rsqrt, seed, step


_LAST_HANDLE = {}


class _SampleSTE(torch.autograd.Function):
    """value = Bernoulli sample (upper-triangle draw mirrored); d value / d edge_probs = identity for EVERY entry —
    the reference applies the estimator to the matrix it was handed (src/models/sampling.py:77-78, 82-85)."""

    @staticmethod
    def forward(ctx, edge_probs, kernel_input, seed, step, explicit_uniforms):
        from .. import kernels
        n = kernel_input.shape[0]
        ld = kernels.padded_ld(n)
        src = kernel_input.detach()
        if not (src.stride(1) == 1 and src.stride(0) == ld and src.data_ptr() % 16 == 0):
            padded = torch.zeros((n, ld), dtype=torch.float32, device=src.device)
            padded[:, :n] = src
            src = padded
        adj, sample, deg, rs = kernels.k1_sample_normalize(src, n, seed, step, u=explicit_uniforms, want_sample=True)
        _LAST_HANDLE["handle"] = SampleHandle(n, adj, deg, rs, seed, step)     # picked up by sample_graph
        return sample

    @staticmethod
    def backward(ctx, grad):
        return grad, None, None, None, None


class FactorSink:
    """Where the hypergradient on theta of FACTORED graphs is collected. With the straight-through estimator every sampled
    graph has d(sample)/d(theta_full) = I, so the contributions of all graphs of an unroll simply add up:
        dL/dtheta_full[i][j] = sum_k a_k[i] . b_k[j]  +  c[i]        (i != j; symmetrised by the update kernel)
    Each differentiable product with a sampled graph deposits one factor pair (a_k, b_k) = (dY, Q) of width <= hidden, each
    differentiable degree one row-constant term — O(N h) per deposit instead of an N x N autograd temporary. The update
    kernel (K3+K4) consumes the K-concatenation of all pairs in ONE pass over theta."""

    def __init__(self):
        self.fa, self.fb, self.scale, self.c = [], [], [], None

    def clear(self):
        self.fa, self.fb, self.scale, self.c = [], [], [], None

    def add_outer(self, a: Tensor, b: Tensor, row_scale: Tensor = None):
        """Deposit the pair (s * a, s * b), s = row_scale[:, None] (or 1): the scaling is applied once, at collect time."""
        self.fa.append(a)
        self.fb.append(b)
        self.scale.append(row_scale)

    def add_row_constant(self, c: Tensor):
        """Row-constant terms are summed once, at collect time (one stacked reduction instead of one add per deposit)."""
        if self.c is None:
            self.c = [c]
        else:
            self.c.append(c)

    def empty(self) -> bool:
        return not self.fa and self.c is None

    def collect(self, n: int, device):
        """(fa [n, d], fb [n, d], c [n]) fp32, contiguous; d >= 1."""
        if self.fa:
            fa = torch.cat(self.fa, dim=1)
            fb = torch.cat(self.fb, dim=1)
            if any(s is not None for s in self.scale):
                cols = torch.cat([(torch.ones(n, dtype=a.dtype, device=a.device) if s is None else s)[:, None].expand(n, a.shape[1])
                                  for a, s in zip(self.fa, self.scale)], dim=1)
                fa, fb = fa * cols, fb * cols
            fa, fb = fa.contiguous(), fb.contiguous()
        else:
            fa = torch.zeros((n, 1), dtype=torch.float32, device=device)
            fb = torch.zeros((n, 1), dtype=torch.float32, device=device)
        if self.c is None:
            c = torch.zeros(n, dtype=torch.float32, device=device)
        elif len(self.c) == 1:
            c = self.c[0].contiguous()
        else:
            c = torch.stack(self.c).sum(dim=0)
        return fa, fb, c


class _FactoredMatmul(torch.autograd.Function):
    """Y = A_tilde Q for a factored sampled graph (K2 on the tensor cores), differentiable to ANY order: the backward is
    the same product (A_tilde is symmetric) applied through this Function again, so a later backward can differentiate
    through it (the unrolled inner steps, src/trainers/inner.py:71). The gradient on the graph, dY Q^T, is never formed:
    when the backward runs as a plain (non-create_graph) pass — the hypergradient pass of outer.py:77 — the factor pair is
    deposited in the graph's FactorSink. A create_graph backward is the inner optimiser differentiating w.r.t. the GCN
    weights only (inner.py:71 -> autograd.grad(loss, params)), which sends nothing to theta, so nothing is deposited."""

    @staticmethod
    def forward(ctx, q, link, fg):
        from .. import kernels
        q = q.contiguous()
        y = kernels.k2_propagate(fg.handle.adj, fg.handle.n, q)
        ctx.save_for_backward(q)
        ctx.fg = fg
        return y

    @staticmethod
    def backward(ctx, dy):
        (q,) = ctx.saved_tensors
        fg = ctx.fg
        if not torch.is_grad_enabled():
            fg.sink.add_outer(dy.detach(), q.detach())
        dq = _FactoredMatmul.apply(dy, fg.link, fg) if ctx.needs_input_grad[0] else None
        return dq, None, None


class _FactoredPropagate(torch.autograd.Function):
    """Z = r * (A_tilde (r * Q)) — the normalised propagation (src/utils/graph.py:148-152 + src/models/layers.py:44) as ONE K2
    launch: r rides in the operand pack and in the epilogue. Differentiable to any order in Q and r:
        dQ = r * (A_tilde (r * dZ))            (this Function again; A_tilde is symmetric)
        dr = (sum_c dZ*Z + sum_c dQ*Q) / r     (ordinary differentiable ops; Z is this node's own output)
    and the gradient on the graph is the pair (r*dZ, r*Q), deposited unscaled together with r (FactorSink scales once)."""

    @staticmethod
    def forward(ctx, q, r, link, fg, own_r=False):
        from .. import kernels
        ctx.own_r = bool(own_r)                              # r is this graph's own deg^-1/2 (FactoredGraph.normalized)
        q, r = q.contiguous(), r.contiguous()
        z = kernels.k2_propagate(fg.handle.adj, fg.handle.n, q, r, r)
        ctx.save_for_backward(q, r, z)
        ctx.fg = fg
        return z

    @staticmethod
    def backward(ctx, dz):
        q, r, z = ctx.saved_tensors
        fg = ctx.fg
        # r = deg^-1/2 leads to theta only (deg is a function of the sample). A create_graph backward is the inner optimiser's
        # gradient w.r.t. the GCN weights, which sends nothing to theta (see _FactoredMatmul): skip dr there; the nodes recorded
        # for dq still carry the dependence on r that the hypergradient pass differentiates.
        need_q, need_r = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and not torch.is_grad_enabled()
        if not torch.is_grad_enabled():
            fg.sink.add_outer(dz.detach(), q.detach(), r.detach())
        dq = _FactoredPropagate.apply(dz, r, fg.link, fg, ctx.own_r) if (need_q or need_r) else None
        dr = None
        if need_r:
            if dz.is_cuda and dz.dtype == torch.float32 and dz.dim() == 2:
                from .. import kernels                       # one launch instead of mul, addcmul, sum, div (need_r only holds in the
                if ctx.own_r:                                # plain hypergradient pass: nothing differentiates this)
                    # r IS this graph's deg^-1/2: it reaches theta only through deg, a row sum of the sample, so the chain
                    # dr -> d deg = -r^3 dr / 2 -> row-constant deposit is closed here: S / (-2 deg) with S = sum dZ Z + dQ Q.
                    # (Saves the rsqrt backward and the gradient accumulation on r and deg: ~5 launches per use.)
                    fg.sink.add_row_constant(kernels.row_dot2(dz.contiguous(), z, dq.contiguous(), q, fg.neg_two_deg()))
                else:
                    dr = kernels.row_dot2(dz.contiguous(), z, dq.contiguous(), q, r)
            else:
                dr = torch.addcmul(dz * z, dq, q).sum(dim=1) / r
        return (dq if need_q else None), dr, None, None, None


class _FactoredDegree(torch.autograd.Function):
    """deg_i = sum_j A_tilde_ij (K1's row sums). d deg_i / d sample_ij = 1 for j != i: a row-constant deposit."""

    @staticmethod
    def forward(ctx, link, fg):
        ctx.fg = fg
        return fg.handle.deg.clone()

    @staticmethod
    def backward(ctx, ddeg):
        if not torch.is_grad_enabled():
            ctx.fg.sink.add_row_constant(ddeg.detach())
        return None, None


class FactoredGraph:
    """A sampled graph for the unrolled bilevel loop that never becomes an N x N autograd tensor: bf16 A_tilde (self
    loops included) + degrees from K1, a `link` tensor (the model's `probs`) that ties it into autograd, and the
    FactorSink its straight-through gradient is deposited in. Quacks like a square matrix for the reference's asserts."""

    def __init__(self, handle: SampleHandle, link: Tensor, sink: FactorSink):
        self.handle, self.link, self.sink = handle, link, sink

    def dim(self):
        return 2

    def size(self, dim=None):
        shape = torch.Size((self.handle.n, self.handle.n))
        return shape if dim is None else shape[dim]

    @property
    def shape(self):
        return self.size()

    @property
    def device(self):
        return self.handle.adj.device

    def matmul(self, q: Tensor) -> Tensor:
        return _FactoredMatmul.apply(q, self.link, self)

    def degree(self) -> Tensor:
        return _FactoredDegree.apply(self.link, self)

    def normalized(self):
        """D^-1/2 (A+I) D^-1/2 (src/utils/graph.py:136-153) in factored, differentiable form."""
        return FactoredNormalizedAdjacency(self, torch.rsqrt(self.degree()), own_r=True)

    def neg_two_deg(self) -> Tensor:
        if getattr(self, "_neg2deg", None) is None:
            self._neg2deg = self.handle.deg * -2.0
        return self._neg2deg


class FactoredNormalizedAdjacency:
    def __init__(self, fg: FactoredGraph, r: Tensor, own_r: bool = False):
        self.fg, self.r, self.own_r = fg, r, own_r            # own_r: r is fg's deg^-1/2 (lets the backward close the chain to deg)

    def propagate(self, embeddings: Tensor) -> Tensor:
        return _FactoredPropagate.apply(embeddings, self.r, self.fg.link, self.fg, self.own_r)


def sample_factored(theta_full: Tensor, n: int, link: Tensor, sink: FactorSink) -> FactoredGraph:
    """The LDS sampling path (same Philox draws, same step counter as `sample_graph`) without the dense fp32 sample."""
    from .. import kernels
    seed, step = PHILOX.next_step()
    adj, _, deg, rs = kernels.k1_sample_normalize(theta_full, n, seed, step, want_sample=False, step_base=PHILOX.capture_base)
    return FactoredGraph(SampleHandle(n, adj, deg, rs, seed, step), link, sink)


def straight_through_estimator(sample: Tensor, parameters: Tensor) -> Tensor:
    """(sample - parameters).detach() + parameters (src/models/sampling.py:82-85)."""
    assert sample.size() == parameters.size()
    return (sample - parameters).detach() + parameters


def sample_graph(edge_probs: Tensor,
                 undirected: bool,
                 embeddings: Optional[Tensor] = None,
                 dense: bool = False,
                 k: Optional[int] = None,
                 sparsification: SPARSIFICATION = SPARSIFICATION.NONE,
                 force_straight_through_estimator: bool = False,
                 eps: Optional[float] = None,
                 knn_metric: str = "cosine",
                 uniforms: Optional[Tensor] = None) -> Tensor:
    """Same signature as the reference plus `uniforms` (explicit draws, parity tests)."""
    assert is_square_matrix(edge_probs)
    assert embeddings is None or edge_probs.size(0) == embeddings.size(0)
    if dense or sparsification != SPARSIFICATION.NONE or not undirected:
        raise NotImplementedError(
            "lds_gnn_b200 implements the LDS sampling path (undirected=True, sparsification NONE, dense=False); "
            "KNN/EPS sparsification, dense=True and directed sampling belong to the GAE/GRCN variants")
    if not edge_probs.is_cuda:
        raise RuntimeError("sample_graph runs on the B200 CUDA path only (edge_probs must be a CUDA tensor); there is no CPU fallback")
    kernel_input = edge_probs.detach()
    if not getattr(edge_probs, "_lds_symmetric", False):
        kernel_input = to_undirected(kernel_input, from_triu_only=True)   # only the upper triangle defines the draw
    seed, step = PHILOX.next_step()
    graph = _SampleSTE.apply(edge_probs, kernel_input, seed, step, uniforms)
    graph._lds_handle = _LAST_HANDLE.pop("handle")
    return graph


class Sampler:
    _ingredient = Ingredient("sampler")
    INGREDIENTS = {"sampler": _ingredient}

    @staticmethod
    @_ingredient.config
    def config():
        undirected: bool = True          # noqa: F841
        k: int = 20                      # noqa: F841
        eps: float = 0.9                 # noqa: F841
        sparsification: str = "NONE"     # noqa: F841
        dense: bool = False              # noqa: F841
        knn_metric: str = "cosine"       # noqa: F841

    @staticmethod
    @_ingredient.capture
    def sample(edge_probs: Tensor,
               undirected: bool,
               sparsification: str,
               k: int,
               eps: float,
               embeddings: Tensor = None,
               dense: bool = False,
               knn_metric: str = "cosine") -> Tensor:
        """Square matrix of Bernoulli parameters -> sampled adjacency with a straight-through gradient
        (src/models/sampling.py:104-138)."""
        assert sparsification in SPARSIFICATION.__members__
        return sample_graph(edge_probs=edge_probs, embeddings=embeddings, undirected=undirected,
                            sparsification=SPARSIFICATION[sparsification], dense=dense, k=k, eps=eps,
                            knn_metric=knn_metric)
