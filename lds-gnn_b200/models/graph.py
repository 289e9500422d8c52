"""Graph generative models with the reference's API (src/models/graph.py:16-78): `GraphGenerativeModel`,
`ParameterClamper`, `BernoulliGraphModel` (LDS: one Bernoulli parameter per potential edge).

State compatibility: `probs` is an `nn.Parameter` with the reference's shape — the (T,) row-major upper
triangle incl. diagonal for the undirected model, (N, N) for `directed=True` — and `state_dict()` has the
single key `probs`, so the in-memory checkpoints of the bilevel loop (src/trainers/bilevel.py:96-98,131-132)
round-trip. The kernels work on a full symmetric N x ld fp32 copy (`theta_full()`): coalesced rows for both
triangles and a natural row-block sharding. The two views are kept consistent lazily:
  * anything that touches `probs` through the Module API (attribute access, parameters(), state_dict(),
    optimizers holding the Parameter) sees fresh values — a fused step only marks `probs` stale and the
    conversion kernel runs on the next access;
  * `theta_full()` re-expands `probs` when the Parameter was modified since the last sync (tensor version
    counter / storage pointer) OR was handed out since (`model.probs`, `parameters()`, `state_dict()`): writes
    through `.data` (`model.probs.data.fill_(..)`, the reference's `ParameterClamper`) do not bump the version
    counter, so a hand-out is treated as a possible write. Code that keeps a long-lived reference to the
    Parameter and writes through `.data` later must call `model.invalidate()`.
"""
from abc import ABC, abstractmethod
from typing import Dict

import torch
from torch import Tensor, nn
from torch.nn import Parameter

from ..utils.graph import get_triu_values, is_square_matrix, triu_values_to_symmetric_matrix
from ..utils.tracking import setup_basic_logger
from .sampling import FactoredGraph, FactorSink, Sampler, sample_factored

logger = setup_basic_logger()


class ParameterClamper(object):
    """Clamp every parameter of a module to [0, 1] in place (src/models/graph.py:16-20)."""

    def __call__(self, module):
        for param in module.parameters():
            param.data.clamp_(0.0, 1.0)


class GraphGenerativeModel(nn.Module, ABC):

    def __init__(self, sample_undirected: bool = True, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.sample_undirected = sample_undirected

    def sample(self, *args, **kwargs) -> Tensor:
        return Sampler.sample(self.forward())

    def project_parameters(self):
        pass

    def refine(self):
        logger.warning("Model called to refine current parameters but method is not implemented. Ignore...")

    @abstractmethod
    def statistics(self) -> Dict[str, float]:
        pass


class BernoulliGraphModel(GraphGenerativeModel):

    def __init__(self, init_matrix: Tensor, directed: bool = False):
        """init_matrix: square matrix of initial edge probabilities (the adjacency for LDS, factory.py:60-62)."""
        super().__init__()
        assert is_square_matrix(init_matrix)
        self.directed = directed
        self.orig_matrix = init_matrix
        self._n = init_matrix.size(0)
        self._full = None            # [n, ld] fp32 device copy (unclamped, symmetric)
        self._full_key = None        # (data_ptr, version) of `probs` when `_full` was built from it
        self._probs_stale = False    # `_full` is newer than `probs` (after fused steps)
        self._handed_out = False     # `probs` left the module since the last expansion: `.data` writes are invisible to `_version`
        self.factor_sink = FactorSink()   # hypergradient deposits of factored graphs (unrolled bilevel loop)
        self.row_block = None        # (row0, rows) when the model owns a row block of theta (`from_row_block`)
        values = init_matrix if directed else get_triu_values(init_matrix)
        self.probs = Parameter(values, requires_grad=True)

    @classmethod
    def from_row_block(cls, theta_rows: Tensor, n: int, row0: int = 0) -> "BernoulliGraphModel":
        """Large graphs (BASELINE.json configs 4 and 5): the model owns rows [row0, row0 + rows) of the symmetric N x N
        probability matrix as ONE device tensor [rows, ld] (ld = round_up(n, 64), padding columns zero) and never builds the
        (T,) upper-triangle vector, an N x N host tensor or the 2 x T index of the reference's forward
        (src/utils/graph.py:166-181) — at N = 65 536 those are 8.6 GB, 17 GB and 34 GB. `probs` IS that tensor (a
        deliberate deviation from the reference's (T,) shape, only for this constructor); with rows < n the outer trainer
        steps it as one rank of the row-block sharded step, the other ranks own the other blocks and stay
        consistent without exchanging theta: the update is symmetric bit for bit."""
        from .. import kernels
        if not theta_rows.is_cuda or theta_rows.dtype != torch.float32 or theta_rows.dim() != 2:
            raise TypeError("from_row_block: theta_rows must be a CUDA float32 matrix [rows, ld]")
        rows, ld = theta_rows.shape
        if ld != kernels.padded_ld(n) or row0 < 0 or row0 + rows > n or row0 % 128:
            raise ValueError(f"from_row_block: need ld == {kernels.padded_ld(n)}, 0 <= row0, row0 + rows <= n and row0 a multiple of 128")
        self = cls.__new__(cls)
        GraphGenerativeModel.__init__(self)
        self.directed = False
        self.orig_matrix = None
        self._n = int(n)
        self._full, self._full_key, self._probs_stale, self._handed_out = None, None, False, False
        self.factor_sink = FactorSink()
        self.row_block = (int(row0), int(rows))
        self.probs = Parameter(theta_rows, requires_grad=True)
        return self

    # ---- lazy consistency between the (T,) Parameter and the full device matrix ------------------
    def _probs_param(self) -> Parameter:
        return self._parameters["probs"]

    def _sync_probs(self):
        if self._probs_stale and self.row_block is None:
            from .. import kernels
            p = self._probs_param()
            with torch.no_grad():
                p.data.copy_(kernels.theta_full_to_triu(self._full, self._n))
            self._probs_stale = False
            self._full_key = (p.data_ptr(), p._version)

    def __getattr__(self, name):
        if name == "probs" and "_parameters" in self.__dict__:
            self._sync_probs()
            self.__dict__["_handed_out"] = True
        return super().__getattr__(name)

    def named_parameters(self, *args, **kwargs):
        self._sync_probs()
        self._handed_out = True
        return super().named_parameters(*args, **kwargs)

    def state_dict(self, *args, **kwargs):
        self._sync_probs()
        self._handed_out = True
        return super().state_dict(*args, **kwargs)

    def invalidate(self):
        """`probs` was written behind the module's back (through `.data` on a reference obtained earlier): the next
        `theta_full()` re-expands it. A device copy that is NEWER than `probs` (after fused steps) is flushed first."""
        self._sync_probs()
        self._full_key = None

    def load_state_dict(self, state_dict, *args, **kwargs):
        src = state_dict.get("probs") if hasattr(state_dict, "get") else None
        if torch.is_tensor(src) and src.data_ptr() == self._probs_param().data_ptr():
            # the dict aliases the live Parameter — the reference's in-memory "checkpoint" (src/trainers/bilevel.py:96-98 keeps
            # `state_dict()` without a copy): loading it is a no-op there, so the newest values (the device copy) must survive
            self._sync_probs()
        else:
            self._probs_stale = False    # the loaded values win over a newer device copy
        return super().load_state_dict(state_dict, *args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._sync_probs()
        self._full, self._full_key = None, None
        return super()._apply(fn, *args, **kwargs)

    def theta_full(self) -> Tensor:
        """Full symmetric [n, ld] fp32 device matrix of the current probabilities (undirected model)."""
        if self.directed:
            raise NotImplementedError("theta_full() is defined for the undirected LDS model")
        p = self._probs_param()
        if self.row_block is not None:
            return p.data                                   # the parameter IS the device matrix (rows of this block)
        if not p.is_cuda:
            raise RuntimeError("BernoulliGraphModel: the B200 path needs the model on a CUDA device (model.to('cuda')); no CPU fallback")
        key = (p.data_ptr(), p._version)
        if self._full is None or (not self._probs_stale and (key != self._full_key or self._handed_out)):
            from .. import kernels
            self._full = kernels.theta_triu_to_full(p.detach(), clamp=False, out=self._full)
            self._full_key = key
            self._handed_out = False
        return self._full

    def mark_full_updated(self):
        """Called by the fused outer step after it updated `theta_full()` in place."""
        self._probs_stale = self.row_block is None

    def sample_factored(self) -> FactoredGraph:
        """`sample()` for callers that understand factored graphs (the trainers of this package): K1 straight from the
        resident full matrix — no (T,) -> N x N rebuild, no dense fp32 sample, no N x N autograd node. Same Philox draw
        as `sample()` would make at this step."""
        return sample_factored(self.theta_full(), self._n, self._probs_param(), self.factor_sink)

    # ---- reference API ---------------------------------------------------------------------------
    def project_parameters(self):
        """clamp(probs, 0, 1) in place (src/models/graph.py:16-20, 63-64). When the device matrix is the newer copy (after
        fused steps) it is the one that gets clamped — `probs` receives the result at its next access."""
        if self._probs_stale and self._full is not None and not self.directed:
            from .. import kernels
            kernels.theta_clamp_(self._full, self._n)
            return
        self.apply(ParameterClamper())
        self._full_key = None

    def forward(self, *args, **kwargs) -> Tensor:
        if self.row_block is not None:
            if self.row_block[1] != self._n:
                raise NotImplementedError("forward() of a row-block model needs every row: only the sharded outer step is defined for it")
            return self.probs[:, :self._n].clamp(0.0, 1.0)
        return self.probs if self.directed else triu_values_to_symmetric_matrix(self.probs)  # type: ignore

    def statistics(self) -> Dict[str, float]:
        """Same keys and definitions as the reference (src/models/graph.py:69-78); one reduction kernel and
        a single device->host copy instead of a rebuild of the N x N matrix plus five `.item()` syncs."""
        if self.row_block is not None:
            return self._row_block_statistics()
        if self.directed or not self._probs_param().is_cuda:
            sample = self.forward()
            probs = self.probs
            total = sample.sum().item()
            return {"expected_num_edges": total, "percentage_edges_expected": total / sample.size(0) ** 2,
                    "mean_prob": torch.mean(probs).item(), "min_prob": torch.min(probs).item(),
                    "max_prob": torch.max(probs).item()}
        from .. import kernels
        s = kernels.theta_stats(self.theta_full(), self._n).tolist()
        t = self._n * (self._n + 1) // 2
        return {"expected_num_edges": s[0], "percentage_edges_expected": s[0] / (self._n ** 2),
                "mean_prob": s[1] / t, "min_prob": s[2], "max_prob": s[3]}

    def _row_block_statistics(self) -> Dict[str, float]:
        """statistics() of a row-block model: local sums over this block (upper-triangle terms: columns >= global row),
        all-reduced over the ranks when torch.distributed is initialised. One device->host transfer."""
        import torch.distributed as dist
        row0, rows = self.row_block
        n = self._n
        th = self.probs.detach()[:, :n]
        gi = torch.arange(row0, row0 + rows, device=th.device)[:, None]
        upper = torch.arange(n, device=th.device)[None, :] >= gi
        big = torch.finfo(torch.float64).max
        vals = th.double()
        acc = torch.stack([vals.clamp(0.0, 1.0).sum(), (vals * upper).sum(),
                           torch.where(upper, vals, torch.full_like(vals, big)).min(), torch.where(upper, vals, torch.full_like(vals, -big)).max()])
        if dist.is_available() and dist.is_initialized() and rows < n:
            sums, mn, mx = acc[:2].clone(), acc[2:3].clone(), acc[3:4].clone()
            dist.all_reduce(sums); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            acc = torch.cat([sums, mn, mx])
        s = acc.tolist()
        t = n * (n + 1) // 2
        return {"expected_num_edges": s[0], "percentage_edges_expected": s[0] / (n ** 2), "mean_prob": s[1] / t, "min_prob": s[2], "max_prob": s[3]}
