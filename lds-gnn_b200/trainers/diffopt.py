"""Differentiable Adam for the unrolled inner problem — a from-scratch stand-in for
`higher.optim.DifferentiableAdam` (reference call sites: src/trainers/inner.py:6, 42-50, 71).

`higher` (git master, unpinned, scripts/install.sh:3-4) is not vendored with the reference and cannot be
installed offline, so NUMERIC PARITY OF THIS UPDATE IS UNPINNED: this file follows the
published algorithm of higher's master — `torch.optim.Adam`'s rule applied out of place so later losses can
differentiate through it: g += wd * p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
p' = p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)  (bias correction on sqrt(v) BEFORE eps is added, exactly
torch.optim.Adam's current form). It is checked against `torch.optim.Adam` on detached copies and against the
live reference's inner steps run over the `higher` stand-in (tests/golden/blk_*.npz).

All parameters are updated as ONE flat vector (per-group hyper-parameters become per-element vectors): the
update is ~10 differentiable ops per step instead of ~14 per parameter tensor, which matters because the
unrolled loop is bound by the host's op-dispatch rate, not by the GPU. The new parameters are views of the
flat result, so the next step's concatenation is free.
"""
import math
from typing import Iterable, List

import numpy as np
import torch


class _FusedAdamStep(torch.autograd.Function):
    """(p, m, v, g) -> (p', m', v') as one launch (csrc/lds_adam.cu) instead of ~12 elementwise kernels, and its backward as one
    instead of ~22. The map is elementwise, so the hyper step needs only its first-order vector-Jacobian product here: the
    second-order terms of the hypergradient come from g's own graph (autograd.grad(..., create_graph=True))."""

    @staticmethod
    def forward(ctx, p, m, v, g, wd, b1, b2, eps, step_size, root_scale):
        from .. import kernels
        p, m, v, g = (t.contiguous() for t in (p, m, v, g))
        ctx.set_materialize_grads(False)                       # an unused output (the last step's state) arrives as None, not as zeros
        ctx.save_for_backward(p, m, v, g)
        ctx.hyper = (wd, b1, b2, eps, step_size, root_scale)
        new_p, new_m, new_v = kernels.adam_step(p, m, v, g, wd, b1, b2, eps, step_size, root_scale)
        return new_p, new_m, new_v

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gp, gm, gv):
        from .. import kernels
        p, m, v, g = ctx.saved_tensors
        as_c = lambda t: None if t is None else t.contiguous()
        dp, dm, dv, dg = kernels.adam_step_backward(as_c(gp), as_c(gm), as_c(gv), p, m, v, g, *ctx.hyper, need=ctx.needs_input_grad[:4])
        return dp, dm, dv, dg, None, None, None, None, None, None


class _SplitFlat(torch.autograd.Function):
    """flat -> the parameter tensors as views. Autograd's own slice backward builds, for EVERY view, a zero vector of the full
    length, copies the slice in and adds it to the running sum (3 launches x 4 parameters per unrolled step); here the
    backward is one concatenation."""

    @staticmethod
    def forward(ctx, flat, offsets, shapes):
        ctx.set_materialize_grads(False)
        ctx.meta = (offsets, shapes)
        return tuple(flat[offsets[i]:offsets[i + 1]].view(shape) for i, shape in enumerate(shapes))

    @staticmethod
    def backward(ctx, *grads):
        offsets, shapes = ctx.meta
        like = next(g for g in grads if g is not None)
        parts = [g.reshape(-1) if g is not None else like.new_zeros(offsets[i + 1] - offsets[i]) for i, g in enumerate(grads)]
        return torch.cat(parts), None, None


FUSED_STEP = [True]            # CUDA fp32, one parameter group's hyper-parameters, every parameter has a gradient: one launch per step


class DifferentiableAdam:
    def __init__(self, optimizer: torch.optim.Adam, reference_params: Iterable[torch.Tensor], track_higher_grads: bool = True):
        reference = list(reference_params)
        position = {id(p): i for i, p in enumerate(reference)}
        self.track = track_higher_grads
        self.param_groups = []
        self._slots = []
        for group in optimizer.param_groups:
            self.param_groups.append({k: v for k, v in group.items() if k != "params"})
            self._slots.append([position[id(p)] for p in group["params"]])
        self._shapes = [p.shape for p in reference]
        self._numels = [p.numel() for p in reference]
        self.state = {"step": 0, "exp_avg": None, "exp_avg_sq": None}
        self._vectors = None          # per-element hyper-parameters, built on the parameters' device at the first step
        self._flat = None             # (flat tensor, views) of the last result: reused when the caller passes the views back
        # CUDA-graph capture of a bilevel block: int64 device tensor [1] = number of steps taken BEFORE the block. The bias
        # correction is then looked up on the device (a replayed graph must not bake the step count in): one gather from a
        # table of (1 / (1 - b1^t), 1 / sqrt(1 - b2^t)), t = 0 .. TABLE-1, computed once in fp64 (both are 1.0f well before the end).
        self.device_step = None
        self.device_offset = 0
        self._correction_table = None
        self._host_correction = None
        self._block_steps = None
        self._block_first = 1

    # ---- per-element hyper-parameter vectors -----------------------------------------------------------------
    def _hyper(self, like: torch.Tensor):
        if self._vectors is None or self._vectors["device"] != like.device or self._vectors["dtype"] != like.dtype:
            total = sum(self._numels)
            offsets = [0]
            for k in self._numels:
                offsets.append(offsets[-1] + k)
            cols = {name: torch.zeros(total, dtype=like.dtype) for name in ("lr", "wd", "b1", "b2", "eps", "in_group")}
            for group, slots in zip(self.param_groups, self._slots):
                for slot in slots:
                    sl = slice(offsets[slot], offsets[slot + 1])
                    cols["lr"][sl], cols["wd"][sl], cols["eps"][sl] = group["lr"], group["weight_decay"], group["eps"]
                    cols["b1"][sl], cols["b2"][sl] = group["betas"]
                    cols["in_group"][sl] = 1.0
            vec = {"device": like.device, "dtype": like.dtype, "offsets": offsets}
            for name, col in cols.items():
                lo, hi = col.min().item(), col.max().item()
                vec[name] = lo if lo == hi else col.to(like.device)            # uniform -> python scalar (no extra op)
            self._vectors = vec
        return self._vectors

    def step(self, loss: torch.Tensor, params: Iterable[torch.Tensor]) -> List[torch.Tensor]:
        params = list(params)
        grads = torch.autograd.grad(loss, params, create_graph=self.track, allow_unused=True)
        hp = self._hyper(params[0])
        if self._flat is not None and len(params) == len(self._flat[1]) and all(a is b for a, b in zip(params, self._flat[1])):
            p = self._flat[0]
        else:
            p = torch.cat([q.reshape(-1) for q in params])
        unused = [g is None for g in grads]
        g = torch.cat([torch.zeros_like(q).reshape(-1) if miss else gr.reshape(-1) for q, gr, miss in zip(params, grads, unused)])
        st = self.state
        if st["exp_avg"] is None:
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros_like(p)
        st["step"] += 1
        t = st["step"]
        fused = (FUSED_STEP[0] and p.is_cuda and p.dtype == torch.float32 and not any(unused) and isinstance(hp["in_group"], float)
                 and all(isinstance(hp[k], float) for k in ("b1", "b2", "eps", "lr")))       # weight decay may differ per group (a vector)
        g_raw = g
        # fused forms (add / addcmul / lerp): a third of the launches of the textbook expressions, same derivatives
        if fused:
            pass
        elif isinstance(hp["wd"], float):
            if hp["wd"] != 0.0:
                g = torch.add(g, p, alpha=hp["wd"])
        else:
            g = torch.addcmul(g, hp["wd"], p)
        b1, b2 = hp["b1"], hp["b2"]
        if fused:
            m = v = root = None
        elif isinstance(b1, float) and isinstance(b2, float):
            m = torch.lerp(st["exp_avg"], g, 1 - b1)
            v = torch.addcmul(st["exp_avg_sq"] * b2, g, g, value=1 - b2)
        else:
            m = st["exp_avg"] * b1 + (1 - b1) * g
            v = st["exp_avg_sq"] * b2 + (1 - b2) * (g * g)
        # sqrt has an infinite derivative at 0: floor exact zeros (higher masks that gradient instead)
        if not fused:
            root = v.clamp_min(1e-30).sqrt()
        if self.device_step is not None:
            self.device_offset += 1
            if not (isinstance(b1, float) and isinstance(b2, float)):
                raise NotImplementedError("device-side step count needs the same betas in every parameter group")
            if self.device_offset == 1 or self._block_steps is None:
                # step sizes of the next BLOCK_MAX steps in one gather (fp32 table entry x fp32(lr), rounded to fp32)
                idx = (self.device_step + torch.arange(1, self.BLOCK_MAX + 1, device=p.device)).clamp_max(self.TABLE - 1)
                rows = self.correction_table(p).index_select(0, idx)
                self._block_steps = (rows[:, 0] * hp["lr"], rows[:, 1])
                self._block_first = self.device_offset
            k = self.device_offset - self._block_first
            if k >= self.BLOCK_MAX:
                raise NotImplementedError(f"more than {self.BLOCK_MAX} unrolled steps per captured block")
            step_size, root_scale = self._block_steps[0][k:k + 1], self._block_steps[1][k:k + 1]
        elif isinstance(b1, float) and isinstance(b2, float):
            if isinstance(hp["lr"], float) and p.dtype == torch.float32:
                # the same fp32 arithmetic as the device-side path above: a step-by-step run and a replayed block agree bitwise
                row = self._host_table()[min(t, self.TABLE - 1)]
                step_size, root_scale = float(np.float32(row[0]) * np.float32(hp["lr"])), float(row[1])
            else:
                step_size, root_scale = hp["lr"] / (1 - b1 ** t), 1.0 / math.sqrt(1 - b2 ** t)
        else:
            step_size = hp["lr"] / (1 - torch.as_tensor(b1) ** t)
            root_scale = 1.0 / torch.sqrt(1 - torch.as_tensor(b2) ** t)
        # higher master / torch.optim.Adam: the bias correction scales sqrt(v) BEFORE eps is added
        if fused:
            new_p, m, v = _FusedAdamStep.apply(p, st["exp_avg"], st["exp_avg_sq"], g_raw, hp["wd"], b1, b2, hp["eps"], step_size, root_scale)
        else:
            new_p = p - step_size * (m / (root * root_scale + hp["eps"]))
        frozen = None
        if any(unused) or not isinstance(hp["in_group"], float):
            # parameters without a gradient (or outside every group) keep value and state, like torch.optim.Adam
            keep = torch.cat([torch.full((k,), 0.0 if miss else 1.0, dtype=p.dtype, device=p.device)
                              for k, miss in zip(self._numels, unused)])
            if not isinstance(hp["in_group"], float):
                keep = keep * hp["in_group"]
            frozen = keep == 0
            new_p = torch.where(frozen, p, new_p)
            m = torch.where(frozen, st["exp_avg"], m)
            v = torch.where(frozen, st["exp_avg_sq"], v)
        st["exp_avg"], st["exp_avg_sq"] = m, v
        offsets = hp["offsets"]
        if new_p.requires_grad:
            views = list(_SplitFlat.apply(new_p, offsets, self._shapes))
        else:
            views = [new_p[offsets[i]:offsets[i + 1]].view(shape) for i, shape in enumerate(self._shapes)]
        self._flat = (new_p, views)
        return views

    TABLE = 1 << 16
    BLOCK_MAX = 64

    def _host_table(self):
        """fp32 bias-correction factors [1 / (1 - b1^t), 1 / sqrt(1 - b2^t)] for t = 0 .. TABLE-1 (numpy; entry 0 unused)."""
        if self._host_correction is None:
            b1, b2 = self.param_groups[0]["betas"]
            t = np.maximum(np.arange(self.TABLE, dtype=np.float64), 1.0)
            self._host_correction = np.stack((1.0 / (1 - b1 ** t), 1.0 / np.sqrt(1 - b2 ** t)), axis=1).astype(np.float32)
        return self._host_correction

    def correction_table(self, like: torch.Tensor) -> torch.Tensor:
        """The same table on `like`'s device."""
        if self._correction_table is None or self._correction_table.device != like.device:
            self._correction_table = torch.from_numpy(self._host_table()).to(like.dtype).to(like.device)
        return self._correction_table

    def detach_(self):
        """Cut the optimiser state from the autograd graph (truncated back-propagation, inner.py:110-125)."""
        for key in ("exp_avg", "exp_avg_sq"):
            if self.state[key] is not None:
                self.state[key] = self.state[key].detach()
        self._flat = None
