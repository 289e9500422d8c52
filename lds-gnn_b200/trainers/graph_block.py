"""A bilevel block — tau `inner_opt_step`s and the `hyper_opt_step` that differentiates through them
(src/trainers/bilevel.py:53-73) — captured ONCE into a CUDA graph and replayed.

Why: on the factored route the block is ~1 ms of GPU work (K1 x (tau+1), ~44 K2 products, one K3+K4 pass and a few hundred
O(N h) elementwise kernels) but ~15 ms of host work (Python autograd over the unrolled double backward). Everything that
varies between blocks is therefore read from device memory, so one captured graph serves every block:
  * the Philox step of each sampled graph  -> lds_k1_sample_normalize_dstep reads a device counter (+ its offset in the block)
  * torch's dropout masks                  -> torch's own graph-safe generator offsets
  * the inner Adam's step count            -> `DifferentiableAdam.device_step` (bias correction gathered from a device table)
  * the outer learning rate (StepLR decay) -> a device scalar folded into the gradient factors (the update is linear in them)
  * fast weights / Adam moments            -> static buffers, written back by the block's last nodes
and everything the host needs afterwards comes back in one transfer: the tau+1 (loss, acc) pairs. The weights after every
inner step are snapshotted inside the graph, and theta is backed up at block entry, so the caller can apply the reference's
per-step early stopping AFTER the replay and roll back exactly to the step at which the reference would have stopped.
"""
from collections import OrderedDict
from typing import List

import torch

from .. import _lib
from ..models.sampling import PHILOX
from . import Metrics

NAMES = ("layer_in.fc.weight", "layer_in.fc.bias", "layer_out.fc.weight", "layer_out.fc.bias")


class BlockCaptureError(RuntimeError):
    """The block could not be captured; the trainers' state is back to what it was at block entry."""


class CapturedBilevelBlock:

    def __init__(self, runner, tau: int):
        self.runner, self.tau = runner, int(tau)
        self.inner, self.outer = runner.inner_trainer, runner.outer_trainer
        gcn = self.inner.model
        dev = self.inner.data.x.device
        self.shapes = [tuple(p.shape) for p in (gcn.layer_in.fc.weight, gcn.layer_in.fc.bias, gcn.layer_out.fc.weight, gcn.layer_out.fc.bias)]
        self.offsets = [0]
        for shape in self.shapes:
            count = 1
            for s in shape:
                count *= s
            self.offsets.append(self.offsets[-1] + count)
        total = self.offsets[-1]
        f32 = dict(dtype=torch.float32, device=dev)
        self.w = torch.zeros(total, **f32)                          # fast weights at block entry / exit
        self.m = torch.zeros(total, **f32)                          # Adam moments at block entry / exit
        self.v = torch.zeros(total, **f32)
        self.snap = torch.zeros((self.tau, total), **f32)           # fast weights after inner step k
        self.t_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.zeros(1, **f32)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.metrics = torch.zeros((self.tau + 1, 2), **f32)
        self.theta_backup = None
        self.graph = None
        self.draws = self.tau + 1
        self.resident = False                                       # True: the static buffers hold the current inner state
        self._lr_saved = None
        self._table = None
        self._keepalive = None
        self._theta_ptr = None

    # ------------------------------------------------------------------------------------------ eligibility
    @staticmethod
    def eligible(runner) -> bool:
        inner, outer = runner.inner_trainer, runner.outer_trainer
        if not getattr(outer, "unroll_plan", None) or not outer.unroll_plan(inner):
            return False
        kind = outer._optimizer_kind()
        if kind is None or kind[0] != _lib.OPT_SGD:
            return False
        names = list(inner.model_params.keys())
        return names == list(NAMES) and all(p.is_cuda and p.dtype == torch.float32 for p in inner.model_params.values())

    # ------------------------------------------------------------------------------------------ state hand-over
    def _is_resident(self) -> bool:
        """True when the trainers' Python-side state IS the static buffers (what `replay` leaves behind). Anything else — fresh
        weights / optimiser after `reset_weights` / `reset_optimizer`, eager steps in between, `store_state` — is detected by
        identity, so a stale buffer can never be replayed."""
        params, state = self.inner.model_params, self.inner.optimizer.state
        first = params.get(NAMES[0]) if hasattr(params, "get") else None
        return (self.resident and first is not None and first.data_ptr() == self.w.data_ptr() and first.grad_fn is None
                and state["exp_avg"] is self.m and state["exp_avg_sq"] is self.v)

    def _publish_state(self):
        """The trainers' Python-side state := the static buffers (views, no history): what a replay leaves behind, and what the
        capture must leave behind too (its own bodies left references to tensors of the graph's pool)."""
        inner = self.inner
        inner.model_params = OrderedDict((name, self.w[self.offsets[i]:self.offsets[i + 1]].view(self.shapes[i]).requires_grad_(True))
                                         for i, name in enumerate(NAMES))          # leaves: an eager step may follow
        inner.optimizer.state["exp_avg"], inner.optimizer.state["exp_avg_sq"] = self.m, self.v
        inner.optimizer._flat = None
        self.resident = True

    def load_state(self):
        """Eager inner state (fast weights, Adam moments, step count) -> static buffers."""
        inner = self.inner
        with torch.no_grad():
            self.w.copy_(torch.cat([inner.model_params[k].detach().reshape(-1) for k in NAMES]))
            st = inner.optimizer.state
            if st["exp_avg"] is None:
                self.m.zero_()
                self.v.zero_()
            else:
                self.m.copy_(st["exp_avg"].detach())
                self.v.copy_(st["exp_avg_sq"].detach())
        self.resident = True

    def _params_from(self, flat: torch.Tensor) -> OrderedDict:
        return OrderedDict((name, flat[self.offsets[i]:self.offsets[i + 1]].view(self.shapes[i]).clone().requires_grad_(True))
                           for i, name in enumerate(NAMES))

    def params_after(self, k: int) -> OrderedDict:
        """Fast weights after inner step k of the last replay (fresh leaves, like `copy_model_params`)."""
        return self._params_from(self.snap[k])

    def store_state(self, steps_done: int):
        """Static buffers -> eager inner state, as if the block had stopped after `steps_done` inner steps. Only the weights are
        exact for steps_done < tau (the reference abandons the optimiser state at that point: it stops and resets)."""
        inner = self.inner
        src = self.w if steps_done >= self.tau else self.snap[steps_done - 1]
        inner.model_params = self._params_from(src)
        opt = inner.optimizer
        opt.state["exp_avg"], opt.state["exp_avg_sq"] = self.m.clone(), self.v.clone()
        opt._flat = None
        self.resident = False

    # ------------------------------------------------------------------------------------------ capture
    def _body(self):
        """The block, written with the trainers' own methods (what gets captured is exactly what the eager loop runs)."""
        inner, outer, runner = self.inner, self.outer, self.runner
        theta = outer.model.theta_full()
        self.theta_backup.copy_(theta)
        leaf = self.w.detach().requires_grad_(True)
        inner.model_params = OrderedDict((name, leaf[self.offsets[i]:self.offsets[i + 1]].view(self.shapes[i])) for i, name in enumerate(NAMES))
        opt = inner.optimizer
        opt.state["exp_avg"], opt.state["exp_avg_sq"] = self.m, self.v
        opt._flat = None
        opt.device_step, opt.device_offset, opt._block_steps = self.t_dev, 0, None
        opt._correction_table = self._table                        # the runner replaces the optimiser object; the table is ours
        results = []
        inner.deferred = results
        outer.deferred = (results, self.lr_dev)
        try:
            for k in range(self.tau):
                runner.inner_opt_step()
                with torch.no_grad():
                    self.snap[k].copy_(opt._flat[0].detach())
            outer.train_step(inner.model_forward, retain_graph=False)
            with torch.no_grad():
                self.w.copy_(opt._flat[0].detach())
                self.m.copy_(opt.state["exp_avg"].detach())
                self.v.copy_(opt.state["exp_avg_sq"].detach())
                self.metrics.copy_(torch.stack(results))
        finally:
            inner.deferred = None
            outer.deferred = None
            opt.device_step, opt.device_offset, opt._block_steps = None, 0, None

    def capture(self):
        inner, outer = self.inner, self.outer
        model = outer.model
        theta = model.theta_full()                                  # in sync BEFORE capture: no layout kernel may end up in the graph
        if self.theta_backup is None or self.theta_backup.shape != theta.shape:
            self.theta_backup = torch.empty_like(theta)
        host = (PHILOX.step, inner.optimizer.state["step"])
        saved = (self.w.clone(), self.m.clone(), self.v.clone(), theta.clone())
        # one eager pass on a side stream first (library handles, workspaces, index caches), then the capture itself; both leave
        # the Python-side counters advanced and (the eager one) the state changed: restore everything afterwards
        self._set_dynamic(host[0], host[1], outer.get_learning_rates()[0])
        self._table = inner.optimizer.correction_table(self.w)      # built (host -> device copy) BEFORE the capture
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                PHILOX.begin_capture(self.step_dev)
                try:
                    self._body()
                finally:
                    self.draws = PHILOX.end_capture()
            torch.cuda.current_stream().wait_stream(side)
            self._restore(saved, theta)
            # capture_begin / capture_end directly: the `torch.cuda.graph` context manager first runs gc.collect() and
            # torch.cuda.empty_cache(), which costs seconds when a previous phase left tens of GB cached, and this graph's pool
            # is a fraction of a GB
            graph = torch.cuda.CUDAGraph()
            capture_stream = torch.cuda.Stream()
            capture_stream.wait_stream(torch.cuda.current_stream())
            PHILOX.begin_capture(self.step_dev)
            try:
                with torch.cuda.stream(capture_stream):
                    graph.capture_begin()
                    try:
                        self._body()
                    except BaseException:
                        try:
                            graph.capture_end()
                        except Exception:
                            pass
                        raise
                    graph.capture_end()
            finally:
                PHILOX.end_capture()
            torch.cuda.current_stream().wait_stream(capture_stream)
        except Exception as exc:
            torch.cuda.synchronize()
            self._restore(saved, theta)
            PHILOX.step = host[0]
            inner.optimizer.state["step"] = host[1]
            self.store_state(self.tau)                             # eager state := the (restored) state at block entry
            raise BlockCaptureError(f"{type(exc).__name__}: {exc}") from exc
        self._restore(saved, theta)
        PHILOX.step = host[0]
        inner.optimizer.state["step"] = host[1]
        model.zero_grad(set_to_none=True)
        self.graph = graph
        self._publish_state()
        # The graph holds raw addresses. Everything it reads that was allocated OUTSIDE its private pool must outlive it: the
        # optimiser whose per-element hyper-parameter vectors were baked in (the runner replaces the optimiser object at every
        # outer iteration), the kernels' cached workspaces (a later, larger request would replace and free them), theta.
        from .. import kernels
        self._keepalive = (inner.optimizer, dict(kernels._ws_cache), theta, inner.data, outer.dataset, outer.opt_mask,
                           dict(inner._rows_cache), outer._opt_rows)
        self._theta_ptr = theta.data_ptr()
        self._captured_signature = self._signature()
        if self.draws != self.tau + 1:                              # undo_hyper_step rewinds the Philox counter on this assumption
            raise BlockCaptureError(f"a block of tau = {self.tau} drew {self.draws} graphs (expected tau + 1)")

    def _signature(self):
        """Everything a captured block bakes in besides theta's address: the masks' row lists, the features, the inner
        optimiser's hyper-parameters (per-element vectors of the capturing optimiser), dropout and the GCN shapes. The eager
        loop would pick a change of any of these up at the next step; a replay must too, so a mismatch recaptures."""
        inner, outer = self.inner, self.outer
        data = inner.data
        groups = tuple((float(g["lr"]), float(g["weight_decay"]), tuple(g["betas"]), float(g["eps"])) for g in inner.optimizer.param_groups)
        return (data.x.data_ptr(), data.y.data_ptr(), data.train_mask.data_ptr(), data.train_mask._version,
                outer.opt_mask.data_ptr(), outer.opt_mask._version, groups, float(inner.model.dropout),
                tuple(tuple(p.shape) for p in inner.model_params.values()))

    def _restore(self, saved, theta):
        with torch.no_grad():
            self.w.copy_(saved[0])
            self.m.copy_(saved[1])
            self.v.copy_(saved[2])
            theta.copy_(saved[3])

    def _set_dynamic(self, philox_step: int, adam_steps: int, lr: float):
        self.step_dev.fill_(int(philox_step))
        self.t_dev.fill_(int(adam_steps))
        self.lr_dev.fill_(float(lr))

    # ------------------------------------------------------------------------------------------ replay
    def replay(self) -> List[Metrics]:
        """Run one block. Returns tau inner Metrics followed by the hyper step's. Host-side bookkeeping (Philox step, Adam step
        count, StepLR, the lazily synchronised `probs`) is advanced as the eager loop would have."""
        inner, outer = self.inner, self.outer
        if self.graph is None:
            if not self._is_resident():
                self.load_state()
            self.capture()
        if outer.model.theta_full().data_ptr() != self._theta_ptr or self._signature() != self._captured_signature:
            self.graph = None                                       # the model was moved / rebuilt, or something the graph baked in changed
            return self.replay()
        if not self._is_resident():
            self.load_state()
        outer.model.train()
        inner.model.train(True)
        self._set_dynamic(PHILOX.step, inner.optimizer.state["step"], outer.get_learning_rates()[0])
        self.graph.replay()
        PHILOX.step += self.draws
        inner.optimizer.state["step"] += self.tau
        outer.model.mark_full_updated()
        outer.last_route = "factored-graph"
        if outer.lr_decayer is not None:
            sched = outer.lr_decayer
            self._lr_saved = (sched.last_epoch, [g["lr"] for g in outer.optimizer.param_groups], list(sched._last_lr), sched._step_count)
            outer.optimizer._opt_called = True
            sched.step()
        self._publish_state()
        rows = self.metrics.tolist()                                # the block's only device->host transfer
        return [Metrics(loss=r[0], acc=r[1]) for r in rows]

    def undo_hyper_step(self, steps_done: int):
        """The reference stopped after `steps_done` < tau inner steps of this block: its hyper step never happened. Restore theta,
        the learning-rate schedule and the counters to that point. (torch's own dropout generator is NOT rewound: with dropout
        > 0 the masks drawn after a rollback differ from the step-by-step loop's, like after any extra draw; the graphs do not.)"""
        outer = self.outer
        with torch.no_grad():
            outer.model.theta_full().copy_(self.theta_backup)
        outer.model.mark_full_updated()
        PHILOX.step -= self.draws - steps_done
        self.inner.optimizer.state["step"] -= self.tau - steps_done
        if outer.lr_decayer is not None and self._lr_saved is not None:
            sched = outer.lr_decayer
            sched.last_epoch, lrs, sched._last_lr, sched._step_count = self._lr_saved
            for group, lr in zip(outer.optimizer.param_groups, lrs):
                group["lr"] = lr
            self._lr_saved = None
