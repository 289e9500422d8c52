"""Outer (graph-parameter) trainer with the reference's API (src/trainers/outer.py:19-161).

`train_step(gcn_predict_fct)` is the hot path. Two routes give the same result:

* FUSED — taken when the step is the plain LDS outer step: an undirected `BernoulliGraphModel` on a CUDA
  device, default sampler config, `regularize=False`, no `refine_embeddings`, an SGD (no momentum / weight
  decay) or Adam (no weight decay / amsgrad) optimiser over exactly `model.probs`, and `gcn_predict_fct`
  being `InnerProblemTrainer.model_forward` of this package (so the GCN weights, features and dropout are
  known) whose fast weights carry no unrolled history (or `trainer.first_order = True`). One C call (`lds_outer_step`) enqueues the whole chain — sample, normalise, GCN forward, masked
  NLL, closed-form hypergradient, optimiser step, clamp — and one small device->host copy returns
  `Metrics`. The optimiser's learning rate is read from `optimizer.param_groups` each step and the
  scheduler is stepped on the host exactly like the reference, so `get_learning_rates()` stays truthful.
  Differences a caller can observe: `model.probs.grad` is not materialised (the gradient is consumed inside
  the update kernel) and dropout masks come from the device Philox stream instead of torch's generator.

* FACTORED — the same plain LDS configuration when the fast weights DO carry unrolled history (every hyper step of
  the real bilevel loop, src/trainers/bilevel.py:53-73): the graphs of the inner steps were sampled as
  `FactoredGraph`s (`sample_for_unroll`), autograd runs over O(N h) tensors only — every product with a sampled
  graph is K2 on the tensor cores, differentiable to any order — and the straight-through hypergradient of ALL
  graphs of the unroll arrives as low-rank factor pairs that ONE K3+K4 pass folds into theta.

* COMPOSABLE — everything else (custom predict functions, regularisers, other optimisers): the reference's
  own sequence `zero_grad / sample / predict / nll / backward / step / decay / project`, with sampling and
  propagation running on the same kernels through autograd Functions.
"""
from typing import Callable, List

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.optim.lr_scheduler import StepLR
from torch.optim.optimizer import Optimizer

from .. import _lib
from ..config import Ingredient
from ..models.graph import BernoulliGraphModel, GraphGenerativeModel
from ..models.sampling import PHILOX, Sampler
from ..utils.evaluation import accuracy
from ..utils.rowops import MaskInfo, fused_nll_ok, masked_nll
from ..utils.graph import DenseData
from ..utils.tracking import get_lr, setup_basic_logger
from . import Metrics

logger = setup_basic_logger()


def graph_regularization(*args, **kwargs):
    raise NotImplementedError("graph regularisers (src/utils/graph.py:195-234) are off by default (outer.py:126) and "
                              "outside the LDS hot path; run with regularize=False")


class OuterProblemTrainer:

    def __init__(self,
                 optimizer: Optimizer,
                 data: DenseData,
                 opt_mask: Tensor,
                 model: GraphGenerativeModel,
                 smoothness_factor: float,
                 disconnection_factor: float,
                 sparsity_factor: float,
                 regularize: float = True,
                 lr_decay: float = None,
                 lr_decay_step_size: int = 1,
                 refine_embeddings: bool = False,
                 pretrain: bool = False,
                 ):
        self.lr_decay = lr_decay
        self.lr_decay_step_size = lr_decay_step_size
        self.dataset = data
        self.opt_mask = opt_mask
        self.model = model
        self.regularize = regularize
        self.smoothness_factor = smoothness_factor
        self.disconnection_factor = disconnection_factor
        self.sparsity_factor = sparsity_factor
        self.optimizer: Optimizer = optimizer
        self.lr_decayer = StepLR(self.optimizer, step_size=self.lr_decay_step_size,
                                 gamma=self.lr_decay) if self.lr_decay is not None else None
        self.refine_embeddings = refine_embeddings
        self.fused_enabled = True          # set False to force the composable route
        self.factored_enabled = True       # set False to run unrolled hyper steps on the dense composable route
        self.first_order = False           # True: drop the hypergradient terms through unrolled inner steps (always fused)
        self.n_samples = 1                 # fused route: Bernoulli samples per outer step (mean of the straight-through gradients)
        self.last_route = None             # "fused" | "composable" (observability / tests)
        self._engine = None
        self._adam_state = None
        self._host_scalars = None
        self._opt_rows = None
        self.deferred = None               # while a bilevel block is captured: (list for the device [loss, acc], device lr scalar)
        if pretrain:
            self.pretrain_model()

    # ------------------------------------------------------------------------------------------ hot path
    def train_step(self,
                   gcn_predict_fct: Callable[[Tensor], Tensor],
                   mask: Tensor = None,
                   retain_graph: bool = True) -> Metrics:
        plan = self._fused_plan(gcn_predict_fct, mask)
        if plan is not None:
            route, inner, opt = plan
            if route == "fused":
                return self._train_step_fused(inner, opt)
            if route == "sharded":
                return self._train_step_sharded(inner, opt)
            return self._train_step_factored(inner, opt, retain_graph)
        return self._train_step_composable(gcn_predict_fct, mask, retain_graph)

    def _train_step_composable(self, gcn_predict_fct, mask, retain_graph) -> Metrics:
        self.last_route = "composable"
        self.model.train()
        self.optimizer.zero_grad()
        graph = self.model.sample()
        predictions = gcn_predict_fct(graph)
        mask = mask or self.opt_mask
        loss = F.nll_loss(predictions[mask], self.dataset.y[mask])
        acc = accuracy(predictions[mask], self.dataset.y[mask])
        if self.regularize:
            loss += graph_regularization(graph=graph, features=self.dataset.x,
                                         smoothness_factor=self.smoothness_factor,
                                         disconnection_factor=self.disconnection_factor,
                                         sparsity_factor=self.sparsity_factor)
        loss.backward(retain_graph=retain_graph)
        self.optimizer.step()
        if self.lr_decayer is not None:
            self.lr_decayer.step()
        self.model.project_parameters()
        if self.refine_embeddings:
            self.model.refine()
        return Metrics(loss=loss.item(), acc=acc)

    # ---- factored route (unrolled hypergradient) ---------------------------------------------------
    def unroll_plan(self, inner) -> bool:
        """True when the inner steps of `inner` may train on FactoredGraphs: the conditions of the fused step."""
        return self.factored_enabled and not self.first_order and self._fused_plan(inner.model_forward, None) is not None

    def sample_for_unroll(self, inner):
        """The graph of one unrolled inner step (src/trainers/bilevel.py:103-107): factored when the hyper step that
        will differentiate through it takes the factored route, else `sample()`."""
        if self.unroll_plan(inner):
            return self.model.sample_factored()
        return self.model.sample()

    def _train_step_factored(self, inner, opt, retain_graph) -> Metrics:
        from .. import kernels
        self.last_route = "factored"
        model = self.model
        model.train()
        self.optimizer.zero_grad()
        sink = model.factor_sink
        sink.clear()
        graph = model.sample_factored()
        key = (self.opt_mask.data_ptr(), self.opt_mask._version)
        if self._opt_rows is None or self._opt_rows[0] != key:          # `tensor[bool_mask]` would sync the host every step
            self._opt_rows = (key, MaskInfo(self.opt_mask, self.dataset.y))
        info = self._opt_rows[1]
        logits = inner.logits_forward(graph)
        if fused_nll_ok(logits):                                        # log-softmax + row selection + NLL + accuracy: one launch
            loss, correct = masked_nll(logits, info)
        else:
            selected = F.log_softmax(logits, dim=1).index_select(0, info.rows)
            loss = F.nll_loss(selected, info.selected_labels)
            correct = (torch.argmax(selected.detach(), dim=-1) == info.selected_labels).float().mean()
        loss.backward(retain_graph=retain_graph)
        n = model._n
        fa, fb, cvec = sink.collect(n, logits.device)
        sink.clear()
        kind, group = opt
        probs = model._probs_param()
        if probs.grad is not None:
            # part of the history was built on dense graphs (`sample()`): their share arrived in probs.grad through autograd.
            # Add the factored share as a dense gradient and finish like the composable route.
            dense = kernels.k3_dense_grad(n, fa, fb, cvec)
            g = kernels.theta_full_to_triu(dense, n, sym_sum=True)
            inside = (probs.detach() >= 0.0) & (probs.detach() <= 1.0)
            probs.grad.add_(g * inside)
            self.optimizer.step()
            model.project_parameters()
        else:
            theta = model.theta_full()
            if self.deferred is not None:
                # captured block: the learning rate lives in device memory (it decays between replays) and is folded into the
                # factors — the gradient is linear in (fa, c) — so the kernel's by-value lr stays 1
                if kind != _lib.OPT_SGD:
                    raise NotImplementedError("captured bilevel blocks support the SGD outer optimiser")
                lr_dev = self.deferred[1]
                kernels.k3k4_theta_update_tc_(theta, n, fa * lr_dev, fb, cvec * lr_dev, 1.0)
                self.deferred[0].append(torch.stack((loss.detach(), correct)))
                return None
            if kind == _lib.OPT_SGD:
                # lr folded into the factors exactly as the captured block does (where it must live in device memory), so a
                # step-by-step run and a replayed block produce the same bits
                lr = float(group["lr"])
                kernels.k3k4_theta_update_tc_(theta, n, fa * lr, fb, cvec * lr, 1.0)
            else:
                st = self._adam_state
                if st is None or st["m"].shape != theta.shape:
                    st = self._adam_state = {"m": torch.zeros_like(theta), "v": torch.zeros_like(theta), "t": 0}
                st["t"] += 1
                kernels.k3k4_theta_update_(theta, n, fa, fb, cvec, group["lr"], opt_kind=kind, adam_m=st["m"], adam_v=st["v"],
                                           betas=group["betas"], eps=group["eps"], t=st["t"])
            model.mark_full_updated()
            self.optimizer._opt_called = True
        if self.lr_decayer is not None:
            self.lr_decayer.step()
        loss_value, acc = torch.stack((loss.detach(), correct)).tolist()          # one device->host transfer
        return Metrics(loss=loss_value, acc=acc)

    # ---- fused route -----------------------------------------------------------------------------
    def _optimizer_kind(self):
        """(kind, hyper-parameters) when the optimiser is one the update kernel implements exactly, else None."""
        opt = self.optimizer
        if len(opt.param_groups) != 1:
            return None
        group = opt.param_groups[0]
        params = group["params"]
        if len(params) != 1 or params[0] is not self.model._probs_param():
            return None
        if type(opt) is torch.optim.SGD:
            if group.get("momentum", 0) or group.get("dampening", 0) or group.get("weight_decay", 0) or group.get("nesterov", False) \
                    or group.get("maximize", False):
                return None
            return _lib.OPT_SGD, group
        if type(opt) is torch.optim.Adam:
            if group.get("weight_decay", 0) or group.get("amsgrad", False) or group.get("maximize", False):
                return None
            return _lib.OPT_ADAM, group
        return None

    def _fused_plan(self, gcn_predict_fct, mask):
        if not self.fused_enabled or mask is not None or self.regularize or self.refine_embeddings:
            return None
        model = self.model
        if type(model) is not BernoulliGraphModel or model.directed or not model._probs_param().is_cuda:
            return None
        cfg = Sampler._ingredient.values
        if not cfg["undirected"] or cfg["sparsification"] != "NONE" or cfg["dense"]:
            return None
        owner = getattr(gcn_predict_fct, "__self__", None)
        from .inner import InnerProblemTrainer
        if not isinstance(owner, InnerProblemTrainer) or getattr(gcn_predict_fct, "__func__", None) is not InnerProblemTrainer.model_forward:
            return None
        gcn = owner.model
        if not getattr(gcn, "normalize_adj", False) or owner.data.x is not self.dataset.x and owner.data.x.data_ptr() != self.dataset.x.data_ptr():
            return None
        # The fused step differentiates the DIRECT term only. If the fast weights still carry the unrolled inner
        # steps' history (src/trainers/inner.py:71-72), the reference's backward also flows through those steps into
        # the graphs they sampled; that needs autograd, so stay composable unless the caller opted for first order.
        opt = self._optimizer_kind()
        if opt is None:
            return None
        if model.row_block is not None:
            # a model that owns a row block of theta (large graphs): the direct step only — its weights must not carry an
            # unrolled history (the unrolled hypergradient of sharded graphs is not implemented)
            if any(p.grad_fn is not None for p in owner.model_params.values()) and not self.first_order:
                raise NotImplementedError("row-block models support the direct outer step (detach the inner trainer or set first_order)")
            return "sharded", owner, opt
        if not self.first_order and any(p.grad_fn is not None for p in owner.model_params.values()):
            return ("factored", owner, opt) if self.factored_enabled else None
        return "fused", owner, opt

    def _get_engine(self, gcn):
        from .. import kernels
        x = self.dataset.x
        h, c = gcn.layer_in.fc.out_features, gcn.layer_out.fc.out_features
        key = (x.data_ptr(), self.opt_mask.data_ptr(), h, c)
        if self._engine is None or self._engine[0] != key:
            eng = kernels.OuterStep(self.model._n, x, self.dataset.y, self.opt_mask, hidden=h, classes=c)
            self._engine = (key, eng)
        return self._engine[1]

    def _train_step_fused(self, inner, opt) -> Metrics:
        self.last_route = "fused"
        model, gcn = self.model, inner.model
        if not model.training:
            model.train()
        if not gcn.training:
            gcn.train(True)                               # side effect of model_forward(graph, is_train=True)
        eng = self._get_engine(gcn)
        params = inner.model_params
        eng.set_weights(params["layer_in.fc.weight"], params["layer_in.fc.bias"],
                        params["layer_out.fc.weight"], params["layer_out.fc.bias"])
        kind, group = opt
        theta = model.theta_full()
        extra = {}
        if kind == _lib.OPT_ADAM:
            st = self._adam_state
            if st is None or st["m"].shape != theta.shape:
                st = self._adam_state = {"m": torch.zeros_like(theta), "v": torch.zeros_like(theta), "t": 0}
            st["t"] += 1
            extra = dict(adam_m=st["m"], adam_v=st["v"], betas=group["betas"], eps=group["eps"], adam_t=st["t"])
        seed, step = PHILOX.next_step()
        # (loss, acc) land in pinned host memory, written by the step's own kernel as soon as they are final (after the
        # second propagation), followed by a tag; the host polls for the tag and returns the metrics while the backward
        # half and the theta update are still running. Everything later is ordered by the stream.
        if self._host_scalars is None:
            self._host_scalars = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._host_view = self._host_scalars.numpy()      # shares the pinned memory
            self._tag = 0
        self._tag = self._tag % 1000000 + 1                   # exactly representable in fp32, never 0
        tag = float(self._tag)
        if self.n_samples > 1:
            if kind != _lib.OPT_SGD:
                raise NotImplementedError("n_samples > 1 is implemented for the SGD outer optimiser (models/factory.py:66-69)")
            eng.run_multi(theta, self.n_samples, lr=group["lr"], seed=seed, step=step, dropout_p=float(gcn.dropout),
                          update=True, scalars_out=self._host_scalars, want_adj=False, scalars_tag=tag)
        else:
            eng.run(theta, lr=group["lr"], seed=seed, step=step, dropout_p=float(gcn.dropout),
                    update=True, opt_kind=kind, scalars_out=self._host_scalars, want_adj=False, scalars_tag=tag, **extra)
        model.mark_full_updated()
        if self.lr_decayer is not None:
            self.optimizer._opt_called = True             # the update ran in the kernel; keeps StepLR's order check quiet
            self.lr_decayer.step()
        hv = self._host_view
        spins = 0
        while hv[2] != tag:
            spins += 1
            if spins > 200000:                            # ~50 ms of polling: let the stream report what happened
                torch.cuda.current_stream().synchronize()
                if hv[2] != tag:
                    raise RuntimeError("lds_outer_step finished without publishing its metrics")
                break
        loss, acc = float(hv[0]), float(hv[1])
        return Metrics(loss=loss, acc=acc)

    # ---- row-block route (BASELINE.json configs 4 and 5 behind the same call) ---------------------
    def _train_step_sharded(self, inner, opt) -> Metrics:
        """`model` owns rows [row0, row0 + rows) of theta (`BernoulliGraphModel.from_row_block`) and `dataset.x / y`, `opt_mask`
        hold the same rows. rows == n: the whole step on this GPU (bit-packed plan); rows < n: this process is one rank of the
        row-block sharded step (lds_gnn_b200/sharded.py) — every rank calls train_step, the exchange runs over
        torch.distributed (NVLink peer-memory push under NCCL). The Philox seed must agree on all ranks (PHILOX.manual_seed)."""
        from .. import kernels, sharded
        self.last_route = "sharded"
        model, gcn = self.model, inner.model
        model.train(); gcn.train(True)
        kind, group = opt
        if kind != _lib.OPT_SGD:
            raise NotImplementedError("row-block models are stepped with the SGD outer optimiser (models/factory.py:66-69)")
        row0, rows = model.row_block
        n = model._n
        h, c = gcn.layer_in.fc.out_features, gcn.layer_out.fc.out_features
        x = self.dataset.x
        key = (x.data_ptr(), self.opt_mask.data_ptr(), h, c, row0, rows)
        if self._engine is None or self._engine[0] != key:
            if x.shape[0] != rows:
                raise ValueError(f"the dataset of a row-block model holds the block's {rows} rows (got {x.shape[0]})")
            if rows == n:
                eng, comm = kernels.OuterStep(n, x, self.dataset.y, self.opt_mask, hidden=h, classes=c), None
            else:
                comm = sharded.make_comm(n, x.device)
                count = comm.all_reduce_sum(self.opt_mask.sum().to(torch.float32).reshape(1).clone())
                eng = sharded.ShardedOuterStep(n, row0, rows, x, self.dataset.y, self.opt_mask, int(count.item()), h, c)
            self._engine = (key, eng, comm)
        _, eng, comm = self._engine
        params = inner.model_params
        eng.set_weights(params["layer_in.fc.weight"], params["layer_in.fc.bias"], params["layer_out.fc.weight"], params["layer_out.fc.bias"])
        seed, step = PHILOX.next_step()
        theta = model.theta_full()
        if comm is None:
            out = eng.run(theta, lr=group["lr"], seed=seed, step=step, dropout_p=float(gcn.dropout), update=True, want_adj=False)
        else:
            out = eng.run(theta, comm, lr=group["lr"], seed=seed, step=step, dropout_p=float(gcn.dropout), update=True, want_adj=False)
        if self.lr_decayer is not None:
            self.optimizer._opt_called = True
            self.lr_decayer.step()
        loss, acc = out[:2].tolist()                          # one device->host transfer
        return Metrics(loss=loss, acc=acc)

    # ------------------------------------------------------------------------------------------ rest of the API
    def sample(self) -> Tensor:
        return self.model.sample()

    def detach(self):
        """The reference reloads both state dicts into themselves (outer.py:92-94) — a no-op for a leaf
        Parameter and a stateless SGD. Nothing here holds an autograd graph across steps, so nothing to cut."""
        return None

    def get_learning_rates(self) -> List[float]:
        if self.optimizer is None:
            raise ValueError("Can't get optimizer learning rate, no optimizer initialized yet.")
        return get_lr(self.optimizer)

    def train(self, mode: bool = True):
        self.model.train(mode=mode)

    def eval(self):
        self.model.eval()

    def pretrain_model(self) -> None:
        raise NotImplementedError("link-prediction pre-training (src/trainers/pretrainer.py) is outside the LDS hot path; "
                                  "construct the trainer with pretrain=False (set 'outer-trainer.pretrain' to False)")


class OuterProblemTrainerFactory:
    _ingredient = Ingredient("outer-trainer")
    INGREDIENTS = {"outer-trainer": _ingredient}

    @staticmethod
    @_ingredient.config
    def _config():
        lr_decay: float = 1.0               # noqa: F841
        lr_decay_step_size: int = 1         # noqa: F841
        refine_embeddings: bool = False     # noqa: F841
        pretrain: bool = True               # noqa: F841  (reference default, outer.py:125; the pretrainer itself is out of scope)
        regularize: bool = False            # noqa: F841
        smoothness_factor: float = 0.0      # noqa: F841
        disconnection_factor: float = 0.0   # noqa: F841
        sparsity_factor: float = 0.0        # noqa: F841

    @staticmethod
    @_ingredient.capture
    def trainer(optimizer: Optimizer,
                data: DenseData,
                opt_mask: Tensor,
                model: GraphGenerativeModel,
                regularize: bool,
                smoothness_factor: float,
                disconnection_factor: float,
                sparsity_factor: float,
                lr_decay: float = None,
                lr_decay_step_size: int = 1,
                refine_embeddings: bool = False,
                pretrain: bool = False) -> OuterProblemTrainer:
        return OuterProblemTrainer(optimizer=optimizer, data=data, opt_mask=opt_mask, model=model, lr_decay=lr_decay,
                                   lr_decay_step_size=lr_decay_step_size, refine_embeddings=refine_embeddings,
                                   pretrain=pretrain, regularize=regularize, smoothness_factor=smoothness_factor,
                                   disconnection_factor=disconnection_factor, sparsity_factor=sparsity_factor)
