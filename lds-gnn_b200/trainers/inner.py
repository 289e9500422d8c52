"""Inner (GCN weights) trainer with the reference's API (src/trainers/inner.py:15-125).

On the outer-step hot path only `model_forward` matters: it is the `gcn_predict_fct` the outer trainer
receives, and the fused outer step recognises it to read the fast weights. `train_step` — the unrolled
inner step the hypergradient later flows back through — is the "next" row of the scope table
(SURVEY.md §8f #2): it is functional here, with sampling on the K1 kernel and the GCN forward written
in differentiable torch ops (O(N^2) elementwise normalisation, no N^3 products) because that forward must
support double backward; its optimiser is `diffopt.DifferentiableAdam` (numeric parity with `higher` unpinned).
"""
from collections import OrderedDict
from typing import Dict, List

import torch
import torch.nn.functional as F
from torch.optim import Adam

from ..models.gcn import MetaDenseGCN, double_backward_path
from ..utils.evaluation import accuracy
from ..utils.graph import DenseData, is_square_matrix
from ..utils.rowops import MaskInfo, fused_nll_ok, masked_nll
from . import Metrics
from .diffopt import DifferentiableAdam


def copy_detach_parameter_dict(parameters: OrderedDict) -> OrderedDict:
    return OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in parameters.items())


class InnerProblemTrainer:
    def __init__(self, model: MetaDenseGCN, data: DenseData, lr: float = 0.01, weight_decay: float = 1e-4):
        self.model = model
        self.lr = lr
        self.weight_decay = weight_decay
        self.model_params: OrderedDict = OrderedDict(model.named_parameters())
        self.optimizer: DifferentiableAdam = None
        self.data = data
        self._rows_cache = {}
        self.deferred = None            # a list while a bilevel block is captured: train_step appends its device [loss, acc] and returns None
        self.reset_optimizer()

    def reset_weights(self):
        self.model.reset_weights()
        self.model_params = OrderedDict(self.model.named_parameters())

    def reset_optimizer(self) -> None:
        """Adam with weight decay on the first layer only (inner.py:42-50)."""
        optimizer = Adam([{"params": self.model.layer_in.parameters(), "weight_decay": self.weight_decay},
                          {"params": self.model.layer_out.parameters()}], lr=self.lr)
        self.optimizer = DifferentiableAdam(optimizer, self.model.parameters())

    def copy_model_params(self) -> Dict:
        return copy_detach_parameter_dict(self.model_params)

    def _masked_rows(self, mask: torch.Tensor):
        """(row indices, their labels, MaskInfo) of a boolean node mask, cached: `tensor[bool_mask]` costs a device->host sync per use."""
        key = (mask.data_ptr(), mask._version, self.data.y.data_ptr())
        hit = self._rows_cache.get(key)
        if hit is None:
            info = MaskInfo(mask, self.data.y)
            hit = self._rows_cache[key] = (info.rows, info.selected_labels, mask, info)       # the mask is kept alive with its key
            if len(self._rows_cache) > 8:
                self._rows_cache.pop(next(iter(self._rows_cache)))
        return hit[0], hit[1], hit[3]

    def train_step(self, graph: torch.Tensor, mask: torch.Tensor = None) -> Metrics:
        """One differentiable optimiser step on the training nodes (inner.py:55-74). One host sync: (loss, acc) together."""
        assert is_square_matrix(graph)
        loss, correct = self.loss_and_accuracy(graph, mask or self.data.train_mask, is_train=True)
        new_params = self.optimizer.step(loss, params=self.model_params.values())
        self._update_model_params(list(new_params))
        if self.deferred is not None:
            self.deferred.append(torch.stack((loss.detach(), correct)))
            return None
        loss_value, acc = torch.stack((loss.detach(), correct)).tolist()
        return Metrics(loss=loss_value, acc=acc)

    def loss_and_accuracy(self, graph, mask: torch.Tensor, is_train: bool = True):
        """(NLL, accuracy) of the GCN at the current fast weights on the rows of `mask`, as 0-dim device tensors; the loss is
        differentiable to second order. On CUDA the log-softmax, the row selection, the NLL and the accuracy are one launch on
        the logits (utils/rowops.py) instead of ~7 (and ~21 more in the two backward orders)."""
        rows, labels, info = self._masked_rows(mask)
        with double_backward_path():
            logits = self.logits_forward(graph, is_train=is_train)
        if fused_nll_ok(logits):
            return masked_nll(logits, info)
        selected = F.log_softmax(logits, dim=1).index_select(0, rows)
        return F.nll_loss(selected, labels), (torch.argmax(selected.detach(), dim=-1) == labels).float().mean()

    def logits_forward(self, graph, is_train: bool = True) -> torch.Tensor:
        """model_forward without the final log-softmax (the fused loss works on the logits)."""
        self.model.train(mode=is_train)
        return self.model.forward_to_last_layer(self.data.x, graph, params=self.model_params)

    def model_forward(self, graph, is_train: bool = True) -> torch.Tensor:
        """The `gcn_predict_fct` of the outer step (inner.py:76-78): sets train/eval mode, runs the GCN at the
        current fast weights."""
        self.model.train(mode=is_train)
        return self.model(self.data.x, graph, params=self.model_params)

    def evaluate(self, graph: torch.Tensor, mask: torch.Tensor = None) -> Metrics:
        self.model.eval()
        with torch.no_grad():
            predictions = self.model_forward(graph, is_train=False)
            mask = mask or self.data.val_mask
            loss = F.nll_loss(predictions[mask], self.data.y[mask])
            acc = accuracy(predictions[mask], self.data.y[mask])
        return Metrics(loss=loss.item(), acc=acc)

    def detach(self) -> None:
        """Truncate the unroll: fast weights and optimiser state become leaves again (inner.py:98-125)."""
        self.model_params = copy_detach_parameter_dict(self.model_params)
        self.detach_optimizer()

    def _update_model_params(self, new_model_params: List[torch.Tensor]) -> None:
        for name, value in zip(list(self.model_params.keys()), new_model_params):
            self.model_params[name] = value

    def detach_optimizer(self):
        self.optimizer.detach_()
