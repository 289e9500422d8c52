"""Evaluation helpers with the reference's API (src/utils/evaluation.py): `accuracy`, `evaluate`,
`empirical_mean_loss` (n_samples x {sample a graph, GCN forward, NLL + accuracy on val and test})."""
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch
from torch.nn import functional as F

from ..trainers import Metrics


def accuracy(predictions: torch.Tensor, labels: torch.Tensor) -> float:
    """Share of rows whose arg-max equals the label (src/utils/evaluation.py:15-22)."""
    return (torch.argmax(predictions, dim=-1) == labels).float().mean().item()


def evaluate(model: torch.nn.Module, data, adj_matrix: torch.Tensor = None) -> Dict:
    model.eval()
    with torch.no_grad():
        out = model(data.x, data.dense_adj if adj_matrix is None else adj_matrix)
        result = {}
        for split, mask in (("val", data.val_mask), ("test", data.test_mask)):
            result[f"{split}.accuracy"] = accuracy(out[mask], data.y[mask])
            result[f"{split}.loss"] = F.nll_loss(out[mask], data.y[mask]).item()
    return result


def empirical_mean_loss(gcn, graph_model, n_samples: int, data, model_parameters: OrderedDict = None) -> Tuple[Metrics, Metrics]:
    """Monte-Carlo estimate of validation / test loss and accuracy under the learned graph distribution
    (src/utils/evaluation.py:51-84). Sampling and propagation run on the CUDA kernels (K1 + K2, forward only);
    the 4 x n_samples scalars are reduced on the device and read back with ONE host sync instead of 64."""
    gcn.eval()
    graph_model.eval()
    rows = []
    with torch.no_grad():
        for _ in range(n_samples):
            graph = graph_model.sample()
            predictions = gcn(data.x, graph, params=model_parameters)
            row = []
            for mask in (data.val_mask, data.test_mask):
                row.append(F.nll_loss(predictions[mask], data.y[mask]))
                row.append((torch.argmax(predictions[mask], dim=-1) == data.y[mask]).float().mean())
            rows.append(torch.stack(row))
        table = torch.stack(rows).double().mean(dim=0).tolist()       # single device->host transfer
    return Metrics(loss=table[0], acc=table[1]), Metrics(loss=table[2], acc=table[3])
