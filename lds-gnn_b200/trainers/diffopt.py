"""Differentiable Adam for the unrolled inner problem — a from-scratch stand-in for
`higher.optim.DifferentiableAdam` (reference call sites: src/trainers/inner.py:6, 42-50, 71).

`higher` (git master, unpinned, scripts/install.sh:3-4) is not vendored with the reference and cannot be
installed offline, so NUMERIC PARITY OF THIS UPDATE IS UNPINNED: this file follows the
published algorithm — `torch.optim.Adam`'s rule of that era applied out of place so later losses can
differentiate through it: g += wd * p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
p' = p - lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps). It is checked against `torch.optim.Adam` on
detached copies in tests/ (same trajectory to 1e-6 while eps placement is immaterial).
"""
import math
from typing import Iterable, List

import torch


class DifferentiableAdam:
    def __init__(self, optimizer: torch.optim.Adam, reference_params: Iterable[torch.Tensor], track_higher_grads: bool = True):
        reference = list(reference_params)
        position = {id(p): i for i, p in enumerate(reference)}
        self.track = track_higher_grads
        self.param_groups = []
        self.state = []
        self._slots = []
        for group in optimizer.param_groups:
            self.param_groups.append({k: v for k, v in group.items() if k != "params"})
            self._slots.append([position[id(p)] for p in group["params"]])
            self.state.append({})

    def step(self, loss: torch.Tensor, params: Iterable[torch.Tensor]) -> List[torch.Tensor]:
        params = list(params)
        grads = torch.autograd.grad(loss, params, create_graph=self.track, allow_unused=True)
        updated = list(params)
        for gi, (group, slots) in enumerate(zip(self.param_groups, self._slots)):
            beta1, beta2 = group["betas"]
            for slot in slots:
                p, g = params[slot], grads[slot]
                if g is None:
                    continue
                st = self.state[gi].setdefault(slot, {})
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                if group["weight_decay"] != 0:
                    g = g + group["weight_decay"] * p
                st["exp_avg"] = m = st["exp_avg"] * beta1 + (1 - beta1) * g
                st["exp_avg_sq"] = v = st["exp_avg_sq"] * beta2 + (1 - beta2) * g * g
                # sqrt has an infinite derivative at 0: floor exact zeros (higher masks that gradient instead)
                root = torch.where(v > 0, v, torch.full_like(v, 1e-30)).sqrt()
                step_size = group["lr"] * math.sqrt(1 - beta2 ** st["step"]) / (1 - beta1 ** st["step"])
                updated[slot] = p - step_size * m / (root + group["eps"])
        return updated

    def detach_(self):
        """Cut the optimiser state from the autograd graph (truncated back-propagation, inner.py:110-125)."""
        for group_state in self.state:
            for st in group_state.values():
                for key, value in st.items():
                    if torch.is_tensor(value):
                        st[key] = value.detach()
