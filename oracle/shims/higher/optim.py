"""DifferentiableAdam restated from higher's published algorithm (facebookresearch/higher, git master — what the
reference installs, scripts/install.sh:3-4; call sites: src/trainers/inner.py:48-50, 71): torch.optim.Adam's update
applied out of place with create_graph=True so later losses can differentiate through it.

    g += wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2        (v == 0 entries get their sqrt-gradient masked)
    denom = sqrt(v) / sqrt(1 - b2^t) + eps;  step_size = lr / (1 - b1^t);  p' = p - step_size * m / denom

i.e. the bias correction is applied to sqrt(v) BEFORE eps is added (the form torch.optim.Adam uses since 1.6 and
higher's master follows). Numeric parity with the real `higher` is UNPINNED: the package is absent and the reference's
tests at this boundary are behavioural only."""
import math

import torch


class DifferentiableOptimizer:
    def __init__(self, other, reference_params, fmodel=None, device=None, override=None, track_higher_grads=True):
        self.param_groups = []
        self.state = []
        self._track = track_higher_grads
        ref = list(reference_params)
        index = {id(p): i for i, p in enumerate(ref)}
        self._group_to_param_list = []
        for g in other.param_groups:
            ng = {k: v for k, v in g.items() if k != "params"}
            ng["params"] = [None] * len(g["params"])
            self.param_groups.append(ng)
            self._group_to_param_list.append([index[id(p)] for p in g["params"]])
            self.state.append({})


class DifferentiableAdam(DifferentiableOptimizer):
    def step(self, loss, params=None):
        params = list(params)
        grads = torch.autograd.grad(loss, params, create_graph=self._track, allow_unused=True)
        new_params = list(params)
        for gi, (group, mapping) in enumerate(zip(self.param_groups, self._group_to_param_list)):
            beta1, beta2 = group["betas"]
            for slot, pidx in enumerate(mapping):
                p, g = params[pidx], grads[pidx]
                if g is None:
                    continue
                st = self.state[gi].setdefault(slot, {})
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p.data)
                    st["exp_avg_sq"] = torch.zeros_like(p.data)
                st["step"] += 1
                bc1 = 1 - beta1 ** st["step"]
                bc2 = 1 - beta2 ** st["step"]
                if group["weight_decay"] != 0:
                    g = g + group["weight_decay"] * p
                st["exp_avg"] = exp_avg = st["exp_avg"] * beta1 + (1 - beta1) * g
                st["exp_avg_sq"] = exp_avg_sq = st["exp_avg_sq"] * beta2 + (1 - beta2) * g * g
                # higher masks the sqrt's infinite gradient where v == 0 (_maybe_mask); a 1e-30 floor does the same job
                safe = exp_avg_sq + (exp_avg_sq == 0).to(exp_avg_sq.dtype) * 1e-30
                denom = safe.sqrt() / math.sqrt(bc2) + group["eps"]
                step_size = group["lr"] / bc1
                new_params[pidx] = p - step_size * (exp_avg / denom)
        return new_params
