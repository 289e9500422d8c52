#!/usr/bin/env python
"""Accuracy parity of the full bilevel LDS loop (north-star check: final validation accuracy within 0.5 pt over 5 seeds
on synthetic data of Cora / Citeseer shape).

Both arms run the SAME host loop (lds_gnn_b200.trainers.BilevelProblemRunner — the reference's control flow,
src/trainers/bilevel.py:34-145) on the same synthetic dataset, splits, hyper-parameters and seeds, on the GPU:

  ours       BernoulliGraphModel + MetaDenseGCN of this package: K1 sampling (Philox), tcgen05 propagation, closed-form
             hypergradient kernels (fused route when the fast weights are leaves, composable autograd route otherwise)
  reference  the reference's algorithm as a torch tensor program (oracle/reference_port.py): torch.bernoulli over the
             full N x N matrix from torch's generator, triu mirror, straight-through estimator, dense normalisation,
             torch.mm propagation, autograd for everything — i.e. what the reference executes, on the same device

Lives under tests/ because it runs the oracle's port of the reference as its second arm (the oracle is test infrastructure:
only tests/, smoke() and bench.py's cpu_baseline leg may execute it). Not collected by pytest; run it by hand on a B200:
    python tests/accuracy_parity.py --shape cora --seeds 5
Writes gpurun_out/accuracy_parity_<shape>.json and prints a markdown table.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
from torch import nn

from lds_gnn_b200.data import make_dataset
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.models.sampling import PHILOX
from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
from lds_gnn_b200.trainers.inner import InnerProblemTrainer
from lds_gnn_b200.trainers.outer import OuterProblemTrainer
from lds_gnn_b200.utils.graph import get_triu_values, split_mask
from oracle import reference_port as P


class PortGraphModel(nn.Module):
    """The reference's BernoulliGraphModel + Sampler as plain torch ops (oracle/reference_port.py)."""

    def __init__(self, init_matrix):
        super().__init__()
        self.probs = nn.Parameter(get_triu_values(init_matrix))

    def forward(self):
        return P.theta_matrix(self.probs)

    def sample(self):
        return P.draw_graph(self.forward())

    def project_parameters(self):
        self.probs.data.clamp_(0.0, 1.0)

    def refine(self):
        pass

    def statistics(self):
        return {}


def run_arm(arm, shape, seed, args, device):
    torch.manual_seed(seed)
    np.random.seed(seed)
    PHILOX.manual_seed(seed)
    data = make_dataset(shape, seed=seed).to(device)
    data.val_mask, opt_mask = split_mask(data.val_mask, ratio=0.5, shuffle=True, device=device)     # scripts/bilevel.py:77
    gcn = MetaDenseGCN(data.num_features, args.hidden, data.num_classes, dropout=args.dropout).to(device)
    inner = InnerProblemTrainer(gcn, data, lr=args.gcn_lr, weight_decay=args.gcn_wd)
    model = (BernoulliGraphModel(data.dense_adj) if arm == "ours" else PortGraphModel(data.dense_adj)).to(device)
    opt = torch.optim.SGD(model.parameters(), lr=args.lds_lr)
    outer = OuterProblemTrainer(optimizer=opt, data=data, opt_mask=opt_mask, model=model, smoothness_factor=0.0, disconnection_factor=0.0,
                                sparsity_factor=0.0, regularize=False, lr_decay=args.lr_decay, pretrain=False)
    runner = BilevelProblemRunner(inner, outer, data, n_samples_empirical_mean=args.samples)
    t0 = time.time()
    runner.train(patience=args.patience, hyper_gradient_interval=args.tau, inner_loop_max_epochs=args.inner_max,
                 outer_loop_max_epochs=args.outer_max)
    out = runner.evaluate()
    out["seconds"] = time.time() - t0
    out["route"] = outer.last_route
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="cora")
    ap.add_argument("--seeds", type=int, default=5)
    ap.add_argument("--hidden", type=int, default=16)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--gcn-lr", type=float, default=0.01)
    ap.add_argument("--gcn-wd", type=float, default=5e-4)
    ap.add_argument("--lds-lr", type=float, default=0.1)
    ap.add_argument("--lr-decay", type=float, default=0.99)
    ap.add_argument("--patience", type=int, default=20)
    ap.add_argument("--tau", type=int, default=5)
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--inner-max", type=int, default=150)
    ap.add_argument("--outer-max", type=int, default=4)
    args = ap.parse_args()
    device = torch.device("cuda")
    rows = []
    for seed in range(args.seeds):
        row = {"seed": seed}
        for arm in ("reference", "ours"):
            r = run_arm(arm, args.shape, seed, args, device)
            row[arm] = r
            print(f"seed {seed} {arm:9s}: val acc {r['acc.val.final']:.4f} test acc {r['acc.test.final']:.4f} "
                  f"val loss {r['loss.val.final']:.4f} ({r['seconds']:.1f} s, last outer route: {r['route']})", flush=True)
        rows.append(row)
    summary = {}
    for key in ("acc.val.final", "acc.test.final", "loss.val.final"):
        a = np.array([r["ours"][key] for r in rows]); b = np.array([r["reference"][key] for r in rows])
        summary[key] = {"ours_mean": float(a.mean()), "ours_std": float(a.std()), "reference_mean": float(b.mean()),
                        "reference_std": float(b.std()), "diff_of_means": float(a.mean() - b.mean())}
    result = {"shape": args.shape, "args": vars(args), "rows": rows, "summary": summary,
              "note": "same host loop, dataset, splits, seeds and hyper-parameters in both arms; the arms differ in the tensor program "
                      "(CUDA kernels + Philox vs the reference's torch op sequence + torch RNG), so individual runs are not bitwise comparable"}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"accuracy_parity_{args.shape}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(result, fh, indent=1)
    print("\n| metric | ours (mean ± std) | reference port (mean ± std) | difference of means |\n|---|---|---|---|")
    for key, v in summary.items():
        print(f"| {key} | {v['ours_mean']:.4f} ± {v['ours_std']:.4f} | {v['reference_mean']:.4f} ± {v['reference_std']:.4f} | {v['diff_of_means']:+.4f} |")


if __name__ == "__main__":
    main()
