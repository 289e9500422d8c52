"""-m gpu: final accuracy of the full bilevel loop against the UNMODIFIED reference.

tests/golden/accuracy_live_reference_<shape>.json holds the final validation / test accuracies of the reference's own
`BilevelProblemRunner.train` + `evaluate` (src/trainers/bilevel.py:34-145), run from /root/reference on the CPU of the build
container by oracle/accuracy_live_reference.py — no code shared with this package except the synthetic dataset generator.
The same datasets (same seeds), hyper-parameters and stopping rule run here through the package on the GPU (captured bilevel
blocks, factored hypergradient, Philox sampling). The two arms draw different random numbers, so runs are not comparable one
by one; the north star asks for the MEAN final validation accuracy within 0.5 pt over 5 seeds — with the 3 (Cora shape) / 2
(Citeseer shape) seeds the CPU reference could afford (10 / 25 minutes per seed) the tolerance here is 1.5 pt on the means and
4 pt on any single seed; measured: +0.10 / -0.21 pt validation, -0.12 / -0.38 pt test (datasets differ by seed far
more than the arms do: 0.92-0.96 across seeds)."""
import argparse
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _parity_module():
    spec = importlib.util.spec_from_file_location("accuracy_parity", os.path.join(HERE, "accuracy_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("shape", ["cora", "citeseer"])
def test_final_accuracy_matches_the_live_reference(shape):
    path = os.path.join(HERE, "golden", f"accuracy_live_reference_{shape}.json")
    if not os.path.exists(path):
        pytest.skip(f"no live-reference run committed for the {shape} shape (oracle/accuracy_live_reference.py)")
    gold = json.load(open(path))
    a = gold["args"]
    args = argparse.Namespace(hidden=a["hidden"], dropout=a["dropout"], gcn_lr=a["gcn_lr"], gcn_wd=a["gcn_wd"], lds_lr=a["lds_lr"],
                              lr_decay=a["lr_decay"], patience=a["patience"], tau=a["tau"], samples=a["samples"],
                              inner_max=a["inner_max"], outer_max=a["outer_max"])
    mod = _parity_module()
    ours, ref = [], []
    for row in gold["rows"]:
        r = mod.run_arm("ours", shape, int(row["seed"]), args, torch.device("cuda"))
        ours.append((r["acc.val.final"], r["acc.test.final"]))
        ref.append((row["acc.val.final"], row["acc.test.final"]))
    ours, ref = np.array(ours), np.array(ref)
    diff_means = ours.mean(0) - ref.mean(0)
    print(f"{shape}: ours val/test {ours.mean(0)}, live reference {ref.mean(0)}, per seed {np.round(ours - ref, 4).tolist()}")
    assert np.all(np.abs(diff_means) <= 0.015), (diff_means, ours, ref)
    assert np.all(np.abs(ours - ref) <= 0.04), (ours - ref)
