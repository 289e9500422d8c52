"""oracle/make_golden.py — TEST INFRASTRUCTURE. Writes tests/golden/*.npz from the LIVE reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Each file holds the inputs of a small outer-step case (theta, explicit uniforms from oracle/philox.py
with the stored seed/step, features, GCN weights, labels, mask, dropout keep-masks, lr) and the
reference's own outputs for it: sampled graph, log-probs, loss, accuracy, probs.grad, updated theta,
statistics() — produced by `OuterProblemTrainer.train_step` of the unmodified reference
(src/trainers/outer.py:57-87) in fp32 AND in fp64 (fp64 = the accuracy yardstick for the 1e-3 bar).
"""
import os

import numpy as np
import torch

from . import live_reference as L
from . import philox as PH

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_inputs(seed, n, f, h, c, theta_kind, p, mask_frac=0.4):
    rng = np.random.default_rng(seed)
    t = n * (n + 1) // 2
    if theta_kind == "uniform":
        theta = rng.random(t).astype(np.float32)
    elif theta_kind == "mixed":              # corners 0/1 plus interior values
        theta = rng.random(t).astype(np.float32)
        theta[rng.random(t) < 0.35] = 0.0
        theta[rng.random(t) < 0.10] = 1.0
    elif theta_kind == "binary":             # theta_0 = adjacency (factory.py:62): deterministic sample
        theta = (rng.random(t) < 0.15).astype(np.float32)
    elif theta_kind == "sparse":             # low-degree graph, most rows have few neighbours
        theta = (rng.random(t) < 3.0 / n).astype(np.float32) * rng.random(t).astype(np.float32)
    else:
        raise ValueError(theta_kind)
    x = (rng.random((n, f)) < 0.2).astype(np.float32)
    x[np.arange(n), rng.integers(0, f, n)] = 1.0
    x /= x.sum(1, keepdims=True)
    lim0, lim1 = np.sqrt(6.0 / (f + h)), np.sqrt(6.0 / (h + c))      # xavier-uniform (layers.py:38-40)
    w0 = rng.uniform(-lim0, lim0, (h, f)).astype(np.float32)
    w1 = rng.uniform(-lim1, lim1, (c, h)).astype(np.float32)
    b0 = (rng.standard_normal(h) * 0.05).astype(np.float32)
    b1 = (rng.standard_normal(c) * 0.05).astype(np.float32)
    y = rng.integers(0, c, n).astype(np.int64)
    mask = rng.random(n) < mask_frac
    mask[0] = True
    return dict(theta_triu=theta, x=x, w0=w0, b0=b0, w1=w1, b1=b1, y=y, mask=mask)


CASES = [
    # name,        seed, n,   f,  h,  c, theta,     p,   steps, lr,  decay
    ("n20_plain",    1,  20,  12,  8, 3, "uniform", 0.0, 1,     1.0, None),
    ("n33_twostep",  2,  33,  17, 16, 7, "mixed",   0.0, 2,     0.5, 0.9),
    ("n64_dropout",  3,  64,  40, 16, 7, "mixed",   0.5, 1,     0.1, 0.99),
    ("n96_binary",   4,  96,  30, 16, 6, "binary",  0.0, 1,     1.0, 1.0),
    ("n130_sparse",  5, 130,  64, 16, 7, "sparse",  0.5, 2,     0.1, 0.99),
    ("n257_h64",     6, 257,  50, 64, 7, "uniform", 0.0, 1,     0.1, 0.99),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, seed, n, f, h, c, kind, p, steps, lr, decay in CASES:
        inp = make_inputs(seed, n, f, h, c, kind, p)
        philox_seed = 0x5EED0000 + seed
        uniforms = [PH.edge_uniforms(n, philox_seed, s) for s in range(steps)]
        keeps = []
        if p > 0:
            for s in range(steps):
                keeps += [PH.dropout_keep_mask(n, f, p, philox_seed, s, PH.STREAM_DROP_X),
                          PH.dropout_keep_mask(n, h, p, philox_seed, s, PH.STREAM_DROP_H)]
        out = dict(inp)
        out.update(n=n, f=f, h=h, c=c, p=np.float64(p), lr=np.float64(lr), steps=steps,
                   lr_decay=np.float64(-1.0 if decay is None else decay), philox_seed=np.int64(philox_seed))
        for s in range(steps):
            out[f"U{s}"] = uniforms[s]
            if p > 0:
                out[f"keep_x{s}"] = keeps[2 * s]
                out[f"keep_h{s}"] = keeps[2 * s + 1]
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            res, stats = L.outer_steps(inp["theta_triu"], uniforms, inp["x"], inp["w0"], inp["b0"], inp["w1"],
                                       inp["b1"], inp["y"], inp["mask"], lr=lr, lr_decay=decay, p=p,
                                       keep_masks=keeps, dtype=dt)
            for s, r in enumerate(res):
                if tag == "f32":
                    out[f"sample{s}"] = r["sample"].astype(np.uint8)
                for k in ("logp", "loss", "acc", "grad_triu", "theta_new", "lr_used"):
                    out[f"{k}{s}_{tag}"] = r[k]
            for k, v in stats.items():
                out[f"stat_{k}_{tag}"] = np.float64(v)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
