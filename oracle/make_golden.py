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


def split_masks(seed, n):
    """Disjoint train / opt (the half of the validation nodes the outer objective uses, scripts/bilevel.py:77) / val / test masks."""
    rng = np.random.default_rng(1000 + seed)
    perm = rng.permutation(n)
    k = max(4, n // 6)
    masks = []
    for b in range(4):
        m = np.zeros(n, dtype=bool)
        m[perm[b * k:(b + 1) * k]] = True
        masks.append(m)
    return masks


# ---- rows (f)1, (f)2 and config 3 of SURVEY.md §8: goldens from the LIVE reference for
#      empirical_mean_loss, the S-sample step and the unrolled bilevel block. Uniforms are not stored: they are
#      PH.edge_uniforms(n, philox_seed, step[, sample]) — the generator is pinned by the Random123 vectors (tests/test_oracle.py).
EVAL_CASES = [
    # name,          seed, n,   f,  h,  c, S, step0
    ("eval_n130_s4",   11, 130, 64, 16, 7, 4, 50),
    ("eval_n257_s16",  12, 257, 50, 64, 7, 16, 7),
]
MULTI_CASES = [
    # name,           seed, n,   f,  h,  c, S,  p,   lr,  step
    # (lr sized so that the largest step is a few 1e-2: the gradients of these small problems are ~1e-5)
    ("multi_n130_s4",   21, 130, 64, 16, 7, 4,  0.5, 600.0, 3),
    ("multi_n300_s16",  22, 300, 40, 16, 6, 16, 0.0, 8000.0, 9),
]
BLOCK_CASES = [
    # name,           seed, n,   f,  h,  c, tau, blocks, outer_lr, decay, inner_lr, wd,   step0
    ("blk_n130_tau3",   31, 130, 64, 16, 7, 3,   2,      0.5,      0.9,   0.01,     5e-4, 100),
    ("blk_n257_tau5",   32, 257, 50, 64, 7, 5,   1,      0.3,      0.99,  0.02,     1e-4, 0),
    ("blk_n96_tau1",    33, 96,  30, 16, 6, 1,   3,      1.0,      1.0,   0.01,     0.0,  5),
]


def interior(theta):
    """Every probability strictly inside (0, 1): every graph of the case is random."""
    return (0.6 * theta + 0.2).astype(np.float32)


def next_rows():
    for name, seed, n, f, h, c, S, step0 in EVAL_CASES:
        inp = make_inputs(seed, n, f, h, c, "mixed", 0.0)
        inp["theta_triu"] = interior(inp["theta_triu"])
        _, _, val_mask, test_mask = split_masks(seed, n)
        philox_seed = 0x5EED0000 + seed
        uniforms = [PH.edge_uniforms(n, philox_seed, step0 + s) for s in range(S)]
        out = dict(inp)
        out.pop("mask")
        out.update(n=n, f=f, h=h, c=c, S=S, step0=np.int64(step0), philox_seed=np.int64(philox_seed), val_mask=val_mask, test_mask=test_mask)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            out[f"metrics_{tag}"] = L.empirical_mean(inp["theta_triu"], uniforms, inp["x"], inp["w0"], inp["b0"], inp["w1"], inp["b1"],
                                                     inp["y"], val_mask, test_mask, dtype=dt)
        yield name, out
    for name, seed, n, f, h, c, S, p, lr, step in MULTI_CASES:
        inp = make_inputs(seed, n, f, h, c, "mixed", p)
        inp["theta_triu"] = interior(inp["theta_triu"])
        inp["theta_triu"][3] = 1.25; inp["theta_triu"][n + 5] = -0.1          # outside [0, 1]: the clamp backward masks them
        philox_seed = 0x5EED0000 + seed
        uniforms = [PH.edge_uniforms(n, philox_seed, step, sample=s) for s in range(S)]
        keeps = []
        if p > 0:
            for s in range(S):
                keeps += [PH.dropout_keep_mask(n, f, p, philox_seed, step, PH.STREAM_DROP_X, sample=s),
                          PH.dropout_keep_mask(n, h, p, philox_seed, step, PH.STREAM_DROP_H, sample=s)]
        out = dict(inp)
        out.update(n=n, f=f, h=h, c=c, S=S, p=np.float64(p), lr=np.float64(lr), step=np.int64(step), philox_seed=np.int64(philox_seed))
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            r = L.multi_sample_step(inp["theta_triu"], uniforms, inp["x"], inp["w0"], inp["b0"], inp["w1"], inp["b1"], inp["y"],
                                    inp["mask"], lr=lr, p=p, keep_masks=keeps, dtype=dt)
            for k, v in r.items():
                out[f"{k}_{tag}"] = v
        yield name, out
    for name, seed, n, f, h, c, tau, blocks, outer_lr, decay, inner_lr, wd, step0 in BLOCK_CASES:
        inp = make_inputs(seed, n, f, h, c, "mixed", 0.0)
        inp["theta_triu"] = interior(inp["theta_triu"])
        train_mask, opt_mask, _, _ = split_masks(seed, n)
        philox_seed = 0x5EED0000 + seed
        uniforms = [PH.edge_uniforms(n, philox_seed, step0 + k) for k in range(blocks * (tau + 1))]
        out = dict(inp)
        out.pop("mask")
        out.update(n=n, f=f, h=h, c=c, tau=tau, blocks=blocks, outer_lr=np.float64(outer_lr), lr_decay=np.float64(decay),
                   inner_lr=np.float64(inner_lr), weight_decay=np.float64(wd), step0=np.int64(step0),
                   philox_seed=np.int64(philox_seed), train_mask=train_mask, opt_mask=opt_mask)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            res = L.bilevel_blocks(inp["theta_triu"], uniforms, inp["x"], inp["w0"], inp["b0"], inp["w1"], inp["b1"], inp["y"],
                                   train_mask, opt_mask, tau=tau, blocks=blocks, outer_lr=outer_lr, lr_decay=decay,
                                   inner_lr=inner_lr, weight_decay=wd, dtype=dt)
            for b, rec in enumerate(res):
                out[f"inner_loss{b}_{tag}"] = np.asarray(rec["inner_loss"], dtype=np.float64)
                out[f"inner_acc{b}_{tag}"] = np.asarray(rec["inner_acc"], dtype=np.float64)
                for k, ws in enumerate(rec["weights"]):
                    for wname, w in zip(("w0", "b0", "w1", "b1"), ws):
                        out[f"{wname}_after{b}_{k}_{tag}"] = w
                out[f"hyper_loss{b}_{tag}"] = np.float64(rec["hyper_loss"])
                out[f"hyper_acc{b}_{tag}"] = np.float64(rec["hyper_acc"])
                out[f"lr_after{b}_{tag}"] = np.float64(rec["lr_after"])
                out[f"theta_new{b}_{tag}"] = rec["theta_new"]
                if tag == "f64":
                    out[f"grad_triu{b}_{tag}"] = rec["grad_triu"]
        yield name, out


def main():
    os.makedirs(OUT, exist_ok=True)
    import sys
    if "--next-only" not in sys.argv:
        outer_cases()
    for name, out in next_rows():
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def outer_cases():
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
