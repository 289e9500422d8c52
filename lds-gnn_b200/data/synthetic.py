"""Planted-partition datasets of Cora / Citeseer shape: uniform labels, an SBM graph with the
real datasets' expected degree and 0.8 homophily, bag-of-words features whose word rates depend on the class
(each class prefers F/C columns at 5x the base rate), row-normalised like `NormalizeFeatures`
(reference src/data/dataloader.py:100-101), Planetoid-style splits (20 train per class, 500 val, 1000 test)."""
import numpy as np
import torch

from ..utils.graph import DenseData

SHAPES = {
    #            nodes  feats classes hidden  density  degree
    "cora":     (2708,  1433, 7,      16,     0.0127,  3.9),
    "citeseer": (3327,  3703, 6,      16,     0.0086,  2.7),
    "tiny":     (300,   64,   4,      16,     0.05,    4.0),
    "n20k":     (20000, 512,  7,      64,     0.02,    4.0),
    "n65k":     (65536, 512,  7,      64,     0.02,    4.0),
}


def make_dataset(name="cora", seed=0, homophily=0.8, n=None, f=None, c=None, dense_adj=True):
    """Returns a CPU `DenseData` (x, y, dense_adj, edge_index, masks, num_classes, name)."""
    nn_, ff, cc, _, rho, degree = SHAPES[name]
    n, f, c = n or nn_, f or ff, c or cc
    rng = np.random.default_rng(seed)
    y = rng.integers(0, c, n)
    # features: class-dependent Bernoulli word rates
    base = rho * f / (f + 4.0 * (f / c))                  # keeps the overall density at rho
    owner = (np.arange(f) * c) // f                        # column -> preferring class
    x = np.zeros((n, f), dtype=np.float32)
    for k in range(c):
        rows = np.nonzero(y == k)[0]
        rate = np.where(owner == k, 5.0 * base, base)
        x[rows] = rng.random((len(rows), f)) < rate
    empty = x.sum(1) == 0
    x[empty, rng.integers(0, f, int(empty.sum()))] = 1.0
    x /= x.sum(1, keepdims=True)
    # SBM edges: each node proposes Poisson(degree/2) partners, same class with probability `homophily`
    by_class = [np.nonzero(y == k)[0] for k in range(c)]
    counts = rng.poisson(degree / 2.0, n)
    src = np.repeat(np.arange(n), counts)
    same = rng.random(len(src)) < homophily
    dst = np.empty(len(src), dtype=np.int64)
    for k in range(c):
        sel = np.nonzero(same & (y[src] == k))[0]
        dst[sel] = rng.choice(by_class[k], len(sel))
    other = np.nonzero(~same)[0]
    dst[other] = rng.integers(0, n, len(other))
    keep = src != dst
    src, dst = src[keep], dst[keep]
    edge_index = torch.as_tensor(np.stack([np.concatenate([src, dst]), np.concatenate([dst, src])]))
    adj = None
    if dense_adj:
        adj = torch.zeros((n, n), dtype=torch.float32)
        adj[edge_index[0], edge_index[1]] = 1.0
    # Planetoid-style split
    order = rng.permutation(n)
    train = np.concatenate([order[y[order] == k][:20] for k in range(c)])
    rest = np.setdiff1d(order, train, assume_unique=False)
    rest = rest[rng.permutation(len(rest))]
    n_val, n_test = min(500, len(rest) // 3), min(1000, len(rest) // 2)
    masks = []
    for idx in (train, rest[:n_val], rest[n_val:n_val + n_test]):
        m = torch.zeros(n, dtype=torch.bool)
        m[torch.as_tensor(idx)] = True
        masks.append(m)
    return DenseData(x=torch.as_tensor(x), y=torch.as_tensor(y, dtype=torch.int64), dense_adj=adj, edge_index=edge_index,
                     train_mask=masks[0], val_mask=masks[1], test_mask=masks[2], num_classes=c, name=f"synthetic-{name}")
