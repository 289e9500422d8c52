"""oracle/restatement.py — TEST INFRASTRUCTURE. numpy restatement of the LDS outer-step path.

Every function cites the reference lines it restates (paths relative to /root/reference). Random
draws are explicit inputs (`U`, dropout keep-masks): the reference's `Bernoulli(probs).sample()`
(src/models/sampling.py:68) is `u < p` with u ~ U[0,1) on the CPU, so feeding the same `U` to the
live reference (oracle/live_reference.py patches torch.bernoulli) and to this file must give the
same mask bit for bit. Arithmetic after the mask runs in float64 unless `dtype` says otherwise.

The backward pass is the closed form of SURVEY.md App. A.2; `tests/test_oracle_vs_reference.py`
checks it against the reference's own autograd, and `make_golden.py` stores the reference's values.
"""
from math import sqrt

import numpy as np


# --------------------------------------------------------------------------------------
# a1/a2: theta storage  (src/models/graph.py:47-67, src/utils/graph.py:41-45, 166-192)
# --------------------------------------------------------------------------------------
def num_nodes_from_triu_shape(n_triu_values):
    """src/utils/graph.py:184-192 — same (slightly odd) formula, kept verbatim in meaning."""
    return int(0.5 * sqrt((8 * n_triu_values + 1) - 1))


def triu_index(i, j, n):
    """Row-major upper triangle incl. diagonal (torch.triu_indices order, src/utils/graph.py:44)."""
    return i * n - (i * (i - 1)) // 2 + (j - i)


def get_triu_values(adj):
    """src/utils/graph.py:41-45."""
    n = adj.shape[0]
    iu = np.triu_indices(n)
    return adj[iu]


def to_undirected(adj, from_triu_only=False):
    """src/utils/graph.py:27-38."""
    if not from_triu_only:
        return np.maximum(adj, adj.T)
    triu = np.triu(adj, 1)
    return triu + triu.T + np.diag(np.diag(adj))


def theta_full_from_triu(theta_triu):
    """src/utils/graph.py:166-181: scatter, mirror the strict upper triangle, keep diag, clamp[0,1]."""
    t = theta_triu.shape[0]
    n = num_nodes_from_triu_shape(t)
    adj = np.zeros((n, n), dtype=theta_triu.dtype)
    adj[np.triu_indices(n)] = theta_triu
    adj = to_undirected(adj, from_triu_only=True)
    return np.clip(adj, 0.0, 1.0)


# --------------------------------------------------------------------------------------
# a5: sampling  (src/models/sampling.py:47-85)
# --------------------------------------------------------------------------------------
def sample_graph(theta_full, U, undirected=True):
    """Bernoulli draw over the full N x N (sampling.py:68), then the upper-triangle draw wins and the
    sampled diagonal is kept (sampling.py:76 -> utils/graph.py:35-37). The straight-through estimator
    (sampling.py:82-85) leaves the VALUE equal to the sample, so it does not appear here."""
    s = (U.astype(np.float32) < theta_full.astype(np.float32)).astype(np.float32)
    return to_undirected(s, from_triu_only=True) if undirected else s


# --------------------------------------------------------------------------------------
# a6/a7: self loops + normalisation  (src/utils/graph.py:123-153)
# --------------------------------------------------------------------------------------
def add_self_loops(adj):
    out = adj.copy()
    np.fill_diagonal(out, 1.0)
    return out


def normalize_adjacency_matrix(adj, dtype=np.float64):
    """Returns (A_hat, A_tilde, deg, r). deg is the ROW sum (utils/graph.py:148)."""
    a_tilde = add_self_loops(adj).astype(dtype)
    deg = a_tilde.sum(axis=1)
    r = 1.0 / np.sqrt(deg)
    a_hat = r[:, None] * a_tilde * r[None, :]
    return a_hat, a_tilde, deg, r


# --------------------------------------------------------------------------------------
# a8: dense GCN  (src/models/gcn.py:23-34, src/models/layers.py:42-44)
# --------------------------------------------------------------------------------------
def log_softmax(z):
    m = z.max(axis=1, keepdims=True)
    e = z - m
    return e - np.log(np.exp(e).sum(axis=1, keepdims=True))


def gcn_forward(a_hat, x, w0, b0, w1, b1, p=0.0, keep_x=None, keep_h=None):
    """H1 = relu(A (drop(X) W0^T + b0)); Z2 = A (drop(H1) W1^T + b1); bias BEFORE propagation
    (layers.py:43-44). Dropout scales kept entries by 1/(1-p) (F.dropout, gcn.py:27,29)."""
    scale = 1.0 / (1.0 - p) if p > 0 else 1.0
    xd = x if keep_x is None else x * keep_x * scale
    p1 = xd @ w0.T + b0
    z1 = a_hat @ p1
    h1 = np.maximum(z1, 0.0)
    h1d = h1 if keep_h is None else h1 * keep_h * scale
    p2 = h1d @ w1.T + b1
    z2 = a_hat @ p2
    return dict(xd=xd, p1=p1, z1=z1, h1=h1, h1d=h1d, p2=p2, z2=z2, logp=log_softmax(z2))


# --------------------------------------------------------------------------------------
# a9: loss + accuracy  (src/trainers/outer.py:65-67, src/utils/evaluation.py:15-22)
# --------------------------------------------------------------------------------------
def nll_and_accuracy(logp, y, mask):
    idx = np.nonzero(mask)[0]
    loss = -logp[idx, y[idx]].mean()
    acc = (logp[idx].argmax(axis=1) == y[idx]).astype(np.float32).mean()
    return float(loss), float(acc)


# --------------------------------------------------------------------------------------
# a10: backward to theta, closed form (reference = autograd, outer.py:77)
# --------------------------------------------------------------------------------------
def backward_theta(fwd, a_hat, deg, r, y, mask, w1, p=0.0, keep_h=None, theta_triu=None):
    n, c = fwd["z2"].shape
    idx = np.nonzero(mask)[0]
    m = float(len(idx))
    dz2 = np.zeros_like(fwd["z2"])
    sm = np.exp(fwd["logp"][idx])
    sm[np.arange(len(idx)), y[idx]] -= 1.0
    dz2[idx] = sm / m
    dp2 = a_hat @ dz2                      # A_hat symmetric
    dh1d = dp2 @ w1
    scale = 1.0 / (1.0 - p) if p > 0 else 1.0
    dh1 = dh1d if keep_h is None else dh1d * keep_h * scale
    dz1 = dh1 * (fwd["z1"] > 0)
    dp1 = a_hat @ dz1
    # G = dL/dA_hat = dZ1 P1^T + dZ2 P2^T (rank h + C)
    g = dz1 @ fwd["p1"].T + dz2 @ fwd["p2"].T
    rho = (dz1 * fwd["z1"]).sum(1) + (dz2 * fwd["z2"]).sum(1)       # row sums of G * A_hat
    kappa = (fwd["p1"] * dp1).sum(1) + (fwd["p2"] * dp2).sum(1)     # column sums of G * A_hat
    cvec = -(rho + kappa) / (2.0 * deg)
    d_atilde = g * r[:, None] * r[None, :] + cvec[:, None]
    np.fill_diagonal(d_atilde, 0.0)                                   # fill_diagonal_ backward
    d_full = d_atilde                                                 # STE: dA/dtheta_full = I
    sym = np.triu(d_full, 1) + np.triu(d_full.T, 1)                   # triu mirror backward; diag -> 0
    iu = np.triu_indices(n)
    d_triu = sym[iu]
    if theta_triu is not None:                                        # clamp backward (closed interval)
        d_triu = d_triu * ((theta_triu >= 0.0) & (theta_triu <= 1.0))
    return dict(dz2=dz2, dp2=dp2, dz1=dz1, dp1=dp1, rho=rho, kappa=kappa, c=cvec,
                d_theta_full=d_full, d_theta_triu=d_triu)


# --------------------------------------------------------------------------------------
# a11: SGD step + StepLR + projection  (src/models/factory.py:66-69, outer.py:78-83, graph.py:16-20)
# --------------------------------------------------------------------------------------
def sgd_project(theta_triu, d_triu, lr):
    return np.clip(theta_triu - lr * d_triu, 0.0, 1.0)


def adam_project(theta_triu, d_triu, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (the north-star's optional optimiser; factory.py:151-180 has the
    selection key for the GAE model). t is the 1-based step count."""
    m = beta1 * m + (1 - beta1) * d_triu
    v = beta2 * v + (1 - beta2) * d_triu * d_triu
    mhat = m / (1 - beta1 ** t)
    vhat = v / (1 - beta2 ** t)
    return np.clip(theta_triu - lr * mhat / (np.sqrt(vhat) + eps), 0.0, 1.0), m, v


# --------------------------------------------------------------------------------------
# a14: statistics  (src/models/graph.py:69-78)
# --------------------------------------------------------------------------------------
def statistics(theta_triu):
    full = theta_full_from_triu(theta_triu)
    n = full.shape[0]
    return {
        "expected_num_edges": float(full.sum()),
        "percentage_edges_expected": float(full.sum()) / (n * n),
        "mean_prob": float(theta_triu.mean()),
        "min_prob": float(theta_triu.min()),
        "max_prob": float(theta_triu.max()),
    }


# --------------------------------------------------------------------------------------
# a12: one direct outer step  (src/trainers/outer.py:57-87)
# --------------------------------------------------------------------------------------
def outer_step(theta_triu, U, x, w0, b0, w1, b1, y, mask, lr, p=0.0, keep_x=None, keep_h=None,
               dtype=np.float64):
    """theta_triu fp32 (T,), U fp32 (N,N) uniforms in [0,1). Returns every intermediate the CUDA path
    is compared on. Mask/degree decisions happen in fp32 exactly like the reference; the rest in `dtype`."""
    theta_triu32 = theta_triu.astype(np.float32)
    theta_full = theta_full_from_triu(theta_triu32)
    sample = sample_graph(theta_full, U)                       # {0,1}, sampled diagonal kept
    a_hat, a_tilde, deg, r = normalize_adjacency_matrix(sample, dtype=dtype)
    cast = lambda t: None if t is None else np.asarray(t, dtype=dtype)
    fwd = gcn_forward(a_hat, cast(x), cast(w0), cast(b0), cast(w1), cast(b1), p=p,
                      keep_x=cast(keep_x), keep_h=cast(keep_h))
    loss, acc = nll_and_accuracy(fwd["logp"], y, mask)
    bwd = backward_theta(fwd, a_hat, deg, r, y, mask, cast(w1), p=p, keep_h=cast(keep_h),
                         theta_triu=theta_triu32)
    theta_new = sgd_project(theta_triu32.astype(dtype), bwd["d_theta_triu"], lr)
    out = dict(theta_full=theta_full, sample=sample, a_tilde=a_tilde, deg=deg, r=r, a_hat=a_hat,
               loss=loss, acc=acc, theta_new=theta_new)
    out.update(fwd)
    out.update(bwd)
    return out
