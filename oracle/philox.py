"""oracle/philox.py — TEST INFRASTRUCTURE. numpy Philox4x32-10 and the element<->draw mapping.

The reference draws its Bernoulli sample with torch's global generator (src/models/sampling.py:68)
and offers no hook for explicit uniforms, so there is no reference RNG stream to match. The CUDA
path uses a counter-based Philox4x32-10 (Salmon et al., SC'11; the same generator cuRAND/torch-CUDA
use) keyed so that any shard can regenerate the draw of any edge. This file restates that mapping so
the oracle can regenerate the uniforms on the CPU and hand them to the reference as explicit `U`.

Mapping (must match lds-gnn_b200/csrc/lds_philox.cuh):
  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (c0, c1, step & 0xffffffff, ((step >> 32) & 0xffff) << 16 | stream << 12 | sample & 0xfff)
  edges    (stream 0): canonical pair (a, b) = (min(i,j), max(i,j)); c0 = b // 2, c1 = a // 2;
           the 4 output words cover the 2x2 block: word = 2 * (a % 2) + (b % 2)
  dropout  (stream 1 = features X, stream 2 = hidden H1): c0 = col // 4, c1 = row; word = col % 4
  uniform  = (word >> 8) * 2^-24   in [0, 1)   (24-bit, like torch's CPU float path)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_EDGES = 0
STREAM_DROP_X = 1
STREAM_DROP_H = 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Inputs broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _c23(step, stream, sample):
    step = int(step)
    c2 = step & 0xFFFFFFFF
    c3 = (((step >> 32) & 0xFFFF) << 16) | ((int(stream) & 0xF) << 12) | (int(sample) & 0xFFF)
    return c2, c3


def to_uniform(words):
    return ((words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def edge_uniforms_elementwise(n, seed, step, sample=0):
    """U[i, j] for the full n x n matrix, one Philox call per ELEMENT — the literal statement of the mapping above."""
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(n, dtype=np.int64)[None, :]
    a = np.minimum(i, j)
    b = np.maximum(i, j)
    c2, c3 = _c23(step, STREAM_EDGES, sample)
    w = philox4x32_10(b // 2, a // 2, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = (2 * (a % 2) + (b % 2)).astype(np.int64)
    words = np.choose(sel, w)
    return to_uniform(words)


def edge_uniforms(n, seed, step, sample=0):
    """U[i, j] for the full n x n matrix, symmetric by construction (U[i,j] == U[j,i]). One Philox call per 2x2 BLOCK
    (a quarter of the work of `edge_uniforms_elementwise`, same values — tests/test_oracle.py checks the equality)."""
    nb = (n + 1) // 2
    P = np.arange(nb, dtype=np.int64)[:, None]
    Q = np.arange(nb, dtype=np.int64)[None, :]
    c2, c3 = _c23(step, STREAM_EDGES, sample)
    w = philox4x32_10(np.maximum(P, Q), np.minimum(P, Q), c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    upper, lower = P < Q, P > Q
    words = np.empty((2 * nb, 2 * nb), dtype=np.uint32)
    for di in range(2):
        for dj in range(2):
            diag = w[2 * min(di, dj) + max(di, dj)]          # canonical pair (min, max) inside a diagonal block
            words[di::2, dj::2] = np.where(upper, w[2 * di + dj], np.where(lower, w[2 * dj + di], diag))
    return to_uniform(words[:n, :n])


def edge_uniform(seed, step, sample, i, j):
    a, b = (i, j) if i <= j else (j, i)
    c2, c3 = _c23(step, STREAM_EDGES, sample)
    w = philox4x32_10(b // 2, a // 2, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return float(to_uniform(np.asarray(w[2 * (a % 2) + (b % 2)]).reshape(1))[0])


def dropout_uniforms(rows, cols, seed, step, stream, sample=0):
    """One Philox call per (row, 4-column group); word = col % 4."""
    r = np.arange(rows, dtype=np.int64)[:, None]
    q = np.arange((cols + 3) // 4, dtype=np.int64)[None, :]
    c2, c3 = _c23(step, stream, sample)
    w = philox4x32_10(q + 0 * r, r + 0 * q, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=2).reshape(rows, -1)[:, :cols]
    return to_uniform(words)


def dropout_keep_mask(rows, cols, p, seed, step, stream, sample=0):
    """keep[i, j] = u < (1 - p) evaluated in fp32 (F.dropout keeps with probability 1 - p)."""
    if p <= 0.0:
        return np.ones((rows, cols), dtype=bool)
    u = dropout_uniforms(rows, cols, seed, step, stream, sample)
    return u < np.float32(1.0 - p)
