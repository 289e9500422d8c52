"""Host-side wrappers of the C ABI on torch CUDA tensors (device memory, streams: plumbing only).

Every function here enqueues hand-written sm_100a kernels from liblds_b200.so on torch's current
stream. Nothing falls back to PyTorch arithmetic: without the library or a B200 these raise.
"""
import ctypes

import torch

from . import _lib

BF16 = torch.bfloat16


def _stream():
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def upload_async(dst, src_pinned):
    """dst (CUDA tensor) <- src_pinned (pinned host tensor, same dtype / numel) on the library's copy stream; work enqueued on
    the current stream afterwards is ordered behind the copy (lds_upload_async). `dst` must not be in use by enqueued work."""
    if not dst.is_cuda or src_pinned.is_cuda or dst.dtype != src_pinned.dtype or dst.numel() != src_pinned.numel() \
            or not dst.is_contiguous() or not src_pinned.is_contiguous():
        raise TypeError("upload_async: contiguous CUDA destination and pinned host source of the same dtype and size")
    lib = _lib.load()
    _lib.check(lib.lds_upload_async(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src_pinned.data_ptr()),
                                    dst.numel() * dst.element_size(), _stream()), "lds_upload_async")


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA float32 tensor (got {t.dtype} on {t.device})")
    return t


def padded_ld(n):
    return int(_lib.load().lds_padded_ld(int(n)))


def num_nodes_from_triu_shape(t):
    """Same formula as the reference (src/utils/graph.py:184-192)."""
    from math import sqrt
    return int(0.5 * sqrt((8 * t + 1) - 1))


# ----------------------------------------------------------------------------- theta layouts (a1/a2)
def theta_triu_to_full(triu, clamp=False, out=None):
    """(T,) upper-triangle vector -> symmetric [n, ld] matrix (src/utils/graph.py:166-181)."""
    _lib.require_device()
    _f32(triu, "triu")
    n = num_nodes_from_triu_shape(triu.numel())
    if n * (n + 1) // 2 != triu.numel():
        raise ValueError(f"{triu.numel()} is not a triangular number")
    ld = padded_ld(n)
    full = out if out is not None else torch.empty((n, ld), dtype=torch.float32, device=triu.device)
    _lib.check(_lib.load().lds_theta_triu_to_full(_ptr(triu.contiguous()), _ptr(full), ld, n, int(bool(clamp)), _stream()),
               "lds_theta_triu_to_full")
    return full


def theta_full_to_triu(full, n, sym_sum=False, ld=None):
    _lib.require_device()
    _f32(full, "full")
    ld = full.stride(0) if ld is None else ld
    triu = torch.empty(n * (n + 1) // 2, dtype=torch.float32, device=full.device)
    _lib.check(_lib.load().lds_theta_full_to_triu(_ptr(full), ld, _ptr(triu), n, int(bool(sym_sum)), _stream()),
               "lds_theta_full_to_triu")
    return triu


def theta_clamp_(full, n):
    """project_parameters (src/models/graph.py:16-20, 63-64) on the full matrix, in place."""
    _lib.require_device()
    _lib.check(_lib.load().lds_theta_clamp(_ptr(_f32(full, "full")), full.stride(0), n, _stream()), "lds_theta_clamp")
    return full


def theta_stats(full, n):
    """Device double[4]: sum(clamp(theta_full)), sum/min/max over the upper triangle (src/models/graph.py:69-78)."""
    _lib.require_device()
    out = torch.empty(4, dtype=torch.float64, device=full.device)
    _lib.check(_lib.load().lds_theta_stats(_ptr(_f32(full, "full")), full.stride(0), n, _ptr(out), _stream()), "lds_theta_stats")
    return out


# ----------------------------------------------------------------------------- K1
def k1_sample_normalize(theta_full, n, seed, step, sample=0, u=None, want_adj=True, want_sample=False,
                        row0=0, rows=None, step_base=None):
    """Fused sample / mirror / self-loop / degree / rsqrt pass. Returns (A_tilde bf16 [rows, ld] or None,
    raw sample fp32 [rows, n] or None, deg [rows], rsqrt [rows]). With `step_base` (int64 device tensor [1]) the kernel
    reads the Philox step from device memory as step_base[0] + step (CUDA-graph replays draw fresh graphs)."""
    _lib.require_device()
    _f32(theta_full, "theta_full")
    rows = n - row0 if rows is None else rows
    dev = theta_full.device
    ld = padded_ld(n)
    adj = torch.empty((rows, ld), dtype=BF16, device=dev) if want_adj else None
    smp = torch.empty((rows, n), dtype=torch.float32, device=dev) if want_sample else None
    deg = torch.empty(rows, dtype=torch.float32, device=dev)
    rs = torch.empty(rows, dtype=torch.float32, device=dev)
    flags = 0
    if step_base is not None:
        if u is not None or want_sample or step_base.dtype != torch.int64 or not step_base.is_cuda:
            raise TypeError("step_base: int64 CUDA tensor, without explicit uniforms / dense sample output")
        _lib.check(_lib.load().lds_k1_sample_normalize_dstep(
            _ptr(theta_full), theta_full.stride(0), n, row0, rows, int(seed), _ptr(step_base), int(step), int(sample),
            _ptr(adj), ld, _ptr(deg), _ptr(rs), _stream()), "lds_k1_sample_normalize_dstep")
        return adj, smp, deg, rs
    if u is not None:
        _f32(u, "u")
        flags |= _lib.K1_EXPLICIT_U
    _lib.check(_lib.load().lds_k1_sample_normalize(
        _ptr(theta_full), theta_full.stride(0), n, row0, rows, int(seed), int(step), int(sample),
        _ptr(u), 0 if u is None else u.stride(0), _ptr(adj), ld, _ptr(smp), n, _ptr(deg), _ptr(rs), flags, _stream()),
        "lds_k1_sample_normalize")
    return adj, smp, deg, rs


def packed_adj_bytes(n, rows=None):
    return int(_lib.load().lds_packed_adj_bytes(int(n), int(n if rows is None else rows)))


def k1_sample_packed(theta_full, n, seed, step, sample=0, u=None, row0=0, rows=None, bits=None):
    """K1 on the bit-packed A_tilde (tile-symmetric: one Philox block / theta read per unordered tile pair of the shard).
    Returns (bits uint8 [packed_adj_bytes], deg [rows], rsqrt [rows]); `unpack_adj` gives the {0,1} matrix."""
    _lib.require_device()
    _f32(theta_full, "theta_full")
    rows = n - row0 if rows is None else rows
    dev = theta_full.device
    nbytes = packed_adj_bytes(n, rows)
    if bits is None:
        bits = torch.zeros(nbytes, dtype=torch.uint8, device=dev)          # zero-filled once: rows / columns beyond the matrix stay zero
    cnt = torch.zeros(rows + 1, dtype=torch.int32, device=dev)
    deg = torch.empty(rows, dtype=torch.float32, device=dev)
    rs = torch.empty(rows, dtype=torch.float32, device=dev)
    if u is not None:
        _f32(u, "u")
    _lib.check(_lib.load().lds_k1_sample_packed(
        _ptr(theta_full), theta_full.stride(0), n, row0, rows, int(seed), int(step), int(sample), _ptr(u), 0 if u is None else u.stride(0),
        _ptr(bits), _ptr(cnt), _ptr(deg), _ptr(rs), _stream()), "lds_k1_sample_packed")
    return bits, deg, rs


def unpack_adj(bits, n, rows=None, dtype=BF16):
    """Bit-packed A_tilde (layout: include/lds_b200.h, lds_k1_sample_packed) -> dense [rows, ld] matrix of 0/1. Tests and
    debugging only: plain torch ops."""
    rows = n if rows is None else rows
    sp, kb = (rows + 255) // 256, (n + 63) // 64
    words = bits[:sp * kb * 2048].view(torch.int32).view(sp, kb, 256, 2)
    q = torch.arange(32, device=bits.device, dtype=torch.int32)
    cells = ((words.unsqueeze(-1) >> q) & 1)                               # [sp, kb, 256, 2 (even / odd), 32 (pair)]
    cells = cells.permute(0, 2, 1, 4, 3).reshape(sp * 256, kb * 64)        # column = 64 kb + 2 pair + parity
    return cells[:rows].to(dtype)


# ----------------------------------------------------------------------------- K2
_ws_cache = {}


def _workspace(nbytes, device, tag, zero=False):
    key = (tag, str(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = (torch.zeros if zero else torch.empty)(nbytes + 1024, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def k2_propagate(adj, n, p, scale_in=None, scale_out=None, flags=0, out=None):
    """z = scale_out * (A_tilde[rows, n] @ (scale_in * p[n, w])) on the tcgen05 tensor cores."""
    _lib.require_device()
    if adj.dtype != BF16 or not adj.is_cuda:
        raise TypeError("adj must be a CUDA bfloat16 tensor")
    _f32(p, "p")
    rows, width = adj.shape[0], p.shape[1]
    if p.shape[0] != n:
        raise ValueError(f"p has {p.shape[0]} rows, expected {n}")
    p = p if p.stride(1) == 1 else p.contiguous()
    z = out if out is not None else torch.empty((rows, width), dtype=torch.float32, device=adj.device)
    need = int(_lib.load().lds_k2_workspace_bytes(n, rows, width))
    if need < 0:
        raise ValueError(f"unsupported propagate shape n={n} rows={rows} width={width}")
    ws = _workspace(need, adj.device, "k2")
    _lib.check(_lib.load().lds_k2_propagate(
        _ptr(adj), adj.stride(0), n, rows, _ptr(p), p.stride(0), width, _ptr(scale_in), _ptr(scale_out),
        _ptr(z), z.stride(0), _ptr(ws), need, int(flags), _stream()), "lds_k2_propagate")
    return z


def k2_propagate_packed(bits, n, rows, p, scale_in=None, scale_out=None, flags=0, out=None):
    """z = scale_out * (A_tilde @ (scale_in * p)) with A_tilde given as bits: expanded on chip, tensor-bound instead of HBM-bound."""
    _lib.require_device()
    _f32(p, "p")
    width = p.shape[1]
    if p.shape[0] != n:
        raise ValueError(f"p has {p.shape[0]} rows, expected {n}")
    p = p if p.stride(1) == 1 else p.contiguous()
    z = out if out is not None else torch.empty((rows, width), dtype=torch.float32, device=p.device)
    need = int(_lib.load().lds_k2_packed_workspace_bytes(n, rows, width))
    if need < 0:
        raise ValueError(f"unsupported propagate shape n={n} rows={rows} width={width}")
    ws = _workspace(need, p.device, "k2p")
    _lib.check(_lib.load().lds_k2_propagate_packed(
        _ptr(bits), n, rows, _ptr(p), p.stride(0), width, _ptr(scale_in), _ptr(scale_out), _ptr(z), z.stride(0), _ptr(ws), need,
        int(flags), _stream()), "lds_k2_propagate_packed")
    return z


# ----------------------------------------------------------------------------- sparse feature products
def spmm_csr(ptr, idx, val, perm, rows, b):
    """y[rows, w] = S @ b for S in CSR (ptr, idx, val[perm]); b [cols, w] fp32, any strides (src/models/layers.py:43)."""
    _lib.require_device()
    _f32(b, "b")
    _f32(val, "val")
    if ptr.dtype != torch.int32 or idx.dtype != torch.int32 or (perm is not None and perm.dtype != torch.int32):
        raise TypeError("CSR index arrays must be int32")
    w = b.shape[1]
    y = torch.empty((rows, w), dtype=torch.float32, device=b.device)
    _lib.check(_lib.load().lds_spmm_csr(_ptr(ptr), _ptr(idx), _ptr(val), _ptr(perm), int(rows), _ptr(b), b.stride(0), b.stride(1),
                                        int(w), _ptr(y), y.stride(0), _stream()), "lds_spmm_csr")
    return y


# ----------------------------------------------------------------------------- skinny dense products
def row_linear(x, w, bias=None):
    """y[n, m] = x[n, k] @ w[m, k]^T (+ bias); w may be any strided view (e.g. a transpose). k, m <= 128."""
    _lib.require_device()
    _f32(x, "x")
    _f32(w, "w")
    x = x if x.stride(1) == 1 else x.contiguous()
    n, k = x.shape
    m = w.shape[0]
    if w.shape[1] != k:
        raise ValueError(f"row_linear: x is {tuple(x.shape)}, w is {tuple(w.shape)}")
    y = torch.empty((n, m), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().lds_row_linear(_ptr(x), x.stride(0), int(k), _ptr(w), w.stride(0), w.stride(1), int(m),
                                          _ptr(None if bias is None else _f32(bias, "bias").contiguous()), _ptr(y), y.stride(0), int(n), _stream()),
               "lds_row_linear")
    return y


def gram_tn(a, b):
    """out[i, j] = sum_n a[n, i] * b[n, j] (deterministic). a [n, p], b [n, q], p, q <= 128."""
    _lib.require_device()
    _f32(a, "a")
    _f32(b, "b")
    a = a if a.stride(1) == 1 else a.contiguous()
    b = b if b.stride(1) == 1 else b.contiguous()
    n, p = a.shape
    q = b.shape[1]
    if b.shape[0] != n:
        raise ValueError(f"gram_tn: a is {tuple(a.shape)}, b is {tuple(b.shape)}")
    out = torch.empty((p, q), dtype=torch.float32, device=a.device)
    need = int(_lib.load().lds_gram_tn_workspace_bytes(int(p), int(q)))
    if need < 0:
        raise ValueError(f"gram_tn: widths {p}, {q} outside 1..128")
    ws = _workspace(max(need, 1 << 20), a.device, "gram", zero=True)[:need]      # one zero-initialised arena for every width
    _lib.check(_lib.load().lds_gram_tn(_ptr(a), a.stride(0), int(p), _ptr(b), b.stride(0), int(q), int(n), _ptr(out), out.stride(0),
                                       _ptr(ws), need, _stream()), "lds_gram_tn")
    return out


def masked_nll_forward(z, rows, y):
    """(loss, acc) device tensor [2] of F.nll_loss(log_softmax(z)[rows], y[rows]) on the logits z [n, c] (csrc/lds_rowops.cu)."""
    _lib.require_device()
    z = _f32(z, "z")
    out = torch.empty(2, dtype=torch.float32, device=z.device)
    _lib.check(_lib.load().lds_masked_nll_forward(_ptr(z), z.stride(0), int(z.shape[1]), _ptr(rows), _ptr(y), int(rows.numel()), _ptr(out), _stream()),
               "lds_masked_nll_forward")
    return out


def masked_nll_grad(z, slot, y, m, grad_loss):
    """d loss / d z, dense [n, c] (zero rows outside the mask), scaled by the device scalar grad_loss."""
    _lib.require_device()
    dz = torch.empty((z.shape[0], z.shape[1]), dtype=torch.float32, device=z.device)
    _lib.check(_lib.load().lds_masked_nll_grad(_ptr(z), z.stride(0), int(z.shape[1]), _ptr(slot), _ptr(y), int(z.shape[0]), int(m),
                                               _ptr(grad_loss), _ptr(dz), dz.stride(0), _stream()), "lds_masked_nll_grad")
    return dz


def masked_nll_grad_grad(u, z, slot, y, m, grad_loss, want_z=True, want_g=False):
    """Backward of masked_nll_grad given u = upstream of dz: (gradient w.r.t. z or None, per-row terms of the gradient w.r.t. grad_loss or None)."""
    _lib.require_device()
    n, c = z.shape
    out_z = torch.empty((n, c), dtype=torch.float32, device=z.device) if want_z else None
    out_g = torch.empty(n, dtype=torch.float32, device=z.device) if want_g else None
    _lib.check(_lib.load().lds_masked_nll_grad_grad(_ptr(u), u.stride(0), _ptr(z), z.stride(0), int(c), _ptr(slot), _ptr(y), int(n), int(m),
                                                    _ptr(grad_loss), None if out_z is None else _ptr(out_z), c,
                                                    None if out_g is None else _ptr(out_g), _stream()), "lds_masked_nll_grad_grad")
    return out_z, out_g


def row_dot2(a1, b1, a2, b2, r):
    """out[i] = (<a1_i, b1_i> + <a2_i, b2_i>) / r[i] for contiguous fp32 [n, w] operands."""
    _lib.require_device()
    n, w = a1.shape
    out = torch.empty(n, dtype=torch.float32, device=a1.device)
    _lib.check(_lib.load().lds_row_dot2(_ptr(a1), _ptr(b1), _ptr(a2), _ptr(b2), a1.stride(0), int(w), _ptr(r), int(n), _ptr(out), _stream()),
               "lds_row_dot2")
    return out


def _scalar_or_ptr(x):
    """(by-value float, device pointer or None) of a hyper-parameter given as a python float or a 1-element device tensor."""
    if isinstance(x, torch.Tensor):
        return 0.0, x.data_ptr()
    return float(x), None


def adam_step(p, m, v, g, wd, b1, b2, eps, step_size, root_scale):
    """One out-of-place Adam step over flat fp32 vectors (csrc/lds_adam.cu). step_size / root_scale: floats or 1-element
    device tensors (both of the same kind); wd: a float or a per-element fp32 vector. Returns (p', m', v')."""
    _lib.require_device()
    outs = [torch.empty_like(p) for _ in range(3)]
    ss, ssp = _scalar_or_ptr(step_size)
    rs, rsp = _scalar_or_ptr(root_scale)
    wdv, wdp = _scalar_or_ptr(wd)                             # a per-element vector (parameter groups) or one float
    _lib.check(_lib.load().lds_adam_step(_ptr(p), _ptr(m), _ptr(v), _ptr(g), p.numel(), wdv, wdp, float(b1), float(b2), float(eps),
                                         ss, rs, ssp, rsp, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _stream()), "lds_adam_step")
    return outs


def adam_step_backward(gp, gm, gv, p, m, v, g, wd, b1, b2, eps, step_size, root_scale, need):
    """Vector-Jacobian product of adam_step: upstream (gp, gm, gv) (gm / gv may be None) -> (dp, dm, dv, dg), None where `need` is False."""
    _lib.require_device()
    outs = [torch.empty_like(p) if flag else None for flag in need]
    ss, ssp = _scalar_or_ptr(step_size)
    rs, rsp = _scalar_or_ptr(root_scale)
    wdv, wdp = _scalar_or_ptr(wd)
    opt = lambda t: None if t is None else _ptr(t)
    _lib.check(_lib.load().lds_adam_step_backward(opt(gp), opt(gm), opt(gv), _ptr(p), _ptr(m), _ptr(v), _ptr(g), p.numel(),
                                                  wdv, wdp, float(b1), float(b2), float(eps), ss, rs, ssp, rsp,
                                                  opt(outs[0]), opt(outs[1]), opt(outs[2]), opt(outs[3]), _stream()), "lds_adam_step_backward")
    return outs


# ----------------------------------------------------------------------------- K3 + K4
def k3k4_theta_update_(theta_full, n, fa, fb, cvec, lr, d=None, opt_kind=_lib.OPT_SGD, adam_m=None, adam_v=None,
                       betas=(0.9, 0.999), eps=1e-8, t=1, row0=0, rows=None):
    _lib.require_device()
    rows = n - row0 if rows is None else rows
    d = fa.shape[1] if d is None else d
    _lib.check(_lib.load().lds_k3k4_theta_update(
        _ptr(_f32(theta_full, "theta_full")), theta_full.stride(0), n, row0, rows, _ptr(_f32(fa, "fa")), _ptr(_f32(fb, "fb")),
        fa.stride(0), d, _ptr(_f32(cvec, "cvec")), float(lr), int(opt_kind), _ptr(adam_m), _ptr(adam_v),
        float(betas[0]), float(betas[1]), float(eps), int(t), None, 0, 0, _stream()), "lds_k3k4_theta_update")
    return theta_full


def k3k4_theta_update_tc_(theta_full, n, fa, fb, cvec, lr, d=None, row0=0, rows=None):
    """SGD update on the tcgen05 kernel (bf16 hi/lo-split factors, fp32 accumulation, TMA-streamed theta)."""
    _lib.require_device()
    rows = n - row0 if rows is None else rows
    d = fa.shape[1] if d is None else d
    need = int(_lib.load().lds_k3_workspace_bytes(n, d))
    ws = _workspace(need, theta_full.device, "k3")
    _lib.check(_lib.load().lds_k3k4_theta_update_tc(
        _ptr(_f32(theta_full, "theta_full")), theta_full.stride(0), n, row0, rows, _ptr(_f32(fa, "fa")), _ptr(_f32(fb, "fb")),
        fa.stride(0), d, _ptr(_f32(cvec, "cvec")), float(lr), _ptr(ws), need, _stream()), "lds_k3k4_theta_update_tc")
    return theta_full


def k3_dense_grad(n, fa, fb, cvec, d=None, out=None, accumulate=False, row0=0, rows=None):
    """grad[i][j] (+)= fa_i . fb_j + c_i (0 on the diagonal): dL/d(sample) of one propagate (composable path)."""
    _lib.require_device()
    rows = n - row0 if rows is None else rows
    d = fa.shape[1] if d is None else d
    g = out if out is not None else torch.empty((rows, n), dtype=torch.float32, device=fa.device)
    flags = _lib.K3_DENSE_GRAD | (_lib.K3_ACCUMULATE if accumulate else 0)
    _lib.check(_lib.load().lds_k3k4_theta_update(
        None, 0, n, row0, rows, _ptr(_f32(fa, "fa")), _ptr(_f32(fb, "fb")), fa.stride(0), d, _ptr(_f32(cvec, "cvec")),
        0.0, 0, None, None, 0.0, 0.0, 0.0, 1, _ptr(g), g.stride(0), flags, _stream()), "lds_k3k4_theta_update(dense)")
    return g


# ----------------------------------------------------------------------------- fused direct outer step
class OuterStep:
    """Resident state + workspace of the fused outer step (lds_outer_step). One instance per (theta, dataset)."""

    BUFFERS = {"adj": 0, "deg": 1, "rsqrt": 2, "p1": 3, "z1": 4, "p2": 5, "z2": 6, "dz2": 7, "dp2": 8, "dz1": 9,
               "dp1": 10, "fa": 11, "fb": 12, "cvec": 13, "operand": 14, "fpack": 15}
    TRANSPOSED = ("p1", "z1", "p2", "z2", "dz2", "dp2", "dz1", "dp1")     # stored [width][state_ld] (coalesced row epilogues)

    SPARSE_DENSITY = 0.25          # below this share of non-zeros the feature GEMM runs from a CSR copy of x

    def __init__(self, n, x, y, mask, hidden, classes, sparse_features=None, row0=0, rows=None, mask_count=None):
        """x, y, mask hold this engine's rows: all n rows, or — for a row-block shard — global rows [row0, row0+rows)
        (then `mask_count` is the GLOBAL number of masked rows and the step is run phase by phase, see sharded.py)."""
        _lib.require_device()
        self.lib = _lib.load()
        self.n, self.f, self.h, self.c = int(n), int(x.shape[1]), int(hidden), int(classes)
        self.row0 = int(row0)
        self.rows = self.n if rows is None else int(rows)
        self.sharded = self.rows < self.n
        if x.shape[0] != self.rows:
            raise ValueError(f"x has {x.shape[0]} rows, expected {self.rows}")
        self._mask_count_override = mask_count
        dev = x.device
        self.device = dev
        self.ld_x = (self.f + 3) // 4 * 4
        self.x = torch.zeros((self.rows, self.ld_x), dtype=torch.float32, device=dev)    # zero-padded copy, 16-byte rows
        self.x[:, :self.f] = x
        nnz = int((x != 0).sum().item())
        self.sparse = (nnz < self.SPARSE_DENSITY * x.numel()) if sparse_features is None else bool(sparse_features)
        self.x_crow = self.x_col = self.x_val = None
        if self.sparse:                                     # one-off layout conversion of static data (setup, not hot path)
            csr = x.detach().to(torch.float32).to_sparse_csr()
            self.x_crow = csr.crow_indices().to(torch.int32).contiguous()
            self.x_col = csr.col_indices().to(torch.int32).contiguous()
            self.x_val = csr.values().to(torch.float32).contiguous()
        self.y = y.to(device=dev, dtype=torch.int64).contiguous()
        self.set_mask(mask)
        self.w0 = self.b0 = self.w1 = self.b1 = None         # references to the caller's current GCN weights
        nbytes = int(self.lib.lds_outer_step_shard_workspace_bytes(self.n, self.rows, self.f, self.h, self.c))
        if nbytes < 0:
            raise ValueError(f"unsupported shape n={n} f={self.f} h={hidden} c={classes}")
        self._ws_raw = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)     # zero-filled once: the library keeps its counters re-armed
        off = (-self._ws_raw.data_ptr()) % 1024
        self.ws = self._ws_raw[off:off + nbytes]
        self.ws_bytes = nbytes
        self.scalars = torch.zeros(4, dtype=torch.float32, device=dev)
        self.args = _lib.OuterStepArgs()
        self._args_ref = ctypes.byref(self.args)
        self._f32_factors = False
        self._fpack_multi = None
        self._plan = -1                                      # launch plan of the last step (lds_outer_step_plan)
        self._init_static_args()

    def set_mask(self, mask):
        m = mask.to(device=self.device)
        self.mask = m.to(torch.uint8).contiguous()
        self.mask_count = int(m.sum().item()) if self._mask_count_override is None else int(self._mask_count_override)
        if getattr(self, "args", None) is not None:
            self.args.mask, self.args.mask_count = self.mask.data_ptr(), self.mask_count

    def set_weights(self, w0, b0, w1, b1):
        """Use these GCN (fast) weights for the next steps. No copies: the step's first kernel stages the layer_in
        weight (padded / transposed) inside the workspace; the other three are read in place."""
        def prep(t, shape):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_cuda or tuple(t.shape) != shape:
                raise TypeError(f"GCN weight of shape {tuple(t.shape)} / {t.dtype}: expected CUDA float32 {shape}")
            return t if t.is_contiguous() else t.contiguous()
        self.w0, self.b0 = prep(w0, (self.h, self.f)), prep(b0, (self.h,))
        self.w1, self.b1 = prep(w1, (self.c, self.h)), prep(b1, (self.c,))

    def buffer(self, name):
        """View of an intermediate buffer of the last step (tests / composable path / the sharded exchange), logical shape
        [rows, width]. The row-local state is stored transposed in the workspace, so those views are non-contiguous.
        "fa" / "fb" (fp32 factor rows r*(dZ1|dZ2), r*(P1|P2)) are materialised by the library only for the CUDA-core
        update; otherwise they are rebuilt here from the state they are defined by."""
        n, h, c = self.rows, self.h, self.c
        ldf = int(self.lib.lds_outer_step_factor_ld(h, c))
        if name in ("fa", "fb") and not self._f32_factors:
            rs = self.buffer("rsqrt")[:, None]
            parts = (self.buffer("dz1"), self.buffer("dz2")) if name == "fa" else (self.buffer("p1"), self.buffer("p2"))
            out = torch.zeros((n, ldf), dtype=torch.float32, device=self.device)
            out[:, :h] = rs * parts[0]
            out[:, h:h + c] = rs * parts[1]
            return out
        which = self.BUFFERS[name]
        ptr = self.lib.lds_outer_step_shard_buffer(_ptr(self.ws), self.n, self.rows, self.f, self.h, self.c, which)
        off = ptr - self.ws.data_ptr()
        if name == "adj":
            shape = (n, padded_ld(self.n))
            if self._plan >= 0 and not (self._plan & 1) and (self._plan & 2):      # bit-packed plan: expand the bits (tests / debugging)
                bptr = self.lib.lds_outer_step_shard_buffer(_ptr(self.ws), self.n, self.rows, self.f, self.h, self.c, 16)
                boff = bptr - self.ws.data_ptr()
                return unpack_adj(self.ws[boff:boff + packed_adj_bytes(self.n, n)], self.n, n)[:, :shape[1]].contiguous()
            return self.ws[off:off + 2 * shape[0] * shape[1]].view(BF16).view(shape)
        if name == "fpack":                                  # this rank's rows of the packed bf16 factor matrix
            kf = int(self.lib.lds_outer_step_packed_k(h, c))
            off += 2 * kf * self.row0
            return self.ws[off:off + 2 * n * kf].view(BF16).view(n, kf)
        if name in self.TRANSPOSED:
            w = h if name in ("p1", "z1", "dz1", "dp1") else c
            ldr = int(self.lib.lds_outer_step_state_ld(n))
            return self.ws[off:off + 4 * w * ldr].view(torch.float32).view(w, ldr)[:, :n].t()
        shapes = {"deg": (n,), "rsqrt": (n,), "fa": (n, ldf), "fb": (n, ldf), "cvec": (n,), "operand": (n, max(h, c))}
        shape = shapes[name]
        numel = 1
        for s in shape:
            numel *= s
        return self.ws[off:off + 4 * numel].view(torch.float32).view(shape)

    def _init_static_args(self):
        """Fields of the argument block that do not change between steps (set once; `run` touches the rest)."""
        a = self.args
        a.struct_bytes = ctypes.sizeof(_lib.OuterStepArgs)
        a.n, a.f, a.h, a.c = self.n, self.f, self.h, self.c
        a.x, a.ld_x = self.x.data_ptr(), self.ld_x
        if self.sparse:
            a.x_crow, a.x_col, a.x_val = self.x_crow.data_ptr(), self.x_col.data_ptr(), self.x_val.data_ptr()
        else:
            a.x_crow = a.x_col = a.x_val = None
        a.reserved_ptr = None
        a.ld_w0 = self.f
        a.y, a.mask, a.mask_count = self.y.data_ptr(), self.mask.data_ptr(), self.mask_count
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws_bytes
        a.row0, a.rows, a.phases, a.reserved2 = (self.row0, self.rows, 0, 0) if self.sharded else (0, 0, 0, 0)
        self._clear_optional_args()

    def run_multi(self, theta_full, num_samples, lr, seed, step, **kw):
        """Multi-sample straight-through estimator: `num_samples` graphs (Philox sample index s) with their own dropout
        masks, theta <- clamp(theta - lr * mean_s dL_s/dtheta) applied once; (loss, acc) are the sample means."""
        s_total = int(num_samples)
        if s_total <= 1:
            return self.run(theta_full, lr=lr, seed=seed, step=step, **kw)
        kf = int(self.lib.lds_outer_step_packed_k(self.h, self.c))
        if self._fpack_multi is None or self._fpack_multi.shape != (self.n, s_total * kf):
            self._fpack_multi = torch.zeros((self.n, s_total * kf), dtype=BF16, device=self.device)
        out = None
        for s in range(s_total):
            out = self.run(theta_full, lr=lr, seed=seed, step=step, sample=s, num_samples=s_total, fpack_multi=self._fpack_multi, **kw)
        return out

    def _clear_optional_args(self):
        a = self.args
        a.u_explicit, a.ld_u, a.keep_x, a.keep_h = None, 0, None, None
        a.adam_m = a.adam_v = None
        a.beta1, a.beta2, a.eps, a.adam_t = 0.9, 0.999, 1e-8, 1
        a.out_logp = None
        a.k2_flags = a.k3_flags = 0
        a.opnd_full = a.fa_full = a.fb_full = a.c_full = a.f_full = a.k2_timeline = None
        a.out_scalars = self.scalars.data_ptr()
        self._optional_set = False

    def run(self, theta_full, lr, seed, step, dropout_p=0.0, update=True, u=None, keep_x=None, keep_h=None,
            opt_kind=_lib.OPT_SGD, adam_m=None, adam_v=None, betas=(0.9, 0.999), eps=1e-8, adam_t=1,
            out_logp=None, k2_flags=0, k3_flags=0, phases=None, opnd_full=None, fa_full=None, fb_full=None, c_full=None,
            f_full=None, k2_timeline=None, scalars_out=None, want_adj=True, sample=0, num_samples=1, fpack_multi=None, opnd_send=None, opnd_rank_rows=0, forward_only=False, scalars_tag=0.0):
        """Enqueue one fused outer step on the current stream. Results: (loss, acc) in `scalars_out[0:2]` (any fp32
        buffer the device can write, e.g. pinned host memory) or, by default, in self.scalars.
        `want_adj`: keep the sampled A_tilde readable through buffer("adj") — small graphs run a fused kernel whose A_tilde
        lives in shared memory only; hot callers (trainer, bench) pass False."""
        a = self.args
        a.theta_full, a.ld_theta = theta_full.data_ptr(), theta_full.stride(0)
        a.w0, a.b0, a.w1, a.b1 = self.w0.data_ptr(), self.b0.data_ptr(), self.w1.data_ptr(), self.b1.data_ptr()
        a.dropout_p = float(dropout_p)
        a.seed, a.step = int(seed), int(step)
        a.lr, a.opt_kind = float(lr), int(opt_kind)
        a.update = int(bool(update))
        if want_adj:
            k2_flags = int(k2_flags) | _lib.K2_DUMP_ADJ
        if forward_only:                                     # sample + forward + (loss, acc) [+ out_logp]: evaluation passes
            k2_flags = int(k2_flags) | _lib.K2_FORWARD_ONLY
        a.num_samples, a.sample_index = int(num_samples), int(sample)
        a.fpack_multi = None if fpack_multi is None else fpack_multi.data_ptr()
        a.opnd_send = None if opnd_send is None else opnd_send.data_ptr()
        a.opnd_rank_rows = int(opnd_rank_rows)
        a.scalars_tag = float(scalars_tag)
        plain = (u is None and keep_x is None and keep_h is None and adam_m is None and out_logp is None and not k2_flags
                 and not k3_flags and opnd_full is None and fa_full is None and c_full is None and f_full is None
                 and k2_timeline is None and scalars_out is None)
        if not plain or self._optional_set:
            if plain:
                self._clear_optional_args()
            else:
                a.u_explicit, a.ld_u = (None, 0) if u is None else (u.data_ptr(), u.stride(0))
                a.keep_x = None if keep_x is None else keep_x.data_ptr()
                a.keep_h = None if keep_h is None else keep_h.data_ptr()
                a.adam_m = None if adam_m is None else adam_m.data_ptr()
                a.adam_v = None if adam_v is None else adam_v.data_ptr()
                a.beta1, a.beta2, a.eps, a.adam_t = float(betas[0]), float(betas[1]), float(eps), int(adam_t)
                a.out_scalars = self.scalars.data_ptr() if scalars_out is None else scalars_out.data_ptr()
                a.out_logp = None if out_logp is None else out_logp.data_ptr()
                a.k2_flags, a.k3_flags = int(k2_flags), int(k3_flags)
                a.opnd_full = None if opnd_full is None else opnd_full.data_ptr()
                a.fa_full = None if fa_full is None else fa_full.data_ptr()
                a.fb_full = None if fb_full is None else fb_full.data_ptr()
                a.c_full = None if c_full is None else c_full.data_ptr()
                a.f_full = None if f_full is None else f_full.data_ptr()
                a.k2_timeline = None if k2_timeline is None else k2_timeline.data_ptr()
                self._optional_set = True
        if self.sharded:
            if phases is None:
                raise ValueError("a row-block shard runs phase by phase (lds_gnn_b200.sharded.ShardedOuterStep)")
            a.phases = int(phases)
        self._f32_factors = not (opt_kind == _lib.OPT_SGD and not (k3_flags & _lib.K3_SIMT))
        _lib.check(self.lib.lds_outer_step(self._args_ref, _stream()), "lds_outer_step")
        if want_adj:
            self._plan = int(self.lib.lds_outer_step_plan(self._args_ref))
        return self.scalars if scalars_out is None else scalars_out
