"""-m gpu tests of the bit-packed A_tilde plan (csrc/lds_k1_packed.cu, lds_k2_packed.cu): tile-symmetric sampling straight to
bits and the propagation that expands them on chip. Oracle: oracle/restatement.py + oracle/philox.py (mask bit-exact,
src/models/sampling.py:47-85, src/utils/graph.py:123-153), fp64 products for the propagation (src/models/layers.py:44)."""
import numpy as np
import pytest
import torch

from oracle import philox as PH
from oracle import restatement as R

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    return (t if dtype is None else t.to(dtype)).cuda()


def theta_matrix(n, seed, kind="mixed"):
    from lds_gnn_b200 import kernels as K
    rng = np.random.default_rng(seed)
    t = n * (n + 1) // 2
    th = rng.random(t).astype(np.float32)
    if kind == "mixed":
        th[rng.random(t) < 0.3] = 0.0
        th[rng.random(t) < 0.1] = 1.0
        if t > 10:
            th[3], th[7] = 1.4, -0.3                                    # outside [0, 1]: clamp semantics of the compare
    return th, K.theta_triu_to_full(dev(th))


@pytest.mark.parametrize("n", [1, 2, 3, 63, 64, 65, 127, 130, 257, 513, 700])
@pytest.mark.parametrize("explicit", [False, True], ids=["philox", "explicit_u"])
def test_packed_sampling_mask_is_bit_exact(n, explicit):
    from lds_gnn_b200 import kernels as K
    th, full = theta_matrix(n, n)
    seed, step, sample = 0xFEED + n, 5, 2
    u = PH.edge_uniforms(n, seed, step, sample=sample)
    bits, deg, rs = K.k1_sample_packed(full, n, seed, step, sample=sample, u=dev(u) if explicit else None)
    a_ref = R.add_self_loops(R.sample_graph(R.theta_full_from_triu(th), u))
    got = K.unpack_adj(bits, n, dtype=torch.float32).cpu().numpy()
    assert got.shape[1] % 64 == 0 and np.all(got[:, n:] == 0), "padding columns must hold zero bits"
    assert np.array_equal(got[:, :n], a_ref.astype(np.float32)), "sampled mask differs from the oracle"
    assert np.array_equal(deg.cpu().numpy(), a_ref.sum(1).astype(np.float32))
    assert np.array_equal(rs.cpu().numpy(), (1.0 / np.sqrt(a_ref.sum(1).astype(np.float32))).astype(np.float32))
    tail = K.unpack_adj(bits, n, rows=((n + 255) // 256) * 256, dtype=torch.float32)[n:]
    assert tail.numel() == 0 or float(tail.abs().max()) == 0.0, "rows beyond the matrix must hold zero bits"


@pytest.mark.parametrize("n,world", [(2708, 1), (3327, 1), (5000, 3), (1500, 4), (900, 7)])
def test_packed_sampling_equals_the_bf16_kernel_incl_row_block_shards(n, world):
    """Same draws as lds_k1_sample_normalize for every shard geometry: no exchange of random bits."""
    from lds_gnn_b200 import kernels as K, sharded as S
    _, full = theta_matrix(n, n + world, kind="uniform")
    seed, step = 77, 9
    adj, _, deg, rs = K.k1_sample_normalize(full, n, seed, step)
    for rank in range(world):
        lo, cnt = S.shard_bounds(n, world, rank)
        if cnt == 0:
            continue
        bits, d, r = K.k1_sample_packed(full[lo:lo + cnt], n, seed, step, row0=lo, rows=cnt)
        got = K.unpack_adj(bits, n, rows=cnt)
        assert torch.equal(got, adj[lo:lo + cnt]), f"rank {rank}"
        assert torch.equal(d, deg[lo:lo + cnt]) and torch.equal(r, rs[lo:lo + cnt])
    bits2, _, _ = K.k1_sample_packed(full, n, seed, step)
    bits3, _, _ = K.k1_sample_packed(full, n, seed, step + 1)
    assert torch.equal(bits2, K.k1_sample_packed(full, n, seed, step)[0]) and not torch.equal(bits2, bits3)


@pytest.mark.parametrize("n,rows,row0,w", [(300, 300, 0, 7), (300, 300, 0, 16), (1000, 1000, 0, 33), (2708, 2708, 0, 64), (1500, 384, 512, 100),
                                           (700, 188, 512, 16), (257, 257, 0, 128), (64, 64, 0, 1)])
def test_packed_propagation_matches_fp64_and_the_bf16_kernel(n, rows, row0, w):
    from lds_gnn_b200 import kernels as K
    _, full = theta_matrix(n, n + w, kind="uniform")
    bits, deg, rs = K.k1_sample_packed(full[row0:row0 + rows], n, 5, 1, row0=row0, rows=rows)
    a = K.unpack_adj(bits, n, rows=rows)
    torch.manual_seed(w)
    p = torch.randn(n, w, device="cuda")
    si, so = torch.rand(n, device="cuda") + 0.5, torch.rand(rows, device="cuda") + 0.5
    z = K.k2_propagate_packed(bits, n, rows, p, scale_in=si, scale_out=so)
    ref = so.double()[:, None] * (a[:, :n].double() @ (si.double()[:, None] * p.double()))
    err = (z.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err
    z_bf16 = K.k2_propagate(a, n, p, scale_in=si, scale_out=so)
    assert (z - z_bf16).abs().max().item() <= 2e-6 * ref.abs().max().item()          # same operands, different split-K order
    assert torch.equal(z, K.k2_propagate_packed(bits, n, rows, p, scale_in=si, scale_out=so))      # bitwise reproducible
    z1 = K.k2_propagate_packed(bits, n, rows, p)
    assert (z1.double() - a[:, :n].double() @ p.double()).abs().max().item() <= 2e-5 * ref.abs().max().item() * 4


@pytest.mark.parametrize("n,f,h,c,p,sparse", [(301, 120, 16, 7, 0.5, False), (700, 64, 64, 7, 0.5, True), (1200, 200, 32, 10, 0.0, False),
                                              (2708, 1433, 16, 7, 0.5, True)])
def test_outer_step_packed_plan_equals_bf16_plan(n, f, h, c, p, sparse):
    """The whole step on both launch plans (NO_FUSE keeps small graphs off the fused kernel): identical graph, theta within
    fp32 summation-order noise."""
    from oracle.make_golden import make_inputs
    from lds_gnn_b200 import _lib, kernels as K
    inp = make_inputs(seed=n + 3, n=n, f=f, h=h, c=c, theta_kind="mixed", p=p, mask_frac=0.2)
    eng = K.OuterStep(n, dev(inp["x"]), dev(inp["y"]), dev(inp["mask"]), hidden=h, classes=c, sparse_features=sparse)
    eng.set_weights(dev(inp["w0"]), dev(inp["b0"]), dev(inp["w1"]), dev(inp["b1"]))
    full = K.theta_triu_to_full(dev(inp["theta_triu"]))
    out = {}
    for name, flags in (("packed", _lib.K2_NO_FUSE), ("bf16", _lib.K2_NO_FUSE | _lib.K2_BF16_ADJ)):
        t = full.clone()
        lp = torch.empty((n, c), device="cuda")
        sc = eng.run(t, lr=0.3, seed=5, step=2, dropout_p=p, update=True, out_logp=lp, k2_flags=flags).clone()
        assert bool(eng._plan & 2) == (name == "packed") and not (eng._plan & 1)
        out[name] = (t, lp, sc, eng.buffer("adj").clone(), eng.buffer("deg").clone())
    assert torch.equal(out["packed"][3], out["bf16"][3]) and torch.equal(out["packed"][4], out["bf16"][4])
    assert (out["packed"][1] - out["bf16"][1]).abs().max().item() <= 1e-5 * out["bf16"][1].abs().max().item()
    assert (out["packed"][0] - out["bf16"][0]).abs().max().item() < 1e-6
    assert abs(out["packed"][2][0].item() - out["bf16"][2][0].item()) < 1e-5
    sym = out["packed"][0][:, :n]
    assert torch.equal(sym, sym.t())
