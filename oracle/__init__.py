"""oracle/ — TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's LDS outer-step path (andreas-grafberger/lds-gnn), used only as
the checker: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it. Nothing under `lds-gnn_b200/` imports it, and the product path fails loudly when
the CUDA extension is missing instead of falling back to anything here.

Contents
--------
restatement.py     numpy restatement (explicit uniforms, closed-form backward), fp64 by default.
reference_port.py  op-for-op torch/CPU port of the reference's code path (what the reference's own
                   CPU run executes, including its O(N^3) normalisation); the timed CPU baseline.
philox.py          numpy Philox4x32-10 + the draw-to-element mapping the CUDA kernels use.
live_reference.py  imports the UNMODIFIED reference from /root/reference through `shims/`
                   (only works in the build container; never used at GPU-box run time).
make_golden.py     runs the live reference and writes `tests/golden/*.npz` (committed).
shims/             stand-ins for sacred / torchmeta / torch_geometric / torch_scatter / higher / seml.

Parity pinning: the restatement and the port are checked (tests/test_oracle_*.py) against
 (1) every known-answer tensor test the reference holds for this path
     (tst/models/test_sampling.py:149-160, tst/utils/test_graph.py:32-52,213-221,232-235,
      tst/models/test_bernoulli_model.py:56-64,113-120, tst/utils/test_evaluation.py:12-18) and
 (2) golden vectors produced by the live reference itself (`make_golden.py`), because the reference
     stores no numeric vectors for A_hat, logits, loss or dL/dtheta.
"""
