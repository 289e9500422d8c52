import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lds_gnn_b200 import kernels as K, _lib
from oracle.make_golden import make_inputs

def run(n, f, h, c, reps=10):
    inp = make_inputs(seed=n, n=n, f=f, h=h, c=c, theta_kind="uniform", p=0.0)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    eng = K.OuterStep(n, dev(inp["x"]), dev(inp["y"]), dev(inp["mask"]), hidden=h, classes=c)
    eng.set_weights(dev(inp["w0"]), dev(inp["b0"]), dev(inp["w1"]), dev(inp["b1"]))
    full0 = K.theta_triu_to_full(dev(inp["theta_triu"]))
    ref = full0.clone()
    eng.run(ref, lr=0.5, seed=1, step=0, dropout_p=0.0, update=True, k3_flags=_lib.K3_SIMT)
    torch.cuda.synchronize()
    kp = 64 * ((3 * (h + c) + 63) // 64)
    bad_runs = 0
    for r in range(reps):
        t = full0.clone()
        eng.run(t, lr=0.5, seed=1, step=0, dropout_p=0.0, update=True)
        torch.cuda.synchronize()
        diff = (t - ref).abs()
        if diff.max().item() > 1e-5:
            bad_runs += 1
            if bad_runs <= 2:
                idx = (diff > 1e-5).nonzero()
                print(f"  n={n} rep {r}: max diff {diff.max().item():.3e} count {idx.shape[0]} first {idx[:5].tolist()} rows {idx[:,0].unique()[:10].tolist()} cols {idx[:,1].unique()[:16].tolist()}")
                i, j = idx[0].tolist()
                print("   vals tc", t[i, j].item(), "simt", ref[i, j].item(), "theta0", full0[i, j].item())
    print(f"n={n} h={h} c={c}: bad runs {bad_runs}/{reps}")

for cfg in [(20, 12, 8, 3), (33, 17, 16, 7), (64, 40, 16, 7), (96, 30, 16, 6), (50, 30, 16, 7), (301, 120, 16, 7), (1200, 200, 32, 10), (257, 50, 64, 7)]:
    run(*cfg)
