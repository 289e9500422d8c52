"""Bitwise reproducibility of the full-size step (debug aid): run the same step twice, compare every intermediate buffer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from lds_gnn_b200 import kernels as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
wl = "n20k"
dev = torch.device("cuda")
d = bench.make_large_rows(wl, dev, 0, 20000, seed=3)
if N != 20000:
    sys.exit("only N=20000")
rng = np.random.default_rng(0)
f, h, c = d["f"], d["h"], d["c"]
w = [torch.as_tensor((rng.standard_normal(s) * 0.2).astype(np.float32), device=dev) for s in ((h, f), (h,), (c, h), (c,))]
eng = K.OuterStep(N, d["x"], d["y"], d["mask"], hidden=h, classes=c); eng.set_weights(*w)
theta0 = d["theta"].clone()
names = ("adj", "deg", "p1", "z1", "p2", "z2", "dz2", "dp2", "dz1", "dp1", "cvec", "fpack")
runs = []
for rep in range(3):
    t = theta0.clone()
    sc = eng.run(t, lr=50.0, seed=5, step=2, dropout_p=0.5, update=True).clone()
    torch.cuda.synchronize()
    runs.append((t, sc, {k: eng.buffer(k).clone() for k in names}))
for rep in (1, 2):
    print("rep", rep, "theta equal", torch.equal(runs[0][0], runs[rep][0]), "scalars", runs[0][1].tolist()[:2], runs[rep][1].tolist()[:2])
    for k in names:
        a, b = runs[0][2][k], runs[rep][2][k]
        if not torch.equal(a, b):
            ne = (a.float() != b.float())
            print("   differs:", k, int(ne.sum()), "elements; first at", ne.nonzero()[:3].tolist(), "max abs diff", float((a.float() - b.float()).abs().max()))
