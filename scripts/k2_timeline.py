"""Device-side timeline of the four K2 launches of one outer step (debug aid): %globaltimer stamps per CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from lds_gnn_b200 import kernels as K
data, weights, opt_mask, shape = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "citeseer", 0)
dev = torch.device("cuda")
data = data.to(dev); opt_mask = opt_mask.to(dev)
n, f, h, c = shape["n"], shape["f"], shape["h"], shape["c"]
eng = K.OuterStep(n, data.x, data.y, opt_mask, hidden=h, classes=c)
eng.set_weights(*(weights[k].to(dev) for k in ("w0", "b0", "w1", "b1")))
iu = torch.triu_indices(n, n)
theta = K.theta_triu_to_full(data.dense_adj[iu[0], iu[1]].contiguous().to(dev))
for i in range(5): eng.run(theta, lr=0.1, seed=1, step=i, dropout_p=0.5)
tl = torch.zeros((4, 512, 8), dtype=torch.int64, device=dev)
eng.run(theta, lr=0.1, seed=1, step=99, dropout_p=0.5, k2_timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().numpy().astype(np.float64)
names = ["start", "acc_done", "partial_written", "counted", "reduced", "epi_done", "end"]
for k in range(4):
    a = t[k]; act = a[:, 0] > 0
    a = a[act]
    t0 = a[:, 0].min()
    print(f"K2 launch {k}: {act.sum()} CTAs")
    for j, nm in enumerate(names):
        col = a[:, j]; col = col[col > 0]
        if len(col): print(f"   {nm:16s} n={len(col):4d}  min {(col.min()-t0)/1e3:7.2f} us  median {(np.median(col)-t0)/1e3:7.2f}  max {(col.max()-t0)/1e3:7.2f}")
