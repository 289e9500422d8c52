import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lds_gnn_b200 import kernels as K
data, weights, opt_mask, shape = bench.make_workload("citeseer", 0)
dev = torch.device("cuda")
data = data.to(dev); opt_mask = opt_mask.to(dev)
n, f, h, c = shape["n"], shape["f"], shape["h"], shape["c"]
eng = K.OuterStep(n, data.x, data.y, opt_mask, hidden=h, classes=c)
eng.set_weights(*(weights[k].to(dev) for k in ("w0", "b0", "w1", "b1")))
iu = torch.triu_indices(n, n)
theta = K.theta_triu_to_full(data.dense_adj[iu[0], iu[1]].contiguous())
for i in range(10): eng.run(theta, lr=0.1, seed=1, step=i, dropout_p=0.5)
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for i in range(N): eng.run(theta, lr=0.1, seed=1, step=100 + i, dropout_p=0.5)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"CPU enqueue per step: {(t1 - t0) / N * 1e6:.1f} us; total per step incl. drain: {(t2 - t0) / N * 1e6:.1f} us")
