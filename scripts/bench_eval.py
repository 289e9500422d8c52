"""Times empirical_mean_loss (16 graphs) at Cora / Citeseer shape. Usage: python scripts/bench_eval.py [cora|citeseer]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.utils.evaluation import empirical_mean_loss

name = sys.argv[1] if len(sys.argv) > 1 else "cora"
dev = torch.device("cuda")
data, weights, opt_mask, shape = bench.make_workload(name, seed=0, knn_on_device=dev)
data = data.to(dev)
gcn = MetaDenseGCN(shape["f"], shape["h"], shape["c"], dropout=0.5).to(dev)
model = BernoulliGraphModel(data.dense_adj).to(dev)
with torch.no_grad():
    model.probs.mul_(0.8).add_(0.05)
for _ in range(3):
    empirical_mean_loss(gcn, model, 16, data)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); r = empirical_mean_loss(gcn, model, 16, data); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(f"{name}: empirical_mean_loss(16 graphs) min {min(ts) * 1e3:.3f} ms median {sorted(ts)[5] * 1e3:.3f} ms  val loss {r[0].loss:.4f}")
import ctypes
from lds_gnn_b200 import _lib
lib = _lib.load()
lib.lds_profile_begin()
empirical_mean_loss(gcn, model, 16, data)
ms = (ctypes.c_float * 64)(); ids = (ctypes.c_int32 * 64)()
k = lib.lds_profile_end(ms, ids, 64)
print("launch ids:", [ids[i] for i in range(k)], "ms:", [round(ms[i], 4) for i in range(k)])
