"""Kernel-level profile (torch.profiler / CUPTI) of ONE replay of the captured bilevel block: kernel count and GPU time by kernel.
usage: python scripts/profile_graph_block.py [cora|citeseer]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
from lds_gnn_b200.trainers.graph_block import CapturedBilevelBlock
from lds_gnn_b200.trainers.inner import InnerProblemTrainer
from lds_gnn_b200.trainers.outer import OuterProblemTrainer

workload = sys.argv[1] if len(sys.argv) > 1 else "citeseer"
dev = torch.device("cuda")
data, weights, opt_mask, shape = bench.make_workload(workload, 0)
data, opt_mask = data.to(dev), opt_mask.to(dev)
gcn = MetaDenseGCN(shape["f"], shape["h"], shape["c"], dropout=0.5).to(dev)
inner = InnerProblemTrainer(gcn, data, lr=0.01, weight_decay=5e-4)
model = BernoulliGraphModel(data.dense_adj).to(dev)
outer = OuterProblemTrainer(optimizer=torch.optim.SGD(model.parameters(), lr=0.1), data=data, opt_mask=opt_mask, model=model,
                            smoothness_factor=0.0, disconnection_factor=0.0, sparsity_factor=0.0, regularize=False, lr_decay=0.99)
runner = BilevelProblemRunner(inner, outer, data)
block = CapturedBilevelBlock(runner, 5)
for _ in range(3):
    block.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    block.replay()
    torch.cuda.synchronize()
events = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
total = sum(e.device_time for e in events)
print(f"{len(events)} device activities in one replay, {total / 1e3:.3f} ms of kernel time")
by = {}
for e in events:
    k = e.name[:90]
    c, tt = by.get(k, (0, 0.0))
    by[k] = (c + 1, tt + e.device_time)
for name, (c, tt) in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{tt / 1e3:8.3f} ms  {c:5d} x  {tt / c:7.1f} us  {name}")
