"""cProfile of the reference-facing e2e step (OuterProblemTrainer.train_step) at Citeseer shape: where the host time goes."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.trainers.inner import InnerProblemTrainer
from lds_gnn_b200.trainers.outer import OuterProblemTrainer
data, weights, opt_mask, shape = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "citeseer", 0)
dev = torch.device("cuda")
data = data.to(dev); opt_mask = opt_mask.to(dev)
n, f, h, c = shape["n"], shape["f"], shape["h"], shape["c"]
gcn = MetaDenseGCN(f, h, c, dropout=0.5).to(dev)
inner = InnerProblemTrainer(gcn, data)
model = BernoulliGraphModel(data.dense_adj).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=0.1)
outer = OuterProblemTrainer(optimizer=opt, data=data, opt_mask=opt_mask, model=model, smoothness_factor=0.0, disconnection_factor=0.0,
                            sparsity_factor=0.0, regularize=False, lr_decay=0.99, pretrain=False)
for _ in range(20):
    outer.train_step(inner.model_forward)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    outer.train_step(inner.model_forward)
torch.cuda.synchronize()
print(f"train_step: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us/step (route {outer.last_route})")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    outer.train_step(inner.model_forward)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
