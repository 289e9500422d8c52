"""Host-side profile (cProfile) of bilevel blocks on the factored route: where the Python / dispatch time goes.
usage: python scripts/profile_bilevel_block.py [cora|citeseer]"""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
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


def block():
    for _ in range(5):
        runner.inner_opt_step()
    runner.hyper_opt_step(0)


for _ in range(3):
    block()
torch.cuda.synchronize()
# GPU time of a block: events around 10 blocks (the host runs ahead only as far as the per-step syncs allow)
prof = cProfile.Profile()
prof.enable()
for _ in range(10):
    block()
torch.cuda.synchronize()
prof.disable()
stats = pstats.Stats(prof)
stats.sort_stats("cumulative").print_stats(45)
stats.sort_stats("tottime").print_stats(25)
