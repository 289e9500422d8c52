"""One replay of the captured bilevel block between cudaProfilerStart/Stop, for an ncu launch list:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/ncu_graph_block.py citeseer"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
from lds_gnn_b200.trainers.graph_block import CapturedBilevelBlock
from lds_gnn_b200.trainers.inner import InnerProblemTrainer
from lds_gnn_b200.trainers.outer import OuterProblemTrainer

workload = sys.argv[1] if len(sys.argv) > 1 else "citeseer"
dev = torch.device("cuda")
data, _, opt_mask, shape = bench.make_workload(workload, 0)
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
torch.cuda.profiler.start()
m = block.replay()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("outer loss", m[-1].loss)
