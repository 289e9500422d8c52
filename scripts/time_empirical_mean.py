"""Time empirical_mean_loss (16 samples, src/utils/evaluation.py:51-84) at Cora shape: fused forward-only route vs the sample loop."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import lds_gnn_b200.utils.evaluation as E
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.trainers.inner import InnerProblemTrainer
data, weights, opt_mask, shape = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cora", 0)
dev = torch.device("cuda")
data = data.to(dev)
gcn = MetaDenseGCN(shape["f"], shape["h"], shape["c"], dropout=0.5).to(dev)
inner = InnerProblemTrainer(gcn, data)
model = BernoulliGraphModel(data.dense_adj).to(dev)
def timed(label):
    for _ in range(3): E.empirical_mean_loss(gcn, model, 16, data, inner.model_params)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): out = E.empirical_mean_loss(gcn, model, 16, data, inner.model_params)
    torch.cuda.synchronize()
    print(f"{label}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call (16 samples)  val={out[0]}")
timed("fused forward-only")
orig = E._fused_eval_engine
E._fused_eval_engine = lambda *a, **k: None
timed("sample-by-sample loop (composable kernels)")
