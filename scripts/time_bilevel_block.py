"""Time one bilevel block (captured CUDA-graph replay, eager factored route, dense composable route)
 (tau = 5 `inner_opt_step` + 1 `hyper_opt_step`, src/trainers/bilevel.py:53-73) on the
factored route (FactoredGraph unroll, K2 for every product, one K3+K4 pass) and on the dense composable route.
usage: python scripts/time_bilevel_block.py [cora|citeseer] [blocks]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lds_gnn_b200.models.gcn import MetaDenseGCN
from lds_gnn_b200.models.graph import BernoulliGraphModel
from lds_gnn_b200.models.sampling import PHILOX
from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
from lds_gnn_b200.trainers.inner import InnerProblemTrainer
from lds_gnn_b200.trainers.outer import OuterProblemTrainer

workload = sys.argv[1] if len(sys.argv) > 1 else "citeseer"
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
TAU = 5
dev = torch.device("cuda")
data, weights, opt_mask, shape = bench.make_workload(workload, 0)
data, opt_mask = data.to(dev), opt_mask.to(dev)
out = {"workload": workload, "shape": shape, "tau": TAU, "blocks": blocks}
for route in ("graph", "factored", "composable"):
    torch.manual_seed(0)
    PHILOX.manual_seed(0)
    gcn = MetaDenseGCN(shape["f"], shape["h"], shape["c"], dropout=0.5).to(dev)
    inner = InnerProblemTrainer(gcn, data, lr=0.01, weight_decay=5e-4)
    model = BernoulliGraphModel(data.dense_adj).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    outer = OuterProblemTrainer(optimizer=opt, data=data, opt_mask=opt_mask, model=model, smoothness_factor=0.0, disconnection_factor=0.0,
                                sparsity_factor=0.0, regularize=False, lr_decay=0.99, pretrain=False)
    outer.factored_enabled = route != "composable"
    runner = BilevelProblemRunner(inner, outer, data)
    runner.logger.disabled = True

    captured = None
    if route == "graph":
        from lds_gnn_b200.trainers.graph_block import CapturedBilevelBlock
        captured = CapturedBilevelBlock(runner, TAU)

    def block():
        if captured is not None:
            return captured.replay()
        for _ in range(TAU):
            runner.inner_opt_step()
        runner.hyper_opt_step(0)

    n_blocks = {"graph": 5 * blocks, "factored": blocks}.get(route, max(3, blocks // 4))
    try:
        for _ in range(3):
            block()
    except torch.OutOfMemoryError as exc:
        out[route] = {"error": "out of memory: " + str(exc)[:120]}
        print(f"{workload} {route}: out of memory", flush=True)
        del runner, outer, inner, model, gcn, opt
        torch.cuda.empty_cache()
        continue
    assert outer.last_route == {"graph": "factored-graph"}.get(route, route), outer.last_route
    torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    for _ in range(n_blocks):
        block()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n_blocks * 1e3
    out[route] = {"ms_per_block": round(ms, 3), "theta_updates_per_s": round(1e3 / ms, 2), "inner_steps_per_s": round(TAU * 1e3 / ms, 1),
                  "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 3)}
    print(f"{workload} {route:10s}: {ms:8.3f} ms per block ({TAU} inner steps + 1 hyper step), peak {out[route]['peak_mem_gb']} GB", flush=True)
    del runner, outer, inner, model, gcn, opt
    torch.cuda.empty_cache()
if "ms_per_block" in out.get("composable", {}):
    out["speedup_graph_vs_composable"] = round(out["composable"]["ms_per_block"] / out["graph"]["ms_per_block"], 2)
out["speedup_graph_vs_eager_factored"] = round(out["factored"]["ms_per_block"] / out["graph"]["ms_per_block"], 2)
print(json.dumps(out))
