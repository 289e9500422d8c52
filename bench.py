#!/usr/bin/env python
"""bench.py — LDS outer steps/s on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload citeseer|cora|cora_knn16|tiny|n20k|n65k] [--impl ours|reference]

A "step" is one direct outer step (OuterProblemTrainer.train_step with model_forward at fixed GCN weights,
SGD, dropout 0.5; reference src/trainers/outer.py:57-87) on synthetic data of the named shape. Default workload:
Citeseer shape (BASELINE.json configs[1]); cora_knn16 = config 3 (kNN theta_0, 16 samples per step), n20k = config 4.
  value     device-resident throughput: K calls of the fused C entry point `lds_outer_step`, theta / X / weights
            already in HBM, each step timed with CUDA events on the launching stream. Before every timed step the L2 is
            flushed (256 MiB write) and one untimed step runs on a second, disjoint problem instance, which brings the
            kernels' code back without touching the timed step's data (`flushed_cold_code` = without that step: the
            round-1 protocol, which also refetches ~200 KB of SASS from HBM; `warm_l2` = back to back, no flush).
  e2e       the same metric through the reference-facing API `OuterProblemTrainer.train_step(inner.model_forward)`:
            every step copies that step's GCN weights from pinned host memory to the device and reads Metrics
            (loss, acc) back to the host (written by the kernel into pinned memory); wall clock between two device
            synchronisations.
  roofline  the dominant kernel of the step (by device time, per-kernel CUDA events recorded inside the library
            on the launching stream): algorithmic bytes per launch / its mean duration vs the measured HBM peak;
            `traffic` = DRAM bytes of that kernel from the committed ncu capture (profiles/ncu_traffic.json).
  cpu_baseline  the oracle's torch/CPU port of the reference's own op sequence, timed on this box's host cores
            (rank 0, N=1 only), a bounded sample of the same workload.
--impl reference times that CPU port for K steps and prints the same line with "impl": "reference".
N > 1 (torchrun): BASELINE.json config 5 — N = 65 536 with theta / A_tilde row-block sharded over the ranks ("scaling":
"strong"; the operand / factor exchange runs over NVLink peer memory, NCCL as fallback; LDS_EXCHANGE=nccl forces NCCL).
Each line carries `scaling_reference`: the same step on one GPU measured in the same run. `--replicas` runs N independent
single-GPU replicas of the default workload instead (weak scaling, no data-path collective).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "lds_outer_steps_per_sec"
UNIT = "steps/s"
KERNEL_NAMES = {0: "k1_sample_normalize", 1: "feat_linear", 3: "k2_layer1", 4: "k2_layer2", 5: "k2_bwd2", 6: "k2_bwd1",
                7: "k3k4_theta_update", 8: "stage_w0", 10: "fused_k1_feat_k2x4"}
K2_IDS = (3, 4, 5, 6)       # the four tcgen05 propagations (each with its fused row epilogue)
LAUNCHES_PER_STEP = 8       # weight staging, K1, feature GEMM, 4 x K2 (+ fused row epilogue), K3+K4 (csrc/lds_outer_step.cu)
HYPER = dict(lr=0.1, lr_decay=0.99, dropout=0.5)            # configs/seml/final/lds.yaml:18-110


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="citeseer")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bilevel-block", action="store_true", help="skip the extra bilevel-block timing (tau inner steps + hyper step)")
    ap.add_argument("--no-scale-ref", action="store_true", help="N > 1: skip the single-GPU run of the same workload on rank 0")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent single-GPU replicas instead of the sharded N=65536 config")
    ap.add_argument("--cpu-steps", type=int, default=3)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_peaks_json():
    return peaks()


def bf16_peak_tflops():
    """Dense bf16 tensor peak for a kernel timed inside a long step: the SUSTAINED cuBLAS figure of MEASURED_PEAKS.json."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh).get("bf16_tflops_sustained", 1397.8))
    return 1397.8


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of
    this workload (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep), or (None, None)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as fh:
            entry = json.load(fh)[workload][kernel]
        return int(entry["dram_bytes"]), entry["source"]
    except (OSError, KeyError, ValueError):
        return None, None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling in the background (started early: it needs a few hundred ms to come up); `stop(t0, t1)` keeps
    the samples whose timestamp falls inside the wall-clock window [t0, t1] of the loaded region."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, clk_max = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            sm.append(clk); mx.append(clk_max)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.file.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------ workload
KNN_WORKLOADS = {"cora_knn16": ("cora", 10, 16)}      # BASELINE.json config 3: kNN theta_0 (k = 10, cosine), 16 samples per outer step


def make_workload(name, seed=0, knn_on_device=None):
    """knn_on_device: CUDA device for the kNN theta_0 of config 3 on OUR kernels (data/utils.py); None = the reference's own
    host path (sklearn kneighbors_graph, src/data/utils.py:165-175) for the CPU arm."""
    from lds_gnn_b200.data import SHAPES, make_dataset
    knn = KNN_WORKLOADS.get(name)
    if knn is not None:
        name = knn[0]
    n, f, c, h, _, _ = SHAPES[name]
    data = make_dataset(name, seed=seed)
    if knn is not None:                                  # theta_0 = symmetrised kNN graph of the features (data/transforms.py:15-37)
        if knn_on_device is not None:
            from lds_gnn_b200.data.utils import knn_init_adjacency
            data.dense_adj = knn_init_adjacency(data.x.to(knn_on_device), k=knn[1], metric="cosine", loop=False).cpu()
        else:
            from oracle.theta0 import knn_connectivity
            a = knn_connectivity(data.x.numpy(), knn[1], "cosine", False)
            data.dense_adj = torch.as_tensor(np.maximum(a, a.T))
    rng = np.random.default_rng(seed + 1)
    lim0, lim1 = np.sqrt(6.0 / (f + h)), np.sqrt(6.0 / (h + c))          # xavier-uniform, zero bias (layers.py:38-40)
    weights = dict(w0=torch.as_tensor(rng.uniform(-lim0, lim0, (h, f)).astype(np.float32)), b0=torch.zeros(h),
                   w1=torch.as_tensor(rng.uniform(-lim1, lim1, (c, h)).astype(np.float32)), b1=torch.zeros(c))
    # outer objective mask = half of the validation nodes (src/scripts/bilevel.py:77)
    val_idx = data.val_mask.nonzero().flatten()
    opt_mask = torch.zeros_like(data.val_mask)
    opt_mask[val_idx[len(val_idx) // 2:]] = True
    return data, weights, opt_mask, dict(n=n, f=f, c=c, h=h)


def small_config(workload, shape, samples, world=1, replicas=False):
    """`config` of a Citeseer / Cora shape line. Built by ONE function and worded arm-neutrally so that the repo arm and the
    `--impl reference` arm print the same dict (the driver compares them)."""
    n, f, c, h = shape["n"], shape["f"], shape["c"], shape["h"]
    return {"workload": f"LDS-GCN direct outer step, {workload} shape (N={n}, F={f}, C={c}, hidden={h}), SGD lr 0.1 decay 0.99, "
                        f"dropout 0.5, {samples} sample(s)/step" + (", theta_0 = kNN graph (k=10, cosine)" if samples > 1 else ""),
            "parallelism": "one device per replica" + (f", {world} independent replicas" if (world > 1 and replicas) else ""),
            "l2": "GPU arm: flushed between timed steps (256 MiB write), then one untimed step on a SECOND problem instance "
                  "(disjoint buffers: brings the kernels' code back, none of the timed step's data); CPU reference arm: host caches as they are",
            "theta_init": "kNN graph of the synthetic features" if samples > 1 else "synthetic SBM adjacency"}


def large_config(n, f, h, c, world):
    """`config` of an N = 20 000 / 65 536 line (same dict in both arms, see small_config)."""
    return {"workload": f"LDS-GCN direct outer step, synthetic N={n} dense theta ~ U(0,1), F={f}, hidden={h}, C={c}, SGD lr 0.1 decay 0.99, "
                        f"dropout 0.5, 1 sample/step",
            "parallelism": "single GPU" if world == 1 else
                           f"theta / A_tilde row-block sharded over {world} GPUs; per step 4 all-gathers of the N x h operand + packed factor rows + c",
            "l2": "not flushed: per-GPU theta rows are %.1f GB, far larger than the 126 MB L2" % (n // world * n * 4 / 1e9)}


def algorithmic_bytes(shape):
    n, f = shape["n"], shape["f"]
    # per-unit figures of SURVEY.md 8(d): K1 6 N^2, each propagation 2 N^2, K3+K4 8 N^2. The fused small-graph kernel (id 10)
    # covers K1 + the four propagations = 14 N^2 algorithmic bytes; its DRAM traffic is far lower (A_tilde stays in smem).
    per_kernel = {0: 6 * n * n, 3: 2 * n * n, 4: 2 * n * n, 5: 2 * n * n, 6: 2 * n * n, 7: 8 * n * n, 10: 14 * n * n}
    step = 6 * n * n + 4 * 2 * n * n + 8 * n * n + 4 * n * f            # the outer step reads X once (no dW)
    return per_kernel, step


def reduce_rank_times(values_ms, device, world):
    """Max over ranks of each timing (the job is as slow as its slowest rank)."""
    t = torch.tensor(list(values_ms), dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args, rank, world, device):
    import torch.distributed as dist
    from lds_gnn_b200 import _lib, kernels as K
    from lds_gnn_b200.models.gcn import MetaDenseGCN
    from lds_gnn_b200.models.graph import BernoulliGraphModel
    from lds_gnn_b200.trainers.inner import InnerProblemTrainer
    from lds_gnn_b200.trainers.outer import OuterProblemTrainer

    clocks = ClockSampler(torch.cuda.current_device())
    data, weights, opt_mask, shape = make_workload(args.workload, seed=rank, knn_on_device=device)
    n, f, h, c = shape["n"], shape["f"], shape["h"], shape["c"]
    data = data.to(device)
    opt_mask = opt_mask.to(device)
    lib = _lib.load()

    # ---- device-resident arm: the fused C entry point ------------------------------------------------
    # Two independent problem instances (own theta, X, masks, weights, workspace). Instance 0 is the one that is timed. Before
    # every timed step the L2 is flushed (256 MiB write) and ONE untimed step runs on instance 1: the flush also evicts the
    # kernels' ~200 KB of SASS, and a step that starts with one dependent instruction miss per code region measures the flush,
    # not the step ("cold start is instruction fetch") — in the training loop the code never leaves L2.
    # Instance 1 shares no buffer with instance 0, so every byte the timed step reads still comes from HBM.
    def make_instance(seed_i, first):
        d_i, w_i, m_i, _ = (data, weights, opt_mask, None) if first else make_workload(args.workload, seed=seed_i, knn_on_device=device)
        d_i = d_i.to(device); m_i = m_i.to(device)
        e_i = K.OuterStep(n, d_i.x, d_i.y, m_i, hidden=h, classes=c)
        e_i.set_weights(*(w_i[k].to(device) for k in ("w0", "b0", "w1", "b1")))
        iu = torch.triu_indices(n, n)
        th_i = K.theta_triu_to_full(d_i.dense_adj[iu[0], iu[1]].contiguous())
        return e_i, th_i

    instances = [make_instance(rank, True), make_instance(1000 + rank, False)]
    eng = instances[0][0]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)        # > 126 MB L2
    lr = HYPER["lr"]
    seed = 1234 + rank

    samples = KNN_WORKLOADS[args.workload][2] if args.workload in KNN_WORKLOADS else 1

    def one(step_idx, lr_now, inst=0):
        e_i, th_i = instances[inst]
        if samples > 1:
            e_i.run_multi(th_i, samples, lr=lr_now, seed=seed, step=step_idx, dropout_p=HYPER["dropout"], update=True, want_adj=False)
        else:
            e_i.run(th_i, lr=lr_now, seed=seed, step=step_idx, dropout_p=HYPER["dropout"], update=True, want_adj=False)

    for w in range(args.warmup):
        one(w, lr); one(w, lr, 1); lr *= HYPER["lr_decay"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall0 = time.time()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                       # evict theta / A_tilde / X (and the code) from L2 between timed steps
        one(args.warmup + k, lr, 1)                 # untimed, other instance: the code comes back, instance 0's data does not
        starts[k].record()
        one(args.warmup + k, lr); lr *= HYPER["lr_decay"]
        ends[k].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    loss_after = float(eng.scalars[0].item())

    # the round-1 protocol for continuity: the timed step directly behind the flush (cold data AND cold code)
    cs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ce = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        cs[k].record()
        one(5_000 + k, lr)
        ce[k].record()
    torch.cuda.synchronize()
    cold_ms = sum(s.elapsed_time(e) for s, e in zip(cs, ce))

    # warm-L2 figure (back-to-back, no flush) for context; run for >= 0.6 s so the clock sampler sees the loaded GPU
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    warm_steps = 0
    wall_sustained = time.time()
    t0.record()
    while True:
        for k in range(args.steps):
            one(10_000 + warm_steps + k, lr)
        warm_steps += args.steps
        if time.time() - wall_sustained > 0.6 or warm_steps > 200_000:
            break
    t1.record()
    torch.cuda.synchronize()
    warm_ms = t0.elapsed_time(t1) * args.steps / warm_steps

    # ---- per-kernel durations (CUDA events inside the library, on the launching stream) ---------------
    per_kernel = {}
    ms_buf = (ctypes.c_float * 64)()
    id_buf = (ctypes.c_int32 * 64)()
    reps = min(args.steps, 20)
    for k in range(reps):
        flush.fill_(k & 0xFF)
        one(20_000 + k, lr, 1)                      # same protocol as the timed loop
        lib.lds_profile_begin()
        one(20_000 + k, lr)
        cnt = lib.lds_profile_end(ms_buf, id_buf, 64)
        for i in range(max(cnt, 0)):
            per_kernel.setdefault(int(id_buf[i]), []).append(float(ms_buf[i]))
    clock_info = clocks.stop(wall0, time.time())

    # ---- e2e arm: the reference-facing API with host buffers ------------------------------------------
    gcn = MetaDenseGCN(f, h, c, dropout=HYPER["dropout"]).to(device)
    inner = InnerProblemTrainer(gcn, data)
    model = BernoulliGraphModel(data.dense_adj).to(device)
    opt = torch.optim.SGD(model.parameters(), lr=HYPER["lr"])
    outer = OuterProblemTrainer(optimizer=opt, data=data, opt_mask=opt_mask, model=model, smoothness_factor=0.0,
                                disconnection_factor=0.0, sparsity_factor=0.0, regularize=False,
                                lr_decay=HYPER["lr_decay"], pretrain=False)
    outer.n_samples = samples
    names = {"w0": "layer_in.fc.weight", "b0": "layer_in.fc.bias", "w1": "layer_out.fc.weight", "b1": "layer_out.fc.bias"}
    # The step's inputs are the current GCN (fast) weights: one flat pinned host buffer, one flat device buffer whose
    # slices ARE the fast-weight tensors handed to the trainer, one host->device copy per step.
    total = sum(v.numel() for v in weights.values())
    host_flat = torch.empty(total, dtype=torch.float32).pin_memory()
    # Two device copies of the weights, alternated: train_step returns as soon as (loss, acc) exist, i.e. while the backward
    # half of the step (which still reads layer_out's weights) and the theta update are running; the next step's upload is
    # ordered behind them by the stream, and writes the OTHER buffer so that user code touching the weights stays simple.
    dev_flats = [torch.empty(total, dtype=torch.float32, device=device) for _ in range(2)]
    param_sets = []
    off = 0
    for k, pname in names.items():
        cnt = weights[k].numel()
        host_flat[off:off + cnt].copy_(weights[k].reshape(-1))
        off += cnt
    for buf in dev_flats:
        views, off = {}, 0
        for k, pname in names.items():
            cnt = weights[k].numel()
            views[pname] = buf[off:off + cnt].view(weights[k].shape)
            off += cnt
        param_sets.append(views)
    h2d_bytes = total * 4
    state = {"i": 0}
    same_stream_upload = os.environ.get("LDS_BENCH_E2E_UPLOAD") == "stream"      # A/B switch: torch copy_ on the compute stream
    # (Round 1: the upload on its own stream with an event through torch's Python stream API cost the host more than the 10 us
    # copy costs the GPU — e2e 8.9 k vs 9.2 k steps/s; as three driver calls inside one C call (lds_upload_async) it pays.
    # Tried: no copy at all — the trainer is handed PINNED HOST tensors and the fused kernel's weight-staging warps read W0
    # straight over PCIe (source-order, full 128-byte lines, under the sampling phase; b0 / W1 / b1 to shared memory):
    # bit-identical results, but SM-initiated reads of host memory reach only ~4 GB/s here — 237 KB took ~55 us, e2e 7.0 k
    # against 10.3 k steps/s with the copy engine (200-step loops, one box). The copy engine stays.)

    def api_step():
        i = state["i"] & 1
        state["i"] += 1
        if same_stream_upload:
            dev_flats[i].copy_(host_flat, non_blocking=True)   # this step's fast weights: pinned host -> device
        else:
            K.upload_async(dev_flats[i], host_flat)            # ... on the library's copy stream (lds_upload_async): the transfer
                                                               # overlaps the previous step's backward half and theta update
        inner.model_params.update(param_sets[i])
        return outer.train_step(inner.model_forward)           # returns host floats (device -> host read inside)

    import gc
    gc.collect()                                           # before the warm-up: a collection is milliseconds of idle GPU, and the first
    gc.disable()                                           # steps after an idle gap run slower; no collector pauses in the timed region
    for _ in range(max(args.warmup, 30)):                      # untimed: the host side (allocator, caches, branch history) settles too
        api_step()
    assert outer.last_route == "fused", "bench e2e must exercise the fused CUDA path"
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_start = time.perf_counter()
    for _ in range(args.steps):
        m = api_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t_start
    gc.enable()
    if world > 1:
        dist.barrier()

    # ---- reduce over ranks -----------------------------------------------------------------------------
    dev_ms, e2e_ms, warm_ms, cold_ms = reduce_rank_times([dev_ms, e2e_s * 1e3, warm_ms, cold_ms], device, world)
    total_steps = args.steps * world
    value = total_steps / (dev_ms / 1e3)
    per_kernel_bytes, step_bytes = algorithmic_bytes(shape)
    hbm_peak, peak_src = peaks()
    kernel_summary = {}
    for kid, vals in per_kernel.items():
        launches = len(vals) / reps
        kernel_summary[KERNEL_NAMES.get(kid, "gap_between_calls" if kid < 0 else str(kid))] = {"launches_per_step": launches, "mean_us": 1e3 * sum(vals) / len(vals),
                                                            "step_share_us": 1e3 * sum(vals) / reps}
    dominant_id = max(per_kernel, key=lambda k: sum(per_kernel[k])) if per_kernel else None
    roofline = None
    if dominant_id is not None:
        mean_s = sum(per_kernel[dominant_id]) / len(per_kernel[dominant_id]) / 1e3
        alg = per_kernel_bytes.get(dominant_id)
        if alg is not None:
            achieved = alg / mean_s / 1e9
            traffic, traffic_src = ncu_traffic(args.workload, KERNEL_NAMES[dominant_id])
            roofline = {"kernel": KERNEL_NAMES[dominant_id], "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak,
                        "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                        "algorithmic_bytes_per_launch": alg, "mean_launch_us": round(mean_s * 1e6, 2), "peak_source": peak_src,
                        "note": "working set fits the 126 MB L2 at this shape: DRAM traffic can be below algorithmic bytes"}
        else:
            roofline = {"kernel": KERNEL_NAMES.get(dominant_id), "bound": "latency", "achieved": None, "peak": hbm_peak,
                        "unit": "GB/s", "frac": None, "traffic": None}
    step_frac = (step_bytes / (dev_ms / 1e3 / args.steps)) / 1e9 / hbm_peak
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dev_ms / args.steps, 5), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 adjacency x (bf16 hi+lo) operands, fp32 accumulate; fp32 theta",
        "data": "synthetic",
        "config": small_config(args.workload, shape, samples, world, args.replicas),
        "scaling_note": "N > 1 lines (torchrun) run BASELINE config 5 (N=65536 row-block sharded, strong scaling): their like-for-like "
                        "single-GPU figure is each line's scaling_reference, not this line",
        "clocks": clock_info,
        "e2e": {"value": round(total_steps / (e2e_ms / 1e3), 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                "api": "OuterProblemTrainer.train_step(InnerProblemTrainer.model_forward)", "l2": "not flushed",
                "warmup_steps": max(args.warmup, 30), "upload": "torch copy_ on the compute stream" if same_stream_upload else "lds_upload_async (copy stream of the library)"},
        # fused small-graph kernel + K3K4 (S samples: S fused launches + one K3K4), else 8 launches per step
        "gpu_launches": ((samples + 1) if 10 in per_kernel else LAUNCHES_PER_STEP * samples) * args.steps,
        "roofline": roofline,
        "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "frac_of_hbm_peak": round(step_frac, 4)},
        "warm_l2": {"value": round(total_steps / (warm_ms / 1e3), 2), "unit": UNIT, "ms_per_step": round(warm_ms / args.steps, 5)},
        "flushed_cold_code": {"value": round(total_steps / (cold_ms / 1e3), 2), "unit": UNIT, "ms_per_step": round(cold_ms / args.steps, 5),
                              "note": "round-1 protocol: timed step directly behind the 256 MiB flush, i.e. the kernels' code is refetched from HBM too"},
        "kernels": kernel_summary,
        "kernels_note": "per-kernel times come from a separate pass with an event recorded between the launches, which disables the "
                        "programmatic dependent launch of the update kernel (its prologue otherwise overlaps the fused kernel's last "
                        "phase): their sum exceeds ms_per_step by that overlap",
        "final_loss": loss_after, "e2e_last_metrics": {"loss": m.loss, "acc": m.acc},
    }
    return line


# ------------------------------------------------------------------------------------------------ bilevel block
def time_bilevel_block(workload, device, tau=5, replays=40, eager_blocks=4):
    """The reference's real training loop never runs an outer step in isolation: it alternates tau = 5 inner steps with one
    hyper step whose backward flows through all of them (src/trainers/bilevel.py:53-73). This times that block through the
    package's own BilevelProblemRunner machinery: one captured CUDA graph per block (trainers/graph_block.py) and, for
    reference, the same block run step by step on the eager factored route. Wall clock around synchronised loops."""
    from lds_gnn_b200.models.gcn import MetaDenseGCN
    from lds_gnn_b200.models.graph import BernoulliGraphModel
    from lds_gnn_b200.trainers.bilevel import BilevelProblemRunner
    from lds_gnn_b200.trainers.graph_block import CapturedBilevelBlock
    from lds_gnn_b200.trainers.inner import InnerProblemTrainer
    from lds_gnn_b200.trainers.outer import OuterProblemTrainer
    data, _, opt_mask, shape = make_workload(workload, seed=0, knn_on_device=device)
    data, opt_mask = data.to(device), opt_mask.to(device)
    gcn = MetaDenseGCN(shape["f"], shape["h"], shape["c"], dropout=0.5).to(device)
    inner = InnerProblemTrainer(gcn, data, lr=0.01, weight_decay=5e-4)
    model = BernoulliGraphModel(data.dense_adj).to(device)
    outer = OuterProblemTrainer(optimizer=torch.optim.SGD(model.parameters(), lr=0.1), data=data, opt_mask=opt_mask, model=model,
                                smoothness_factor=0.0, disconnection_factor=0.0, sparsity_factor=0.0, regularize=False, lr_decay=0.99)
    runner = BilevelProblemRunner(inner, outer, data)
    runner.logger.disabled = True
    if not CapturedBilevelBlock.eligible(runner):
        return {"error": "configuration not eligible for captured blocks"}
    block = CapturedBilevelBlock(runner, tau)
    for _ in range(3):
        block.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(replays):
        metrics = block.replay()                                     # includes the block's device->host read of its 6 (loss, acc)
    torch.cuda.synchronize()
    graph_ms = (time.perf_counter() - t0) / replays * 1e3
    block.store_state(tau)

    def eager_block():
        for _ in range(tau):
            runner.inner_opt_step()
        runner.hyper_opt_step(0)

    eager_block()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(eager_blocks):
        eager_block()
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t0) / eager_blocks * 1e3
    return {"unit": f"one block = {tau} inner_opt_step + 1 hyper_opt_step (hypergradient through the {tau} unrolled steps)",
            "ms_per_block": round(graph_ms, 3), "theta_updates_per_s": round(1e3 / graph_ms, 1),
            "inner_steps_per_s": round(tau * 1e3 / graph_ms, 1), "route": "factored-graph",
            "eager_factored_ms_per_block": round(eager_ms, 3), "replays": replays,
            "last_outer_metrics": {"loss": metrics[-1].loss, "acc": metrics[-1].acc}}


# ------------------------------------------------------------------------------------------------ large-N / sharded arm
def hashed_theta(n, ld, row0, rows, device, seed):
    """Dense symmetric theta ~ U[0,1), generated row block by row block on the device from a
    hash of the canonical pair (min(i,j), max(i,j)): every rank builds its own rows, theta_ij == theta_ji across ranks."""
    out = torch.zeros((rows, ld), dtype=torch.float32, device=device)
    cols = torch.arange(n, device=device, dtype=torch.int64)[None, :]
    m35, m32 = (1 << 35) - 1, (1 << 32) - 1
    for r in range(0, rows, 1024):
        i = torch.arange(row0 + r, row0 + min(r + 1024, rows), device=device, dtype=torch.int64)[:, None]
        key = (torch.minimum(i, cols) * n + torch.maximum(i, cols) + seed) * -7046029254386353131       # 0x9E3779B97F4A7C15
        key = (key ^ ((key >> 29) & m35)) * -4658895280553007687                                        # 0xBF58476D1CE4E5B9
        key = key ^ ((key >> 32) & m32)
        out[r:r + i.shape[0], :n] = ((key >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)
    return out


def make_large_rows(name, device, row0, rows, seed=0):
    """Row block [row0, row0+rows) of the synthetic large-N workload: theta, sparse bag-of-words X, labels, objective mask."""
    from lds_gnn_b200.data import SHAPES
    n, f, c, h, rho, _ = SHAPES[name]
    ld = (n + 63) // 64 * 64
    theta = hashed_theta(n, ld, row0, rows, device, seed)
    g = torch.Generator(device=device); g.manual_seed(1000 + seed * 977 + row0)
    x = (torch.rand((rows, f), device=device, generator=g) < rho).to(torch.float32)
    x[torch.arange(rows, device=device), torch.randint(0, f, (rows,), device=device, generator=g)] = 1.0
    x /= x.sum(1, keepdim=True)
    gl = torch.Generator(device="cpu"); gl.manual_seed(seed)
    y_all = torch.randint(0, c, (n,), generator=gl)
    mask_all = torch.zeros(n, dtype=torch.bool)
    mask_all[torch.randperm(n, generator=gl)[:250]] = True
    return dict(n=n, f=f, c=c, h=h, theta=theta, x=x, y=y_all[row0:row0 + rows].to(device), mask=mask_all[row0:row0 + rows].to(device),
                mask_count=250)


def run_large(args, rank, world, device, workload):
    """N = 20 000 / 65 536 on one GPU, or N = 65 536 row-block sharded over `world` GPUs with NCCL all-gathers of the
    N x h operand (lds_gnn_b200/sharded.py). Inputs are far larger than L2, so no flush between timed steps."""
    import torch.distributed as dist
    from lds_gnn_b200 import _lib, kernels as K, sharded as S
    from lds_gnn_b200.data import SHAPES
    n, f, c, h, _, _ = SHAPES[workload]
    clocks = ClockSampler(torch.cuda.current_device())
    lo, cnt = S.shard_bounds(n, world, rank) if world > 1 else (0, n)
    d = make_large_rows(workload, device, lo, cnt, seed=0)
    rng = np.random.default_rng(1)
    lim0, lim1 = np.sqrt(6.0 / (f + h)), np.sqrt(6.0 / (h + c))
    host_w = [torch.as_tensor(rng.uniform(-lim0, lim0, (h, f)).astype(np.float32)), torch.zeros(h),
              torch.as_tensor(rng.uniform(-lim1, lim1, (c, h)).astype(np.float32)), torch.zeros(c)]
    total = sum(w.numel() for w in host_w)
    host_flat = torch.cat([w.reshape(-1) for w in host_w]).pin_memory()
    dev_flat = host_flat.to(device)
    views, off = [], 0
    for w in host_w:
        views.append(dev_flat[off:off + w.numel()].view(w.shape)); off += w.numel()
    theta = d["theta"]
    lib = _lib.load()
    if world > 1:
        eng = S.ShardedOuterStep(n, lo, cnt, d["x"], d["y"], d["mask"], d["mask_count"], h, c)
        exchange = "nccl all-gather"
        comm = None
        if os.environ.get("LDS_EXCHANGE", "peer") == "peer":      # NVLink peer-memory push (symmetric memory); NCCL if it cannot be set up
            try:
                comm = S.SymmComm(n)
                eng._setup_comm(comm)
                exchange = "NVLink peer-memory push (one 128-bit store kernel per exchange + signal-pad barrier)"
            except Exception as exc:                              # noqa: BLE001  (report, then use NCCL)
                print(f"[bench] symmetric-memory exchange unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
                comm = None
                eng._comm = None
        if comm is None:
            comm = S.DistComm(n)
        eng.set_weights(*views)
        step_fn = lambda k, lr_now: eng.run(theta, comm, lr=lr_now, seed=1234, step=k, dropout_p=HYPER["dropout"], update=True)
        scal = lambda: None
    else:
        eng = K.OuterStep(n, d["x"], d["y"], d["mask"], hidden=h, classes=c)
        eng.set_weights(*views)
        step_fn = lambda k, lr_now: eng.run(theta, lr=lr_now, seed=1234, step=k, dropout_p=HYPER["dropout"], update=True, want_adj=False)
    lr = HYPER["lr"]
    for w in range(args.warmup):
        step_fn(w, lr); lr *= HYPER["lr_decay"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall0 = time.time()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for k in range(args.steps):
        out = step_fn(args.warmup + k, lr); lr *= HYPER["lr_decay"]
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = t0.elapsed_time(t1)
    # per-kernel breakdown (rank-local)
    per_kernel = {}
    ms_buf = (ctypes.c_float * 64)(); id_buf = (ctypes.c_int32 * 64)()
    reps = min(args.steps, 5)
    for k in range(reps):
        lib.lds_profile_begin()
        step_fn(20_000 + k, lr)
        cntm = lib.lds_profile_end(ms_buf, id_buf, 64)
        for i in range(max(cntm, 0)):
            per_kernel.setdefault(int(id_buf[i]), []).append(float(ms_buf[i]))
    # e2e: through the reference-facing API — a row-block BernoulliGraphModel behind OuterProblemTrainer.train_step(model_forward)
    # (src/trainers/outer.py:57-87) — with the step's weights from pinned host memory and Metrics back on the host, every step
    from lds_gnn_b200.models.gcn import MetaDenseGCN
    from lds_gnn_b200.models.graph import BernoulliGraphModel
    from lds_gnn_b200.models.sampling import PHILOX
    from lds_gnn_b200.trainers.inner import InnerProblemTrainer
    from lds_gnn_b200.trainers.outer import OuterProblemTrainer
    from lds_gnn_b200.utils.graph import DenseData
    api_data = DenseData(x=d["x"], y=d["y"], train_mask=d["mask"], val_mask=d["mask"], test_mask=d["mask"], num_classes=c)
    api_gcn = MetaDenseGCN(f, h, c, dropout=HYPER["dropout"]).to(device)
    api_inner = InnerProblemTrainer(api_gcn, api_data)
    api_inner.model_params = type(api_inner.model_params)(zip(api_inner.model_params.keys(), views))      # the step's weights: views of dev_flat
    api_model = BernoulliGraphModel.from_row_block(theta, n, lo)
    api_outer = OuterProblemTrainer(optimizer=torch.optim.SGD(api_model.parameters(), lr=lr), data=api_data, opt_mask=d["mask"], model=api_model,
                                    smoothness_factor=0.0, disconnection_factor=0.0, sparsity_factor=0.0, regularize=False,
                                    lr_decay=HYPER["lr_decay"], pretrain=False)
    if world > 1:
        api_outer._engine = ((d["x"].data_ptr(), d["mask"].data_ptr(), h, c, lo, cnt), eng, comm)           # reuse the buffers of the timed engine
    PHILOX.manual_seed(1234)
    PHILOX.step = 30_000
    api_outer.train_step(api_inner.model_forward)                # builds the trainer's engine (untimed)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ts = time.perf_counter()
    for k in range(args.steps):
        dev_flat.copy_(host_flat, non_blocking=True)
        metrics = api_outer.train_step(api_inner.model_forward)
        loss_acc = [metrics.loss, metrics.acc]
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - ts) * 1e3
    clock_info = clocks.stop(wall0, time.time())
    dev_ms, e2e_ms = reduce_rank_times([dev_ms, e2e_ms], device, world)
    value = args.steps / (dev_ms / 1e3)
    # strong-scaling reference: the SAME workload on one GPU (rank 0), so the sharded number can be read against it
    scaling_reference = None
    if world > 1 and not args.no_scale_ref:
        if rank == 0:
            d1 = make_large_rows(workload, device, 0, n, seed=0)
            eng1 = K.OuterStep(n, d1["x"], d1["y"], d1["mask"], hidden=h, classes=c)
            eng1.set_weights(*views)
            th1 = d1["theta"]
            for k in range(2):
                eng1.run(th1, lr=0.1, seed=1234, step=k, dropout_p=HYPER["dropout"], update=True)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            ref_steps = 4
            for k in range(ref_steps):
                eng1.run(th1, lr=0.1, seed=1234, step=10 + k, dropout_p=HYPER["dropout"], update=True)
            r1.record()
            torch.cuda.synchronize()
            ref_ms = r0.elapsed_time(r1) / ref_steps
            scaling_reference = {"n_gpus": 1, "workload": f"same N={n} step on one GPU (rank 0, unsharded)", "value": round(1e3 / ref_ms, 3),
                                 "unit": UNIT, "ms_per_step": round(ref_ms, 4), "steps": ref_steps,
                                 "speedup": round(value / (1e3 / ref_ms), 3), "efficiency": round(value / (1e3 / ref_ms) / world, 4)}
            del eng1, d1, th1
            torch.cuda.empty_cache()
        dist.barrier()
    hbm_peak, peak_src = peaks()
    bf16_peak = bf16_peak_tflops()
    rows_local = cnt
    packed = not os.environ.get("LDS_BF16_PLAN")               # the bit-packed plan is the default of every step this function times
    hp1, hp2 = int(lib.lds_outer_step_operand_hp(h, c, _lib.PHASE_LAYER1)), int(lib.lds_outer_step_operand_hp(h, c, _lib.PHASE_LAYER2))
    # Bytes each kernel of THIS launch plan has to move per launch on this rank. Bit-packed plan: sampling reads the
    # rank's theta rows once per unordered tile pair (the rank's own diagonal block only once) and writes bits; a propagation
    # reads the bits (it is tensor-bound: reported as TFLOP/s, hi + lo MMAs counted); the update is 6 N^2 unsharded
    # (tile-symmetric: upper tiles read, both triangles written), 8 rows n for a row-block shard. SURVEY 8(d)'s bf16 figures
    # (6 / 2 / 8 N^2) are kept next to them as `survey_bytes`.
    k1_bytes = (4 * (rows_local * n - rows_local * rows_local // 2) + rows_local * n // 8) if packed else 6 * rows_local * n
    k2_bytes = (rows_local * n // 8) if packed else 2 * rows_local * n
    k3_bytes = 6 * n * n if world == 1 else 8 * rows_local * n
    alg = {0: k1_bytes, 3: k2_bytes, 4: k2_bytes, 5: k2_bytes, 6: k2_bytes, 7: k3_bytes}
    survey = {0: 6 * rows_local * n, 3: 2 * rows_local * n, 4: 2 * rows_local * n, 5: 2 * rows_local * n, 6: 2 * rows_local * n, 7: 8 * rows_local * n}
    k2_flops = {3: 4 * rows_local * n * hp1, 6: 4 * rows_local * n * hp1, 4: 4 * rows_local * n * hp2, 5: 4 * rows_local * n * hp2}
    kernel_summary = {}
    for kid, vals in per_kernel.items():
        if kid < 0:
            kernel_summary["exchange_and_host_gaps"] = {"step_share_us": 1e3 * sum(vals) / reps}
            continue
        mean_ms = sum(vals) / len(vals)
        entry = {"mean_us": 1e3 * mean_ms, "step_share_us": 1e3 * sum(vals) / reps}
        if kid in alg:
            entry["bytes"] = alg[kid]
            entry["survey_bytes"] = survey[kid]
            entry["GBps"] = round(alg[kid] / (mean_ms / 1e3) / 1e9, 1)
            entry["frac_of_hbm_peak"] = round(entry["GBps"] / hbm_peak, 4)
        if kid in k2_flops:
            entry["tflops"] = round(k2_flops[kid] / (mean_ms / 1e3) / 1e12, 1)
            entry["frac_of_bf16_peak"] = round(entry["tflops"] / bf16_peak, 4)
        kernel_summary[KERNEL_NAMES.get(kid, str(kid))] = entry
    cand = {k: v for k, v in per_kernel.items() if k in alg}
    dom = max(cand, key=lambda k: sum(cand[k]))
    mean_s = sum(cand[dom]) / len(cand[dom]) / 1e3
    achieved = alg[dom] / mean_s / 1e9
    step_bytes = sum(alg.values())
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "bf16 adjacency x (bf16 hi+lo) operands, fp32 accumulate; fp32 theta", "data": "synthetic",
        "config": large_config(n, f, h, c, world),
        "exchange": exchange if world > 1 else None,
        "clocks": clock_info,
        "e2e": {"value": round(args.steps / (e2e_ms / 1e3), 3), "unit": UNIT, "h2d_bytes_per_step": total * 4, "d2h_bytes_per_step": 8,
                "api": "OuterProblemTrainer.train_step(InnerProblemTrainer.model_forward) on BernoulliGraphModel.from_row_block"},
        "gpu_launches": (LAUNCHES_PER_STEP + (4 if world > 1 else 0)) * args.steps,
        "roofline": {"kernel": KERNEL_NAMES[dom], "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(achieved / hbm_peak, 4), "traffic": ncu_traffic(workload if world == 1 else None, KERNEL_NAMES[dom])[0],
                     "traffic_source": ncu_traffic(workload if world == 1 else None, KERNEL_NAMES[dom])[1], "algorithmic_bytes_per_launch": alg[dom],
                     "mean_launch_us": round(mean_s * 1e6, 1), "peak_source": peak_src, "scope": "per GPU (rank 0)"},
        "step_roofline": {"algorithmic_bytes_per_step_per_gpu": step_bytes,
                          "frac_of_hbm_peak": round(step_bytes / (dev_ms / 1e3 / args.steps) / 1e9 / hbm_peak, 4)},
        "kernels": kernel_summary, "last_metrics": {"loss": loss_acc[0], "acc": loss_acc[1]},
        "scaling_reference": scaling_reference,
        "sharding": None if world == 1 else {
            # what row-block sharding costs by construction and what the exchange costs on top, separated: the unsharded plan uses
            # theta's symmetry (sampling: one draw / one theta read per unordered tile pair; update: 6 N^2 instead of 8 N^2); a
            # row-block shard can do that only inside its own diagonal block — the mirrored tiles live on other ranks, and moving
            # theta over NVLink would cost more than it saves — so the ranks together move more bytes than one GPU does
            "bytes_all_ranks_over_unsharded": round(world * step_bytes / (2.125 * n * n + 4 * n * n / 8 + 6 * n * n), 3),
            "compute_us_per_step": round(sum(v["step_share_us"] for k, v in kernel_summary.items() if k != "exchange_and_host_gaps"), 1),
            "exchange_and_gaps_us_per_step": round(kernel_summary.get("exchange_and_host_gaps", {}).get("step_share_us", 0.0), 1),
            "parallel_efficiency": round(sum(v["step_share_us"] for k, v in kernel_summary.items() if k != "exchange_and_host_gaps")
                                         / max(1e-9, sum(v["step_share_us"] for v in kernel_summary.values())), 4),
            "note": "parallel_efficiency = rank-0 kernel time / (kernel time + exchange + launch gaps); scaling_reference.efficiency also "
                    "contains bytes_all_ranks_over_unsharded (the symmetric single-GPU plan does less work than the shards together)"},
        "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"not runnable: the reference keeps ~25 dense N x N fp32 tensors ({25 * n * n * 4 / 1e9:.0f} GB) and does "
                                   f"6 N^3 SGEMMs per step at N={n}"},
    }
    return line


# ------------------------------------------------------------------------------------------------ reference / CPU baseline
def time_cpu_port(workload, steps, warmup):
    """The oracle's op-for-op torch/CPU port of the reference outer step on this box's host cores."""
    from oracle.reference_port import ReferenceOuterStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    data, weights, opt_mask, shape = make_workload(workload, seed=0)
    n = shape["n"]
    iu = torch.triu_indices(n, n)
    theta = data.dense_adj[iu[0], iu[1]].clone()
    ref = ReferenceOuterStep(theta, data.x, data.y, opt_mask, weights["w0"], weights["b0"], weights["w1"], weights["b1"],
                             lr=HYPER["lr"], lr_decay=HYPER["lr_decay"], dropout=HYPER["dropout"])
    for _ in range(warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, acc = ref.step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, torch.get_num_threads(), shape, loss


def run_reference_large(args, workload):
    """The reference cannot run N >= 20 000 (25 dense N x N fp32 temporaries, 6 N^3 SGEMMs per step): no value, and no
    extrapolation either — a made-up denominator would only produce a made-up ratio."""
    from lds_gnn_b200.data import SHAPES
    n, f, c, h, _, _ = SHAPES[workload]
    why = (f"the reference cannot run this config: ~25 dense N x N fp32 tensors ({25 * n * n * 4 / 1e9:.0f} GB) and 6 N^3 SGEMMs per "
           f"step at N={n} on the host")
    return {"impl": "reference", "metric": METRIC, "value": None, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": large_config(n, f, h, c, args.gpus), "unavailable": why,
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": why},
            "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def run_reference(args):
    workload = args.workload
    if args.gpus > 1 and workload in ("citeseer", "cora", "tiny") and not args.replicas:
        workload = "n65k"
    if workload in ("n20k", "n65k"):
        return run_reference_large(args, workload)
    rate, ms, threads, shape, loss = time_cpu_port(args.workload, args.steps, args.warmup)
    samples = KNN_WORKLOADS[args.workload][2] if args.workload in KNN_WORKLOADS else 1
    rate, ms = rate / samples, ms * samples                  # the port runs single-sample steps: an S-sample step is S of them
    cpu = {"value": round(rate, 4), "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{args.steps} full single-sample outer steps of the same workload (oracle/reference_port.py, torch CPU, {threads} threads)"
                     + (f", rate divided by {samples} samples per outer step" if samples > 1 else "")}
    return {"impl": "reference", "metric": METRIC, "value": round(rate, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": small_config(args.workload, shape, samples, args.gpus, args.replicas),
            "cpu_baseline": cpu, "e2e": {"value": round(rate, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "final_loss": loss}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the single JSON line only
        if os.environ.get("NCCL_DEBUG", "").upper() in ("WARN", "VERSION"):
            os.environ.pop("NCCL_DEBUG")                # WARN / VERSION print "NCCL version ..." on stdout
        dist.init_process_group("nccl", device_id=device)
    workload = args.workload
    if world > 1 and workload in ("citeseer", "cora", "tiny") and not args.replicas:
        workload = "n65k"                        # the multi-GPU config of BASELINE.json: N = 65 536 row-block sharded
    if workload in ("n20k", "n65k"):
        line = run_large(args, rank, world, device, workload)
    else:
        line = run_ours(args, rank, world, device)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and "cpu_baseline" not in line:
            rate, ms, threads, _, _ = time_cpu_port(args.workload, args.cpu_steps, 1)
            note = ""
            if args.workload in KNN_WORKLOADS:           # the port runs single-sample steps; an S-sample step is S of them
                rate /= KNN_WORKLOADS[args.workload][2]
                note = f"; single-sample steps timed, rate divided by {KNN_WORKLOADS[args.workload][2]} samples per outer step"
            line["cpu_baseline"] = {"value": round(rate, 4), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_steps} full outer steps of the same workload after 1 warm-up "
                                              f"(oracle/reference_port.py: the reference's torch op sequence on CPU){note}"}
        if world == 1 and not args.no_bilevel_block and workload in ("citeseer", "cora", "tiny"):
            try:
                line["bilevel_block"] = time_bilevel_block(workload, device)
            except Exception as exc:                     # an extra, never the headline: report instead of failing the line
                line["bilevel_block"] = {"error": f"{type(exc).__name__}: {exc}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
