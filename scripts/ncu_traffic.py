"""Turn an `ncu --set full` report into the committed evidence bench.py cites:
   python scripts/ncu_traffic.py <report.ncu-rep> <workload> <profiles/summary.md>
 * profiles/ncu_traffic.json[workload][kernel] = {"dram_bytes": read + write per launch (mean over captured launches), ...}
 * a markdown table with per-kernel duration, DRAM bytes, DRAM / L2 / tensor-pipe utilisation (ncu's own peaks).
Run here (no GPU needed): `ncu -i` only reads the report."""
import csv, io, json, os, re, subprocess, sys

rep, workload, md_path = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
EPI = {"1": "k2_layer1", "2": "k2_layer2", "3": "k2_bwd2", "4": "k2_bwd1", "0": "k2_plain"}


def bench_name(kname):
    if "k1_sample_kernel" in kname: return "k1_sample_normalize"
    if "k1p_sample_kernel" in kname: return "k1p_sample_normalize_packed"
    if "k2p_mma_kernel" in kname: return "k2p_propagate_packed"
    if "k3_tc_kernel" in kname: return "k3k4_theta_update"
    if "fused_small_kernel" in kname: return "fused_k1_feat_k2x4"
    if "feat_sparse_kernel" in kname or "feat_linear_kernel" in kname: return "feat_linear"
    m = re.search(r"k2_mma_kernel<\(int\)(\d+), \(int\)(\d+)>", kname) or re.search(r"k2_mma_kernel<(\d+), (\d+)>", kname)
    if m: return EPI.get(m.group(2), "k2")
    return kname.split("(")[0]


def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    v = float(v)
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(u, 1)


cols = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "grid": "launch__grid_size", "block": "launch__block_size", "hit": "lts__t_sector_hit_rate.pct"}
agg = {}
for r in data:
    name = bench_name(r[idx["Kernel Name"]])
    e = agg.setdefault(name, {"n": 0, "dur": 0.0, "rd": 0.0, "wr": 0.0, "dram": 0.0, "lts": 0.0, "tensor": 0.0, "hit": 0.0, "sass": r[idx["Kernel Name"]].split("(")[0]})
    e["n"] += 1
    e["dur"] += to_us(r[idx[cols["dur"]]], units[idx[cols["dur"]]])
    e["rd"] += to_bytes(r[idx[cols["rd"]]], units[idx[cols["rd"]]])
    e["wr"] += to_bytes(r[idx[cols["wr"]]], units[idx[cols["wr"]]])
    for k in ("dram", "lts", "tensor", "hit"):
        e[k] += float(r[idx[cols[k]]])
    e["regs"], e["grid"], e["block"] = r[idx[cols["regs"]]], r[idx[cols["grid"]]], r[idx[cols["block"]]]

jpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
db = json.load(open(jpath)) if os.path.exists(jpath) else {}
db.setdefault(workload, {})
lines = [f"# ncu --set full, workload `{workload}`, report `{os.path.basename(rep)}` (cache flushed before every replay, --clock-control none)", "",
         "| kernel | launches | mean us | DRAM read MB | DRAM write MB | DRAM % of ncu peak | L2 % | L2 hit % | tensor pipe % | regs | grid x block |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
for name, e in agg.items():
    n = e["n"]
    db[workload][name] = {"dram_bytes": int((e["rd"] + e["wr"]) / n), "dram_read_bytes": int(e["rd"] / n), "dram_write_bytes": int(e["wr"] / n),
                          "duration_us_under_ncu": round(e["dur"] / n, 2), "launches": n, "source": os.path.relpath(md_path, ROOT)}
    lines.append(f"| `{e['sass']}` ({name}) | {n} | {e['dur'] / n:.1f} | {e['rd'] / n / 1e6:.1f} | {e['wr'] / n / 1e6:.1f} | {e['dram'] / n:.1f} | "
                 f"{e['lts'] / n:.1f} | {e['hit'] / n:.1f} | {e['tensor'] / n:.1f} | {e['regs']} | {e['grid']} x {e['block']} |")
json.dump(db, open(jpath, "w"), indent=1, sort_keys=True)
open(md_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
