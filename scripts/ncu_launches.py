"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) as a markdown table of our kernels:
   python scripts/ncu_launches.py <launches.csv> <steps in the run> <out.md> "<title line>" """
import collections, csv, sys
path, steps, out, title = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
rows = list(csv.reader(open(path)))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
idx = {h: j for j, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v, u = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
    v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    name = r[idx["Kernel Name"]]
    if "lds::" not in name:
        continue
    agg.setdefault(name.split("(")[0], []).append(v)
total = sum(sum(v) for v in agg.values())
lines = [title, "", "| kernel | launches | mean us | per step us | share of our kernels |", "|---|---|---|---|---|"]
for k, v in agg.items():
    if len(v) < steps:        # setup-only kernels (theta layout conversion)
        continue
    lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {sum(v) / steps:.2f} | {100 * sum(v) / total:.1f}% |")
lines.append("")
lines.append(f"Sum of our kernels per step: {sum(sum(v) for v in agg.values() if len(v) >= steps) / steps:.1f} us "
             f"({steps} steps in the capture: warm-up + timed + profile + e2e passes of bench.py).")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
