#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_packed.py -q -x > gpurun_out/r2i_packed.log 2>&1; echo "packed tests rc=$?"; tail -4 gpurun_out/r2i_packed.log
python scripts/bench_packed.py 20000 > gpurun_out/r2i_bench_packed_20k.log 2>&1; cat gpurun_out/r2i_bench_packed_20k.log
python scripts/bench_packed.py 65536 8192 > gpurun_out/r2i_bench_packed_65k.log 2>&1; cat gpurun_out/r2i_bench_packed_65k.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2i_launches_packed_20k.csv python scripts/bench_packed.py 20000 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2i_launches_packed_20k.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:70]].append(float(r[-1]))
for k, v in agg.items():
    if "lds::" in k:
        print(f"{k:72s} n={len(v):3d} min={min(v)/1e3:9.1f} us median={sorted(v)[len(v)//2]/1e3:9.1f} us")
PY
