#!/bin/bash
mkdir -p gpurun_out
LDS_K3_FULL=1 python bench.py --workload n65k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_n65k_full.json 2>gpurun_out/r2w_err.log; echo rc=$?
python bench.py --workload n65k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_n65k_sym.json 2>>gpurun_out/r2w_err.log; echo rc=$?
LDS_K3_FULL=1 ncu --set full --clock-control none --import-source on -k regex:k3_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2w_k3_n65k python bench.py --workload n65k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_ncu.log 2>&1
ls -la gpurun_out/r2w_k3_n65k.ncu-rep
python - <<'PY'
import json
for f in ("full","sym"):
    d=json.loads(open(f"gpurun_out/r2w_n65k_{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], {k:(round(v["mean_us"],1), v.get("frac_of_hbm_peak"), v.get("frac_of_bf16_peak")) for k,v in d["kernels"].items()})
PY
