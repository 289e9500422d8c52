#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2j_pytest.log
python bench.py --workload n20k --steps 10 --warmup 3 > gpurun_out/r2j_bench_n20k.json 2> gpurun_out/r2j_bench_n20k.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench_n20k.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e")})
for k,v in d.get("kernels",{}).items(): print(k, {kk:round(vv,1) if isinstance(vv,float) else vv for kk,vv in v.items()})
PY
tail -3 gpurun_out/r2j_bench_n20k.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_citeseer.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench_citeseer.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","roofline")})
PY
