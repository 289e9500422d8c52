#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -3
python bench.py --workload n20k --steps 10 --warmup 3 > gpurun_out/r2o_bench_n20k.json 2> gpurun_out/r2o_bench_n20k.err; tail -2 gpurun_out/r2o_bench_n20k.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2o_bench_n20k.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e")})
PY
