#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r3a_bench_citeseer.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3a_bench_citeseer.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","steps","warmup")}, d["e2e"], d["roofline"]["frac"], d["kernels"], d["cpu_baseline"]["value"], d.get("bilevel_block",{}).get("ms_per_block"))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 | cut -c1-600
