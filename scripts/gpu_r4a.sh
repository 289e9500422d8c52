#!/bin/bash
# round-2 session 3: full GPU suite on HEAD, bench with the two-instance protocol, K2 grid cap A/B inside the bilevel block
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r4a_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r4a_pytest.log
timeout 600 python bench.py > gpurun_out/r4a_bench_citeseer.json 2> gpurun_out/r4a_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4a_bench_citeseer.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['warm_l2'], d['flushed_cold_code'], d['roofline']['frac'], d['roofline']['mean_launch_us'], d['kernels'], d.get('bilevel_block',{}).get('ms_per_block'))
P
for w in citeseer cora; do for i in 1 2; do
  echo -n "$w cap148 "; timeout 600 python scripts/time_bilevel_block.py $w 30 2>/dev/null | head -1
  echo -n "$w grid296 "; LDS_K2_GRID_MAX=296 timeout 600 python scripts/time_bilevel_block.py $w 30 2>/dev/null | head -1
done; done
