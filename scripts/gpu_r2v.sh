#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_outer_step.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest.log
run() { timeout 300 python bench.py $2 --steps 20 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"; }
run sym_citeseer ""
LDS_K3_FULL=1 run full_citeseer ""
run sym_n20k "--workload n20k"
LDS_K3_FULL=1 run full_n20k "--workload n20k"
