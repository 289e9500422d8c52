#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2x_pytest.log
run() { timeout 300 python bench.py $2 --steps $3 --warmup 3 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], {k: (round(v['mean_us'],1), v.get('frac_of_hbm_peak')) for k,v in d['kernels'].items() if 'k3' in k})"; }
run sym_n65k "--workload n65k" 3
LDS_K3_FULL=1 run full_n65k "--workload n65k" 3
run sym_n20k "--workload n20k" 10
LDS_K3_FULL=1 run full_n20k "--workload n20k" 10
run sym_citeseer "" 20
