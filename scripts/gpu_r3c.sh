#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py $2 --steps $3 --warmup 3 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k: (round(v['mean_us'],1), v.get('frac_of_hbm_peak')) for k,v in d['kernels'].items() if 'k3' in k})"; }
run persist_sym_n65k "--workload n65k" 3
LDS_K3_PERSIST=0 run nopersist_sym_n65k "--workload n65k" 3
LDS_K3_FULL=1 run persist_full_n65k "--workload n65k" 3
LDS_K3_FULL=1 LDS_K3_PERSIST=0 run nopersist_full_n65k "--workload n65k" 3
run citeseer_copystream "" 30
run citeseer_copystream "" 30
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
