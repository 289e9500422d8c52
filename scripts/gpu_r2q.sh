#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_outer_step.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest.log
for v in ts ss; do
  if [ $v = ss ]; then export LDS_K3_SS=1; fi
  echo "== K3 $v"
  python bench.py --workload n20k --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('n20k', d['value'], d['ms_per_step'], 'k3', d['kernels']['k3k4_theta_update'])"
  python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('citeseer', d['value'], d['ms_per_step'], {k:v for k,v in d.get('kernels',{}).items() if 'k3' in k})"
done
