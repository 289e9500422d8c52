#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_api.py tests/test_golden_next.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2l_pytest.log
for v in "" 1; do
  echo "== LDS_FUSED_V1=$v"
  LDS_FUSED_V1=$v timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['mean_launch_us'])"
  LDS_FUSED_V1=$v timeout 300 python bench.py --workload cora --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cora', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['mean_launch_us'])"
done
