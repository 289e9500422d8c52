#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2s_pytest.log
echo "== warm"; timeout 120 python scripts/fused_timeline.py citeseer 2>/dev/null
echo "== cold"; COLD=1 timeout 120 python scripts/fused_timeline.py citeseer 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('citeseer', d['value'], d['ms_per_step'], d['e2e']['value'], d['warm_l2'], d['kernels'])"
