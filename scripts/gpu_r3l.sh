#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_packed.py tests/test_gpu_fullsize.py tests/test_gpu_kernels.py tests/test_golden_next.py -m gpu -q -x > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3l_pytest.log
run() { timeout 300 python bench.py --workload $2 --steps $3 --warmup 3 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', '$2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"; }
run pdl n20k 20
LDS_NO_PDL=1 run nopdl n20k 20
run pdl n20k 20
LDS_NO_PDL=1 run nopdl n20k 20
run pdl cora_knn16 20
LDS_NO_PDL=1 run nopdl cora_knn16 20
