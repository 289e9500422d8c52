#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_outer_step.py tests/test_gpu_block.py tests/test_gpu_script.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 300 python bench.py --steps $2 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'])"; }
run steady 200
LDS_NO_STEADY=1 run nosteady 200
run steady 200
LDS_NO_STEADY=1 run nosteady 200
run steady 50
python scripts/profile_e2e.py 2>/dev/null | head -14
