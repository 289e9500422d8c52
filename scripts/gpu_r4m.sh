#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 300 python bench.py --steps $2 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'])"; }
run side 200
LDS_BENCH_E2E_UPLOAD=stream run stream 200
run side 200
LDS_BENCH_E2E_UPLOAD=stream run stream 200
run side 50
LDS_BENCH_E2E_UPLOAD=stream run stream 50
