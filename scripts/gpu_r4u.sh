#!/bin/bash
run() { timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'warm', d['warm_l2']['ms_per_step'], {k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"; }
run grid148
LDS_K3_GRID=117 run grid117
LDS_K3_GRID=130 run grid130
LDS_K3_GRID=140 run grid140
run grid148
LDS_K3_GRID=117 run grid117
