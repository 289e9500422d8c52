#!/bin/bash
run() { timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'], {k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"; }
run cluster
LDS_FUSED_NO_CLUSTER=1 run no_cluster
run cluster
LDS_FUSED_NO_CLUSTER=1 run no_cluster
