#!/bin/bash
# A/B of two builds of the library in ONE box: bench lines of the new build, the old build (ab/liblds_old.so), then the new one again
run() { timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'], {k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"; }
cp lds-gnn_b200/lib/liblds_b200.so /tmp/new.so
run new
cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; run old
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so; run new
cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; run old
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so
