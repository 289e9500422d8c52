#!/bin/bash
# dry run of the leader's row epilogues under the MMAs (instruction-cache warm-up): tests, timelines, A/B against the previous build
timeout 900 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -2
for d in 18 6 2; do echo "== LDS_FUSED_DRY_FROM=$d"; LDS_FUSED_DRY_FROM=$d python scripts/fused_timeline.py citeseer 2>/dev/null | grep -E "epi._done|after.|barrier2|end|ph0" ; done
run() { timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'], 'coldcode', d['flushed_cold_code']['ms_per_step'], {k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"; }
cp lds-gnn_b200/lib/liblds_b200.so /tmp/new.so
run new6
LDS_FUSED_DRY_FROM=2 run new2
LDS_FUSED_DRY_FROM=18 run new18
cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; run old
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so; run new6
LDS_FUSED_DRY_FROM=2 run new2
