#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_api.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -2
cp lds-gnn_b200/lib/liblds_b200.so /tmp/new.so
for v in old new; do
  if [ $v = old ]; then cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; else cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so; fi
  echo "== $v"; python scripts/fused_timeline.py citeseer 2>/dev/null | grep -E "sampled|barrier1|feat_done|barrier2|end |tiles_done"
done
run() { timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-bilevel-block $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'], 'coldcode', d['flushed_cold_code']['ms_per_step'], {k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"; }
run new
cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; run old
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so; run new
cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; run old
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so; run new_cora "--workload cora"
cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; run old_cora "--workload cora"
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so
