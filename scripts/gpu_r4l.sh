#!/bin/bash
# weights in pinned host memory read by the fused kernel (no staging copy): parity tests + e2e A/B against the explicit copy
timeout 900 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'], 'coldcode', d['flushed_cold_code']['ms_per_step'], {k: round(v['mean_us'],1) for k,v in d['kernels'].items()})"; }
run zerocopy
LDS_BENCH_E2E_COPY=1 run copy
run zerocopy
LDS_BENCH_E2E_COPY=1 run copy
cp lds-gnn_b200/lib/liblds_b200.so /tmp/new.so
cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; LDS_BENCH_E2E_COPY=1 run oldlib_copy
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so
timeout 300 python bench.py --workload cora_knn16 --steps 50 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('knn16 zerocopy', d['value'], 'e2e', d['e2e']['value'])"
LDS_BENCH_E2E_COPY=1 timeout 300 python bench.py --workload cora_knn16 --steps 50 --warmup 5 --no-cpu-baseline --no-bilevel-block 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('knn16 copy', d['value'], 'e2e', d['e2e']['value'])"
