#!/bin/bash
run() { timeout 300 python bench.py --workload $1 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'fused_us', d['roofline']['mean_launch_us'])"; }
timeout 600 python -m pytest tests/test_gpu_outer_step.py tests/test_golden_next.py -m gpu -q -x 2>&1 | tail -2
echo "== first plan + dry-run epilogues"; run citeseer; run cora; run citeseer
echo "== COLD timeline (citeseer)"; COLD=1 python scripts/fused_timeline.py citeseer 2>/dev/null | tail -14
