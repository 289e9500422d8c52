#!/bin/bash
# round-2 session 3: full GPU suite + bench lines + ncu evidence of the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r4i_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r4i_pytest.log
timeout 600 python bench.py > gpurun_out/r4i_bench_citeseer.json 2> gpurun_out/r4i_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload cora_knn16 --no-bilevel-block > gpurun_out/r4i_bench_cora_knn16.json 2>> gpurun_out/r4i_bench.err; echo "bench knn16 rc=$?"
timeout 600 python bench.py --workload cora --no-bilevel-block --no-cpu-baseline > gpurun_out/r4i_bench_cora.json 2>> gpurun_out/r4i_bench.err; echo "bench cora rc=$?"
B="--no-cpu-baseline --no-bilevel-block"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4i_launches_citeseer.csv python bench.py --steps 3 --warmup 3 $B > gpurun_out/r4i_ncu1.log 2>&1; echo "launch list citeseer rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'fused_small_kernel|k3_tc_kernel' -s 8 -c 2 -f -o gpurun_out/r4i_full_citeseer python bench.py --steps 3 --warmup 3 $B > gpurun_out/r4i_ncu2.log 2>&1; echo "full citeseer rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4i_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4i_ncu6.log 2>&1; echo "smoke under ncu rc=$?"; tail -2 gpurun_out/r4i_ncu6.log
python - <<'P'
import json
for f in ("citeseer","cora_knn16","cora"):
    d=json.loads(open(f'gpurun_out/r4i_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'], 'cold', d['flushed_cold_code']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['mean_launch_us'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'block', (d.get('bilevel_block') or {}).get('ms_per_block'))
P
