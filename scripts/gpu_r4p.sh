#!/bin/bash
# final verification of round 2: full GPU suite, smoke, default bench (as the driver runs it) and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r4p_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r4p_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r4p_bench_driver_style.json 2> gpurun_out/r4p_bench.err; echo "bench(20,5) rc=$?"
timeout 600 python bench.py > gpurun_out/r4p_bench_citeseer.json 2>> gpurun_out/r4p_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r4p_bench_reference.json 2>> gpurun_out/r4p_bench.err; echo "reference rc=$?"
python - <<'P'
import json
for f in ("driver_style","citeseer"):
    d=json.loads(open(f'gpurun_out/r4p_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'warm', d['warm_l2']['ms_per_step'], 'cold', d['flushed_cold_code']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['mean_launch_us'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'block', (d.get('bilevel_block') or {}).get('ms_per_block'), 'clocks', d['clocks'])
r=json.loads(open('gpurun_out/r4p_bench_reference.json').read().strip().splitlines()[-1])
o=json.loads(open('gpurun_out/r4p_bench_citeseer.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['steps'], r['warmup'], 'same config:', r['config']==o['config'])
P
