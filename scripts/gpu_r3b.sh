#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r3b_bench_2gpu.json 2> gpurun_out/r3b_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r3b_bench_2gpu.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3b_bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","e2e","scaling_reference")})
for k,v in d.get("kernels",{}).items(): print(k, {kk:round(vv,1) if isinstance(vv,float) else vv for kk,vv in v.items()})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 8 --warmup 3 | cut -c1-400
