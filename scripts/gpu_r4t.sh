#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r4s_bench_8gpu.json 2> gpurun_out/r4s_bench_8gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r4s_bench_8gpu.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4s_bench_8gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","e2e","scaling_reference","sharding")})
PY
