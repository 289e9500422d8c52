#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r3f_bench_${N}gpu.json 2> gpurun_out/r3f_bench_${N}gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r3f_bench_${N}gpu.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r3f_bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","scaling_reference","sharding")})
print({k:round(v.get("step_share_us",0),1) for k,v in d.get("kernels",{}).items()})
PY
