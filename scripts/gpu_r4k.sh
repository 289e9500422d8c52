#!/bin/bash
# 2-GPU evidence with the final code: multi-rank parity tests (skipped on the driver's 1-GPU test box) + the sharded bench line
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -v -x 2>&1 | tail -15 > gpurun_out/r4k_pytest_multi_2gpu.log; cat gpurun_out/r4k_pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r4k_bench_2gpu.json 2> gpurun_out/r4k_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r4k_bench_2gpu.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r4k_bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","e2e","scaling_reference","sharding")})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | cut -c1-600
