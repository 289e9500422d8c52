#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_packed.py -q -x > gpurun_out/r2d_packed.log 2>&1; echo "packed tests rc=$?"; tail -15 gpurun_out/r2d_packed.log
python scripts/bench_packed.py 20000 > gpurun_out/r2d_bench_packed_20k.log 2>&1; cat gpurun_out/r2d_bench_packed_20k.log
python scripts/bench_packed.py 65536 8192 > gpurun_out/r2d_bench_packed_65k.log 2>&1; cat gpurun_out/r2d_bench_packed_65k.log
python scripts/bench_packed.py 3327 > gpurun_out/r2d_bench_packed_3327.log 2>&1; cat gpurun_out/r2d_bench_packed_3327.log
