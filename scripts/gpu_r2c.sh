#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_packed.py -q -x > gpurun_out/r2c_packed.log 2>&1; echo "packed tests rc=$?"; tail -25 gpurun_out/r2c_packed.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_packed.py > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2c_pytest.log
python scripts/bench_packed.py 20000 > gpurun_out/r2c_bench_packed_20k.log 2>&1; cat gpurun_out/r2c_bench_packed_20k.log
python scripts/bench_packed.py 65536 8192 > gpurun_out/r2c_bench_packed_65k.log 2>&1; cat gpurun_out/r2c_bench_packed_65k.log
python bench.py --workload n20k --steps 10 --warmup 3 > gpurun_out/r2c_bench_n20k.json 2> gpurun_out/r2c_bench_n20k.err; tail -c 1500 gpurun_out/r2c_bench_n20k.json; tail -3 gpurun_out/r2c_bench_n20k.err
