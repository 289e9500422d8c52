#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_golden_next.py tests/test_gpu_api.py tests/test_gpu_script.py -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
python scripts/bench_eval.py cora 2>&1 | tail -2
python scripts/bench_eval.py citeseer 2>&1 | tail -2
