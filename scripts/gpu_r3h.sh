#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_kernels.py tests/test_golden_next.py tests/test_gpu_api.py -m gpu -q -x > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3h_pytest.log
bash scripts/gpu_ab.sh
