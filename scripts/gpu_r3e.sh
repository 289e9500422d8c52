#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_outer_step.py tests/test_golden_next.py -m gpu -q -x > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3e_pytest.log
echo "== warm"; timeout 120 python scripts/fused_timeline.py citeseer 2>/dev/null | sed -n 2,16p
bash scripts/gpu_ab.sh
