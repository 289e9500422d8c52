#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_block.py tests/test_gpu_api.py tests/test_golden_next.py -m gpu -q > gpurun_out/r3g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3g_pytest.log
timeout 600 python scripts/time_bilevel_block.py citeseer 20 2>/dev/null | head -1
LDS_NO_PDL=1 timeout 600 python scripts/time_bilevel_block.py citeseer 20 2>/dev/null | head -1
timeout 600 python scripts/time_bilevel_block.py citeseer 20 2>/dev/null | head -1
LDS_NO_PDL=1 timeout 600 python scripts/time_bilevel_block.py citeseer 20 2>/dev/null | head -1
