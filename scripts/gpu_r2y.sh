#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_block.py -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2y_pytest.log
python scripts/ncu_graph_block.py --help 2>/dev/null | head -3
