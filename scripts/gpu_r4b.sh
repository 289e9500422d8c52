#!/bin/bash
# leader epilogues with preloaded row state + per-k-block operand barriers: tests, timeline, A/B against the previous build
timeout 900 python -m pytest tests/test_gpu_outer_step.py tests/test_gpu_api.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
python scripts/fused_timeline.py citeseer 2>/dev/null | tail -32
bash scripts/gpu_ab.sh
