#!/bin/bash
mkdir -p gpurun_out
echo "== warm"; python scripts/fused_timeline.py citeseer 2>/dev/null
echo "== cold"; COLD=1 python scripts/fused_timeline.py citeseer 2>/dev/null
