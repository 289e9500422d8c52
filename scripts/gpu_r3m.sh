#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tests/accuracy_parity.py --shape cora --seeds 5 2>/dev/null | tail -14
timeout 1200 python tests/accuracy_parity.py --shape citeseer --seeds 5 2>/dev/null | tail -14
