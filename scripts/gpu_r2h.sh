#!/bin/bash
for v in 0 3 7 11 15 31 4 8 16; do echo "== LDS_K2P_DEBUG=$v"; LDS_K2P_DEBUG=$v python scripts/bench_packed.py 20000 2>&1 | grep -E "k2_packed"; done
