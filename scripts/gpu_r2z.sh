#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_graph_block.csv python scripts/ncu_graph_block.py citeseer > gpurun_out/r2z_ncu.log 2>&1; echo rc=$?
wc -l gpurun_out/r2z_launches_graph_block.csv
