#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 > gpurun_out/r2n_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_small_kernel -s 4 -c 1 -f -o gpurun_out/r2n_fused python bench.py --steps 6 --warmup 3 > gpurun_out/r2n_ncu.log 2>&1
ls -la gpurun_out/r2n_fused.ncu-rep; tail -2 gpurun_out/r2n_ncu.log
