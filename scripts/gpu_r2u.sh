#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fused_small_kernel -s 4 -c 1 -f -o gpurun_out/r2u_fused python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-bilevel-block > gpurun_out/r2u_ncu.log 2>&1
ls -la gpurun_out/r2u_fused.ncu-rep; tail -2 gpurun_out/r2u_ncu.log | cut -c1-300
