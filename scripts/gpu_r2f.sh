#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_packed.py 20000 > gpurun_out/r2f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1p_sample_kernel -s 1 -c 1 -f -o gpurun_out/r2f_k1p python scripts/bench_packed.py 20000 > gpurun_out/r2f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2p_mma_kernel -s 1 -c 1 -f -o gpurun_out/r2f_k2p python scripts/bench_packed.py 20000 > gpurun_out/r2f_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2f_ncu1.log gpurun_out/r2f_ncu2.log
