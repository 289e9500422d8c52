#!/bin/bash
# round-2 final evidence: launch lists + ncu --set full of the dominant kernels (each only after the same command ran clean)
mkdir -p gpurun_out
B="--no-cpu-baseline --no-bilevel-block"
python bench.py --steps 3 --warmup 3 $B > gpurun_out/r3i_plain_citeseer.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3i_launches_citeseer.csv python bench.py --steps 3 --warmup 3 $B > gpurun_out/r3i_ncu1.log 2>&1; echo "launch list citeseer rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'fused_small_kernel|k3_tc_kernel' -s 4 -c 2 -f -o gpurun_out/r3i_full_citeseer python bench.py --steps 3 --warmup 3 $B > gpurun_out/r3i_ncu2.log 2>&1; echo "full citeseer rc=$?"
python bench.py --workload n20k --steps 3 --warmup 3 $B > gpurun_out/r3i_plain_n20k.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3i_launches_n20k.csv python bench.py --workload n20k --steps 3 --warmup 3 $B > gpurun_out/r3i_ncu3.log 2>&1; echo "launch list n20k rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k1p_sample_kernel|k2p_mma_kernel|k3_tc_kernel' -s 6 -c 6 -f -o gpurun_out/r3i_full_n20k python bench.py --workload n20k --steps 3 --warmup 3 $B > gpurun_out/r3i_ncu4.log 2>&1; echo "full n20k rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3i_launches_graph_block.csv python scripts/ncu_graph_block.py citeseer > gpurun_out/r3i_ncu5.log 2>&1; echo "block launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3i_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3i_ncu6.log 2>&1; echo "smoke under ncu rc=$?"; tail -2 gpurun_out/r3i_ncu6.log
ls -la gpurun_out/r3i_*
