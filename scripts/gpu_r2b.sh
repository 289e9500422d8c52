#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_smoke.csv \
   python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2b_ncu_smoke.log 2>&1; echo "ncu smoke rc=$?" | tee -a gpurun_out/r2b_ncu_smoke.log
grep -o 'lds::[a-z0-9_]*' gpurun_out/r2b_launches_smoke.csv | sort | uniq -c
tail -15 gpurun_out/r2b_pytest.log
