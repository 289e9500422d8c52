#!/bin/bash
# round 2, call A: GPU test-suite + smoke under ncu diagnostics
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2a_smoke.log
env | sort > gpurun_out/r2a_env_plain.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_smoke.csv \
   python -c 'import os; open("gpurun_out/r2a_env_ncu.txt","w").write("\n".join(sorted(f"{k}={v}" for k,v in os.environ.items()))); import __graft_entry__ as g; g.smoke()' > gpurun_out/r2a_ncu_smoke.log 2>&1; echo "ncu smoke rc=$?" | tee -a gpurun_out/r2a_ncu_smoke.log
LDS_FUSED_NO_CLUSTER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_smoke_nocluster.csv \
   python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2a_ncu_smoke_nocluster.log 2>&1; echo "ncu smoke nocluster rc=$?" | tee -a gpurun_out/r2a_ncu_smoke_nocluster.log
tail -5 gpurun_out/r2a_pytest.log
