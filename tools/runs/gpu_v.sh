#!/bin/bash
# final ncu evidence (code as committed at the end of round 2)
set -x
mkdir -p gpurun_out/v
timeout 300 python tools/ncu_target.py > gpurun_out/v/ncu_target_plain.log 2>&1; echo "target rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/v/ncu_r2_final python tools/ncu_target.py > gpurun_out/v/ncu_full.log 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/v/launches_c3_steps20_final.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/v/ncu_launches.log 2>&1; echo "rc=$?"
ls -la gpurun_out/v
