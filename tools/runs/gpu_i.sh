#!/bin/bash
# 1-GPU call: ncu evidence of the round (launch list of the driver-like bench, one --set full capture of every kernel)
# + wave-quantisation probe of k_step
set -x
mkdir -p gpurun_out/i
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/i/bench_steps20.json 2> gpurun_out/i/bench_steps20.err; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/i/launches_c3_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/i/ncu_launches.log 2>&1; echo "rc=$?"
timeout 300 python tools/ncu_target.py > gpurun_out/i/ncu_target_plain.log 2>&1; echo "target rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/i/ncu_r2_full python tools/ncu_target.py > gpurun_out/i/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/i/
for E in 37888 65536 75776; do timeout 300 python tools/kstep_ab.py --envs $E --label "E=$E" >> gpurun_out/i/kstep_wave_probe.txt 2>&1; done
cat gpurun_out/i/kstep_wave_probe.txt
