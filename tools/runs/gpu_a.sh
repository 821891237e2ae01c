#!/bin/bash
# round-2 GPU call A (1 GPU): tests, driver-like bench, default bench, reference arm
set -x
mkdir -p gpurun_out/a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a/smi.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a/pytest.log
tail -5 gpurun_out/a/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a/bench_driverlike.json 2> gpurun_out/a/bench_driverlike.err; echo "rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a/bench_reference.json 2> gpurun_out/a/bench_reference.err; echo "rc=$?"
timeout 900 python bench.py > gpurun_out/a/bench_default.json 2> gpurun_out/a/bench_default.err; echo "rc=$?"
tail -c 1500 gpurun_out/a/bench_driverlike.err
head -c 3000 gpurun_out/a/bench_driverlike.json
