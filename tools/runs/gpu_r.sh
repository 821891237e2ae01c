#!/bin/bash
set -x
mkdir -p gpurun_out/r
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r/pytest.log
tail -6 gpurun_out/r/pytest.log
timeout 600 python tools/soak.py --envs 4096 --episodes 25 > gpurun_out/r/soak_small.txt 2>&1; echo "soak rc=$?"
tail -1 gpurun_out/r/soak_small.txt
