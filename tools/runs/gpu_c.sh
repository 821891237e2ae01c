#!/bin/bash
# 2-GPU call: multi-GPU correctness test, driver-like bench at N=2 (c4 strong + c3 weak secondary), K4b probe
set -x
mkdir -p gpurun_out/c
nvidia-smi topo -m > gpurun_out/c/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_moments.py -m gpu -x -q > gpurun_out/c/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c/pytest_multi.log
tail -15 gpurun_out/c/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c/bench_n2_driverlike.json 2> gpurun_out/c/bench_n2_driverlike.err; echo "rc=$?"
tail -c 1200 gpurun_out/c/bench_n2_driverlike.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 600 --warmup 5 --no-k4b > gpurun_out/c/bench_n2_600.json 2> gpurun_out/c/bench_n2_600.err; echo "rc=$?"
timeout 300 python tools/k4b_probe.py > gpurun_out/c/k4b_probe.json 2> gpurun_out/c/k4b_probe.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c/bench_n2_driverlike.json","gpurun_out/c/bench_n2_600.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.4f closes %s merge %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["rollout_closes_timed"], d["merge_check"]), d["exchange"], d.get("secondary"), d["config"]["workload"][:20])
    except Exception as e:
        print(f, "ERR", e)
PY
head -c 1500 gpurun_out/c/k4b_probe.json
