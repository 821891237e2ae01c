#!/bin/bash
# 2-GPU confirmation of the final code: multi-GPU test, driver-like bench at N=2, soak on GPU 0
set -x
mkdir -p gpurun_out/q
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/q/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q/pytest_multi.log
tail -3 gpurun_out/q/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/q/bench_n2_steps20.json 2> gpurun_out/q/bench_n2_steps20.err; echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --impl reference --steps 20 --warmup 5 > gpurun_out/q/bench_n2_reference.json 2> gpurun_out/q/bench_n2_reference.err; echo "rc=$?"
timeout 900 python tools/soak.py --episodes 13 > gpurun_out/q/soak.txt 2>&1; echo "soak rc=$?"
tail -2 gpurun_out/q/soak.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/q/bench_n2_steps20.json").read().strip().splitlines()[-1])
s=d["secondary"]["c3_weak"]
print("N=2 value %.4g steady %.4g e2e %.4g ms %.4f closes %s merge %s" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"], d["rollout_closes_timed"], d["merge_check"]), d["exchange"]["peer_us"], d["exchange"]["nccl_us"], "c3w %.4g e2e %.4g" % (s["value"], s["e2e_value"]))
d=json.loads(open("gpurun_out/q/bench_n2_reference.json").read().strip().splitlines()[-1])
print("ref", d["value"], d["config"]["workload"][:20])
PY
