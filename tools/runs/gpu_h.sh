#!/bin/bash
# 8-GPU call: multi-GPU correctness at world 8, driver-like bench (c4 strong + c3 weak), longer windows at N=8 and N=4,
# config 5 at scale (PPO rollout, 131,072 envs per GPU)
set -x
mkdir -p gpurun_out/h
nvidia-smi topo -m > gpurun_out/h/topo.txt 2>&1
nproc > gpurun_out/h/nproc.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/h/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h/pytest_multi.log
tail -6 gpurun_out/h/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/h/bench_n8_steps20.json 2> gpurun_out/h/bench_n8_steps20.err; echo "rc=$?"
tail -c 800 gpurun_out/h/bench_n8_steps20.err
timeout 600 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 600 --warmup 5 --no-k4b > gpurun_out/h/bench_n8_steps600.json 2> gpurun_out/h/bench_n8_steps600.err; echo "rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 600 --warmup 5 --no-k4b > gpurun_out/h/bench_n4_steps600.json 2> gpurun_out/h/bench_n4_steps600.err; echo "rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29524 examples/ppo_rollout.py --envs 131072 --horizon 16 --chunk 32768 --minibatches 64 --iters 2 > gpurun_out/h/ppo_8gpu_131072.json 2> gpurun_out/h/ppo_8gpu_131072.err; echo "rc=$?"
tail -c 1500 gpurun_out/h/ppo_8gpu_131072.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/h/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=(d.get("secondary") or {}).get("c3_weak",{})
        print(f, "value %.4g steady %.4g e2e %.4g ms %.4f closes %s merge %s" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"], d["rollout_closes_timed"], d["merge_check"]), d["exchange"]["peer_us"], d["exchange"]["nccl_us"], "c3w %.4g e2e %.4g" % (s.get("value",0), s.get("e2e_value",0)))
    except Exception as e:
        print(f, "ERR", e)
try:
    d=json.loads(open("gpurun_out/h/ppo_8gpu_131072.json").read().strip().splitlines()[-1])
    print(d["config"], d["agent_steps_per_s_rollout"], d["iters"][-1])
except Exception as e:
    print("ppo ERR", e)
PY
