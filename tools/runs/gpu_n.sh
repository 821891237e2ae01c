#!/bin/bash
# final 8-GPU call: multi-GPU test, driver-like bench, c4 e2e A/B (NVML sampler off / zero-copy actions off)
set -x
mkdir -p gpurun_out/n
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/n/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n/pytest_multi.log
tail -3 gpurun_out/n/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n/bench_n8_steps20.json 2> gpurun_out/n/bench_n8_steps20.err; echo "rc=$?"
timeout 600 $TR --master-port 29532 bench.py --gpus 8 --steps 600 --warmup 5 --no-k4b --no-secondary > gpurun_out/n/bench_n8_600_default.json 2> gpurun_out/n/bench_n8_600_default.err; echo "rc=$?"
timeout 600 $TR --master-port 29533 bench.py --gpus 8 --steps 600 --warmup 5 --no-k4b --no-secondary --clock-period 0 > gpurun_out/n/bench_n8_600_nosampler.json 2> gpurun_out/n/bench_n8_600_nosampler.err; echo "rc=$?"
RT_B200_ZEROCOPY=0 timeout 600 $TR --master-port 29534 bench.py --gpus 8 --steps 600 --warmup 5 --no-k4b --no-secondary > gpurun_out/n/bench_n8_600_h2dcopy.json 2> gpurun_out/n/bench_n8_600_h2dcopy.err; echo "rc=$?"
timeout 600 $TR --master-port 29535 bench.py --gpus 8 --impl reference --steps 20 --warmup 5 > gpurun_out/n/bench_n8_reference.json 2> gpurun_out/n/bench_n8_reference.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl")=="reference":
            print(f, d["value"], d["cpu_baseline"]["cores"], json.dumps(d.get("reference_python"))[:600]); continue
        s=(d.get("secondary") or {}).get("c3_weak",{})
        print(f, "value %.4g steady %.4g e2e %.4g steady-e2e %.4g ms %.4f closes %s merge %s" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["steady_state"].get("e2e_value",0), d["ms_per_step"], d["rollout_closes_timed"], d["merge_check"]), "c3w %.4g e2e %.4g" % (s.get("value",0), s.get("e2e_value",0)))
    except Exception as e:
        print(f, "ERR", e)
PY
