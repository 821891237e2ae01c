#!/bin/bash
set -x
mkdir -p gpurun_out/l
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/l/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l/pytest.log
tail -6 gpurun_out/l/pytest.log
for m in 1 0 1 0; do RT_B200_HANDOFF=$m timeout 300 python tools/kstep_ab.py --label "handoff=$m" >> gpurun_out/l/handoff_ab.txt 2>&1; done
for m in 1 0; do RT_B200_OBS_COMPRESS=0 RT_B200_HANDOFF=$m timeout 300 python tools/kstep_ab.py --label "plain obs, handoff=$m" >> gpurun_out/l/handoff_ab.txt 2>&1; done
cat gpurun_out/l/handoff_ab.txt
for m in 1 0; do
RT_B200_HANDOFF=$m timeout 600 python bench.py --steps 600 --warmup 5 --no-cpu-baseline --no-k4b > gpurun_out/l/bench_600_handoff$m.json 2> gpurun_out/l/bench_600_handoff$m.err; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/l/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.5f" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,5) for k,v in d["kernels"].items() if k!="k_step_ncu"}, "c2 %.4g c4 %.4g" % (d["secondary"]["c2"]["value"], d["secondary"]["c4_one_gpu"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
