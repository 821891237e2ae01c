#!/bin/bash
set -x
mkdir -p gpurun_out/b
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b/pytest.log
tail -4 gpurun_out/b/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b/bench_driverlike.json 2> gpurun_out/b/bench_driverlike.err; echo "rc=$?"
timeout 600 python bench.py --steps 600 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/b/bench_600.json 2> gpurun_out/b/bench_600.err; echo "rc=$?"
timeout 600 python tools/k4b_probe.py > gpurun_out/b/k4b_probe.json 2> gpurun_out/b/k4b_probe.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/b/bench_driverlike.json","gpurun_out/b/bench_600.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.4f closes %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["rollout_closes_timed"]), d.get("secondary"))
    except Exception as e:
        print(f, "ERR", e)
PY
cat gpurun_out/b/k4b_probe.json
