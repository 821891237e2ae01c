#!/bin/bash
set -x
mkdir -p gpurun_out/k
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k/pytest.log
tail -4 gpurun_out/k/pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/k/bench_reference.json 2> gpurun_out/k/bench_reference.err; echo "rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/k/bench_steps20.json 2> gpurun_out/k/bench_steps20.err; echo "rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k/smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/k/smoke.log
python - <<'PY'
import json
for f in ("bench_reference","bench_steps20"):
    d=json.loads(open("gpurun_out/k/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], json.dumps(d.get("cpu_baseline")), json.dumps(d.get("reference_python"))[:900])
PY
