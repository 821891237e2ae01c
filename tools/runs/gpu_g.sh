#!/bin/bash
set -x
mkdir -p gpurun_out/g
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g/pytest.log
tail -4 gpurun_out/g/pytest.log
for epb in 1 2 4 8; do
  RT_B200_EPB=$epb timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/g/bench_c2_epb$epb.json 2> gpurun_out/g/bench_c2_epb$epb.err; echo "rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-k4b > gpurun_out/g/bench_driverlike.json 2> gpurun_out/g/bench_driverlike.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.5f merge %s" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"], d["merge_check"]), (d.get("secondary") or {}).get("c2",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
