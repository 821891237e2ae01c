#!/bin/bash
set -x
mkdir -p gpurun_out/d
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d/pytest.log
tail -4 gpurun_out/d/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d/bench_driverlike.json 2> gpurun_out/d/bench_driverlike.err; echo "rc=$?"
timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/d/bench_c2.json 2> gpurun_out/d/bench_c2.err; echo "rc=$?"
RT_B200_RASTER=sparse_lsu timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/d/bench_c2_lsu.json 2> gpurun_out/d/bench_c2_lsu.err; echo "rc=$?"
RT_B200_RASTER=small timeout 600 python bench.py --steps 240 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/d/bench_c3_zero_patch.json 2> gpurun_out/d/bench_c3_zero_patch.err; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/d/launches_c2.csv python bench.py --workload c2 --steps 40 --warmup 3 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/d/ncu_c2.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_driverlike","bench_c2","bench_c2_lsu","bench_c3_zero_patch"):
    try:
        d=json.loads(open("gpurun_out/d/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.4f" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"]), {k:v for k,v in d["kernels"].items() if k!="k_step_ncu"}, (d.get("secondary") or {}).get("c2",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
