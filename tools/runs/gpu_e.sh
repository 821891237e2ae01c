#!/bin/bash
set -x
mkdir -p gpurun_out/e
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e/pytest.log
tail -4 gpurun_out/e/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-k4b > gpurun_out/e/bench_driverlike.json 2> gpurun_out/e/bench_driverlike.err; echo "rc=$?"
timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/e/bench_c2_graph_small.json 2> gpurun_out/e/bench_c2_graph_small.err; echo "rc=$?"
RT_B200_RASTER=sparse_lsu timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/e/bench_c2_graph_lsu.json 2> gpurun_out/e/bench_c2_graph_lsu.err; echo "rc=$?"
RT_B200_PDL=0 timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/e/bench_c2_graph_small_nopdl.json 2> gpurun_out/e/bench_c2_graph_small_nopdl.err; echo "rc=$?"
timeout 600 python bench.py --workload c2 --steps 1200 --graph off --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/e/bench_c2_eager_small.json 2> gpurun_out/e/bench_c2_eager_small.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_driverlike","bench_c2_graph_small","bench_c2_graph_lsu","bench_c2_graph_small_nopdl","bench_c2_eager_small"):
    try:
        d=json.loads(open("gpurun_out/e/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.4f merge %s" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"], d["merge_check"]), {k:v for k,v in d["kernels"].items() if k!="k_step_ncu"}, (d.get("secondary") or {}).get("c2",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/e/bench_c2_graph_small.err
