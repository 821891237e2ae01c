#!/bin/bash
set -x
mkdir -p gpurun_out/w
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w/pytest.log
tail -3 gpurun_out/w/pytest.log
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/w/bench_20_$i.json 2> gpurun_out/w/bench_20_$i.err; done
RT_B200_RASTER=dense timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/w/bench_20_dense.json 2> gpurun_out/w/bench_20_dense.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/w/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.5f close %.4f e2e-close %.4f" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"], d["close_ms"], d["e2e"]["close_ms"]))
    except Exception as e:
        print(f, "ERR", e)
PY
