#!/bin/bash
set -x
mkdir -p gpurun_out/j
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j/pytest.log
tail -4 gpurun_out/j/pytest.log
for m in 1 0 1 0; do RT_B200_RASTER_MASK=$m timeout 300 python tools/kstep_ab.py --label "mask=$m" >> gpurun_out/j/raster_mask_ab.txt 2>&1; done
for m in 1 0; do RT_B200_OBS_COMPRESS=0 RT_B200_RASTER_MASK=$m timeout 300 python tools/kstep_ab.py --label "plain obs, mask=$m" >> gpurun_out/j/raster_mask_ab.txt 2>&1; done
cat gpurun_out/j/raster_mask_ab.txt
for m in 1 0; do
RT_B200_RASTER_MASK=$m timeout 600 python bench.py --steps 600 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/j/bench_600_mask$m.json 2> gpurun_out/j/bench_600_mask$m.err; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/j/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.5f" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,5) for k,v in d["kernels"].items() if k!="k_step_ncu"}, "fill frac", d["roofline"]["compressible"]["store_ceiling"]["frac"], "plain frac", d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
