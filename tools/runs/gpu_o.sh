#!/bin/bash
set -x
mkdir -p gpurun_out/o
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/o/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o/pytest.log
tail -4 gpurun_out/o/pytest.log
for lib in lib_before_episode_tags lib_episode_tags lib_before_episode_tags lib_episode_tags; do RT_B200_LIB=ab_libs/$lib.so timeout 300 python tools/kstep_ab.py --label $lib >> gpurun_out/o/episode_tags_ab.txt 2>&1; done
cat gpurun_out/o/episode_tags_ab.txt
for lib in lib_before_episode_tags lib_episode_tags; do
RT_B200_LIB=ab_libs/$lib.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/o/bench_20_$lib.json 2> gpurun_out/o/bench_20_$lib.err; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/o/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.5f close %.4f" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"], d["close_ms"]), {k:round(v,5) for k,v in d["kernels"].items() if k!="k_step_ncu"})
    except Exception as e:
        print(f, "ERR", e)
PY
