#!/bin/bash
set -x
mkdir -p gpurun_out/f
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f/pytest.log
tail -4 gpurun_out/f/pytest.log
for lib in lib_single_candidate lib_two_candidates; do
  for i in 1 2; do RT_B200_LIB=ab_libs/$lib.so timeout 300 python tools/kstep_ab.py --label $lib >> gpurun_out/f/kstep_ab.txt 2>&1; done
done
cat gpurun_out/f/kstep_ab.txt
for lib in lib_single_candidate lib_two_candidates; do
  RT_B200_LIB=ab_libs/$lib.so timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/f/bench_c2_$lib.json 2> gpurun_out/f/bench_c2_$lib.err; echo "rc=$?"
done
for epb in 8 16; do
  RT_B200_EPB=$epb timeout 600 python bench.py --workload c2 --steps 1200 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/f/bench_c2_epb$epb.json 2> gpurun_out/f/bench_c2_epb$epb.err; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f/bench_c2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.5f merge %s" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"], d["merge_check"]))
    except Exception as e:
        print(f, "ERR", e)
PY
