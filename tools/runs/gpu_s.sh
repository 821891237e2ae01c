#!/bin/bash
set -x
mkdir -p gpurun_out/s
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s/pytest.log
tail -4 gpurun_out/s/pytest.log
for rep in 1 2; do
RT_B200_STEP96=0 timeout 300 python tools/kstep_ab.py --label "generic k_step (74 regs, 8 CTAs/SM)" >> gpurun_out/s/step96_ab.txt 2>&1
for n in 9 10 11; do RT_B200_LIB=ab_libs/lib_step96_$n.so timeout 300 python tools/kstep_ab.py --label "k_step96 min_ctas=$n" >> gpurun_out/s/step96_ab.txt 2>&1; done
done
cat gpurun_out/s/step96_ab.txt
for v in 0 1; do RT_B200_STEP96=$v timeout 600 python bench.py --steps 600 --warmup 5 --no-cpu-baseline --no-k4b --no-secondary > gpurun_out/s/bench_600_step96_$v.json 2> gpurun_out/s/bench_600_step96_$v.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g steady %.4g e2e %.4g ms %.5f" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["ms_per_step"]), {k:round(v,5) for k,v in d["kernels"].items() if k!="k_step_ncu"})
    except Exception as e:
        print(f, "ERR", e)
PY
