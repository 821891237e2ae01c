#!/bin/bash
set -x
mkdir -p gpurun_out/p
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/p/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p/pytest.log
tail -4 gpurun_out/p/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/p/bench_reference.json 2> gpurun_out/p/bench_reference.err; echo "rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p/bench_steps20.json 2> gpurun_out/p/bench_steps20.err; echo "rc=$?"
timeout 900 python bench.py > gpurun_out/p/bench_default.json 2> gpurun_out/p/bench_default.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_reference","bench_steps20","bench_default"):
    d=json.loads(open("gpurun_out/p/%s.json"%f).read().strip().splitlines()[-1])
    if d.get("impl")=="reference":
        print(f, d["value"], json.dumps(d.get("reference_python"))[:400]); continue
    print(f, "value %.4g steady %.4g e2e %.4g steady-e2e %.4g ms %.5f close %.4f merge %s" % (d["value"], d["steady_state"]["value"], d["e2e"]["value"], d["steady_state"]["e2e_value"], d["ms_per_step"], d["close_ms"], d["merge_check"]), {k:round(v,5) for k,v in d["kernels"].items() if k!="k_step_ncu"}, "frac %.3f"%d["roofline"]["frac"], {k:(v.get("value"),v.get("e2e_value")) for k,v in d["secondary"].items()}, d["normalisation_kernels"]["k_moments_update"]["frac"], d["normalisation_kernels"]["k_moments_standardize"]["frac"], d["cpu_baseline"]["value"], (d.get("reference_python") or {}).get("composed_agent_steps_per_s_all_cores_est"))
PY
