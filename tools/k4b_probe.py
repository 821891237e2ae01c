#!/usr/bin/env python
"""Why does rt_moments_update run at 0.25 ms per GiB on the env's own readings and 0.18 ms on Poisson(200) noise?
Times the same kernel on several 1 GiB fp32 buffers in plain device memory, plus the D2H copy of one step's results."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402
from rad_team_b200.math_utils import RunningMoments  # noqa: E402

dev = torch.device("cuda", 0)
n = 256 * 1024 * 1024


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


E, A, T = 65536, 3, 120
env = BatchedRadSearchEnv(n_envs=E, n_agents=A, grid=32, max_steps=T, poly_min=1, poly_max=5, alloc_maps=False)
roll = torch.zeros((T, E, A), dtype=torch.float32, device=dev)
env.set_rollout(roll)
env.reset()
acts = torch.randint(1, 5, (T, E, A), dtype=torch.int32, device=dev)
for t in range(T):
    env.step(acts[t], rasterize=False)
torch.cuda.synchronize()
rv = roll.view(-1)
bufs = {}
x = torch.empty(n, dtype=torch.float32, device=dev)
for o in range(0, n, rv.numel()):
    m = min(rv.numel(), n - o)
    x[o:o + m].copy_(rv[:m])
bufs["env_readings_tiled"] = x
bufs["env_readings_shuffled"] = x[torch.randperm(n, device=dev)]
bufs["poisson200"] = torch.poisson(torch.full((n,), 200.0, device=dev))
bufs["poisson39"] = torch.poisson(torch.full((n,), 39.0, device=dev))
bufs["zeros"] = torch.zeros(n, dtype=torch.float32, device=dev)
bufs["clamped_readings"] = x.clamp(max=500.0)
out = {"readings_stats": {"mean": float(x.mean()), "max": float(x.max()), "p999": float(x[:1 << 22].quantile(0.999))}}
y = torch.empty(n, dtype=torch.float32, device=dev)
for k, b in bufs.items():
    mm = RunningMoments()
    out[k] = {"update_ms": timeit(lambda: mm.update(b)), "standardize_ms": timeit(lambda: mm.standardize(b, out=y)),
              "torch_sum_ms": timeit(lambda: b.sum())}
# D2H of one step's results (2.56 MB) and of 16 bytes
h = torch.empty(E * A * 13, dtype=torch.uint8).pin_memory()
d = torch.empty(E * A * 13, dtype=torch.uint8, device=dev)
out["d2h_2p5MB_ms"] = timeit(lambda: h.copy_(d, non_blocking=True), 50)
out["d2h_16B_ms"] = timeit(lambda: h[:16].copy_(d[:16], non_blocking=True), 50)
print(json.dumps(out, indent=1))
