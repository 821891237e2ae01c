#!/usr/bin/env python
"""Can the latency-bound k_step and the bandwidth-bound rasteriser overlap?  Two independent env handles,
one stream each: time (a) N rasterise launches alone, (b) N k_step launches alone, (c) both at once."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402

E, A, N = 65536, 3, 40
kw = dict(n_agents=A, grid=32, max_steps=120, poly_min=1, poly_max=5)
ra = BatchedRadSearchEnv(n_envs=E, seed=1, **kw)
st = BatchedRadSearchEnv(n_envs=E, seed=2, alloc_maps=False, **kw)
acts = torch.randint(1, 5, (N, E, A), dtype=torch.int32, device="cuda")
ra.reset(); st.reset()
for t in range(20):
    ra.step(acts[t])
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_r, do_s):
    st.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(N):
        if do_r:
            with torch.cuda.stream(s1):
                ra.rasterize()
        if do_s:
            with torch.cuda.stream(s2):
                st.step(acts[t], rasterize=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e3


for _ in range(2):
    r, s, b = run(True, False), run(False, True), run(True, True)
print(f"per launch: rasterise alone {r:.4f} ms, k_step alone {s:.4f} ms, both streams at once {b:.4f} ms "
      f"(sum {r + s:.4f}, max {max(r, s):.4f})")
