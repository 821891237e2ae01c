#!/usr/bin/env python
"""Soak: many whole episodes at full c3 size with resets, a sampled block of envs checked bit-exactly
against the CPU oracle every step, and the sticky error flags checked at the end.

    python tools/soak.py [--episodes 10] [--envs 65536]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import oracle as O  # noqa: E402  (checker only)
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=10)
ap.add_argument("--envs", type=int, default=65536)
args = ap.parse_args()
E, A, T = args.envs, 3, 120
kw = dict(O.DEFAULTS)
kw.update(n_envs=E, n_agents=A, grid=32, max_steps=T, poly_min=1, poly_max=5, seed=77)
env = BatchedRadSearchEnv(**kw)
bases = [7, E // 2 + 13, E - 40]
orcs = []
for b in bases:
    k2 = dict(kw); k2.update(n_envs=32, env_id_base=b)
    orcs.append(O.OracleEnv(**k2))
gen = torch.Generator(device=env.device); gen.manual_seed(1)
t0 = time.time()
mism = 0
for ep in range(args.episodes):
    acts = torch.randint(1, 5, (T, E, A), dtype=torch.int32, device=env.device, generator=gen)
    xy, _ = env.reset()
    for b, o in zip(bases, orcs):
        xy_o, m_o = o.reset()
        assert np.array_equal(xy[b:b + 32].cpu().numpy(), xy_o), (ep, "reset")
        assert np.array_equal(env.obs_maps[b:b + 32].cpu().numpy(), m_o)
    for t in range(T):
        xy, rew, done, _, _ = env.step(acts[t])
        for b, o in zip(bases, orcs):
            xy_o, m_o, rew_o, done_o = o.step(acts[t, b:b + 32].cpu().numpy())
            ok = (np.array_equal(xy[b:b + 32].cpu().numpy(), xy_o) and np.array_equal(rew[b:b + 32].cpu().numpy(), rew_o)
                  and np.array_equal(done[b:b + 32].cpu().numpy(), done_o)
                  and np.array_equal(env.buffer("LAST_COUNT")[b:b + 32].cpu().numpy(), o.export("LAST_COUNT"))
                  and np.array_equal(env.obs_maps[b:b + 32].cpu().numpy(), m_o))
            mism += 0 if ok else 1
    for b, o in zip(bases, orcs):
        for k in ("WELFORD", "EST_MAXMIN", "VISIT", "INTENSITY"):
            if not np.array_equal(env.buffer(k)[b:b + 32].cpu().numpy(), o.export(k)):
                mism += 1
print(f"soak: {args.episodes} episodes x {T} steps x {E} envs x {A} agents, {len(bases) * 32} envs checked per step; "
      f"mismatching checks: {mism}; error flags: {env.errors():#x}; {time.time() - t0:.1f} s")
sys.exit(1 if mism or env.errors() else 0)
