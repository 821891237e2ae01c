#!/usr/bin/env python
"""Workload for ONE `ncu --set full` capture of every kernel of the hot path (B200_PROFILING.md recipe):

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/ncu_r2 python tools/ncu_target.py

After an untimed warm-up (a whole 120-step episode, so record lists and bucket chains are in their steady state) the
profiled region launches, once each and in this order:
  c3 (65,536 x 3): k_step, k_raster_sparse_lsu -> COMPRESSIBLE tensor; k_step, k_raster_sparse_lsu -> PLAIN tensor;
                   k_moments_partial, k_moments_final (1 GiB plain readings), k_moments_standardize, k_episode_stats,
                   k_stats_exchange, k_reset_state, k_zero_maps, k_raster_sparse_lsu (after reset)
  c2 (4,096 x 1):  k_step, k_raster_small
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rad_team_b200.distributed import RolloutStatsExchange  # noqa: E402
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402
from rad_team_b200.math_utils import RunningMoments  # noqa: E402

dev = torch.device("cuda", 0)
E, A, T = 65536, 3, 120
env = BatchedRadSearchEnv(n_envs=E, n_agents=A, grid=32, max_steps=T, poly_min=1, poly_max=5)
acts = torch.randint(1, 5, (T, E, A), dtype=torch.int32, device=dev)
roll = torch.zeros((T, E, A), dtype=torch.float32, device=dev)
env.set_rollout(roll)
plain = torch.empty((E, 3, 32, 32), dtype=torch.float32, device=dev)
env.reset()
for t in range(T):
    env.step(acts[t])
env.reset()
for t in range(60):   # mid-episode: ~40 visited cells per env, chains of a few buckets
    env.step(acts[t])
n = 256 * 1024 * 1024
x = torch.empty(n, dtype=torch.float32, device=dev)
rv = roll.view(-1)
for o in range(0, n, rv.numel()):
    m = min(rv.numel(), n - o)
    x[o:o + m].copy_(rv[:m])
y = torch.empty_like(x)
loc, glob = RunningMoments(), RunningMoments()
ex = RolloutStatsExchange(backend="peer")
ep = torch.zeros(3, dtype=torch.float64, device=dev)
loc.update(x); glob.standardize(x, out=y); env.episode_stats(ep); ex.exchange(loc, glob, ep)
small = BatchedRadSearchEnv(n_envs=4096, n_agents=1, grid=32, max_steps=T)
sa = torch.randint(1, 5, (T, 4096, 1), dtype=torch.int32, device=dev)
small.reset()
for t in range(60):
    small.step(sa[t])
torch.cuda.synchronize()

torch.cuda.profiler.start()
env.step(acts[60])                                   # k_step, k_raster_sparse_lsu (compressible)
env.step(acts[61], rasterize=False)                  # k_step
env.rasterize(out_maps=plain)                        # k_raster_sparse_lsu (plain)
loc.update(x)                                        # k_moments_partial, k_moments_final
glob.standardize(x, out=y)                           # k_moments_standardize
env.episode_stats(ep)                                # k_episode_stats
ex.exchange(loc, glob, ep)                           # k_stats_exchange
env.reset()                                          # k_reset_state, k_zero_maps, rasteriser
small.step(sa[60])                                   # k_step (small batch), k_raster_small
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu target done")
