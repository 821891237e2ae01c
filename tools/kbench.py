#!/usr/bin/env python
"""Kernel micro-benchmarks (not the driver's bench): each hot kernel alone, CUDA events, inputs larger
than L2, next to plain torch copy / fill / sum of the same footprint as live HBM ceilings.

    python tools/kbench.py [--envs 65536] [--agents 3] [--reps 30] [--moments-mb 1024]
Prints one JSON object.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402
from rad_team_b200.math_utils import RunningMoments  # noqa: E402


def timeit(fn, reps, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--agents", type=int, default=3)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--moments-mb", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda")
    peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] \
        if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6650.0
    out = {"peak_gbs": peak, "envs": args.envs, "agents": args.agents}

    # live ceilings: same footprint as one raster launch (12 KB written, 6 KB read per env)
    E = args.envs
    n_w, n_r = E * 3072, E * 1536  # fp32 elements written / "read-equivalent" fp32 elements
    w = torch.empty(n_w, dtype=torch.float32, device=dev)
    r = torch.ones(n_w, dtype=torch.float32, device=dev)
    ms = timeit(lambda: w.copy_(r), args.reps)
    out["torch_copy_1r1w_gbs"] = 2 * n_w * 4 / ms / 1e6
    ms = timeit(lambda: w.zero_(), args.reps)
    out["torch_fill_0r1w_gbs"] = n_w * 4 / ms / 1e6
    ms = timeit(lambda: r.sum(), args.reps)
    out["torch_sum_1r0w_gbs"] = n_w * 4 / ms / 1e6
    # 1 read : 2 write mix with torch: out[:, :2] = a ; out[:, 2] = 0  approximated by add of half-size
    half = r[: n_w // 2]
    w2 = w.view(2, -1)
    def mix():
        torch.add(half, 1.0, out=w2[0])
        w2[1].zero_()

    ms = timeit(mix, args.reps)
    out["torch_1r2w_two_kernels_gbs"] = (n_w * 4 + n_w // 2 * 4) / ms / 1e6
    del w, r, w2, half

    env = BatchedRadSearchEnv(n_envs=E, n_agents=args.agents, grid=32, max_steps=120, poly_min=1, poly_max=5)
    env.reset()
    acts = torch.randint(1, 5, (120, E, args.agents), dtype=torch.int32, device=dev)
    for t in range(30):
        env.step(acts[t])
    ms = timeit(lambda: env.rasterize(), args.reps)
    dense = os.environ.get("RT_B200_RASTER", "") == "dense"
    rb = 18432 if dense else 12288 + 1024  # bytes per env: dense maps re-read, or record prefix only
    out["k_raster_vec"] = {"kernel": "k_raster_vec" if dense else "k_raster_sparse*", "bytes_per_env": rb, "ms": ms,
                           "gbs": E * rb / ms / 1e6, "frac": E * rb / ms / 1e6 / peak}
    state = {"t": 30}

    def st(raster):
        env.step(acts[state["t"] % 120], rasterize=raster)
        state["t"] += 1
        if state["t"] % 120 == 0:
            env.reset()

    ms = timeit(lambda: st(False), 60)
    out["k_step"] = {"ms": ms, "agent_steps_per_s": E * args.agents / ms * 1e3}
    ms = timeit(lambda: st(True), 60)
    b = E * (rb + args.agents * 254)
    out["step_group"] = {"ms": ms, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak,
                            "agent_steps_per_s": E * args.agents / ms * 1e3}
    env.set_incremental(True)   # opt-in: same tensor every step -> only the changed cells are rewritten
    ms = timeit(lambda: st(True), 60)
    out["step_group_incremental_obs"] = {"ms": ms, "agent_steps_per_s": E * args.agents / ms * 1e3,
                                         "note": "opt-in for inference loops; NOT what bench.py measures"}
    env.set_incremental(False)
    ms = timeit(lambda: env.reset(), 10)
    out["reset_3_kernels_ms"] = ms
    del env

    # K4b on a rollout-sized buffer (HBM regime)
    n = args.moments_mb * 1024 * 1024 // 4
    x = torch.poisson(torch.full((n,), 200.0, device=dev))
    y = torch.empty_like(x)
    m = RunningMoments()
    ms = timeit(lambda: m.update(x), args.reps)
    out["k_moments_update"] = {"ms": ms, "mb": args.moments_mb, "gbs": n * 4 / ms / 1e6, "frac": n * 4 / ms / 1e6 / peak}
    ms = timeit(lambda: m.standardize(x, out=y), args.reps)
    out["k_moments_standardize"] = {"ms": ms, "gbs": n * 8 / ms / 1e6, "frac": n * 8 / ms / 1e6 / peak}
    ms = timeit(lambda: x.sum(), args.reps)
    out["torch_sum_same_buffer_gbs"] = n * 4 / ms / 1e6
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
