#!/usr/bin/env python
"""A/B timing of k_step alone and of the step group (k_step + rasteriser) for one build of the library.

    RT_B200_LIB=ab_libs/lib_x.so python tools/kstep_ab.py [--envs 65536] [--agents 3]
Prints one line: label, k_step ms, step-group ms (CUDA events, 3 x 120 steps each, resets inside).
"""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--agents", type=int, default=3)
    ap.add_argument("--label", default=os.environ.get("RT_B200_LIB", "default"))
    args = ap.parse_args()
    dev = torch.device("cuda")
    env = BatchedRadSearchEnv(n_envs=args.envs, n_agents=args.agents, grid=32, max_steps=120, poly_min=1, poly_max=5)
    acts = torch.randint(1, 5, (120, args.envs, args.agents), dtype=torch.int32, device=dev)

    def episode(raster):
        env.reset()
        for t in range(120):
            env.step(acts[t], rasterize=raster)

    res = []
    for raster in (False, True):
        episode(raster)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            episode(raster)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 120)
        res.append(best)
    print(f"{args.label}: k_step(+reset/120) {res[0]:.4f} ms | step group {res[1]:.4f} ms | flags {env.errors()}")


if __name__ == "__main__":
    main()
