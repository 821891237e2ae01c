#!/usr/bin/env python
"""Config 5: PPO rollout loop — the batched GPU env feeding a small CNN actor-critic
(plain PyTorch, the consumer) — reporting the env's share of the step time.

    python examples/ppo_rollout.py [--envs 8192] [--agents 3] [--horizon 32] [--iters 3] [--dtype bf16]
    torchrun --nproc-per-node N examples/ppo_rollout.py ...      (envs are per GPU; grads all-reduced by DDP)
    config 5 at scale (8 x B200, 131,072 envs per GPU = 1,048,576 envs):
        torchrun --nproc-per-node 8 examples/ppo_rollout.py --envs 131072 --horizon 16 --chunk 32768 --minibatches 64

The policy's forward pass walks the envs in chunks (`--chunk`), so its activations never exceed one chunk; the
observation slots themselves ([horizon + 1, E, 3, 32, 32], compressible HBM) are the env's output tensors.
NVTX ranges (`rt_env_step`, `policy`, `close_rollout`, `ppo_update`) make an nsys timeline readable.

Every env step rasterises the observation straight into its slot of the rollout buffer
(`env.step(actions, out_maps=obs[t])`): no copy between the env and the network's input tensor.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rad_team_b200.distributed import RolloutStatsExchange  # noqa: E402
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402
from rad_team_b200.memory import obs_tensor, readings_tensor  # noqa: E402
from rad_team_b200.math_utils import RunningMoments  # noqa: E402


class ActorCritic(nn.Module):
    """Shared-map CNN trunk; per-agent head on (trunk features, agent x/y, standardised reading)."""

    def __init__(self, grid: int, n_actions: int = 4):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                                  nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2))
        feat = 32 * (grid // 4) ** 2
        self.trunk = nn.Sequential(nn.Flatten(), nn.Linear(feat, 128), nn.ReLU())
        self.head = nn.Sequential(nn.Linear(128 + 3, 64), nn.ReLU())
        self.pi = nn.Linear(64, n_actions)
        self.v = nn.Linear(64, 1)
        self.grid = grid

    def forward(self, maps, xy, reading):
        # maps [B,3,G,G]; xy int32 [B,A,2]; reading [B,A] (standardised counts)
        z = self.trunk(self.conv(maps))                                   # [B,128]
        A = xy.shape[1]
        extra = torch.cat([xy.to(z.dtype) / self.grid, reading.to(z.dtype).unsqueeze(-1)], -1)  # [B,A,3]
        h = self.head(torch.cat([z.unsqueeze(1).expand(-1, A, -1), extra], -1))
        return self.pi(h), self.v(h).squeeze(-1)                          # [B,A,4], [B,A]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--agents", type=int, default=3)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=0, help="envs per policy forward pass (0 = all at once)")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    E, A, T, G = args.envs, args.agents, args.horizon, 32
    env = BatchedRadSearchEnv(n_envs=E, n_agents=A, grid=G, max_steps=120, poly_min=1, poly_max=5,
                              env_id_base=rank * E, seed=7, alloc_maps=False)
    net = ActorCritic(G).to(dev)
    model = nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    amp = torch.bfloat16 if args.dtype == "bf16" else torch.float32

    obs = obs_tensor((T + 1, E, 3, G, G), device=dev, zero=False)   # rollout slots = env outputs, compressible HBM
    xy = torch.empty((T + 1, E, A, 2), dtype=torch.int32, device=dev)
    rd = torch.zeros((T + 1, E, A), dtype=torch.float32, device=dev)
    act = torch.empty((T, E, A), dtype=torch.int64, device=dev)
    logp = torch.empty((T, E, A), dtype=torch.float32, device=dev)
    val = torch.empty((T + 1, E, A), dtype=torch.float32, device=dev)
    rew = torch.empty((T, E, A), dtype=torch.int32, device=dev)    # written by the env, slot by slot
    dn = torch.empty((T, E, A), dtype=torch.uint8, device=dev)
    readings = readings_tensor((120, E, A), device=dev)
    env.set_rollout(readings)
    local_m, moments = RunningMoments(), RunningMoments()   # this rollout's readings / the global running normaliser
    exchange = RolloutStatsExchange()                        # once per rollout: peer-memory kernel (NCCL as the fallback)
    chunk = args.chunk if args.chunk > 0 else E
    nvtx = torch.cuda.nvtx

    def policy(maps, pos, reading):
        """Forward pass in env chunks: activations of one chunk at a time."""
        if chunk >= E:
            return model(maps, pos, reading)
        outs = [model(maps[i:i + chunk], pos[i:i + chunk], reading[i:i + chunk]) for i in range(0, E, chunk)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    ev = lambda: torch.cuda.Event(enable_timing=True)
    env_ms = policy_ms = update_ms = 0.0
    steps_in_episode = 0
    x0, _ = env.reset(out_maps=obs[0])
    xy[0].copy_(x0)
    log = []
    for it in range(args.iters):
        e_pairs, p_pairs = [], []
        for t in range(T):
            a0, a1, a2 = ev(), ev(), ev()
            a0.record()
            nvtx.range_push("policy")
            with torch.no_grad(), torch.autocast("cuda", dtype=amp, enabled=amp != torch.float32):
                logits, v = policy(obs[t], xy[t], rd[t])
            dist = torch.distributions.Categorical(logits=logits.float())
            a = dist.sample()
            act[t], logp[t], val[t] = a, dist.log_prob(a), v.float()
            nvtx.range_pop()
            a1.record()
            nvtx.range_push("rt_env_step")
            # the env writes observation maps, positions, rewards and done flags straight into the
            # rollout buffer's slots; the step's readings land in row `tick` of the readings sink
            env.step((a + 1).to(torch.int32), out_maps=obs[t + 1], out_xy=xy[t + 1], out_reward=rew[t],
                     out_done=dn[t])
            moments.standardize(readings[steps_in_episode], out=rd[t + 1])   # (count - mean) / max(std, 1)
            nvtx.range_pop()
            a2.record()
            p_pairs.append((a0, a1))
            e_pairs.append((a1, a2))
            steps_in_episode += 1
            if steps_in_episode == 120:   # rollout of readings complete: fold into the running normaliser
                nvtx.range_push("close_rollout")
                local_m.update(readings.view(-1))
                exchange.exchange(local_m, moments, env.episode_stats())   # all ranks: global <- chan(global, merged rollout)
                x0, _ = env.reset(out_maps=obs[t + 1])
                xy[t + 1].copy_(x0)
                steps_in_episode = 0
                nvtx.range_pop()
        u0, u1 = ev(), ev()
        u0.record()
        nvtx.range_push("ppo_update")
        with torch.no_grad(), torch.autocast("cuda", dtype=amp, enabled=amp != torch.float32):
            _, v_last = policy(obs[T], xy[T], rd[T])
        val[T] = v_last.float()
        adv = torch.zeros((T, E, A), dtype=torch.float32, device=dev)
        last = torch.zeros((E, A), device=dev)
        for t in reversed(range(T)):   # GAE(0.99, 0.95); `done` is informational in the reference (no auto-reset)
            delta = rew[t].float() + 0.99 * val[t + 1] - val[t]
            last = delta + 0.99 * 0.95 * last
            adv[t] = last
        ret = adv + val[:T]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        idx = torch.randperm(T * E, device=dev).chunk(args.minibatches)
        fo, fx, fr = obs[:T].reshape(T * E, 3, G, G), xy[:T].reshape(T * E, A, 2), rd[:T].reshape(T * E, A)
        fa, fl, fadv, fret = (z.reshape(T * E, A) for z in (act, logp, adv, ret))
        for mb in idx:
            with torch.autocast("cuda", dtype=amp, enabled=amp != torch.float32):
                logits, v = model(fo[mb], fx[mb], fr[mb])
            dist = torch.distributions.Categorical(logits=logits.float())
            ratio = torch.exp(dist.log_prob(fa[mb]) - fl[mb])
            pg = -torch.min(ratio * fadv[mb], torch.clamp(ratio, 0.8, 1.2) * fadv[mb]).mean()
            loss = pg + 0.5 * F.mse_loss(v.float(), fret[mb]) - 0.01 * dist.entropy().mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        nvtx.range_pop()
        u1.record()
        torch.cuda.synchronize()
        e = sum(p.elapsed_time(q) for p, q in e_pairs) / T
        p = sum(p.elapsed_time(q) for p, q in p_pairs) / T
        u = u0.elapsed_time(u1)
        log.append({"iter": it, "env_ms_per_step": e, "policy_ms_per_step": p, "update_ms": u,
                    "env_fraction_of_rollout_step": e / (e + p),
                    "env_fraction_of_iteration": e * T / (e * T + p * T + u), "loss": float(loss.detach())})
        obs[0].copy_(obs[T]); xy[0].copy_(xy[T]); rd[0].copy_(rd[T])
    if rank == 0:
        print(json.dumps({"config": {"envs_per_gpu": E, "agents": A, "horizon": T, "gpus": world, "dtype": args.dtype,
                                     "policy_chunk": chunk, "minibatches": args.minibatches,
                                     "stats_exchange": exchange.backend},
                          "agent_steps_per_s_rollout": world * E * A / ((log[-1]["env_ms_per_step"] + log[-1]["policy_ms_per_step"]) * 1e-3),
                          "iters": log, "env_error_flags": env.errors()}))
    exchange.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
