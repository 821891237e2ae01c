#!/usr/bin/env python
"""bench.py — agent-env-steps/s (observations rasterised) of the batched radiation-search env.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is ONE env step of every env of the workload: motion/collision/reward, measurement (line of
sight + Poisson), per-env statistics and the rasterisation of the [E,3,32,32] observation tensor.  After the
120th step of an episode the rollout CLOSES — batch Welford moments over the [120,E,A] readings, episode
statistics, the cross-GPU exchange (one peer-memory kernel per rank; NCCL all-gather as the alternative
transport) and the rank-ordered merge — and all envs reset.  Whatever --steps is, at least one close lies
INSIDE each timed region: for K < 120 the episode is phased so that tick 119 -> 0 falls in the middle of
the K timed steps (`rollout_closes_timed`, `close_ms`, `allgather_merge_ms` say what was timed).

Workloads (BASELINE.json configs):
  c3  65,536 envs x 3 agents per GPU, 1-5 rectangular obstructions, G = 32, T = 120 — configs[2], the
      3-agent obstructed scenario the north-star target is quoted on.  Default at N = 1; weak scaling.
  c4  1,048,576 envs x 3 agents in total, split over the N ranks — configs[3].  Default at N > 1 (strong
      scaling); the c3-weak figure is then reported as `secondary.c3_weak`.
  c2  4,096 envs x 1 agent, no obstructions — configs[1], L2-resident; reported as `secondary.c2` at N = 1.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM, CUDA-event time, max over ranks.
`e2e`: the same loop through the host-buffer C-ABI entry points (rt_env_step_host): pinned actions read
and obs_xy / reward / done returned to host memory EVERY step inside the timed region, wall clock, max over
ranks.  `roofline`: the rasteriser (dominant kernel) timed with CUDA events around its own launch inside the
step loop, into PLAIN device memory (DRAM traffic = algorithmic bytes); the default compressible tensor is
`roofline.compressible`.  `cpu_baseline`: the CPU oracle (C port of the reference algorithm) on this box's
host cores, bounded sample; `reference_python`: the unmodified Python reference on the same cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

T_EPISODE = 120
GRID = 32
METRIC = "agent-env-steps/sec (obs rasterised)"

WORKLOADS = {
    "c2": dict(envs=4096, agents=1, poly_min=0, poly_max=0, scaling="weak",
               name="c2: 4,096 envs x 1 agent, 1 source, no obstructions, G=32, T=120 (BASELINE configs[1])"),
    "c3": dict(envs=65536, agents=3, poly_min=1, poly_max=5, scaling="weak",
               name="c3: 65,536 envs x 3 agents per GPU, 1-5 polygon obstructions (LOS attenuation), shared "
                    "location/intensity/visit maps, G=32, T=120 (BASELINE configs[2]; N GPUs = N x 65,536 envs)"),
    "c4": dict(envs=1048576, agents=3, poly_min=1, poly_max=5, scaling="strong",
               name="c4: 1,048,576 envs x 3 agents sharded over the ranks, 1-5 polygon obstructions, G=32, T=120 "
                    "(BASELINE configs[3])"),
}


def env_cfg(w, n_envs, base, seed=0):
    return dict(n_envs=n_envs, n_agents=w["agents"], grid=GRID, max_steps=T_EPISODE, poly_min=w["poly_min"],
                poly_max=w["poly_max"], rect_min=2, rect_max=6, lognorm_step=2, fixed_scenario=0, env_id_base=base,
                seed=seed, cell_pitch_cm=100.0, min_dist_cm=50.0, transmission=0.0, src_lo=1e6, src_hi=1e7,
                bkg_lo=10.0, bkg_hi=51.0)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML every 100 ms while running (one NVML query per
    100 ms: it shares this process's cores with the step loop, so it is kept rare)."""

    def __init__(self, index: int, period: float = 0.1):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.period = index, [], set(), None, period
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        try:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in names.items():
                if bit and (r & bit):
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            self.sample()
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def measured_peak_gbs():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic_per_launch(n_envs, key="plain_memory_dram_bytes_per_env"):
    """dram bytes (read + write) per rasteriser launch from the committed ncu --set full captures
    (profiles/raster_traffic.json), scaled per env: `plain_memory_dram_bytes_per_env` for a plain cudaMalloc
    observation tensor (what roofline.frac is measured on), `dram_bytes_per_env` for the compressible one."""
    p = REPO / "profiles" / "raster_traffic.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d[key]) * n_envs
        except Exception:
            return None
    return None


def ncu_kstep_alu():
    """SM / ALU utilisation of k_step from the committed ncu --set full capture (profiles/kstep_alu.json): the
    kernel that holds the line-of-sight tests is judged on instruction issue, not on HBM."""
    p = REPO / "profiles" / "kstep_alu.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return None


def pin_cores(local_rank: int, world: int):
    """Give every rank of the box its own slice of the cores this job may use: eight step loops that share one
    affinity mask migrate and queue behind each other (SCALE_r01: e2e efficiency 0.86 at N = 8)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // world
        if world > 1 and per >= 1:
            mine = cpus[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, mine)
            return mine
        return cpus
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(w, n_envs, steps, warmup, threads, seconds=None):
    """Time the CPU oracle (all `threads` host threads) on `n_envs` envs of the workload: at least `steps` steps
    AND at least `seconds` of wall clock.  Returns (agent-steps/s, steps actually timed, threads used)."""
    import numpy as np

    from oracle import oracle as O

    A = w["agents"]
    env = O.OracleEnv(**env_cfg(w, n_envs, 0))
    used = env.set_threads(threads)
    rng = np.random.default_rng(0)
    acts = rng.integers(1, 5, size=(T_EPISODE, n_envs, A)).astype(np.int32)
    xy = np.zeros((n_envs, A, 2), np.int32)
    rew = np.zeros((n_envs, A), np.int32)
    done = np.zeros((n_envs, A), np.uint8)
    maps = np.zeros((n_envs, 3, GRID, GRID), np.float32)
    s = 0

    def one():
        nonlocal s
        if s % T_EPISODE == 0:
            env.reset(maps=False)
        env.step_into(acts[s % T_EPISODE], xy, maps, rew, done)
        s += 1

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    n = 0
    while n < steps or (seconds is not None and time.perf_counter() - t0 < seconds):
        one()
        n += 1
    dt = time.perf_counter() - t0
    return n_envs * A * n / dt, n, used


def reference_python_rates(seconds: float):
    """The UNMODIFIED Python reference on this box's cores (oracle/ref_py.py: mounted source tree, or the
    byte-code compiled from it into oracle/_ref/refpy/ that ships with the snapshot)."""
    try:
        from oracle import ref_py

        if not ref_py.available():
            return {"available": False, "why": "oracle/_ref/refpy byte-code absent and /root/reference not mounted "
                                               "(run `make -C oracle ref_py` in the build container)"}
        origin = ref_py.load_reference()[-1]
        one, _ = ref_py.time_simplegrid(seconds)
        allc, procs, _, _ = ref_py.time_simplegrid_all_cores(seconds)
        pipe, _ = ref_py.time_pipeline(seconds)
        composed = 1.0 / (1.0 / one + 1.0 / pipe)
        return {"available": True, "kind": "reference", "origin": origin, "cores": procs,
                "simplegrid_env_steps_per_s_1core": one, "simplegrid_env_steps_per_s_all_cores": allc,
                "math_utils_pipeline_readings_per_s_1core": pipe,
                "composed_agent_steps_per_s_1core": composed,
                "composed_agent_steps_per_s_all_cores_est": composed * allc / one,
                "sample": f"SimpleGrid(start=(5,5), terminal=(0,0), size=10) reset + 120 random-action steps per episode "
                          f"for {seconds:g} s on 1 process and on {procs} processes (c1); estimator -> Welford -> "
                          f"Normalizer + BensLog per reading for {seconds:g} s on 1 process; composed = one env step + one "
                          f"reading's statistics (the reference has no physics, line of sight, Poisson or rasteriser to time)"}
    except Exception as exc:  # the GPU arm must not die on the CPU leg
        return {"available": False, "why": f"{type(exc).__name__}: {exc}"}


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation on this box's host cores.  `value` is the
    oracle's C port of the reference algorithm on THIS arm's config (OpenMP over envs, every host thread, timed
    for at least 1 s) — the conservative baseline: the unmodified Python reference, timed beside it
    (`reference_python`), is ~30x slower per core and has no physics or rasteriser."""
    if rank != 0:
        return
    wname = args.workload or ("c3" if args.gpus <= 1 else "c4")
    w = WORKLOADS[wname]
    threads = os.cpu_count() or 1
    n_envs = min(args.ref_envs, w["envs"])
    rate, n, used = cpu_oracle_rate(w, n_envs, args.steps, max(args.warmup, 3), threads, seconds=args.ref_seconds)
    cpu = {"value": rate, "unit": "agent-env-steps/s", "cores": used, "kind": "port",
           "sample": f"{n_envs} envs x {w['agents']} agents x {n} steps (>= {args.ref_seconds:.0f} s) of the workload's "
                     f"scenario, C oracle, {used} OpenMP threads"}
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "agent-env-steps/s", "n_gpus": args.gpus,
        "steps": n, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * n_envs * w["agents"] / rate,
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "sample_envs": n_envs, "agents": w["agents"], "grid": GRID,
                   "episode_steps": T_EPISODE},
        "cpu_baseline": cpu,
        "reference_python": None if args.no_ref_python else reference_python_rates(args.ref_py_seconds),
        "e2e": {"value": rate, "unit": "agent-env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
class Ctx:
    """What every measured workload of one run shares."""

    def __init__(self, torch, dist, rank, world, dev, exchange):
        self.torch, self.dist, self.rank, self.world, self.dev, self.exchange = torch, dist, rank, world, dev, exchange

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def phase_plan(K: int, W: int):
    """(warm-up steps, tick at which the timed region starts).  For K < T the timed region starts at tick
    T - ceil(K / 2), so the episode's last step, the rollout close and the reset fall in its middle; the warm-up
    runs whole extra episodes when W asks for more steps than that."""
    T = T_EPISODE
    t0 = 0 if K >= T else T - (K + 1) // 2
    warm = t0
    while warm < W:
        warm += T
    return warm, t0


def measure(cx: Ctx, w, E, base, K, W, full: bool, e2e_steps: int = 0, graph: bool = False):
    """Device-resident loop + host-buffer loop of one workload on this rank's shard.  Returns a dict.
    graph=True (K a multiple of 120): the device-resident loop is ONE captured CUDA graph of a whole episode — 120
    steps, the rollout close and the reset — replayed K / 120 times: for small batches the Python loop cannot issue
    two launches per ~15 us step, and the device-resident number would measure the host."""
    torch, dev, world = cx.torch, cx.dev, cx.world
    from rad_team_b200.envs import BatchedRadSearchEnv
    from rad_team_b200.math_utils import RunningMoments
    from rad_team_b200.memory import readings_tensor

    A, T = w["agents"], T_EPISODE
    env = BatchedRadSearchEnv(**env_cfg(w, E, base))
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + cx.rank)
    actions = torch.randint(1, 5, (T, E, A), dtype=torch.int32, device=dev, generator=gen)
    rollout = readings_tensor((T, E, A), device=dev)
    env.set_rollout(rollout)
    local_m, global_m = RunningMoments(), RunningMoments()   # this rollout's readings / everything so far, all ranks
    ep_stats = torch.zeros(3, dtype=torch.float64, device=dev)
    exch_launches = 1 if cx.exchange.backend == "peer" else 2  # our kernels only (NCCL's own are not counted)
    st = {"tick": 0, "closes": 0, "extra": 0}
    close_ev = []   # (start, before exchange, after exchange, after reset) CUDA events of the closes of the timed regions
    probes = []     # (before k_step, between, after rasteriser) CUDA events of the sampled steps

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def close_rollout(host: bool, timed: bool):
        e = [ev() for _ in range(4)] if timed else None
        if e:
            e[0].record()
        local_m.update(rollout.view(-1))          # k_moments_partial + k_moments_final -> this rollout's moments
        env.episode_stats(ep_stats)               # k_episode_stats
        if e:
            e[1].record()
        cx.exchange.exchange(local_m, global_m, ep_stats)   # all ranks: gather rows, rank-ordered merge into global_m
        if e:
            e[2].record()
        env.reset(host=host)                      # k_reset_state + k_zero_maps + rasteriser
        if e:
            e[3].record()
            close_ev.append(e)
        st["closes"] += 1
        st["extra"] += 3 + exch_launches
        st["tick"] = 0

    def dev_step(probe=False, out_maps=None, timed=False):
        t = st["tick"]
        if probe:  # same two kernels, launched through the split calls with events around each
            e = [ev() for _ in range(3)]
            e[0].record()
            env.step(actions[t], rasterize=False)
            e[1].record()
            env.rasterize(out_maps=out_maps)
            e[2].record()
            probes.append(e)
        else:
            env.step(actions[t], out_maps=out_maps)
        st["tick"] = t + 1
        if st["tick"] == T:
            close_rollout(False, timed)

    graph = graph and K % T == 0
    warm, t0 = phase_plan(K, W)
    # ---- device-resident throughput ("value") -------------------------------------------------
    env.reset()
    close_rollout(False, False)  # warm the close path (the first NCCL call builds its channels lazily); resets again
    for _ in range(warm):
        dev_step()
    assert st["tick"] == t0
    if graph:
        while st["tick"] != 0:
            dev_step()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            l0, x0 = env.launches, st["extra"]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):   # 120 x (k_step + rasteriser), close (moments, stats, exchange), reset
                for _ in range(T):
                    dev_step()
            per_episode = (env.launches - l0) + (st["extra"] - x0)
            g.replay()                               # the capture executed nothing: run the episode once
            side.synchronize()
            cx.barrier()
            c0 = st["closes"]
            ev0, ev1 = ev(), ev()
            ev0.record()
            for _ in range(K // T):
                g.replay()
            ev1.record()
            side.synchronize()
        torch.cuda.current_stream().wait_stream(side)
        st["closes"] += K // T                       # the capture counted one close (= the warm replay); add the timed ones
        cx.barrier()
        ms_total = cx.max_over_ranks(ev0.elapsed_time(ev1))
        launches = per_episode * (K // T)
        for i in range(16):                          # kernel samples, eagerly, right after the timed region
            dev_step(probe=True)
        torch.cuda.synchronize()
    else:
        cx.barrier()
        l0, x0, c0 = env.launches, st["extra"], st["closes"]
        ev0, ev1 = ev(), ev()
        ev0.record()
        for i in range(K):
            dev_step(probe=(i % 8 == 3), timed=True)  # every 8th step also times its two kernels, live, on this stream
        ev1.record()
        cx.barrier()
        ms_total = cx.max_over_ranks(ev0.elapsed_time(ev1))
        launches = (env.launches - l0) + (st["extra"] - x0)
    closes_dev = st["closes"] - c0
    n_probe_c = len(probes)
    kstep_ms = sum(e[0].elapsed_time(e[1]) for e in probes) / max(1, n_probe_c)
    raster_c_ms = sum(e[1].elapsed_time(e[2]) for e in probes) / max(1, n_probe_c)
    close_ms = [e[0].elapsed_time(e[3]) for e in close_ev]
    exch_ms = [e[1].elapsed_time(e[2]) for e in close_ev]
    del close_ev[:]

    out = {"E": E, "ms_total": ms_total, "ms_per_step": ms_total / K, "launches": int(launches),
           "closes_dev": closes_dev, "close_ms": close_ms, "exchange_ms": exch_ms, "kstep_ms": kstep_ms,
           "raster_compressible_ms": raster_c_ms, "probe_samples": n_probe_c,
           "mode": "one CUDA graph per episode (120 steps + rollout close + reset), replayed" if graph
                   else "eager launches from the Python loop"}

    if full:
        # ---- a longer window of the same loop: whole episodes, so the close is amortised as in production -------
        n_long = int(math.ceil(max(K, 0.25e3 / max(out["ms_per_step"], 1e-6)) / T)) * T
        while st["tick"] != 0:
            dev_step()
        cx.barrier()
        la, lb = ev(), ev()
        c1 = st["closes"]
        if graph:
            with torch.cuda.stream(side):
                la.record()
                for _ in range(n_long // T):
                    g.replay()
                lb.record()
                side.synchronize()
            st["closes"] += n_long // T
        else:
            la.record()
            for _ in range(n_long):
                dev_step()
            lb.record()
        cx.barrier()
        out["long"] = {"steps": n_long, "ms_per_step": cx.max_over_ranks(la.elapsed_time(lb)) / n_long,
                       "closes": st["closes"] - c1}
        # ---- the rasteriser IN THE STEP LOOP into plain device memory (DRAM traffic = algorithmic bytes) -------
        del probes[:]
        plain = torch.empty((E, 3, GRID, GRID), dtype=torch.float32, device=dev)
        for i in range(48):
            dev_step(probe=(i >= 8), out_maps=plain)
        torch.cuda.synchronize()
        out["raster_plain_ms"] = sum(e[1].elapsed_time(e[2]) for e in probes) / len(probes)
        out["kstep_plain_ms"] = sum(e[0].elapsed_time(e[1]) for e in probes) / len(probes)
        out["plain_samples"] = len(probes)
        del plain
        del probes[:]
        # live store ceiling of the compressible observation tensor: a pure fill of the same memory
        fe0, fe1 = ev(), ev()
        for _ in range(3):
            env.obs_maps.zero_()
        fe0.record()
        for _ in range(10):
            env.obs_maps.zero_()
        fe1.record()
        torch.cuda.synchronize()
        out["fill_ms"] = fe0.elapsed_time(fe1) / 10
        out["obs_compressible"] = bool(getattr(env.obs_maps, "rt_compressible", False))
        env.rasterize()  # the fill destroyed the observation: bring it back before the host loop

    # ---- end to end through the host-buffer C ABI ("e2e") ----------------------------------------
    h_actions = torch.randint(1, 5, (T, E, A), dtype=torch.int32).pin_memory()

    def host_step(timed=False):
        t = st["tick"]
        env.step(h_actions[t])  # pinned actions read in place, kernels, results written to pinned host memory, wait
        st["tick"] = t + 1
        if st["tick"] == T:
            close_rollout(True, timed)

    K2 = min(K, e2e_steps) if e2e_steps else K
    warm2, t02 = phase_plan(K2, W)
    env.reset(host=True)
    st["tick"] = 0
    for _ in range(warm2):
        host_step()
    assert st["tick"] == t02
    cx.barrier()
    c2 = st["closes"]
    t_0 = time.perf_counter()
    for _ in range(K2):
        host_step(timed=True)
    torch.cuda.synchronize()
    e2e_s = cx.max_over_ranks(time.perf_counter() - t_0)
    cx.barrier()
    out.update(e2e_s=e2e_s, e2e_steps=K2, closes_e2e=st["closes"] - c2,
               close_ms_e2e=[e[0].elapsed_time(e[3]) for e in close_ev],
               exchange_ms_e2e=[e[1].elapsed_time(e[2]) for e in close_ev])
    if full:  # the same host loop over whole episodes (>= 0.25 s): the close amortised as in production
        n_long = out["long"]["steps"]
        while st["tick"] != 0:
            host_step()
        cx.barrier()
        t_0 = time.perf_counter()
        for _ in range(n_long):
            host_step()
        torch.cuda.synchronize()
        out["long"]["e2e_s"] = cx.max_over_ranks(time.perf_counter() - t_0)
        cx.barrier()

    # ---- merge check: the global moments are bit-identical on every rank and hold every reading once ----------
    torch.cuda.synchronize()
    bits = global_m.state.view(torch.int64).clone()
    same = True
    if world > 1:
        allb = torch.empty((world, 3), dtype=torch.int64, device=dev)
        cx.dist.all_gather_into_tensor(allb.view(-1), bits)
        same = bool((allb == allb[0:1]).all().item())
    total_envs = int(round(cx.sum_over_ranks(float(E))))
    expect = float(st["closes"]) * T * total_envs * A
    count = float(global_m.state[0].item())
    out["merge_check"] = bool(same and count == expect and cx.exchange.errors() == 0)
    out["merge_detail"] = {"bit_identical_on_all_ranks": same, "count": count, "expected_count": expect,
                           "closes_total": st["closes"], "mean": float(global_m.state[1].item()),
                           "episode_stats_last_close": [float(x) for x in cx.exchange.ep_merged.tolist()]}
    out["flags"] = env.errors()
    out["total_agents"] = total_envs * A
    out["rollout"] = rollout
    out["env"] = env
    return out


def time_exchange(cx: Ctx, reps: int = 20):
    """The exchange alone, both transports, CUDA events on the launching stream: us per call, max over ranks."""
    torch = cx.torch
    from rad_team_b200.math_utils import RunningMoments

    a, b = RunningMoments(), RunningMoments()
    ep = torch.zeros(3, dtype=torch.float64, device=cx.dev)
    res = {}
    for be in (["peer"] if cx.exchange.backend == "peer" else []) + ["nccl"]:
        for _ in range(3):
            cx.exchange.exchange(a, b, ep, backend=be)
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            cx.exchange.exchange(a, b, ep, backend=be)
        e1.record()
        cx.barrier()
        res[be + "_us"] = cx.max_over_ranks(e0.elapsed_time(e1)) / reps * 1e3
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from rad_team_b200.distributed import RolloutStatsExchange, shard_range
    from rad_team_b200.math_utils import RunningMoments

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference)")
    cores = pin_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    exchange = RolloutStatsExchange(backend=args.exchange)
    cx = Ctx(torch, dist, rank, world, dev, exchange)

    wname = args.workload or ("c3" if world == 1 else "c4")
    w = WORKLOADS[wname]
    A = w["agents"]
    if args.envs:
        E, base = args.envs, rank * args.envs
    elif w["scaling"] == "strong":
        base, E = shard_range(w["envs"], rank, world)
    else:
        E, base = w["envs"], rank * w["envs"]
    K, W = args.steps, max(args.warmup, 3)

    sampler = ClockSampler(local_rank, period=args.clock_period if args.clock_period > 0 else 3600.0)
    sampler.start()  # (a period of 0 leaves one sample at the start: an A/B knob for the sampler's own cost)
    use_graph = args.graph == "on" or (args.graph == "auto" and wname == "c2")
    m = measure(cx, w, E, base, K, W, full=True, e2e_steps=args.e2e_steps, graph=use_graph)
    clocks = sampler.stop()
    env, rollout = m.pop("env"), m.pop("rollout")

    total_agents = m["total_agents"]
    value = total_agents * K / (m["ms_total"] * 1e-3)
    e2e_value = total_agents * m["e2e_steps"] / m["e2e_s"]
    xt = time_exchange(cx)

    peak, peak_src = measured_peak_gbs()
    # Algorithmic bytes per launch.  Default rasteriser k_raster_sparse_lsu: 12,288 B stored +
    # 1,024 B loaded (record prefix) per env-step.  SURVEY §8d's K3 figure, 18,432 B/env-step, assumes the dense
    # maps are re-read every step (that is k_raster_vec, RT_B200_RASTER=dense); it is reported as
    # `survey_equiv_*` so the two designs can be compared on one scale.
    dense = os.environ.get("RT_B200_RASTER", "") == "dense"
    survey_bytes = E * (3 * GRID * GRID * 4 + GRID * GRID * (2 + 4))
    raster_bytes = survey_bytes if dense else E * (3 * GRID * GRID * 4 + 128 * 8)
    rname = "k_raster_vec (K3, dense maps)" if dense else "k_raster_sparse_lsu (K3, visited-cell records -> smem tile -> coalesced 16 B stores)"
    plain_ms, comp_ms = m["raster_plain_ms"], m["raster_compressible_ms"]
    achieved = raster_bytes / (plain_ms * 1e-3) / 1e9
    step_bytes = E * A * (150 + 104)  # K1+K2 150, K4a 104 B / agent-step
    fused_ms = m["long"]["ms_per_step"]
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu_traffic_per_launch(E), "kernel": rname, "ms_per_launch": plain_ms,
        "algorithmic_bytes_per_launch": raster_bytes, "peak_source": peak_src,
        "memory": "PLAIN device memory (torch allocator = cudaMalloc): the DRAM moves exactly the algorithmic bytes, so "
                  "this is the fraction to read against the copy-measured HBM peak",
        "timing": "CUDA events around the rasteriser's own launch on the launching stream, inside the step loop "
                  "(k_step -> rasteriser back to back, %d consecutive steps right after the timed region)" % m["plain_samples"],
        "regime": "HBM (per-launch footprint %.0f MB >> 126 MB L2)" % (raster_bytes / 1e6)
        if raster_bytes > 4 * 126e6 else "L2-resident footprint (%.0f MB)" % (raster_bytes / 1e6),
        "traffic_note": "ncu dram__bytes_read + write per launch of the plain-memory capture, scaled per env "
                        "(profiles/raster_traffic.json); the compressible capture is compressible.traffic",
        "compressible": {
            "what": "the product default: observation tensor in compressible HBM (rt_obs_alloc, cuMemCreate COMP_GENERIC); "
                    "the L2 compresses the mostly-zero maps, DRAM traffic < algorithmic bytes, so this is NOT an HBM "
                    "fraction — the binding limit is the store path into the L2 (see store_ceiling)",
            "granted": m["obs_compressible"], "ms_per_launch": comp_ms,
            "traffic": ncu_traffic_per_launch(E, "dram_bytes_per_env"),
            "achieved": raster_bytes / (comp_ms * 1e-3) / 1e9,
            "frac_compressible": raster_bytes / (comp_ms * 1e-3) / 1e9 / peak,
            "timing": "every 8th step of the timed region (%d samples)" % m["probe_samples"],
            "store_ceiling": {"what": "torch fill (zero_) of the same tensor, timed live",
                              "fill_gbs": E * 3 * GRID * GRID * 4 / (m["fill_ms"] * 1e-3) / 1e9,
                              "raster_store_gbs": E * 3 * GRID * GRID * 4 / (comp_ms * 1e-3) / 1e9,
                              "frac": m["fill_ms"] / comp_ms}},
        "survey_equiv_achieved": survey_bytes / (plain_ms * 1e-3) / 1e9,
        "survey_equiv_frac": survey_bytes / (plain_ms * 1e-3) / 1e9 / peak,
        "whole_step": {"ms": fused_ms, "algorithmic_bytes": raster_bytes + step_bytes,
                       "achieved": (raster_bytes + step_bytes) / (fused_ms * 1e-3) / 1e9,
                       "frac": (raster_bytes + step_bytes) / (fused_ms * 1e-3) / 1e9 / peak,
                       "what": "long window (%d steps, close amortised over whole episodes), compressible tensor"
                               % m["long"]["steps"]}}

    # ---- K4b (normalisation kernels) on a fully written rollout-sized buffer in PLAIN memory, rank 0 only -------
    k4 = None
    if rank == 0 and not args.no_k4b:
        nbuf = 256 * 1024 * 1024  # fp32 readings = 1 GiB >> 126 MB L2
        xb = torch.empty(nbuf, dtype=torch.float32, device=dev)  # plain memory: DRAM traffic = algorithmic bytes
        rv = rollout.view(-1)  # every row written: the loops above ran whole episodes
        for o in range(0, nbuf, rv.numel()):
            n_ = min(rv.numel(), nbuf - o)
            xb[o:o + n_].copy_(rv[:n_])
        nonzero_frac = float((xb[:min(nbuf, 1 << 24)] != 0).float().mean().item())
        yb = torch.empty_like(xb)
        mm = RunningMoments()

        def t_ev(fn, reps=10):
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        mu, ms_ = t_ev(lambda: mm.update(xb)), t_ev(lambda: mm.standardize(xb, out=yb))
        k4 = {"buffer": "1 GiB fp32 Poisson readings of the env's own rollouts (every row written; %.1f %% non-zero), "
                        "PLAIN device memory in and out (HBM regime)" % (100 * nonzero_frac),
              "k_moments_update": {"ms": mu, "achieved": nbuf * 4 / mu / 1e6, "frac": nbuf * 4 / mu / 1e6 / peak,
                                   "algorithmic_bytes_per_reading": 4},
              "k_moments_standardize": {"ms": ms_, "achieved": nbuf * 8 / ms_ / 1e6, "frac": nbuf * 8 / ms_ / 1e6 / peak,
                                        "algorithmic_bytes_per_reading": 8}}
        del xb, yb, mm
    del env, rollout
    torch.cuda.empty_cache()
    cx.barrier()

    # ---- secondary workloads ------------------------------------------------------------------------------------
    secondary = {}
    if not args.no_secondary:
        if world > 1 and wname == "c4":
            w3 = WORKLOADS["c3"]
            # (collective: every rank runs it; a failure on one rank would hang the others, so no try/except here)
            s = measure(cx, w3, w3["envs"], rank * w3["envs"], K, W, full=False, e2e_steps=args.e2e_steps)
            s.pop("env"), s.pop("rollout")
            secondary["c3_weak"] = {
                "workload": w3["name"], "scaling": "weak", "value": s["total_agents"] * K / (s["ms_total"] * 1e-3),
                "ms_per_step": s["ms_per_step"], "e2e_value": s["total_agents"] * s["e2e_steps"] / s["e2e_s"],
                "rollout_closes_timed": s["closes_dev"], "merge_check": s["merge_check"], "k_step_ms": s["kstep_ms"],
                "raster_ms": s["raster_compressible_ms"]}
        if world == 1 and wname == "c3":
            # the strong-scaling base of configs[3]: all 1,048,576 envs on ONE GPU (43 GB of state + 12.9 GB of maps)
            w4 = WORKLOADS["c4"]
            try:
                s = measure(cx, w4, w4["envs"], 0, K, W, full=False, e2e_steps=min(K, 120))
                s.pop("env"), s.pop("rollout")
                secondary["c4_one_gpu"] = {
                    "workload": w4["name"], "value": s["total_agents"] * K / (s["ms_total"] * 1e-3),
                    "ms_per_step": s["ms_per_step"], "e2e_value": s["total_agents"] * s["e2e_steps"] / s["e2e_s"],
                    "rollout_closes_timed": s["closes_dev"], "merge_check": s["merge_check"],
                    "k_step_ms": s["kstep_ms"], "raster_ms": s["raster_compressible_ms"],
                    "what": "divide the N-GPU c4 `value` by N x this for the strong-scaling efficiency of configs[3]"}
            except Exception as exc:  # e.g. a smaller device: the headline must not die on a secondary figure
                secondary["c4_one_gpu"] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
            w2 = WORKLOADS["c2"]
            K2s = -(-max(K, 1200) // T_EPISODE) * T_EPISODE
            try:
                s = measure(cx, w2, w2["envs"], 0, K2s, W, full=False, e2e_steps=0, graph=True)
                s.pop("env"), s.pop("rollout")
                secondary["c2"] = {
                    "workload": w2["name"], "steps": K2s, "value": s["total_agents"] * K2s / (s["ms_total"] * 1e-3),
                    "ms_per_step": s["ms_per_step"], "e2e_value": s["total_agents"] * s["e2e_steps"] / s["e2e_s"],
                    "rollout_closes_timed": s["closes_dev"], "merge_check": s["merge_check"],
                    "k_step_ms": s["kstep_ms"], "raster_ms": s["raster_compressible_ms"], "mode": s["mode"],
                    "regime": "L2-resident (50 MB observation tensor): launch / latency bound, not HBM"}
            except Exception as exc:  # the headline must not die on a secondary figure
                secondary["c2"] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()

    cpu = refpy = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s = min(args.ref_envs, E)
        rate, n, used = cpu_oracle_rate(w, n_s, 0, 3, threads, seconds=args.cpu_seconds)
        cpu = {"value": rate, "unit": "agent-env-steps/s", "cores": used, "kind": "port",
               "sample": f"{n_s} envs x {A} agents x {n} steps (~{args.cpu_seconds:.0f} s) of the same scenario, "
                         f"C oracle of the reference algorithm, {used} OpenMP threads"}
        if not args.no_ref_python:
            refpy = reference_python_rates(args.ref_py_seconds)

    if rank == 0:
        mean_or_none = lambda v: (sum(v) / len(v)) if v else None  # noqa: E731
        line = {
            "metric": METRIC, "value": value, "unit": "agent-env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "envs_per_gpu": E, "agents": A, "grid": GRID, "episode_steps": T_EPISODE,
                       "obs_memory": "compressible HBM" if m["obs_compressible"] else "plain device memory",
                       "l2": "inputs larger than L2: each step streams %.0f MB per GPU (obs tensor %.0f MB alone)"
                             % (raster_bytes / 1e6, E * 3 * GRID * GRID * 4 / 1e6),
                       "warmup_steps_run": phase_plan(K, W)[0],
                       "episode_phase": "timed region starts at tick %d of %d: %d rollout close(s) inside it"
                                        % (phase_plan(K, W)[1], T_EPISODE, m["closes_dev"]),
                       "rollout_close": "after tick 119: batch Welford moments + episode stats + exchange ("
                                        + exchange.backend + (" transport, %d ranks" % world) + ") + reset",
                       "arithmetic": "fp64 physics/statistics, fp32 geometry and maps, int32 indices/counts",
                       "device_loop": m["mode"],
                       "cores_of_rank0": None if cores is None else len(cores)},
            "rollout_closes_timed": m["closes_dev"], "close_ms": mean_or_none(m["close_ms"]),
            "allgather_merge_ms": mean_or_none(m["exchange_ms"]),
            "exchange": dict(xt, backend=exchange.backend, peer_error=exchange.peer_error,
                             what="one rollout-stats exchange alone, CUDA events, max over ranks; peer = "
                                  "rt_stats_allgather_merge (one kernel over NVLink-mapped mailboxes), nccl = "
                                  "rt_stats_pack + all_gather_into_tensor + rt_stats_merge_gathered"),
            "merge_check": m["merge_check"], "merge_detail": m["merge_detail"],
            "steady_state": {"steps": m["long"]["steps"], "ms_per_step": m["long"]["ms_per_step"],
                             "value": total_agents / (m["long"]["ms_per_step"] * 1e-3),
                             "rollout_closes_timed": m["long"]["closes"],
                             "e2e_value": total_agents * m["long"]["steps"] / m["long"]["e2e_s"],
                             "what": "the same device-resident loop over whole episodes (>= 0.25 s), close amortised"},
            "roofline": roofline,
            "kernels": {"step_ms": m["ms_per_step"], "k_step_ms": m["kstep_ms"], "raster_ms": comp_ms,
                        "raster_plain_ms": plain_ms, "k_step_ncu": ncu_kstep_alu()},
            "normalisation_kernels": k4,
            "secondary": secondary or None,
            "cpu_baseline": cpu,
            "reference_python": refpy,
            "e2e": {"value": e2e_value, "unit": "agent-env-steps/s", "h2d_bytes_per_step": E * A * 4,
                    "d2h_bytes_per_step": E * A * 13 + 12, "steps": m["e2e_steps"],
                    "rollout_closes_timed": m["closes_e2e"], "close_ms": mean_or_none(m["close_ms_e2e"]),
                    "allgather_merge_ms": mean_or_none(m["exchange_ms_e2e"]),
                    "note": "rt_env_step_host per step: pinned int32 actions cross PCIe every step (k_step reads the mapped "
                            "buffer in place); obs_xy, reward, done and the flag word are written into pinned host memory by "
                            "ONE kernel on an auxiliary stream under the rasteriser and the call returns when its sequence "
                            "word lands (polled, no driver call); maps are rasterised into the device-resident CNN input "
                            "tensor in stream order (the timed region ends with a device synchronise)"},
            "gpu_launches": int(m["launches"]),
            "clocks": clocks,
            "error_flags": m["flags"],
        }
        emit(line)
    exchange.close()
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries also write there (NCCL prints its version banner
    on first use when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the whole run and the
    JSON line alone goes to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1200)
    ap.add_argument("--warmup", type=int, default=120)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c3 at N = 1, c4 (strong scaling, BASELINE configs[3]) at N > 1")
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"])
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="device-resident loop as one CUDA graph per episode (auto: c2 only; needs steps % 120 == 0)")
    ap.add_argument("--ref-envs", type=int, default=4096, help="envs in the CPU sample")
    ap.add_argument("--ref-seconds", type=float, default=1.0, help="--impl reference: minimum timed wall clock")
    ap.add_argument("--ref-py-seconds", type=float, default=2.0, help="per leg of the Python-reference timing")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-python", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-k4b", action="store_true")
    ap.add_argument("--clock-period", type=float, default=0.1, help="seconds between NVML clock samples")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (see the module docstring)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
