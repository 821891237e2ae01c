"""Host-side multi-rank logic on CPU: shard arithmetic, and the pack -> all-gather -> rank-ordered merge
under a real world_size-2 gloo group (two processes).  The merge itself runs on the device in the
product (rt_moments_merge); here the gathered rows are merged with the ORACLE's Chan merge to show that
every rank ends with bit-identical global moments equal to the single-pass result."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

from rad_team_b200.distributed import shard_range

REPO = Path(__file__).resolve().parent.parent


def test_shard_range_covers_everything_once():
    for n, w in ((1_048_576, 8), (1_048_576, 2), (10, 4), (7, 8), (0, 3), (65_536 * 3, 3)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert shard_range(1_048_576, 7, 8) == (917_504, 131_072)
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["RT_REPO"])
    from rad_team_b200.distributed import shard_range, pack_stats, gather_stats
    from oracle import oracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 4001
    rng = np.random.default_rng(123)
    readings = rng.poisson(300.0, n_total).astype(np.float32)      # the same global buffer on all ranks
    base, count = shard_range(n_total, rank, world)
    local = O.moments_update(np.zeros(3), readings[base:base + count])   # this rank's shard only
    ep = np.array([-5.0 * (rank + 1), 120.0 * count, float(rank)])
    gathered = gather_stats(pack_stats(torch.from_numpy(local), torch.from_numpy(ep)))
    assert gathered.shape == (world, 6)
    merged = np.zeros(3)
    for r in range(world):
        merged = O.moments_merge(merged, gathered[r, :3].numpy())
    whole = O.moments_update(np.zeros(3), readings)
    out = dict(rank=rank, merged=[float(x).hex() for x in merged], whole=[float(x) for x in whole],
               ep=gathered[:, 3:].sum(0).tolist(), rows=gathered.tolist(), base=base, count=count)
    print("RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()
""")


def test_gloo_world2_gather_and_rank_ordered_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), RT_REPO=str(REPO), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    import json

    res = []
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        res.append(json.loads(next(l for l in out.splitlines() if l.startswith("RESULT "))[7:]))
    res.sort(key=lambda r: r["rank"])
    assert res[0]["merged"] == res[1]["merged"]           # bit-identical on every rank
    assert res[0]["rows"] == res[1]["rows"]               # row r is rank r's contribution everywhere
    merged = [float.fromhex(x) for x in res[0]["merged"]]
    np.testing.assert_allclose(merged, res[0]["whole"], rtol=1e-12)
    assert merged[0] == 4001
    assert (res[0]["base"], res[0]["count"], res[1]["base"], res[1]["count"]) == (0, 2001, 2001, 2000)
    assert res[0]["ep"] == [-15.0, 120.0 * 4001, 1.0]


WORKER2 = textwrap.dedent("""
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["RT_REPO"])
    from rad_team_b200.distributed import shard_range, pack_stats, gather_stats
    from oracle import oracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rollouts, n_total = 3, 3001
    rng = np.random.default_rng(7)
    data = [rng.poisson(100.0 * (r + 1), n_total).astype(np.float32) for r in range(rollouts)]
    base, count = shard_range(n_total, rank, world)
    glob = np.zeros(3)           # persistent global state, identical on every rank
    naive = np.zeros(3)          # the round-1 scheme: ONE state that is both accumulator and merge result
    for r in range(rollouts):
        local = O.moments_update(np.zeros(3), data[r][base:base + count])      # this rollout, this rank only
        g = gather_stats(pack_stats(torch.from_numpy(local), torch.zeros(3, dtype=torch.float64)))
        delta = np.zeros(3)
        for k in range(world):
            delta = O.moments_merge(delta, g[k, :3].numpy().copy())
        glob = O.moments_merge(glob, delta)                                     # global <- chan(global, merged rollout)
        naive = O.moments_update(naive, data[r][base:base + count])            # accumulate into the running state ...
        gn = gather_stats(pack_stats(torch.from_numpy(naive), torch.zeros(3, dtype=torch.float64)))
        naive = np.zeros(3)
        for k in range(world):
            naive = O.moments_merge(naive, gn[k, :3].numpy().copy())           # ... and overwrite it with the merge
    whole = O.moments_update(np.zeros(3), np.concatenate(data))
    print("RESULT " + json.dumps(dict(rank=rank, glob=[float(x).hex() for x in glob], whole=[float(x) for x in whole],
                                      naive_count=float(naive[0]))), flush=True)
    dist.destroy_process_group()
""")


def test_gloo_world2_three_rollouts_count_every_reading_once(tmp_path):
    """The protocol of rt_stats_allgather_merge / RolloutStatsExchange on the host: per-rollout LOCAL moments are
    gathered and merged in rank order, the result is folded into a persistent GLOBAL state.  Equal to one pass over
    all the data; the single-state scheme of round 1 double-counts from the second rollout on (shown alongside)."""
    script = tmp_path / "worker2.py"
    script.write_text(WORKER2)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), RT_REPO=str(REPO), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    import json

    res = []
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        res.append(json.loads(next(l for l in out.splitlines() if l.startswith("RESULT "))[7:]))
    assert res[0]["glob"] == res[1]["glob"]                      # bit-identical on every rank
    glob = [float.fromhex(x) for x in res[0]["glob"]]
    assert glob[0] == 3 * 3001
    np.testing.assert_allclose(glob, res[0]["whole"], rtol=1e-12)
    assert res[0]["naive_count"] > 3 * 3001                      # what the separate global state fixes


def test_bench_phase_plan_puts_a_close_inside_any_window():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", REPO / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    T = bench.T_EPISODE
    for K in (1, 2, 3, 20, 21, 119, 120, 121, 600, 1200):
        for W in (0, 3, 5, 120, 500):
            warm, t0 = bench.phase_plan(K, W)
            assert warm >= W and warm % T == t0 % T and 0 <= t0 < T
            closes = (t0 + K) // T          # a close follows the step that brings the tick to T
            assert closes >= 1, (K, W)
            if K < T:
                assert closes == 1 and 0 <= (T - t0) <= K   # it falls inside the window, not at its edge for K > 1
