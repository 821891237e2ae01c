"""Worker of tests/test_gpu_multi.py: one rank of a sharded run (one process per GPU, or — `RT_SAME_GPU=1` — every
rank on cuda:0 with a gloo group, which exercises the peer-mailbox kernel across processes on a one-GPU box).

Checks, on hardware:
  * per-env outputs of the sharded run (obs_xy, Poisson counts, obstruction counts, rewards, done flags, observation
    maps) are bit-equal to a ONE-GPU run over the same global env ids;
  * the merged moments are bit-identical on all ranks, bit-equal to the oracle's rank-ordered Chan merge of the
    gathered per-rank rows folded into the running global state, and agree with a sequential pass over ALL readings;
  * the peer-memory transport and the library all-gather transport give the same bits;
  * count == E_total * A * T * rollouts (no double counting over rollouts); episode statistics == the plain sums;
  * rt_comm_create_shm (no communication library at all) connects the same ranks.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.environ["RT_REPO"])
from oracle import oracle as O  # noqa: E402  (the checker)
from rad_team_b200 import _lib  # noqa: E402
from rad_team_b200.distributed import RolloutStatsExchange, shard_range  # noqa: E402
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402
from rad_team_b200.math_utils import RunningMoments  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    same_gpu = os.environ.get("RT_SAME_GPU", "0") == "1"
    local = 0 if same_gpu else int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)

    E_total, A, G, T, rollouts = 1530, 3, 32, 12, 2
    cfg = dict(O.DEFAULTS)
    cfg.update(n_agents=A, grid=G, max_steps=T, poly_min=1, poly_max=5, seed=11)
    base, count = shard_range(E_total, rank, world)
    acts_all = np.random.default_rng(5).integers(1, 5, size=(rollouts * T, E_total, A)).astype(np.int32)

    env = BatchedRadSearchEnv(**dict(cfg, n_envs=count, env_id_base=base))
    readings = torch.zeros((T, count, A), dtype=torch.float32, device=dev)
    env.set_rollout(readings)
    ex_peer = RolloutStatsExchange(backend="peer")
    ex_lib = RolloutStatsExchange(backend="nccl")   # the library all-gather (NCCL; gloo in the same-GPU mode)
    loc_p, glob_p, loc_l, glob_l = RunningMoments(), RunningMoments(), RunningMoments(), RunningMoments()
    ep = torch.zeros(3, dtype=torch.float64, device=dev)

    rec = {k: [] for k in ("xy", "cnt", "los", "rew", "done")}
    maps_last, rows, ep_merged, all_readings = [], [], [], []
    for r in range(rollouts):
        env.reset()
        for t in range(T):
            a = torch.as_tensor(acts_all[r * T + t, base:base + count], device=dev)
            xy, rw, dn, _, _ = env.step(a)
            rec["xy"].append(xy.cpu().numpy().copy())
            rec["rew"].append(rw.cpu().numpy().copy())
            rec["done"].append(dn.cpu().numpy().copy())
            rec["cnt"].append(env.buffer("LAST_COUNT").cpu().numpy().copy())
            rec["los"].append(env.buffer("LAST_LOS").cpu().numpy().copy())
        maps_last.append(env.obs_maps.cpu().numpy().copy())
        all_readings.append(readings.cpu().numpy().copy())
        loc_p.update(readings.view(-1))
        loc_l.update(readings.view(-1))
        env.episode_stats(ep)
        ex_peer.exchange(loc_p, glob_p, ep)
        rows.append(ex_peer.gathered.cpu().numpy().copy())
        ep_merged.append(ex_peer.ep_merged.cpu().numpy().copy())
        ex_lib.exchange(loc_l, glob_l, ep)
        assert np.array_equal(ex_lib.gathered.cpu().numpy(), rows[-1]), "transports gathered different rows"
        assert np.array_equal(ex_lib.ep_merged.cpu().numpy(), ep_merged[-1])
        assert np.array_equal(loc_p.state.cpu().numpy(), np.zeros(3)), "the per-rollout accumulator must be cleared"
    assert ex_peer.errors() == 0 and env.errors() == 0
    g_peer, g_lib = glob_p.state.cpu().numpy(), glob_l.state.cpu().numpy()
    assert np.array_equal(g_peer, g_lib), "peer-memory and library transports disagree"

    # oracle: rank-ordered Chan merge of the gathered rows, folded into the running global state
    want = np.zeros(3)
    for r in range(rollouts):
        delta = np.zeros(3)
        for k in range(world):
            delta = O.moments_merge(delta, rows[r][k, :3].copy())
        want = O.moments_merge(want, delta)
    assert np.array_equal(g_peer, want), (g_peer, want)
    assert g_peer[0] == float(E_total * A * T * rollouts)

    # rt_comm_create_shm: the same mailboxes without any communication library
    L = _lib.lib()
    import ctypes as C
    comm = C.c_void_p()
    name = os.environ["RT_SHM_NAME"].encode()
    _lib.check(L.rt_comm_create_shm(name, rank, world, C.byref(comm)), "rt_comm_create_shm")
    a_, b_ = RunningMoments(), RunningMoments()
    a_.update(readings.view(-1))
    out6 = torch.zeros((world, 6), dtype=torch.float64, device=dev)
    _lib.check(L.rt_stats_allgather_merge(comm, a_._h, b_._h, ep.data_ptr(), None, out6.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(out6.cpu().numpy(), rows[-1]), "shm-connected mailboxes gathered different rows"
    L.rt_comm_disconnect(comm)
    dist.barrier()
    L.rt_comm_destroy(comm)

    # everything to rank 0 for the comparison with the one-GPU run
    payload = dict(rank=rank, base=base, count=count, rec={k: np.stack(v) for k, v in rec.items()},
                   maps=np.stack(maps_last), readings=np.stack(all_readings), glob=g_peer.tobytes())
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    result = {"rank": rank, "ok": True, "backend_peer": ex_peer.backend, "count": float(g_peer[0])}
    if rank == 0:
        gathered.sort(key=lambda p: p["rank"])
        assert all(p["glob"] == gathered[0]["glob"] for p in gathered), "merged moments differ between ranks"
        one = BatchedRadSearchEnv(**dict(cfg, n_envs=E_total, env_id_base=0))
        rd1 = torch.zeros((T, E_total, A), dtype=torch.float32, device=dev)
        one.set_rollout(rd1)
        sums = []
        for r in range(rollouts):
            one.reset()
            for t in range(T):
                xy, rw, dn, _, _ = one.step(torch.as_tensor(acts_all[r * T + t], device=dev))
                i = r * T + t
                for key, got in (("xy", xy), ("rew", rw), ("done", dn), ("cnt", one.buffer("LAST_COUNT")),
                                 ("los", one.buffer("LAST_LOS"))):
                    sharded = np.concatenate([p["rec"][key][i] for p in gathered], axis=0)
                    assert np.array_equal(sharded, got.cpu().numpy()), f"{key} differs at rollout {r} step {t}"
            sharded_maps = np.concatenate([p["maps"][r] for p in gathered], axis=0)
            assert np.array_equal(sharded_maps, one.obs_maps.cpu().numpy()), f"observation maps differ (rollout {r})"
            sharded_rd = np.concatenate([p["readings"][r] for p in gathered], axis=1)
            assert np.array_equal(sharded_rd, rd1.cpu().numpy())
            sums.append([float(one.buffer("EP_RETURN").sum().item()), float(one.buffer("TICK").sum().item()) * A,
                         float(one.buffer("EP_DONE").sum().item())])
            np.testing.assert_array_equal(ep_merged[r], np.array(sums[-1]))  # episode stats == the plain sums
        # sequential pass over ALL readings in global order agrees with the merged moments (rounding only)
        seq = np.zeros(3)
        for r in range(rollouts):
            seq = O.moments_update(seq, np.concatenate([p["readings"][r] for p in gathered], axis=1).reshape(-1))
        np.testing.assert_allclose(g_peer, seq, rtol=1e-12)
        result["episode_stats"] = sums
    print("RESULT " + json.dumps(result), flush=True)
    ex_peer.close()
    ex_lib.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
