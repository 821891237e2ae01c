"""Multi-GPU plumbing: one process per GPU, environments sharded in contiguous blocks, no cross-GPU
traffic in `step`.  The only exchange is once per rollout: every rank contributes six doubles — the
moments of ITS readings of this rollout (count, mean, M2) and its episode statistics (sum of returns, sum
of lengths, agents that reached the source) — and every rank merges the rows in rank order, so all
ranks hold bit-identical results.  Moments are never all-reduce-summed as
(sum x, sum x^2): that cancels catastrophically and its order is not deterministic.

The exchange itself lives behind the C ABI (include/rt_b200.h, csrc/rt_comm.cu):

  * ``backend="peer"`` — ``rt_stats_allgather_merge``: ONE kernel per rank that stores the row into every
    peer's mailbox over NVLink (CUDA-IPC-mapped device memory), waits on sequence flags and merges.  The
    mailbox handles are exchanged once, at construction, through the process group (64 bytes per rank).
  * ``backend="nccl"`` — ``rt_stats_pack`` -> ``all_gather_into_tensor`` (NCCL over NVLink on GPUs, gloo in
    the CPU tests) -> ``rt_stats_merge_gathered``.

Both fold the merged moments of the rollout into a persistent GLOBAL state and clear the per-rollout LOCAL
accumulator: a running state that doubled as the accumulator would be counted once per rank at every later
merge.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

PACK = 6  # count, mean, M2, sum_return, sum_len, n_done


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [base, base + count) of global env ids owned by `rank`; the first
    n_total % world ranks own one extra env."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    q, r = divmod(n_total, world)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def pack_stats(moments_state: torch.Tensor, episode_stats: torch.Tensor) -> torch.Tensor:
    """[6] float64 on the tensors' device (host-side reference of rt_stats_pack; used by the CPU tests)."""
    return torch.cat([moments_state.reshape(3), episode_stats.reshape(3)]).contiguous()


def gather_stats(packed: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank [6] rows into [world, 6] (row r = rank r).  Works for any backend."""
    if not (dist.is_available() and dist.is_initialized()):
        return packed.reshape(1, PACK).clone()
    world = dist.get_world_size(group)
    out = torch.empty((world, PACK), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out.view(-1), packed.contiguous(), group=group)
    return out


def _world_rank(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class RolloutStatsExchange:
    """The once-per-rollout exchange.  ``exchange(local, global_, ep_stats)`` merges every rank's per-rollout
    ``RunningMoments`` `local` into `global_` (identically on all ranks), clears `local`, and returns the
    episode statistics summed over the ranks as a float64 CUDA tensor [3].  Nothing synchronises the host."""

    def __init__(self, backend: str = "auto", group: Optional[dist.ProcessGroup] = None, device=None) -> None:
        t = _lib.require_cuda()
        self._L = _lib.lib()
        self.group = group
        self.world, self.rank = _world_rank(group)
        self.device = t.device("cuda", t.cuda.current_device() if device is None else device)
        self._comm = C.c_void_p()
        self.ep_merged = t.zeros(3, dtype=t.float64, device=self.device)
        self.gathered = t.zeros((self.world, PACK), dtype=t.float64, device=self.device)
        self._packed = t.zeros(PACK, dtype=t.float64, device=self.device)
        self.peer_error: Optional[str] = None
        _lib.require(backend in ("auto", "peer", "nccl"), "backend must be auto, peer or nccl")
        _lib.require(self.world <= _lib.COMM_MAX_RANKS, "at most %d ranks per box" % _lib.COMM_MAX_RANKS)
        want_peer = backend in ("auto", "peer")
        if want_peer:
            ok = self._connect_peers()
            if not ok and backend == "peer":
                raise _lib.RtError(f"peer mailboxes unavailable: {self.peer_error}")
            want_peer = ok
        self.backend = "peer" if want_peer else "nccl"

    def _connect_peers(self) -> bool:
        """Create the mailbox, exchange the 64-byte handles through the process group, map the peers.  All ranks
        agree on the outcome (a MIN all-reduce), so either every rank uses the peer kernel or none does."""
        t = torch
        handle = (C.c_ubyte * _lib.COMM_HANDLE_BYTES)()
        with t.cuda.device(self.device):
            rc = self._L.rt_comm_create(self.rank, self.world, C.byref(self._comm), handle)
        ok = rc == _lib.RT_OK
        if not ok:
            self.peer_error = "rt_comm_create: " + _lib.status_string(rc)
        if self.world > 1:
            backend = dist.get_backend(self.group)
            dev = self.device if backend == "nccl" else t.device("cpu")
            mine = t.tensor(list(bytes(handle)), dtype=t.uint8, device=dev)
            allh = t.empty(self.world * _lib.COMM_HANDLE_BYTES, dtype=t.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=self.group)
            if ok:
                buf = (C.c_ubyte * (self.world * _lib.COMM_HANDLE_BYTES)).from_buffer_copy(bytes(allh.cpu().tolist()))
                with t.cuda.device(self.device):
                    rc = self._L.rt_comm_connect(self._comm, buf)
                if rc != _lib.RT_OK:
                    ok = False
                    self.peer_error = "rt_comm_connect: " + _lib.status_string(rc)
            flag = t.tensor([1 if ok else 0], dtype=t.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if ok and int(flag.item()) == 0:
                self.peer_error = "a peer rank could not map the mailboxes"
            ok = bool(int(flag.item()))
        if not ok and self._comm.value:
            self._L.rt_comm_destroy(self._comm)
            self._comm = C.c_void_p()
        return ok

    def close(self) -> None:
        """Collective when world > 1: every rank unmaps its peers, the ranks meet at a barrier, then every rank
        frees its own mailbox (an exported mailbox must outlive its mappings in the other processes)."""
        if getattr(self, "_comm", None) is not None and self._comm.value:
            self._L.rt_comm_disconnect(self._comm)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
            self._L.rt_comm_destroy(self._comm)
            self._comm = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover  (no collective from a finaliser: this rank's side only)
        try:
            if self._comm.value:
                self._L.rt_comm_destroy(self._comm)
                self._comm = C.c_void_p()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def exchange(self, local, global_, ep_stats: Optional[torch.Tensor] = None, backend: Optional[str] = None):
        """One exchange on torch's current stream.  `local`, `global_`: RunningMoments; `ep_stats`: float64 CUDA
        tensor [3] (rt_env_episode_stats) or None.  Returns ``self.ep_merged`` (overwritten by the next call);
        ``self.gathered`` holds the gathered rows [world, 6]."""
        be = self.backend if backend is None else backend
        ep = None if ep_stats is None else ep_stats.data_ptr()
        if be == "peer":
            _lib.require(bool(self._comm.value), "peer backend not connected")
            _lib.check(self._L.rt_stats_allgather_merge(self._comm, local._h, global_._h, ep, self.ep_merged.data_ptr(),
                                                        self.gathered.data_ptr(), self._stream()),
                       "rt_stats_allgather_merge")
        else:
            _lib.check(self._L.rt_stats_pack(local._h, ep, self._packed.data_ptr(), self._stream()), "rt_stats_pack")
            if self.world == 1:
                self.gathered.view(-1).copy_(self._packed)
            elif dist.get_backend(self.group) == "nccl":
                dist.all_gather_into_tensor(self.gathered.view(-1), self._packed, group=self.group)
            else:  # a host-memory group (gloo, in tests): stage through the host
                host = torch.empty(self.world * PACK, dtype=torch.float64)
                dist.all_gather_into_tensor(host, self._packed.cpu(), group=self.group)
                self.gathered.view(-1).copy_(host)
            _lib.check(self._L.rt_stats_merge_gathered(local._h, global_._h, self.gathered.data_ptr(), self.world,
                                                       self.ep_merged.data_ptr(), self._stream()),
                       "rt_stats_merge_gathered")
        return self.ep_merged

    def errors(self) -> int:
        """Sticky error word of the peer transport (1 = a peer did not arrive within the timeout); synchronises."""
        if not self._comm.value:
            return 0
        e = C.c_int32(0)
        _lib.check(self._L.rt_comm_errors(self._comm, C.byref(e), self._stream()), "rt_comm_errors")
        return e.value
