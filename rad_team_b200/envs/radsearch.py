"""Batched radiation-search environment on one B200 — the Python face of rt_env_* (include/rt_b200.h).

Mirrors the Gymnasium protocol of the reference env (src/envs/simple_gridworld.py:114-167):
``reset(seed=None, options=None) -> (obs, info)`` and
``step(actions) -> (obs, reward, terminated, truncated, info)``, with a leading ``[E, A]`` on everything.

Two ways to call it:
  * device path — ``actions`` is a CUDA int32 tensor: kernels are enqueued on torch's current stream and
    CUDA tensors come back; nothing synchronises.
  * host path — ``actions`` is a NumPy array / CPU tensor: the C ABI copies it up, steps, copies
    obs / reward / done back (pinned) and returns NumPy arrays.  The heat-maps stay on the device: they
    are written straight into ``obs_maps``, the CNN's input tensor.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import asdict, dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import _lib
from ..memory import obs_tensor


@dataclass
class RadSearchConfig:
    """Field-for-field rt_env_cfg (include/rt_b200.h)."""
    n_envs: int = 1
    n_agents: int = 1
    grid: int = 32
    max_steps: int = 120
    poly_min: int = 0
    poly_max: int = 0
    rect_min: int = 2
    rect_max: int = 6
    lognorm_step: int = 2
    fixed_scenario: int = 0
    env_id_base: int = 0
    seed: int = 0
    cell_pitch_cm: float = 100.0
    min_dist_cm: float = 50.0
    transmission: float = 0.0
    src_lo: float = 1e6
    src_hi: float = 1e7
    bkg_lo: float = 10.0
    bkg_hi: float = 51.0

    def to_c(self) -> _lib.EnvCfg:
        return _lib.EnvCfg(**asdict(self))


_BUF_SPEC = {
    # name: (torch dtype name, shape lambda(E, A, cells, cap))
    "AGENT_XY": ("int32", lambda E, A, c, cap: (E, A, 2)),
    "PREV_DIST": ("int32", lambda E, A, c, cap: (E, A)),
    "START_XY": ("int32", lambda E, A, c, cap: (E, A, 2)),
    "SRC_XY": ("int32", lambda E, A, c, cap: (E, 2)),
    "SRC_INT": ("float64", lambda E, A, c, cap: (E,)),
    "BKG": ("float64", lambda E, A, c, cap: (E,)),
    "N_POLY": ("int32", lambda E, A, c, cap: (E,)),
    "EDGES": ("float32", lambda E, A, c, cap: (E, 20, 4)),
    "TICK": ("int32", lambda E, A, c, cap: (E,)),
    "EPISODE": ("int32", lambda E, A, c, cap: (E,)),
    "VISIT": ("uint16", lambda E, A, c, cap: (E, c)),
    "INTENSITY": ("float32", lambda E, A, c, cap: (E, c)),
    "CELLREC": ("uint16", lambda E, A, c, cap: (E, c, 4)),
    "LOG": ("int32", lambda E, A, c, cap: (cap, E, 8)),
    "WELFORD": ("float64", lambda E, A, c, cap: (E, 6)),
    "WELFORD_N": ("int32", lambda E, A, c, cap: (E,)),
    "EST_MAXMIN": ("float64", lambda E, A, c, cap: (E, 2)),
    "FLAGS": ("int32", lambda E, A, c, cap: (E,)),
    "LAST_COUNT": ("int32", lambda E, A, c, cap: (E, A)),
    "LAST_LOS": ("uint8", lambda E, A, c, cap: (E, A)),
    "LAST_LAMBDA": ("float64", lambda E, A, c, cap: (E, A)),
    "LAST_EST": ("float64", lambda E, A, c, cap: (E, A)),
    "LAST_Z": ("float64", lambda E, A, c, cap: (E, A)),
    "LAST_VALUE": ("float64", lambda E, A, c, cap: (E, A)),
    "VISIT_LUT": ("float32", lambda E, A, c, cap: (cap + 1,)),
    "EP_RETURN": ("int32", lambda E, A, c, cap: (E, A)),
    "EP_DONE": ("uint8", lambda E, A, c, cap: (E, A)),
    "VREC": ("int32", lambda E, A, c, cap: (E, (3 + min(cap, c) + 1) & ~1, 2)),
}
_ITEMSIZE = {"int32": 4, "float64": 8, "float32": 4, "uint16": 2, "uint8": 1}


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL_CTX = _NullCtx()


class BatchedRadSearchEnv:
    """E independent radiation-search envs x A agents, resident on one GPU."""

    metadata = {"render_modes": [], "render_fps": 4}

    def __init__(self, cfg: Optional[RadSearchConfig] = None, device: Optional[int] = None,
                 alloc_maps: bool = True, **kw: Any) -> None:
        torch = _lib.require_cuda()
        self._torch = torch
        self.cfg = cfg if cfg is not None else RadSearchConfig(**kw)
        self.E, self.A, self.G = self.cfg.n_envs, self.cfg.n_agents, self.cfg.grid
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._L = _lib.lib()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            c = self.cfg.to_c()
            _lib.check(self._L.rt_env_create(C.byref(c), C.byref(self._h)), "rt_env_create")
        E, A, G = self.E, self.A, self.G
        dev = self.device
        # caller-visible outputs (device).  obs_maps is the CNN input tensor; pass alloc_maps=False
        # and hand your own tensor to step(out_maps=...) to write into a rollout slot directly.
        self.obs_xy = torch.zeros((E, A, 2), dtype=torch.int32, device=dev)
        self.reward = torch.zeros((E, A), dtype=torch.int32, device=dev)
        self.done = torch.zeros((E, A), dtype=torch.uint8, device=dev)
        self.obs_maps = obs_tensor((E, 3, G, G), device=dev) if alloc_maps else None
        # host-path outputs (pinned, so the C ABI copies straight into them)
        self._h_xy = self._h_rw = self._h_dn = None
        self._inj = None
        self._rollout = None
        self._views: Dict[str, Any] = {}   # zero-copy tensor views of the handle's arrays, built once
        self._info_dev: Optional[Dict[str, Any]] = None

    # -- plumbing ------------------------------------------------------------------------------
    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._views.clear()      # the views alias memory the handle is about to free
            self._info_dev = None
            self._L.rt_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _host_bufs(self):
        if self._h_xy is None:
            # ONE pinned block laid out obs_xy | reward | done: rt_env_step_host then returns all three with one copy
            t, n = self._torch, self.E * self.A
            self._h_block = t.zeros(n * 13, dtype=t.uint8).pin_memory()
            self._h_xy = self._h_block[:n * 8].view(t.int32).view(self.E, self.A, 2)
            self._h_rw = self._h_block[n * 8:n * 12].view(t.int32).view(self.E, self.A)
            self._h_dn = self._h_block[n * 12:].view(self.E, self.A)
        return self._h_xy, self._h_rw, self._h_dn

    def buffer(self, name: str):
        """Zero-copy CUDA tensor view of one state array (rt_buf) — get_state / set_state."""
        v = self._views.get(name)
        if v is None:
            v = self._views[name] = self._make_view(name)
        elif name in ("INTENSITY", "VISIT"):  # materialised from the records on every request (sparse mode)
            _lib.check(self._L.rt_env_buffer(self._h, _lib.BUF[name], None, None), name)
        return v

    def _make_view(self, name: str):
        torch = self._torch
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _lib.check(self._L.rt_env_buffer(self._h, _lib.BUF[name], C.byref(ptr), C.byref(nbytes)), name)
        dt, shp = _BUF_SPEC[name]
        shape = shp(self.E, self.A, self.G * self.G, self.cfg.max_steps * self.A)
        n = int(np.prod(shape))
        _lib.require(n * _ITEMSIZE[dt] == nbytes.value, f"{name}: shape {shape} does not match {nbytes.value} bytes")

        class _Arr:  # __cuda_array_interface__ carrier
            pass

        a = _Arr()
        typestr = {"int32": "<i4", "float64": "<f8", "float32": "<f4", "uint16": "<u2", "uint8": "|u1"}[dt]
        a.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr.value, False),
                                      "version": 2, "strides": None}
        return torch.as_tensor(a, device=self.device)

    def state_dict(self) -> Dict[str, Any]:
        """Snapshot of every persistent array (deterministic replay / checkpoint)."""
        skip = {"VISIT_LUT"}
        return {k: self.buffer(k).clone() for k in _BUF_SPEC if k not in skip}

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        for k, v in sd.items():
            self.buffer(k).copy_(v)

    def get_state(self, out=None):
        """The handle's whole persistent state as ONE opaque uint8 CUDA tensor (rt_env_get_state): checkpoint /
        deterministic replay without walking the named buffers."""
        t = self._torch
        n = C.c_size_t()
        _lib.check(self._L.rt_env_state_bytes(self._h, C.byref(n)))
        out = t.empty(n.value, dtype=t.uint8, device=self.device) if out is None else out
        _lib.require(out.dtype == t.uint8 and out.is_contiguous() and out.numel() == n.value, "state blob size / dtype")
        with self._on_device():
            _lib.check(self._L.rt_env_get_state(self._h, out.data_ptr(), self._stream()), "rt_env_get_state")
        return out

    def set_state(self, blob) -> None:
        """Restore a blob from get_state() of an env of the same shape (ValueError otherwise)."""
        t = self._torch
        _lib.require(blob.dtype == t.uint8 and blob.is_contiguous(), "state blob must be a contiguous uint8 tensor")
        with self._on_device():
            _lib.check(self._L.rt_env_set_state(self._h, blob.data_ptr(), self._stream()), "rt_env_set_state")

    def set_scenario(self, src_xy=None, src_int=None, bkg=None, start_xy=None, n_poly=None, edges=None) -> None:
        E, A = self.E, self.A

        def prep(a, dt, shape):
            return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=dt).reshape(shape))

        arrs = [prep(src_xy, np.int32, (E, 2)), prep(src_int, np.float64, (E,)), prep(bkg, np.float64, (E,)),
                prep(start_xy, np.int32, (E, A, 2)), prep(n_poly, np.int32, (E,)), prep(edges, np.float32, (E, 20, 4))]
        ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
        with self._torch.cuda.device(self.device):
            _lib.check(self._L.rt_env_set_scenario(self._h, *ptrs), "rt_env_set_scenario")

    def set_incremental(self, enable: bool = True) -> None:
        """Opt in to incremental observation updates: a step that is handed the same ``obs_maps`` tensor as
        the previous rasterisation rewrites only the cells the step changed (the caller must not modify
        the tensor in between).  Meant for inference / evaluation loops; rollout buffers use distinct slots
        and are always rasterised in full."""
        _lib.check(self._L.rt_env_set_incremental(self._h, 1 if enable else 0))

    def inject_uniforms(self, u) -> None:
        """Test hook: Poisson sampler reads u[E, A, K] instead of Philox (None switches back)."""
        if u is None:
            self._inj = None
            _lib.check(self._L.rt_env_inject_uniforms(self._h, None, 0))
            return
        t = self._torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64)).reshape(self.E, self.A, -1)
        self._inj = t.to(self.device).contiguous()
        _lib.check(self._L.rt_env_inject_uniforms(self._h, self._inj.data_ptr(), self._inj.shape[2]))

    def set_rollout(self, readings) -> None:
        """Attach a float32 CUDA tensor [T, E, A]: each step stores the agents' counts at row `tick`."""
        if readings is None:
            self._rollout = None
            _lib.check(self._L.rt_env_set_rollout(self._h, None))
            return
        t = self._torch
        _lib.require(readings.is_cuda and readings.dtype == t.float32 and readings.is_contiguous()
                     and tuple(readings.shape) == (self.cfg.max_steps, self.E, self.A),
                     "rollout readings must be a contiguous float32 CUDA tensor [max_steps, E, A]")
        self._rollout = readings
        _lib.check(self._L.rt_env_set_rollout(self._h, readings.data_ptr()))

    def episode_stats(self, out=None):
        """float64 CUDA tensor [3]: sum of returns, sum of episode lengths, agents that reached the source."""
        t = self._torch
        out = t.empty(3, dtype=t.float64, device=self.device) if out is None else out
        with t.cuda.device(self.device):
            _lib.check(self._L.rt_env_episode_stats(self._h, out.data_ptr(), self._stream()))
        return out

    def errors(self) -> int:
        """OR of all envs' sticky RT_FLAG_* bits (synchronises)."""
        f = C.c_int32(0)
        _lib.check(self._L.rt_env_errors(self._h, C.byref(f), self._stream()))
        return f.value

    @property
    def launches(self) -> int:
        return int(self._L.rt_env_launch_count(self._h))

    def _maps_ptr(self, out_maps):
        m = self.obs_maps if out_maps is None else out_maps
        if m is None:
            return None, None
        _lib.require(m.is_cuda and m.dtype == self._torch.float32 and m.is_contiguous()
                     and m.numel() == self.E * 3 * self.G * self.G and m.device == self.device,
                     "obs maps must be a contiguous float32 tensor [E, 3, G, G] on the env's device")
        return m, m.data_ptr()

    def _info(self, xy, maps) -> Dict[str, Any]:
        # the reference's oracle view {agent, target, distance} (simple_gridworld.py:192-195); the values
        # are views of persistent device arrays, so the dict is built once and only `maps` changes
        if self._info_dev is None:
            self._info_dev = {"agent": xy, "target": self.buffer("SRC_XY"), "distance": self.buffer("PREV_DIST"),
                              "count": self.buffer("LAST_COUNT"), "maps": maps}
        self._info_dev["maps"] = maps
        return self._info_dev

    def _on_device(self):
        """Context that makes this env's GPU current — a no-op object when it already is."""
        t = self._torch
        return _NULL_CTX if t.cuda.current_device() == self.device.index else t.cuda.device(self.device)

    # -- Gymnasium protocol --------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None, mask=None, out_maps=None,
              host: bool = False):
        """SimpleGrid.reset (simple_gridworld.py:114-129), batched; ``mask`` selects envs."""
        torch = self._torch
        maps, mptr = self._maps_ptr(out_maps)
        with self._on_device():
            if host or isinstance(mask, np.ndarray):
                hxy, _, _ = self._host_bufs()
                mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
                _lib.check(self._L.rt_env_reset_host(self._h, None if mk is None else mk.ctypes.data_as(C.c_void_p),
                                                     hxy.data_ptr(), mptr, self._stream()), "rt_env_reset_host")
                xy = hxy.numpy().copy()
                return xy, {"agent": xy, "maps": maps}
            mk = None
            if mask is not None:
                mk = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            _lib.check(self._L.rt_env_reset(self._h, None if mk is None else mk.data_ptr(), self.obs_xy.data_ptr(),
                                            mptr, self._stream()), "rt_env_reset")
        return self.obs_xy, self._info(self.obs_xy, maps)

    def step(self, actions, out_maps=None, rasterize: bool = True, out_xy=None, out_reward=None,
             out_done=None) -> Tuple[Any, Any, Any, Any, Dict[str, Any]]:
        """SimpleGrid.step (simple_gridworld.py:131-167), batched, plus measurement and map rasterisation
        (``rasterize=False`` leaves the maps to a later ``rasterize()`` call).  On the device path
        ``out_xy`` (int32 [E,A,2]), ``out_reward`` (int32 [E,A]) and ``out_done`` (uint8 [E,A]) may name
        caller-owned CUDA tensors — e.g. slots of a rollout buffer — to be written instead of the env's own."""
        torch = self._torch
        maps, mptr = self._maps_ptr(out_maps) if rasterize else (None, None)
        with self._on_device():
            if isinstance(actions, torch.Tensor) and actions.is_cuda:
                _lib.require(actions.dtype == torch.int32 and actions.is_contiguous()
                             and actions.numel() == self.E * self.A and actions.device == self.device,
                             "actions must be a contiguous int32 tensor [E, A] on the env's device")
                xy = self.obs_xy if out_xy is None else out_xy
                rw = self.reward if out_reward is None else out_reward
                dn = self.done if out_done is None else out_done
                for t_, dt_, n_ in ((xy, torch.int32, 2), (rw, torch.int32, 1), (dn, torch.uint8, 1)):
                    _lib.require(t_.is_cuda and t_.dtype == dt_ and t_.is_contiguous()
                                 and t_.numel() == self.E * self.A * n_ and t_.device == self.device,
                                 "step outputs must be contiguous CUDA tensors of the documented dtype and size")
                _lib.check(self._L.rt_env_step(self._h, actions.data_ptr(), xy.data_ptr(), mptr, rw.data_ptr(),
                                               dn.data_ptr(), self._stream()), "rt_env_step")
                info = self._info(self.obs_xy, maps)
                if out_xy is not None:
                    info = dict(info, agent=xy)
                return xy, rw, dn, False, info
            # host path
            if isinstance(actions, torch.Tensor):
                _lib.require(actions.dtype == torch.int32 and actions.is_contiguous(),
                             "host actions must be a contiguous int32 tensor")
                aptr, keep = actions.data_ptr(), actions
            else:
                keep = np.ascontiguousarray(actions, dtype=np.int32)
                aptr = keep.ctypes.data
            _lib.require((keep.numel() if hasattr(keep, "numel") else keep.size) == self.E * self.A,
                         "actions must hold E * A entries")
            hxy, hrw, hdn = self._host_bufs()
            rc = self._L.rt_env_step_host(self._h, aptr, hxy.data_ptr(), hrw.data_ptr(), hdn.data_ptr(), mptr,
                                          self._stream())
            if rc == _lib.RT_ERR_BAD_ACTION:
                bad = np.asarray(keep).reshape(-1)
                bad = bad[(bad < 1) | (bad > 4)]
                raise ValueError(f"{int(bad[0])} is not a valid Action")  # IntEnum's message (:138)
            _lib.check(rc, "rt_env_step_host")
        xy = hxy.numpy()
        return xy, hrw.numpy(), hdn.numpy(), False, {"agent": xy, "maps": maps}

    def rasterize(self, out_maps=None):
        maps, mptr = self._maps_ptr(out_maps)
        with self._on_device():
            _lib.check(self._L.rt_env_rasterize(self._h, mptr, self._stream()), "rt_env_rasterize")
        return maps

    def render(self, env_index: int = 0, window_size: int = 512, out_maps=None):
        """``rgb_array`` frame of ONE env built from the DEVICE maps (the reference draws its
        single agent with pygame, simple_gridworld.py:197-263): the observation tensor's own channels — intensity
        in red, visits in green, agent cells in blue — plus the source cell (the reference's red target square) and
        the obstruction rectangles, upsampled with nearest-neighbour to ``window_size`` pixels.  Row 0 of the frame
        is the TOP of the world (y grows upward, as in the reference's ``ActionDirectionMap``).  Returns a uint8
        NumPy array [window_size, window_size, 3]; only that one frame crosses PCIe."""
        t = self._torch
        maps = self.obs_maps if out_maps is None else out_maps
        _lib.require(maps is not None, "render needs the observation tensor (alloc_maps=True or out_maps=...)")
        _lib.require(0 <= env_index < self.E, "env_index out of range")
        G = self.G
        m = maps.view(self.E, 3, G, G)[env_index]                     # [3, G, G], map[row = y][col = x]
        inten = m[1].clamp(0, 1)
        visit = m[2].clamp(0, 1)
        img = t.ones((3, G, G), dtype=t.float32, device=maps.device)
        img[1] -= 0.8 * inten                                         # intensity: white -> red
        img[2] -= 0.8 * inten
        img[0] -= 0.5 * visit * (1 - inten)                           # visited, low intensity: towards green
        img[2] -= 0.5 * visit * (1 - inten)
        # obstructions: grey cells
        n_poly = int(self.buffer("N_POLY")[env_index].item())
        edges = self.buffer("EDGES")[env_index]
        for p in range(n_poly):
            e = edges[4 * p:4 * p + 4]
            x0, x1 = int(e[:, [0, 2]].min().item()), int(e[:, [0, 2]].max().item())
            y0, y1 = int(e[:, [1, 3]].min().item()), int(e[:, [1, 3]].max().item())
            if x1 > x0 and y1 > y0:
                img[:, y0:y1, x0:x1] = 0.45
        sx, sy = (int(v) for v in self.buffer("SRC_XY")[env_index].tolist())
        img[:, sy, sx] = t.tensor([1.0, 0.0, 0.0], device=maps.device)   # the source: the reference's red target
        agents = m[0] > 0.5
        img[0][agents], img[1][agents], img[2][agents] = 0.0, 0.0, 1.0  # agents: the reference's blue
        img = img.flip(1)                                             # y up
        scale = max(1, window_size // G)
        img = img.repeat_interleave(scale, 1).repeat_interleave(scale, 2)
        frame = t.zeros((3, window_size, window_size), dtype=t.float32, device=maps.device)
        h = min(window_size, img.shape[1])
        frame[:, :h, :h] = img[:, :h, :h]
        frame[:, ::scale, :] = 0.0                                    # grid lines
        frame[:, :, ::scale] = 0.0
        return (frame.permute(1, 2, 0) * 255).round().to(t.uint8).cpu().numpy()
