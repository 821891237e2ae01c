"""Gymnasium-style vector-env adapter over the batched GPU env.

``RadSearchVectorEnv`` follows the ``gymnasium.vector.VectorEnv`` protocol — ``num_envs``,
``reset(seed=None, options=None) -> (obs, info)``, ``step(actions) -> (obs, rewards, terminations,
truncations, infos)`` with NumPy arrays carrying a leading ``num_envs`` — so code written against the
reference's Gymnasium env (src/envs/__init__.py:3-6, ``gymnasium.make("SimpleGrid", ...)``) can drive
65,536 envs with the same loop.  It subclasses ``gymnasium.vector.VectorEnv`` when gymnasium is
importable and is duck-typed otherwise (gymnasium is absent from this image).

Observations: ``obs`` is the reference's observation, the agents' grid coordinates, int32
``[num_envs, A, 2]`` on the host; the heat-maps stay on the GPU (``infos["maps"]``, the CNN's input
tensor) unless ``maps_to_host=True``.  Episodes end when an agent reaches the source (terminated) or
after ``max_steps`` steps (truncated); finished envs are reset in the same call (the final position is
reported in ``infos["final_observation"]`` / ``infos["_final_observation"]``).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from .radsearch import BatchedRadSearchEnv, RadSearchConfig

try:  # pragma: no cover - gymnasium is absent from this image
    from gymnasium.vector import VectorEnv as _Base  # type: ignore
except Exception:
    _Base = object


class RadSearchVectorEnv(_Base):
    metadata = {"render_modes": [], "autoreset_mode": "same_step"}

    def __init__(self, num_envs: int, n_agents: int = 1, grid: int = 32, max_steps: int = 120,
                 maps_to_host: bool = False, **cfg: Any) -> None:
        self.num_envs = int(num_envs)
        self.n_agents = int(n_agents)
        self.max_steps = int(max_steps)
        self.maps_to_host = bool(maps_to_host)
        self.env = BatchedRadSearchEnv(RadSearchConfig(n_envs=self.num_envs, n_agents=self.n_agents, grid=int(grid),
                                                       max_steps=self.max_steps, **cfg))
        self._tick = np.zeros(self.num_envs, np.int32)
        self.single_observation_shape = (self.n_agents, 2)
        self.single_action_shape = (self.n_agents,)
        self.closed = False

    def _infos(self) -> Dict[str, Any]:
        maps = self.env.obs_maps
        return {"maps": maps.cpu().numpy() if self.maps_to_host else maps}

    def reset(self, *, seed: Optional[int] = None, options: Optional[Dict] = None) -> Tuple[np.ndarray, Dict[str, Any]]:
        xy, _ = self.env.reset(host=True)
        self._tick[:] = 0
        return xy.copy(), self._infos()

    def step(self, actions) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, Dict[str, Any]]:
        act = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.num_envs, self.n_agents)
        # The reference raises before touching any state (simple_gridworld.py:138).  Checked here, on the host, so a
        # bad action advances NO env (the device would void only the offending envs and step the others).
        bad = (act < 1) | (act > 4)
        if bad.any():
            raise ValueError(f"{int(act[bad][0])} is not a valid Action")
        xy, rew, done, _, _ = self.env.step(act)
        self._tick += 1
        terminated = done.astype(bool).any(axis=1)
        truncated = (self._tick >= self.max_steps) & ~terminated
        obs, rewards = xy.copy(), rew.astype(np.float64).sum(axis=1)
        infos: Dict[str, Any] = {"agent_reward": rew.copy(), "agent_done": done.astype(bool)}
        finished = terminated | truncated
        if finished.any():
            infos["final_observation"] = obs.copy()
            infos["_final_observation"] = finished.copy()
            new_xy, _ = self.env.reset(mask=finished.astype(np.uint8), host=True)
            obs[finished] = new_xy[finished]
            self._tick[finished] = 0
        infos.update(self._infos())
        return obs, rewards, terminated, truncated, infos

    def errors(self) -> int:
        """OR of the envs' sticky RT_FLAG_* bits (include/rt_b200.h); 0 in a healthy run.  Synchronises."""
        return self.env.errors()

    def close(self, **kwargs: Any) -> None:
        if not self.closed:
            self.env.close()
            self.closed = True
