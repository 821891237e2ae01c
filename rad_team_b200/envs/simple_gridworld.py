"""``SimpleGrid`` — the reference's single-agent gridworld API (src/envs/simple_gridworld.py:34-195),
served by the B200 kernels: one env, one agent, fixed start / terminal, every ``reset`` / ``step`` going
through the C ABI (rt_env_reset_host / rt_env_step_host).  Same names, argument meaning, return types and
exceptions as the reference, so the reference's tests/test_envs.py reads unchanged against it.
"""
from __future__ import annotations

from enum import IntEnum
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import _lib
from .radsearch import BatchedRadSearchEnv, RadSearchConfig

try:  # the reference subclasses gymnasium.Env; do the same when gymnasium is installed
    from gymnasium import Env as _Base  # type: ignore
except Exception:  # pragma: no cover - gymnasium is absent from this image
    _Base = object


class Action(IntEnum):
    """simple_gridworld.py:15-19 — 1-based; 0 and >= 5 are invalid."""
    UP = 1
    RIGHT = 2
    LEFT = 3
    DOWN = 4


#: action -> (dx, dy); y grows upward (simple_gridworld.py:26-31)
ActionDirectionMap: Dict[int, np.ndarray] = {
    Action.UP: np.array([0, 1], dtype=np.int32),
    Action.RIGHT: np.array([1, 0], dtype=np.int32),
    Action.LEFT: np.array([-1, 0], dtype=np.int32),
    Action.DOWN: np.array([0, -1], dtype=np.int32),
}


class SimpleGrid(_Base):
    metadata: dict = {"render_modes": ["human", "rgb_array"], "render_fps": 4}
    window_size: int = 512
    seed: Optional[int] = None
    spec = None

    def __init__(self, start, terminal, size: int = 10, render_mode: Optional[str] = None,
                 max_steps: int = 65534) -> None:
        if _Base is not object:
            super().__init__()
        self.start = start
        self.terminal = terminal
        self.size = size
        assert render_mode is None or render_mode in self.metadata["render_modes"]
        self.render_mode = render_mode
        self.reward_range = (-np.inf, np.inf)
        # spaces as in the reference (:103-104) when gymnasium is importable; plain descriptions otherwise
        try:
            from gymnasium import spaces as _sp  # type: ignore

            self.action_space = _sp.MultiDiscrete([a.value for a in Action])
            self.observation_space = _sp.Box(low=0, high=self.size - 1, shape=(2,), dtype=np.int32)
        except Exception:  # pragma: no cover - gymnasium is absent from this image
            self.action_space = {"type": "MultiDiscrete", "nvec": [a.value for a in Action]}
            self.observation_space = {"type": "Box", "low": 0, "high": self.size - 1, "shape": (2,), "dtype": np.int32}
        cfg = RadSearchConfig(n_envs=1, n_agents=1, grid=int(size), max_steps=int(max_steps), fixed_scenario=1)
        self._env = BatchedRadSearchEnv(cfg)
        self._env.set_scenario(src_xy=np.asarray(terminal, np.int32), src_int=[1e6], bkg=[10.0],
                               start_xy=np.asarray(start, np.int32), n_poly=[0])
        # like the reference, no reset() here; the agent sits on the start cell (:112)
        self.agent = np.asarray(self.start, dtype=np.int32)
        self.agent_previous_dist = self._get_distance()

    # -- Gymnasium API -------------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
        xy, _ = self._env.reset(host=True)
        self.agent = xy[0, 0].astype(np.int32)
        self.agent_previous_dist = self._get_distance()
        return self._get_obs(), self._get_info()

    def step(self, action: int) -> Tuple[np.ndarray, int, bool, bool, Dict[str, Any]]:
        action = Action(action)  # ValueError("0 is not a valid Action") like the reference (:138)
        xy, rew, done, trunc, _ = self._env.step(np.array([[int(action)]], dtype=np.int32))
        flags = self._env.errors()
        if flags & _lib.FLAG_LOG_FULL:
            # the reference env has no episode length limit; the device keeps a bounded per-env reading log
            raise MemoryError(f"episode longer than max_steps={self._env.cfg.max_steps} steps without reset(): the device "
                              "reading log is bounded (max_steps * agents <= 65534); the step was not executed")
        if flags:
            raise Exception(f"Something has gone wrong in the step function (device flags {flags:#x})")
        self.agent = xy[0, 0].astype(np.int32)
        distance = self._get_distance()
        self.agent_previous_dist = distance
        return self._get_obs(), int(rew[0, 0]), bool(done[0, 0]), False, self._get_info(distance=distance)

    def render(self):
        if self.render_mode == "rgb_array":
            return self._render_frame()
        return None

    def close(self) -> None:
        self._env.close()

    # -- helpers with the reference's names ------------------------------------------------------
    def _get_distance(self):
        return np.linalg.norm(self.agent - np.asarray(self.terminal), ord=1)  # np.float64, as at runtime

    def _get_oracle_obs(self) -> Dict[str, np.ndarray]:
        return {"agent": self.agent, "target": self.terminal}

    def _get_obs(self) -> np.ndarray:
        return self.agent

    def _get_info(self, distance=None) -> Dict[str, Any]:
        distance = distance if distance else self._get_distance()
        return {**self._get_oracle_obs(), "distance": distance}

    def _render_frame(self):
        """rgb_array frame in the reference's colours (simple_gridworld.py:197-263), drawn with
        NumPy from the agent / terminal cells — no pygame dependency."""
        n = self.window_size
        pix = n / self.size
        img = np.full((n, n, 3), 255, dtype=np.uint8)

        def fill(cell, colour):
            x0, y0 = int(pix * cell[0]), int(pix * cell[1])
            img[y0:int(y0 + pix), x0:int(x0 + pix)] = colour

        fill(np.asarray(self.terminal), (255, 0, 0))
        cx, cy = (np.asarray(self.agent) + 0.5) * pix
        yy, xx = np.mgrid[0:n, 0:n]
        img[(xx - cx) ** 2 + (yy - cy) ** 2 <= (pix / 3) ** 2] = (0, 0, 255)
        for k in range(self.size + 1):
            p = min(int(pix * k), n - 1)
            img[max(p - 1, 0):p + 2, :] = 0
            img[:, max(p - 1, 0):p + 2] = 0
        return img
