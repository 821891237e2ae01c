"""Environments — mirrors the reference's ``src/envs`` package (src/envs/__init__.py:3-6)."""
from .radsearch import BatchedRadSearchEnv, RadSearchConfig, obs_tensor
from .simple_gridworld import Action, ActionDirectionMap, SimpleGrid
from .vector import RadSearchVectorEnv

__all__ = ["BatchedRadSearchEnv", "RadSearchConfig", "obs_tensor", "RadSearchVectorEnv", "SimpleGrid", "Action", "ActionDirectionMap"]

try:  # same registry id as the reference when gymnasium is present
    from gymnasium.envs.registration import register  # type: ignore

    register(id="SimpleGrid", entry_point="rad_team_b200.envs.simple_gridworld:SimpleGrid")
except Exception:  # pragma: no cover - gymnasium is absent from this image
    pass
