"""rad_team_b200 — B200-native (sm_100a) hot path of bentotten/RAD-TEAM: the batched radiation-search
environment step and its observation pipeline, behind the C ABI of include/rt_b200.h.

Sub-packages mirror the reference's layout: ``envs`` (src/envs) and ``math_utils`` (src/math_utils).
Importing the package needs no GPU; every compute call does (there is no CPU fallback).
"""
__version__ = "0.1.0"
