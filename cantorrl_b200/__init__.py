"""cantorrl_b200 -- B200-native (sm_100a) implementation of CantorRL's data-parallel hot path.

Scope: the batched option-hedging environment step, the path simulation and the
Black-Scholes repricing that feed it, behind the reference's gym-style reset/step API.  All compute is
in hand-written CUDA kernels reached through the C ABI of ``include/cantor_hedge.h``; there is no CPU
fallback and importing the env without the built library raises.
"""
from ._lib import CantorError, build, lib  # noqa: F401
from .data import ReplayData, load_schema_b, save_schema_b  # noqa: F401
from .env import Box, HedgingEnv, HedgingVecEnv, VecInfo  # noqa: F401

__all__ = ["CantorError", "build", "lib", "ReplayData", "HedgingVecEnv", "HedgingEnv", "VecInfo", "Box", "load_schema_b",
           "save_schema_b"]
