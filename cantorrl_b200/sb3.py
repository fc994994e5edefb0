"""Stable-Baselines3 adapter: the host-buffer env as a real ``VecEnv`` subclass.

The reference wraps N ``HedgingEnv`` objects in ``SubprocVecEnv`` / ``DummyVecEnv`` and then ``VecNormalize``
(``src/agents/train_ppo_v2.py:127-141, 204-208``); SB3's wrappers and algorithms insist on an instance of their own
``VecEnv`` base class.  ``cantor_vec_env_class()`` builds that subclass over ``HostVecEnv`` (NumPy in / NumPy out through
``cantor_vecenv_step_host``).  SB3 is not a dependency of this package: the base class is imported when the factory is
called, or passed in (the tests pass a stand-in with SB3's abstract interface).

    from cantorrl_b200.sb3 import cantor_vec_env_class
    CantorVecEnv = cantor_vec_env_class()                       # needs stable_baselines3
    venv = CantorVecEnv(DATA_FILE, lambda_cost=1e-4, slippage_bps=1.0, num_envs=4096)
    venv = VecNormalize(venv, norm_obs=True, norm_reward=True, gamma=0.99)       # train_ppo_v2.py:204-208
"""
from __future__ import annotations

import numpy as np

from .host_env import HostVecEnv


class _LazyInfos(list):
    """``infos`` as SB3 consumes it: a real ``list`` of per-env dicts (``VecEnvWrapper`` code indexes, iterates and copies it).
    Every env shares ONE empty dict unless it finished this step; SB3 only reads ``infos[i]`` (``.get(...)``, ``in``)."""


def cantor_vec_env_class(base=None):
    """Returns ``CantorVecEnv``, a subclass of ``base`` (default: ``stable_baselines3.common.vec_env.VecEnv``) and ``HostVecEnv``."""
    if base is None:
        from stable_baselines3.common.vec_env import VecEnv as base          # noqa: N813  (raises ImportError without SB3)

    class CantorVecEnv(HostVecEnv, base):
        # plain attributes (SB3's VecEnv.__init__ assigns them); they shadow HostVecEnv's read-only properties
        observation_space = None
        action_space = None

        def __init__(self, *args, observation_space=None, action_space=None, **kwargs):
            HostVecEnv.__init__(self, *args, **kwargs)
            obs_space = observation_space if observation_space is not None else HostVecEnv.observation_space.fget(self)
            act_space = action_space if action_space is not None else HostVecEnv.action_space.fget(self)
            base.__init__(self, self.num_envs, obs_space, act_space)
            self._actions = None
            self._empty = {}

        # -- SB3's abstract interface -------------------------------------------------------------------------
        def reset(self):
            return HostVecEnv.reset(self).copy()                  # SB3 keeps references to returned arrays across steps

        def step_async(self, actions):
            self._actions = np.asarray(actions, np.float32)

        def step(self, actions):                                 # SB3's VecEnv.step; HostVecEnv.step would win the MRO otherwise
            self.step_async(actions)
            return self.step_wait()

        def step_wait(self):
            obs, reward, done, _ = HostVecEnv.step(self, self._actions)
            infos = _LazyInfos([self._empty] * self.num_envs)
            for i in np.nonzero(done)[0]:                        # truncated is always False in this env (hedging_env_v2.py:221)
                infos[i] = {"TimeLimit.truncated": False}
            return obs.copy(), reward.astype(np.float32, copy=True), done.copy(), infos

        def close(self):
            HostVecEnv.close(self)

        def get_attr(self, attr_name, indices=None):
            return HostVecEnv.get_attr(self, attr_name, indices)

        def set_attr(self, attr_name, value, indices=None):
            return HostVecEnv.set_attr(self, attr_name, value, indices)

        def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
            return HostVecEnv.env_method(self, method_name, *method_args, indices=indices, **method_kwargs)

        def env_is_wrapped(self, wrapper_class, indices=None):
            return HostVecEnv.env_is_wrapped(self, wrapper_class, indices)

        def seed(self, seed=None):
            return HostVecEnv.seed(self, seed)

    return CantorVecEnv
