"""Stand-in for ``stable_baselines3.common.vec_env.VecEnv`` (SB3 2.6.0 is not installed here): the same abstract interface
(``base_vec_env.py``: reset, step_async, step_wait, close, get_attr, set_attr, env_method, env_is_wrapped), the same
constructor ``(num_envs, observation_space, action_space)`` and the concrete ``step`` = ``step_async`` + ``step_wait``;
plus the part of ``VecNormalize.step_wait`` that touches ``infos``.  TEST INFRASTRUCTURE."""
from abc import ABC, abstractmethod

import numpy as np


class VecEnv(ABC):
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]

    @abstractmethod
    def reset(self): ...

    @abstractmethod
    def step_async(self, actions): ...

    @abstractmethod
    def step_wait(self): ...

    @abstractmethod
    def close(self): ...

    @abstractmethod
    def get_attr(self, attr_name, indices=None): ...

    @abstractmethod
    def set_attr(self, attr_name, value, indices=None): ...

    @abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs): ...

    @abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None): ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


class VecNormalizeLike:
    """What SB3's VecNormalize / on-policy collection do with a step's outputs (shape, dtype and infos access pattern)."""

    def __init__(self, venv):
        assert isinstance(venv, VecEnv)
        self.venv = venv
        self.returns = np.zeros(venv.num_envs)

    def step(self, actions):
        obs, rewards, dones, infos = self.venv.step(actions)
        assert obs.shape == (self.venv.num_envs,) + tuple(self.venv.observation_space.shape) and obs.dtype == np.float32
        assert rewards.shape == (self.venv.num_envs,) and dones.dtype == np.bool_ and isinstance(infos, list)
        self.returns = self.returns * 0.99 + rewards
        for idx, done in enumerate(dones):
            if not done:
                continue
            if "terminal_observation" in infos[idx]:
                infos[idx]["terminal_observation"] = infos[idx]["terminal_observation"] * 1.0
            assert infos[idx].get("TimeLimit.truncated", False) is False
        self.returns[dones] = 0
        return obs, rewards, dones, infos
