"""CPU ORACLE for VecNormalize / Monitor  --  TEST INFRASTRUCTURE ONLY.

The reference wraps its env in Stable-Baselines3's ``Monitor`` and ``VecNormalize``
(src/agents/train_ppo_v2.py:119, 204-208, 305-309).  stable-baselines3 (2.6.0 per
quantconnect/model_files/final_model.zip:system_info.txt) is an un-vendored dependency that is NOT installed here, so
this restates its published algorithm (``common/running_mean_std.py``, ``common/vec_env/vec_normalize.py``):

  RunningMeanStd(epsilon=1e-4): mean 0, var 1, count eps; update(x): batch mean, population variance, count folded in with
      delta = bmean - mean; tot = count + n; mean += delta n / tot; M2 = var count + bvar n + delta^2 count n / tot; var = M2 / tot
  VecNormalize.step_wait: obs_rms.update(obs) [training]; obs = clip((obs - mean) / sqrt(var + eps), +-clip_obs) -> float32;
      returns = returns gamma + reward; ret_rms.update(returns) [training]; reward = clip(reward / sqrt(ret_var + eps), +-clip_reward);
      terminal observations normalised too; returns[dones] = 0
  VecNormalize.reset: returns = 0; obs_rms.update(obs) [training]; normalise
  Monitor: info["episode"] = {"r": sum of rewards, "l": number of steps} at the end of an episode.

Parity status: the update TRAJECTORY is unpinned (no SB3 here, no stored statistics trajectory in the reference).  Two
reference anchors exist and are checked: the application formula quantconnect/model_wrapper.py:131, which ``normalize_obs``
follows, and the shipped SB3 pickle quantconnect/model_files/final_vecnormalize.pkl, readable without SB3
(cantorrl_b200/vecnorm.py: read_sb3_vecnormalize), which confirms count0 = 1e-4, the reset batch going to the observation
statistics only, clip_obs = clip_reward = 10 and epsilon = 1e-8 (tests/test_oracle_policy.py).
"""
from __future__ import annotations

import numpy as np


class RunningMeanStd:
    def __init__(self, shape=(), epsilon=1e-4):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, x):
        x = np.asarray(x, np.float64)
        bmean, bvar, n = x.mean(axis=0), x.var(axis=0), x.shape[0]
        delta = bmean - self.mean
        tot = self.count + n
        new_mean = self.mean + delta * n / tot
        m2 = self.var * self.count + bvar * n + np.square(delta) * self.count * n / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class VecNormalizeOracle:
    def __init__(self, n_envs, obs_dim=13, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0,
                 gamma=0.99, epsilon=1e-8):
        self.obs_rms, self.ret_rms = RunningMeanStd((obs_dim,)), RunningMeanStd(())
        self.returns = np.zeros(n_envs)
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon

    def normalize_obs(self, obs):
        if not self.norm_obs:
            return np.asarray(obs, np.float32)
        x = (np.asarray(obs, np.float64) - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon)   # model_wrapper.py:131
        return np.clip(x, -self.clip_obs, self.clip_obs).astype(np.float32)

    def normalize_reward(self, r):
        if not self.norm_reward:
            return np.asarray(r)
        return np.clip(np.asarray(r, np.float64) / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)

    def reset(self, obs):
        self.returns[:] = 0
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        return self.normalize_obs(obs)

    def step(self, obs, reward, done, terminal_obs=None):
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        out_obs = self.normalize_obs(obs)
        if self.training and self.norm_reward:
            self.returns = self.returns * self.gamma + np.asarray(reward, np.float64)
            self.ret_rms.update(self.returns)
        out_r = self.normalize_reward(reward)
        out_t = None if terminal_obs is None else self.normalize_obs(terminal_obs)
        self.returns[np.asarray(done, bool)] = 0
        return out_obs, out_r, out_t
