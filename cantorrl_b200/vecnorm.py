"""Device-resident ``VecNormalize``: the wrapper the reference puts around its env, without a device->host round trip.

Reference use: ``VecNormalize(env, norm_obs=True, norm_reward=True, clip_obs=10.0, gamma=gamma)``
(``src/agents/train_ppo_v2.py:204-208, 305-309``), saved / loaded with the model (``:315-317, 449-455``), its statistics
exported for deployment (``quantconnect/extract_model.py:62-79``; applied as ``(obs - mean) / sqrt(var + 1e-8)``,
``quantconnect/model_wrapper.py:131``).  Stable-Baselines3 is an un-vendored dependency, so this mirrors its public
behaviour (attribute and method names included); all arithmetic runs in ``cantor_vecnorm_step`` (two chained kernels).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def read_sb3_vecnormalize(path) -> dict:
    """Statistics and hyper-parameters out of a pickle written by Stable-Baselines3's ``VecNormalize.save`` -- the file the
    reference keeps next to its model (``train_ppo_v2.py:315-317``, loaded at ``:449-455``; shipped as
    ``quantconnect/model_files/final_vecnormalize.pkl``) -- WITHOUT SB3 / gymnasium installed: their classes are replaced by
    attribute bags while unpickling, everything outside NumPy / builtins / collections is refused."""
    import pickle

    class _Bag:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, state):
            self.__dict__.update(state if isinstance(state, dict) else {"_state": state})

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            root = module.split(".")[0]
            if root in ("stable_baselines3", "sb3_contrib", "gymnasium", "gym"):
                return type(name, (_Bag,), {"__module__": module})
            if root in ("numpy", "collections", "builtins", "copyreg", "_codecs"):
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}")

    with open(path, "rb") as f:
        o = _Unpickler(f).load()
    try:
        out = {"obs_mean": np.asarray(o.obs_rms.mean, np.float64), "obs_var": np.asarray(o.obs_rms.var, np.float64),
               "obs_count": float(o.obs_rms.count), "ret_mean": float(o.ret_rms.mean), "ret_var": float(o.ret_rms.var),
               "ret_count": float(o.ret_rms.count)}
    except AttributeError as e:
        raise ValueError(f"{path}: not a VecNormalize pickle ({e})") from None
    if out["obs_mean"].shape != (13,):
        raise ValueError(f"{path}: observation statistics of shape {out['obs_mean'].shape}, expected (13,)")
    for k, default in (("clip_obs", 10.0), ("clip_reward", 10.0), ("gamma", 0.99), ("epsilon", 1e-8)):
        out[k] = float(getattr(o, k, default))
    for k in ("norm_obs", "norm_reward", "training"):
        out[k] = bool(getattr(o, k, True))
    return out


class RunningMeanStdView:
    """``mean`` / ``var`` / ``count`` of one running statistic (views of the device buffer, float64)."""

    def __init__(self, mean, var, count):
        self.mean, self.var, self._count = mean, var, count

    @property
    def count(self):
        return float(self._count)


class VecNormalize:
    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8, keep_original=False, fuse=True):
        _lib.lib()
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.training, self.norm_obs, self.norm_reward = bool(training), bool(norm_obs), bool(norm_reward)
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = float(clip_obs), float(clip_reward), float(gamma), float(epsilon)
        self.device = venv.device
        if hasattr(venv, "_rule"):          # the observations are read again on the device right away: keep them in L2
            venv._rule.flags |= _lib.STEP_KEEP_OBS_IN_L2
        self._rms = torch.zeros(_lib.VECNORM_DOUBLES, dtype=torch.float64, device=self.device)
        self.returns = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cantor_vecnorm_init(self._rms.data_ptr(), self.returns.data_ptr(), self.num_envs,
                                                      _lib.current_stream_ptr(self.device)), "cantor_vecnorm_init")
        self._rms_ptr, self._returns_ptr, self._fn = self._rms.data_ptr(), self.returns.data_ptr(), _lib.lib().cantor_vecnorm_step
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.obs_rms = RunningMeanStdView(self._rms[0:13], self._rms[13:26], self._rms[26])
        self.ret_rms = RunningMeanStdView(self._rms[27], self._rms[28], self._rms[29])
        self.keep_original = bool(keep_original)
        self.old_obs = self.old_reward = None
        self._zero_done = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        self._dummy_reward = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        # Fusion with the step kernel (cantor_vecnorm_step_fused): a HedgingVecEnv without record_info computes the batch moments
        # while its observation tile is still in shared memory, so the wrapper's own pass over the batch (the moments kernel)
        # disappears: step -> two-level fold of the per-CTA records -> in-place apply.  fuse=False (or any other venv) keeps the two-kernel form.
        self._fuse = self._fuse_ptr = None
        if fuse and hasattr(venv, "_state") and getattr(venv, "_info", None) is None and hasattr(_lib.EnvState, "vecnorm"):
            n_cta = (self.num_envs + 127) // 128
            self._partial = torch.zeros(28 * (n_cta + (n_cta + 127) // 128), dtype=torch.float64, device=self.device)
            self._fuse = _lib.VecNormFuse(self._partial.data_ptr(), self.returns.data_ptr(), self.gamma, n_cta,
                                          int(self.norm_obs), int(self.norm_reward))
            self._fuse_ptr = C.cast(C.pointer(self._fuse), C.c_void_p)
            self._fused_fn = _lib.lib().cantor_vecnorm_step_fused

    # ------------------------------------------------------------------------------------------------ core
    def _apply(self, obs, reward, done_u8, terminal_obs, norm_reward):
        args = (self._rms_ptr, self._returns_ptr, self.num_envs, obs.data_ptr(), reward.data_ptr(),
                _lib.F64 if reward.dtype is torch.float64 else _lib.F32, done_u8.data_ptr(), _lib.ptr(terminal_obs),
                self.gamma, self.clip_obs, self.clip_reward, self.epsilon, int(self.training), int(self.norm_obs), int(norm_reward))
        if torch.cuda.current_device() == self._dev_index:      # no context switch on the hot path (two launches per env-step)
            status = self._fn(*args, torch.cuda.current_stream().cuda_stream)
        else:
            with torch.cuda.device(self.device):
                status = self._fn(*args, _lib.current_stream_ptr(self.device))
        if status != 0:
            _lib.check(status, "cantor_vecnorm_step")

    def reset(self, *args, **kwargs):
        obs = self.venv.reset(*args, **kwargs)
        if self.keep_original:
            self.old_obs = obs.clone()
        self.returns.zero_()
        self._apply(obs, self._dummy_reward, self._zero_done, None, False)
        return obs

    def step(self, actions):
        """``venv.step`` followed by SB3's ``VecNormalize.step_wait``; obs / rewards are normalised in place."""
        fused = self._fuse is not None and self.training
        if self._fuse is not None:          # the step kernel produces the batch moments itself while training
            self._fuse.gamma, self._fuse.norm_obs, self._fuse.norm_reward = self.gamma, int(self.norm_obs), int(self.norm_reward)
            self.venv._state.vecnorm = self._fuse_ptr if fused else None
        obs, reward, done, infos = self.venv.step(actions)
        if self._fuse is not None:
            self.venv._state.vecnorm = None     # a direct venv.step / step_many by the caller must not pay for (or trip over) the moments
        if self.keep_original:
            self.old_obs, self.old_reward = obs.clone(), reward.clone()
        term = infos["terminal_observation"] if hasattr(infos, "__getitem__") else None
        done_u8 = self.venv._done if getattr(self.venv, "_done_bool", None) is done else done.view(torch.uint8)
        if fused:
            args = (self._rms_ptr, C.byref(self._fuse), self.num_envs, obs.data_ptr(), reward.data_ptr(),
                    _lib.F64 if reward.dtype is torch.float64 else _lib.F32, done_u8.data_ptr(), _lib.ptr(term),
                    self.clip_obs, self.clip_reward, self.epsilon)
            if torch.cuda.current_device() == self._dev_index:
                status = self._fused_fn(*args, torch.cuda.current_stream().cuda_stream)
            else:
                with torch.cuda.device(self.device):
                    status = self._fused_fn(*args, _lib.current_stream_ptr(self.device))
            if status != 0:
                _lib.check(status, "cantor_vecnorm_step_fused")
        else:
            self._apply(obs, reward, done_u8, term, self.norm_reward)
        return obs, reward, done, infos

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        return self.step(self._pending)

    # ------------------------------------------------------------------------------------- SB3 conveniences
    def normalize_obs(self, obs: torch.Tensor) -> torch.Tensor:
        if not self.norm_obs:
            return obs
        x = (obs.double() - self.obs_rms.mean) / torch.sqrt(self.obs_rms.var + self.epsilon)
        return torch.clamp(x, -self.clip_obs, self.clip_obs).float()

    def normalize_reward(self, reward: torch.Tensor) -> torch.Tensor:
        if not self.norm_reward:
            return reward
        return torch.clamp(reward.double() / torch.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward).to(reward.dtype)

    def get_original_obs(self):
        if self.old_obs is None:
            raise RuntimeError("construct VecNormalize(keep_original=True) to keep the unnormalised observations")
        return self.old_obs

    def get_original_reward(self):
        if self.old_reward is None:
            raise RuntimeError("construct VecNormalize(keep_original=True) to keep the unnormalised rewards")
        return self.old_reward

    def policy_normalisation(self):
        """(mean, var) float32 tensors for ``pack_mlp(..., obs_mean=, obs_var=)``: the fused rollout applies them itself."""
        return self.obs_rms.mean.float(), self.obs_rms.var.float()

    def get_attr(self, name, indices=None):
        return self.venv.get_attr(name, indices)

    def close(self):
        self.venv.close()

    # --------------------------------------------------------------------------------------- checkpoint / resume
    def state_dict(self):
        return {"rms": self._rms[:32].cpu(), "returns": self.returns.cpu(), "clip_obs": self.clip_obs, "clip_reward": self.clip_reward,
                "gamma": self.gamma, "epsilon": self.epsilon, "norm_obs": self.norm_obs, "norm_reward": self.norm_reward,
                "training": self.training}                     # SB3's pickle keeps the training flag too

    def load_state_dict(self, sd):
        self._rms[:32].copy_(sd["rms"])
        if sd["returns"].numel() == self.returns.numel():
            self.returns.copy_(sd["returns"])
        for k in ("clip_obs", "clip_reward", "gamma", "epsilon", "norm_obs", "norm_reward"):
            setattr(self, k, sd[k])
        self.training = bool(sd.get("training", self.training))        # files written before the flag was saved keep the default

    def save(self, path):
        """Like ``VecNormalize.save`` (train_ppo_v2.py:315-317, save_vecnormalize=True): statistics only, not the env."""
        torch.save(self.state_dict(), path)

    @classmethod
    def load(cls, path, venv):
        """``VecNormalize.load(path, venv)`` as the reference calls it (train_ppo_v2.py:449-455).  Accepts this class's own
        ``save`` files and the pickles SB3's ``VecNormalize.save`` wrote (``read_sb3_vecnormalize``)."""
        with open(path, "rb") as f:
            magic = f.read(2)
        if magic != b"PK":                                    # not a torch.save archive: an SB3 pickle
            st = read_sb3_vecnormalize(path)
            self = cls(venv, training=st["training"], norm_obs=st["norm_obs"], norm_reward=st["norm_reward"], clip_obs=st["clip_obs"],
                       clip_reward=st["clip_reward"], gamma=st["gamma"], epsilon=st["epsilon"])
            rms = torch.zeros(32, dtype=torch.float64)
            rms[0:13], rms[13:26], rms[26] = torch.from_numpy(st["obs_mean"]), torch.from_numpy(st["obs_var"]), st["obs_count"]
            rms[27], rms[28], rms[29] = st["ret_mean"], st["ret_var"], st["ret_count"]
            self._rms[:32].copy_(rms)
            return self
        sd = torch.load(path, weights_only=False)
        self = cls(venv, training=sd.get("training", True), norm_obs=sd["norm_obs"], norm_reward=sd["norm_reward"],
                   clip_obs=sd["clip_obs"], clip_reward=sd["clip_reward"], gamma=sd["gamma"], epsilon=sd["epsilon"])
        self.load_state_dict(sd)
        return self

    def export_stats(self):
        """NumPy dict in the shape the reference ships (quantconnect/extract_model.py:62-79)."""
        return {"obs_mean": self.obs_rms.mean.cpu().numpy().astype(np.float32), "obs_var": self.obs_rms.var.cpu().numpy().astype(np.float32),
                "ret_mean": float(self.ret_rms.mean), "ret_var": float(self.ret_rms.var), "clip_obs": self.clip_obs,
                "clip_reward": self.clip_reward, "gamma": self.gamma, "epsilon": self.epsilon}
