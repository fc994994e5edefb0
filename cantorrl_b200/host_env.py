"""NumPy-in / NumPy-out vectorised env over the host-buffer C ABI (``cantor_vecenv_*``), no torch involved.

This is the class a maintainer of the reference would put where ``SubprocVecEnv([create_env_fn(...)] * N)`` is
built today (``src/agents/train_ppo_v2.py:127-141``): same constructor keywords as ``HedgingEnv``
(``src/env/hedging_env_v2.py:10-22``), ``reset() -> obs[N, 13]``, ``step(actions[N, 2]) -> obs, rewards, dones,
infos`` with SB3's auto-reset convention.  All arithmetic runs in the CUDA library; if it cannot be loaded or
there is no GPU the constructor raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_KEYS = ("paths", "volatilities", "call_prices_atm", "put_prices_atm")


class _EmptyInfos:
    """``infos`` of a step as SB3 indexes it (``infos[i].get(...)``, ``len(infos)``) without building N dicts per step."""

    def __init__(self, n):
        self._n = n

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [{} for _ in range(*i.indices(self._n))]
        if not -self._n <= int(i) < self._n:
            raise IndexError(i)
        return {}

    def __iter__(self):
        return ({} for _ in range(self._n))


class HostVecEnv:
    def __init__(self, data_file_path=None, transaction_cost_per_contract=0.65, lambda_cost=1.0, pnl_penalty_weight=0.01,
                 theta_weight=0.0, slippage_bps=0.0, loss_type="abs", initial_cash=0.0, shares_to_hedge=10000,
                 max_contracts_held_per_type=200, max_trade_per_step=15, profile_print_interval=0, record_metrics=True,
                 *, num_envs=1, data=None, simulate=None, device=0, precision="fp32", episode_sampler="same_path",
                 seed=0, env_offset=0, n_chunks=0, pin=True):
        L = _lib.lib()
        self.num_envs = int(num_envs)
        self.precision = precision
        self._prec = _lib.F64 if precision == "fp64" else _lib.F32
        self.loss_type = loss_type
        self.max_contracts_held = max_contracts_held_per_type
        self.max_trade_per_step = max_trade_per_step
        self.shares_held_fixed = shares_to_hedge
        self.option_contract_multiplier = 100
        params = _lib.EnvParams(float(transaction_cost_per_contract), float(lambda_cost), float(pnl_penalty_weight),
                                float(theta_weight), float(slippage_bps), float(initial_cash), 0.04, 30 / 252,
                                _lib.LOSS_MSE if loss_type == "mse" else _lib.LOSS_ABS, int(shares_to_hedge),
                                int(max_contracts_held_per_type), int(max_trade_per_step), 100, int(bool(record_metrics)))
        self._h = C.c_void_p()
        _lib.check(L.cantor_vecenv_create(C.byref(self._h), C.byref(params), self._prec, self.num_envs, int(device),
                                          int(n_chunks)), "cantor_vecenv_create")
        try:
            if simulate is not None:                       # dict(num_paths=, n_steps=, model=, seed=, ...)
                sp = _lib.SimParams(_lib.MODEL_HESTON if simulate.get("model", "gbm") == "heston" else _lib.MODEL_GBM, 1,
                                    simulate.get("s0", 100.0), simulate.get("v0", 0.04), 0.04, 1 / 252,
                                    simulate.get("kappa", 2.0), simulate.get("theta", 0.04), simulate.get("sigma_v", 0.5),
                                    simulate.get("rho", -0.7), 30 / 252, int(simulate.get("seed", 42)),
                                    int(simulate.get("path_offset", env_offset)))
                _lib.check(L.cantor_vecenv_simulate_book(self._h, C.byref(sp), int(simulate["num_paths"]),
                                                         int(simulate.get("n_steps", 252))), "cantor_vecenv_simulate_book")
            else:
                if data is None:
                    try:
                        with np.load(data_file_path) as z:
                            data = {k: z[k] for k in _KEYS}
                    except Exception as e:                 # hedging_env_v2.py:42-43
                        raise FileNotFoundError(f"Could not load or parse data from {data_file_path}. Error: {e}")
                dt = np.float64 if any(np.asarray(data[k]).dtype == np.float64 for k in _KEYS) else np.float32
                arrs = [np.ascontiguousarray(np.asarray(data[k], dtype=dt)) for k in _KEYS]
                S, V, Cc, Pp = arrs
                if not (S.ndim == 2 and S.shape == V.shape and Cc.ndim == 2 and Pp.ndim == 2
                        and S.shape[0] == Cc.shape[0] == Pp.shape[0] and S.shape[1] == Cc.shape[1] + 1 == Pp.shape[1] + 1):
                    raise ValueError("Data shapes are inconsistent.")                      # :45-48
                _lib.check(L.cantor_vecenv_load_book_host(self._h, *(a.ctypes.data for a in arrs),
                                                          _lib.F64 if dt == np.float64 else _lib.F32, S.shape[0],
                                                          S.shape[1] - 1), "cantor_vecenv_load_book_host")
        except Exception:
            L.cantor_vecenv_destroy(self._h)
            self._h = None
            raise
        self.episode_length = L.cantor_vecenv_episode_length(self._h)
        self.num_episodes = L.cantor_vecenv_num_paths(self._h)
        mode = {"same_path": _lib.RESET_SAME_PATH, "philox": _lib.RESET_PHILOX, "array": _lib.RESET_FROM_ARRAY}[episode_sampler]
        self._reset_mode, self._env_offset = mode, int(env_offset)
        if mode != _lib.RESET_FROM_ARRAY:
            _lib.check(L.cantor_vecenv_set_reset_rule(self._h, mode, int(seed) & (2 ** 64 - 1), int(env_offset)))
        n = self.num_envs
        self._obs = np.empty((n, _lib.OBS_DIM), np.float32)
        self._reward = np.empty(n, np.float64 if precision == "fp64" else np.float32)
        self._done = np.empty(n, np.uint8)
        self._pinned = []
        if pin:
            for a in (self._obs, self._reward, self._done):
                self.pin(a)

    # -- pinned host memory --------------------------------------------------------------------------------
    def pin(self, array: np.ndarray) -> np.ndarray:
        """Page-lock a caller-owned NumPy array (e.g. the action buffer) for full-speed asynchronous copies."""
        _lib.check(_lib.lib().cantor_host_register(array.ctypes.data, array.nbytes), "cantor_host_register")
        self._pinned.append(array)
        return array

    # -- gym / VecEnv surface -------------------------------------------------------------------------------
    def reset(self, path_idx=None):
        p = None if path_idx is None else np.ascontiguousarray(path_idx, np.int32)
        _lib.check(_lib.lib().cantor_vecenv_reset_host(self._h, None if p is None else p.ctypes.data, self._obs.ctypes.data),
                   "cantor_vecenv_reset_host")
        return self._obs

    def step(self, actions, next_path=None):
        a = np.ascontiguousarray(actions, np.float32)
        if a.shape != (self.num_envs, 2):
            raise ValueError(f"actions must have shape ({self.num_envs}, 2)")
        nxt = None if next_path is None else np.ascontiguousarray(next_path, np.int32)
        _lib.check(_lib.lib().cantor_vecenv_step_host(self._h, a.ctypes.data, self._obs.ctypes.data, self._reward.ctypes.data,
                                                      self._done.ctypes.data, None if nxt is None else nxt.ctypes.data),
                   "cantor_vecenv_step_host")
        return self._obs, self._reward, self._done.view(np.bool_), _EmptyInfos(self.num_envs)

    # -- the rest of SB3's VecEnv protocol (stable_baselines3.common.vec_env.base_vec_env.VecEnv) ---------------
    @property
    def observation_space(self):
        from .env import OBS_HIGH, OBS_LOW, Box                # hedging_env_v2.py:62-68 (needs torch, like every device class)
        return Box(OBS_LOW, OBS_HIGH, (13,), np.float32)

    @property
    def action_space(self):
        from .env import Box
        return Box(-1.0, 1.0, (2,), np.float32)

    def step_async(self, actions):
        self._pending_actions = actions

    def step_wait(self):
        return self.step(self._pending_actions)

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        raise AttributeError(f"{attr_name}: the envs of a HostVecEnv share one configuration; construct a new one")

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        raise AttributeError(f"{method_name}: there are no per-env Python objects behind a HostVecEnv")

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def seed(self, seed=None):
        if seed is not None:
            _lib.check(_lib.lib().cantor_vecenv_set_reset_rule(self._h, self._reset_mode, int(seed) & (2 ** 64 - 1), self._env_offset))
        return [seed] * self.num_envs

    def render(self, mode=None):
        return None

    def close(self):
        if getattr(self, "_h", None):
            L = _lib.lib()
            for a in self._pinned:
                L.cantor_host_unregister(a.ctypes.data)
            self._pinned = []
            L.cantor_vecenv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_copy_probe(device=0, d2h_bytes=57 << 20, h2d_bytes=8 << 20, n_chunks=8, sync_each_round=True, seconds=0.5):
    """Raw copy ceiling of the host-buffer face (``cantor_host_copy_probe``): the page-locked H2D / D2H traffic of one
    ``HostVecEnv.step`` with no kernel in between.  Returns ``dict(d2h_gbs, h2d_gbs, rounds_per_s)``."""
    d2h, h2d, rps = C.c_double(), C.c_double(), C.c_double()
    _lib.check(_lib.lib().cantor_host_copy_probe(int(device), int(d2h_bytes), int(h2d_bytes), int(n_chunks),
                                                 int(bool(sync_each_round)), float(seconds), C.byref(d2h), C.byref(h2d),
                                                 C.byref(rps)), "cantor_host_copy_probe")
    return dict(d2h_gbs=d2h.value, h2d_gbs=h2d.value, rounds_per_s=rps.value)
