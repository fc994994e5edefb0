"""Vectorised, GPU-resident replacement for the reference's ``HedgingEnv``.

``HedgingVecEnv`` keeps the reference constructor keywords and defaults
(``src/env/hedging_env_v2.py:10-22``; ``version="v1"`` gives ``src/env/hedging_env.py:10-20``) and the
gym ``reset/step`` contract, batched over ``num_envs`` environments with the auto-reset convention of
Stable-Baselines3's ``VecEnv``.  Every call forwards to one hand-written sm_100a kernel through the C ABI
in ``include/cantor_hedge.h``; torch tensors are only the buffers.  There is no CPU path.

``HedgingEnv`` (bottom of the file) is the single-environment, NumPy-facing adapter with exactly the
reference's signatures and exceptions, for callers such as ``src/agents/baselines.py:132-141``.
"""
from __future__ import annotations

import ctypes as C
import inspect
from typing import Optional

import numpy as np
import torch

from . import _lib
from .data import ReplayData

OBS_LOW = np.array([0.1, -1.0, -1.0, -1.0, -1.0, 0.0, 0.0, -1.0, 0.0, -1.0, 0.0, -1.0, -1.0], np.float32)   # :62-64
OBS_HIGH = np.array([10.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 50.0, 1.0, 50.0, 1.0, 1.0], np.float32)       # :65-67


class Box:
    """The two attributes-and-``sample`` subset of ``gymnasium.spaces.Box`` that the reference's callers use."""

    def __init__(self, low, high, shape, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def _ref_signature_v2(data_file_path=None, transaction_cost_per_contract=0.65, lambda_cost=1.0, pnl_penalty_weight=0.01,
                      theta_weight=0.0, slippage_bps=0.0, loss_type="abs", initial_cash=0.0, shares_to_hedge=10000,
                      max_contracts_held_per_type=200, max_trade_per_step=15, profile_print_interval=0, record_metrics=True):
    """Positional order, names and defaults of ``HedgingEnv.__init__`` in src/env/hedging_env_v2.py:10-22."""
    return locals()


def _ref_signature_v1(data_file_path=None, transaction_cost_per_contract=0.05, lambda_cost=1.0, pnl_penalty_weight=0.01,
                      loss_type="abs", initial_cash=0.0, shares_to_hedge=10000, max_contracts_held_per_type=200,
                      max_trade_per_step=15, profile_print_interval=0, record_metrics=True):
    """The same for src/env/hedging_env.py:10-20 (v1): no theta_weight / slippage_bps, so ``loss_type`` is the 5th
    positional argument -- ``HedgingEnv(DATA_FILE, 0.05, 1.0, 0.0, 10000, 200)`` (src/agents/test_rand_ppo.py:26-27)
    binds ``loss_type=10000, initial_cash=200`` there, exactly as it does here."""
    return dict(locals(), theta_weight=0.0, slippage_bps=0.0)


def bind_reference_arguments(version, args, kwargs):
    """``(positional, keyword)`` arguments of a reference-style constructor call -> dict of the 13 v2 keywords.
    Raises TypeError like Python would for an unknown keyword (e.g. ``theta_weight`` with ``version="v1"``)."""
    fn = _ref_signature_v1 if version == "v1" else _ref_signature_v2
    try:
        return fn(*args, **kwargs)
    except TypeError as e:
        raise TypeError(str(e).replace(fn.__name__, "HedgingEnv.__init__")) from None


def philox_episode_draw(seed, global_env, counter, num_episodes):
    """Host mirror of the step kernel's auto-reset draw (csrc/hedge_step.cu: next_episode_path): first word of
    Philox4x32-10(counter = (env lo, env hi, counter lo, counter hi ^ "RESE"), key = seed) scaled to [0, num_episodes).
    ``global_env`` is an int64 array, ``counter`` one int (two's complement for negative values)."""
    g = np.asarray(global_env, np.uint64)
    cnt = int(counter) & (2 ** 64 - 1)
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    c0, c1 = g & mask, g >> np.uint64(32)
    c2 = np.full_like(c0, cnt & 0xFFFFFFFF)
    c3 = np.full_like(c0, ((cnt >> 32) ^ 0x52455345) & 0xFFFFFFFF)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & mask, p0 & mask
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return ((c0 * np.uint64(num_episodes)) >> np.uint64(32)).astype(np.int32)


class VecInfo:
    """Struct-of-arrays view of the per-step ``info`` dicts (``hedging_env_v2.py:268-293``).

    ``info["per_share_step_pnl"]`` is a device tensor over envs; ``info[i]`` materialises the dict of env ``i``
    (one device->host copy) the way SB3 callers index ``infos[i]``.
    """

    def __init__(self, env: "HedgingVecEnv", done: torch.Tensor, terminal_obs: Optional[torch.Tensor]):
        self._env, self._done, self._terminal_obs = env, done, terminal_obs

    def keys(self):
        ks = ["terminal_observation"] + (["episode"] if self._env.monitor else [])
        if self._env._info_f64 is not None:
            ks += list(_lib.INFO_F64_KEYS) + list(_lib.INFO_I32_KEYS) + ["loss_type_used"]
        return ks

    def __getitem__(self, key):
        env = self._env
        if isinstance(key, str):
            if key == "terminal_observation":
                return self._terminal_obs
            if key == "loss_type_used":
                return env.loss_type
            if key == "episode":          # Monitor: return / length of the episode that just ended, valid where done
                if not env.monitor:
                    raise KeyError("'episode': construct HedgingVecEnv(monitor=True)")
                return {"r": env._ep_return, "l": env._ep_length}
            if env._info_f64 is None:
                raise KeyError(f"{key!r}: construct HedgingVecEnv(record_info=True) to materialise the info dict")
            if key in _lib.INFO_F64_KEYS:
                return env._info_f64[_lib.INFO_F64_KEYS.index(key)]
            if key in _lib.INFO_I32_KEYS:
                return env._info_i32[_lib.INFO_I32_KEYS.index(key)]
            raise KeyError(key)
        i = int(key)
        out = {}
        if env._info_f64 is not None:
            f = env._info_f64[:, i].cpu().numpy()
            q = env._info_i32[:, i].cpu().numpy()
            out.update({k: np.float64(f[j]) for j, k in enumerate(_lib.INFO_F64_KEYS)})
            out.update({k: np.int64(q[j]) for j, k in enumerate(_lib.INFO_I32_KEYS)})
            for k in ("raw_action_call", "raw_action_put", "scaled_float_call", "scaled_float_put", "initial_S0_for_episode"):
                out[k] = np.float32(out[k])
            out["loss_type_used"] = env.loss_type
        if self._terminal_obs is not None and bool(self._done[i]):
            out["terminal_observation"] = self._terminal_obs[i].cpu().numpy()
        if env.monitor and bool(self._done[i]):
            out["episode"] = {"r": float(env._ep_return[i]), "l": int(env._ep_length[i]), "t": 0.0}
        return out

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __len__(self):
        return self._env.num_envs


class HedgingVecEnv:
    """``num_envs`` copies of the reference ``HedgingEnv`` stepped by one CUDA kernel per ``step``.

    Reference keywords keep their names, order and defaults -- positionally too, in the order of the chosen ``version``
    (``bind_reference_arguments``).  Additions (keyword-only):

    num_envs         number of environments (SB3 ``VecEnv.num_envs``)
    data             a ``ReplayData`` or a dict with the npz keys, instead of ``data_file_path``
    simulate         no data at all: ``dict(model="gbm" | "heston", seed=, n_steps=, s0=, v0=, kappa=, theta=, sigma_v=, rho=, r=, dt=,
                     tenor=)`` -- the ON-THE-FLY mode (``cantor_env_step_sim``): every env carries {S, v} and the step kernel generates
                     the day's move and the ATM marks itself, bit-identical to replaying the book ``sim.generate_paths_and_options``
                     writes for the same parameters.  Episode e of global env g runs global path ``e * total_envs + g``.
    total_envs       global env population of a sharded on-the-fly run (default ``env_offset + num_envs``)
    precision        "fp32" (float cash/reward, 137 B per env-step) or "fp64" (the reference's exact
                     float32/float64 ledger: integers bit-exact, floats <= 1e-6 relative; 157 B)
    version          "v2" (default) or "v1" (commission default 0.05; slippage/theta must stay 0)
    episode_sampler  how a finished env picks its next path (reference: ``np_random.integers``, :150)
                     "pcg64"     env i owns ``Generator(PCG64(SeedSequence(seed + i)))`` like N seeded reference envs
                     "philox"    counter-based draw on the device (no host work; default for num_envs > 4096)
                     "same_path" env i replays path ``(env_offset + i) % n_paths`` forever (sharded-by-path runs)
    record_info      materialise the numeric ``info`` keys each step (costs 160 extra bytes per env-step)
    auto_reset       VecEnv convention (default).  False keeps finished envs at the terminal state.
    env_offset       global index of env 0 on this rank (multi-GPU sharding by path index)
    monitor          SB3 ``Monitor`` fused into the step kernel: per-env episode return / length, reported as
                     ``infos["episode"]`` (``{"r", "l"}`` tensors, valid where done) and ``infos[i]["episode"]``
    stats            an ``EpisodeStats``: finished episodes are reduced into it inside the kernel (needs ``monitor=True``)
    """

    metadata = {"render_modes": [], "render_fps": 1}

    def __init__(self, *args, num_envs=1, data=None, simulate=None, device="cuda", precision="fp32", version="v2",
                 episode_sampler=None, seed=None, record_info=False, auto_reset=True, env_offset=0, total_envs=None,
                 monitor=False, stats=None, **kwargs):
        if version not in ("v1", "v2"):
            raise ValueError("version must be 'v1' or 'v2'")
        # reference arguments, positional or keyword, in the order of the chosen version (hedging_env_v2.py:10-22 / hedging_env.py:10-20)
        ref = bind_reference_arguments(version, args, kwargs)
        data_file_path = ref["data_file_path"]
        transaction_cost_per_contract = ref["transaction_cost_per_contract"]
        lambda_cost, pnl_penalty_weight = ref["lambda_cost"], ref["pnl_penalty_weight"]
        theta_weight, slippage_bps, loss_type = ref["theta_weight"], ref["slippage_bps"], ref["loss_type"]
        initial_cash, shares_to_hedge = ref["initial_cash"], ref["shares_to_hedge"]
        max_contracts_held_per_type, max_trade_per_step = ref["max_contracts_held_per_type"], ref["max_trade_per_step"]
        record_metrics = ref["record_metrics"]
        if precision not in ("fp32", "fp64"):
            raise ValueError("precision must be 'fp32' or 'fp64'")
        _lib.lib()                                   # fail loudly before anything else if the .so is missing
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CantorError("HedgingVecEnv runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:                # "cuda" -> "cuda:<current>": tensors report an indexed device
            self.device = torch.device("cuda", torch.cuda.current_device())

        # -- data (hedging_env_v2.py:36-51), or no data at all: the on-the-fly mode generates each day inside the step kernel ------
        self.simulate = None
        if simulate is not None:
            if data is not None or data_file_path is not None:
                raise ValueError("simulate= (on-the-fly paths) excludes data / data_file_path")
            sm = dict(simulate)
            model = sm.pop("model", "gbm")
            if model not in ("gbm", "heston"):
                raise ValueError("simulate['model'] must be 'gbm' or 'heston'")
            self.episode_length = int(sm.pop("n_steps", 252))
            self._sim = _lib.SimParams(
                _lib.MODEL_GBM if model == "gbm" else _lib.MODEL_HESTON, 1, float(sm.pop("s0", 100.0)), float(sm.pop("v0", 0.04)),
                float(sm.pop("r", 0.04)), float(sm.pop("dt", 1 / 252)), float(sm.pop("kappa", 2.0)), float(sm.pop("theta", 0.04)),
                float(sm.pop("sigma_v", 0.5)), float(sm.pop("rho", -0.7)), float(sm.pop("tenor", 30 / 252)),
                int(sm.pop("seed", 42)) & (2 ** 64 - 1), int(env_offset))
            if sm:
                raise TypeError(f"unknown simulate keys: {sorted(sm)}")
            self.simulate = dict(simulate, model=model)
            self.data = None
            self.num_episodes = 2 ** 31 - 1          # episodes are numbered, not stored: the stream never repeats
        elif isinstance(data, ReplayData):
            self.data = data
        elif isinstance(data, dict):
            try:
                arrs = [data[k] for k in ("paths", "volatilities", "call_prices_atm", "put_prices_atm")]
            except Exception as e:
                raise FileNotFoundError(f"Could not load or parse data from {data!r}. Error: {e}")
            self.data = ReplayData.from_arrays(*arrs, device=self.device)
        elif data_file_path is not None:
            self.data = ReplayData.from_npz(data_file_path, device=self.device)
        else:
            raise FileNotFoundError("Could not load or parse data from None. Error: no data_file_path / data given")
        if self.data is not None:
            self.num_episodes = self.data.n_paths
            self.episode_length = self.data.episode_length

        # -- reference attributes (:26-33, :53-58) --------------------------------------------------------
        self.pnl_penalty_weight = pnl_penalty_weight
        self.lambda_cost = lambda_cost
        self.theta_weight = theta_weight
        self.slippage_bps = slippage_bps
        self.loss_type = loss_type
        self.record_metrics = record_metrics
        self.initial_cash = initial_cash
        self._max_trade_per_step_internal = max_trade_per_step
        self.max_trade_per_step = max_trade_per_step          # read by baselines.py:100 / delta_and_nothing.py:79
        self.transaction_cost_per_contract = transaction_cost_per_contract
        self.max_contracts_held = max_contracts_held_per_type
        self.shares_held_fixed = shares_to_hedge
        self.option_contract_multiplier = 100
        self.risk_free_rate = 0.04
        self.option_tenor_years = 30 / 252
        self.action_space = Box(-1.0, 1.0, (2,), np.float32)
        self.observation_space = Box(OBS_LOW, OBS_HIGH, (13,), np.float32)

        self.num_envs = int(num_envs)
        self.precision = precision
        self.version = version
        self.auto_reset = bool(auto_reset)
        self.env_offset = int(env_offset)
        self._prec = _lib.F64 if precision == "fp64" else _lib.F32
        self.total_envs = int(total_envs) if total_envs is not None else self.env_offset + self.num_envs
        if self.total_envs < self.env_offset + self.num_envs:
            raise ValueError("total_envs < env_offset + num_envs")
        if self.simulate is not None:
            if episode_sampler not in (None, "sequential"):
                raise ValueError("on-the-fly mode numbers its episodes: episode_sampler must be None or 'sequential'")
            episode_sampler = "sequential"           # episode e of global env g = global path e * total_envs + g
        elif episode_sampler is None:
            episode_sampler = "pcg64" if self.num_envs <= 4096 else "philox"
        if episode_sampler not in ("pcg64", "philox", "same_path", "sequential") or (episode_sampler == "sequential" and self.simulate is None):
            raise ValueError("episode_sampler must be 'pcg64', 'philox' or 'same_path'")
        self.episode_sampler = episode_sampler

        self._params = _lib.EnvParams(
            float(transaction_cost_per_contract), float(lambda_cost), float(pnl_penalty_weight), float(theta_weight),
            float(slippage_bps), float(initial_cash), self.risk_free_rate, self.option_tenor_years,
            _lib.LOSS_MSE if loss_type == "mse" else _lib.LOSS_ABS,            # :246-253
            int(shares_to_hedge), int(max_contracts_held_per_type), int(max_trade_per_step),
            self.option_contract_multiplier, int(bool(record_metrics)))
        self._book = self.data.book() if self.data is not None else None

        # -- caller-owned buffers ---------------------------------------------------------------------------
        n, dev = self.num_envs, self.device
        ftype = torch.float64 if precision == "fp64" else torch.float32
        self._core = torch.zeros((n, 4), dtype=torch.int32, device=dev)
        self._cash = torch.zeros(n, dtype=ftype, device=dev)
        self._pv_prev = torch.zeros(n, dtype=torch.float64, device=dev) if precision == "fp64" else None
        self._state = _lib.EnvState(self._core.data_ptr(), self._cash.data_ptr(), _lib.ptr(self._pv_prev))
        self._sv = self._source = None
        if self.simulate is not None:                # carried {S, v} of every env: the whole "book" of the on-the-fly mode
            self._sv = torch.zeros((n, 2), dtype=torch.float32, device=dev)
            self._source = _lib.EnvSim(C.pointer(self._sim), self._sv.data_ptr(), self.total_envs, self.episode_length, 0)
        self.monitor = bool(monitor)
        self.stats = stats
        self._ep_acc = self._ep_return = self._ep_length = self._stats_c = None
        if stats is not None and not monitor:
            raise ValueError("stats needs monitor=True")
        if monitor:
            self._ep_acc = torch.zeros((n, 4), dtype=ftype, device=dev)
            self._ep_return = torch.zeros(n, dtype=ftype, device=dev)
            self._ep_length = torch.zeros(n, dtype=torch.int32, device=dev)
            self._state.episode_acc = self._ep_acc.data_ptr()
            self._state.episode_return = self._ep_return.data_ptr()
            self._state.episode_length = self._ep_length.data_ptr()
            if stats is not None:
                # the struct is owned by the EpisodeStats object and rewritten in place when its transport changes
                # (enable_fused_all_reduce after this constructor), so the pointer stays valid and current
                self._stats_c = stats.c_struct()
                self._state.stats = C.cast(C.pointer(self._stats_c), C.c_void_p)
        self._obs = torch.zeros((n, _lib.OBS_DIM), dtype=torch.float32, device=dev)
        self._reward = torch.zeros(n, dtype=ftype, device=dev)
        self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._done_bool = self._done.view(torch.bool)
        self._terminal_obs = torch.zeros((n, _lib.OBS_DIM), dtype=torch.float32, device=dev)
        self._next_path = torch.zeros(n, dtype=torch.int32, device=dev)
        self._info_f64 = self._info_i32 = None
        self._info = None
        if record_info:
            # float keys in the ledger's own precision: float32 arrays in fp32 mode (92 instead of 160 info bytes per env-step)
            self._info_f64 = torch.zeros((len(_lib.INFO_F64_KEYS), n), dtype=ftype, device=dev)
            self._info_i32 = torch.zeros((len(_lib.INFO_I32_KEYS), n), dtype=torch.int32, device=dev)
            if precision == "fp64":
                self._info = _lib.InfoOut(self._info_f64.data_ptr(), self._info_i32.data_ptr(), None)
            else:
                self._info = _lib.InfoOut(None, self._info_i32.data_ptr(), self._info_f64.data_ptr())
        self._rule = _lib.ResetRule()
        self._call_cache = None                      # ctypes byref objects / pointers that never change between steps
        self._rule.mode = {"pcg64": _lib.RESET_FROM_ARRAY, "philox": _lib.RESET_PHILOX,
                           "same_path": _lib.RESET_SAME_PATH, "sequential": _lib.RESET_SAME_PATH}[episode_sampler]
        self._rule.next_path = self._next_path.data_ptr()
        self._rule.env_offset = self.env_offset
        self._seed = seed
        self._rngs = None                # pcg64 sampler: one generator per env
        self._host_steps = None          # pcg64 sampler: host mirror of current_step (no device sync needed)
        self._global_step = 0
        self._n_resets = 0               # philox sampler: every reset() draws a fresh set of first episodes
        self._pending_actions = None
        self._was_reset = False

    # ------------------------------------------------------------------------------------------ sampling
    def _seed_generators(self, seed):
        if self.episode_sampler == "pcg64":
            if seed is None:
                ss = np.random.SeedSequence().spawn(self.num_envs)
                self._rngs = [np.random.Generator(np.random.PCG64(s)) for s in ss]
            else:   # env i is seeded seed + i, as create_vec_env does (train_ppo_v2.py:129)
                self._rngs = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(seed) + i)))
                              for i in range(self.num_envs)]
        self._rule.seed = int(seed) & (2 ** 64 - 1) if seed is not None else 0x5EED5EED

    def _draw_paths(self, which=None) -> np.ndarray:
        """One ``np_random.integers(num_episodes)`` per selected env (hedging_env_v2.py:150)."""
        n = self.num_envs
        if self.episode_sampler == "pcg64":
            idx = range(n) if which is None else which
            return np.array([self._rngs[i].integers(self.num_episodes) for i in idx], dtype=np.int32)
        if self.episode_sampler == "same_path":
            ids = np.arange(n, dtype=np.int64) if which is None else np.asarray(which, np.int64)
            return ((self.env_offset + ids) % self.num_episodes).astype(np.int32)
        if self.episode_sampler == "sequential":         # on-the-fly mode: every env starts with episode number 0
            return np.zeros(n if which is None else len(which), np.int32)
        # philox: the device's own counter scheme, keyed by (seed; GLOBAL env index, reset number) -- so two resets, and two
        # ranks of a sharded run, draw different first episodes, and reset(seed=s) reproduces them
        ids = np.arange(n, dtype=np.int64) if which is None else np.asarray(which, np.int64)
        return philox_episode_draw(self._rule.seed, self.env_offset + ids, -(1 + self._n_resets), self.num_episodes)

    # --------------------------------------------------------------------------------------------- reset
    def seed(self, seed=None):
        self._seed = seed
        self._seed_generators(seed)

    def reset(self, seed=None, options=None, *, path_idx=None):
        """Reset every env; returns the observation tensor ``[num_envs, 13]`` (float32, on the device).

        ``path_idx`` (array of ``num_envs`` ints) overrides the sampler, e.g. with indices exported from the reference.
        """
        if seed is not None or not self._was_reset:
            self._seed_generators(seed if seed is not None else self._seed)
            self._global_step = 0                  # re-seeding restarts the device-side counters too: reset(seed=s) is reproducible
            self._n_resets = 0
        idx = np.asarray(path_idx, np.int32) if path_idx is not None else self._draw_paths()
        self._n_resets += 1
        if idx.shape != (self.num_envs,) or idx.min() < 0 or idx.max() >= self.num_episodes:
            raise IndexError("path_idx out of range")
        idx_dev = torch.from_numpy(idx).to(self.device)
        with torch.cuda.device(self.device):
            if self._source is not None:
                _lib.check(_lib.lib().cantor_env_reset_sim(
                    C.byref(self._params), C.byref(self._source), C.byref(self._state), self.num_envs, self._prec,
                    None, idx_dev.data_ptr(), self._obs.data_ptr(), _lib.current_stream_ptr(self.device)), "cantor_env_reset_sim")
            else:
                _lib.check(_lib.lib().cantor_env_reset(
                    C.byref(self._params), C.byref(self._book), C.byref(self._state), self.num_envs, self._prec,
                    None, idx_dev.data_ptr(), self._obs.data_ptr(), _lib.current_stream_ptr(self.device)), "cantor_env_reset")
        if self.episode_sampler == "pcg64":
            self._host_steps = np.zeros(self.num_envs, np.int64)
            self._next_path.copy_(torch.from_numpy(self._draw_paths()))
        self._was_reset = True
        return self._obs

    def set_next_paths(self, next_path):
        """Supply the episode index each env will take at its next auto-reset (overrides the sampler once)."""
        self._next_path.copy_(torch.as_tensor(np.asarray(next_path, np.int32)))

    # ---------------------------------------------------------------------------------------------- step
    def step(self, actions, *, obs_out=None, reward_out=None, done_out=None):
        """One env-step for all envs.  ``actions`` is ``[num_envs, 2]`` float32 (device tensor, or host array).

        Returns ``(obs, rewards, dones, infos)`` like SB3 ``VecEnv.step``; the tensors are reused every call
        unless ``*_out`` buffers are given (rollout storage can be written in place).
        """
        if not self._was_reset:
            raise RuntimeError("reset() must be called before step()")
        # the call below is ~20 us of CPU time against a 17 us kernel at 2^20 envs: keep the host side of a step minimal
        if (isinstance(actions, torch.Tensor) and actions.dtype is torch.float32 and actions.device == self.device
                and actions.is_contiguous() and actions.numel() == 2 * self.num_envs):
            a = actions
        else:
            if not isinstance(actions, torch.Tensor):
                actions = torch.as_tensor(np.asarray(actions, np.float32))
            a = actions.to(device=self.device, dtype=torch.float32).reshape(self.num_envs, 2).contiguous()
        obs = self._obs if obs_out is None else obs_out
        reward = self._reward if reward_out is None else reward_out
        done = self._done if done_out is None else done_out
        self._rule.episode_counter = self._global_step
        c = self._call_cache
        if c is None:
            c = self._call_cache = dict(
                fn=_lib.lib().cantor_env_step if self._source is None else _lib.lib().cantor_env_step_sim,
                params=C.byref(self._params), book=C.byref(self._book) if self._book is not None else None,
                source=C.byref(self._source) if self._source is not None else None, state=C.byref(self._state),
                rule=C.byref(self._rule), info=C.byref(self._info) if self._info is not None else None,
                term=self._terminal_obs.data_ptr(), dev=self.device.index if self.device.index is not None else torch.cuda.current_device())
        if self._source is None:
            args = (c["params"], c["book"], c["state"], self.num_envs, self._prec, a.data_ptr(), obs.data_ptr(), reward.data_ptr(),
                    done.data_ptr(), c["term"], int(self.auto_reset), c["rule"], c["info"])
        else:       # on-the-fly mode: the day's path step is generated inside the kernel
            args = (c["params"], c["source"], c["state"], self.num_envs, self._prec, a.data_ptr(), obs.data_ptr(), reward.data_ptr(),
                    done.data_ptr(), c["term"], int(self.auto_reset), c["info"],
                    int(self._rule.flags) | (_lib.STEP_WALK_BACKWARD if (self._global_step & 1) else 0))     # alternate the walk: L2 reuse
        if torch.cuda.current_device() == c["dev"]:
            status = c["fn"](*args, torch.cuda.current_stream().cuda_stream)
        else:
            with torch.cuda.device(self.device):
                status = c["fn"](*args, _lib.current_stream_ptr(self.device))
        if status != 0:
            _lib.check(status, "cantor_env_step")
        self._global_step += 1
        if self.episode_sampler == "pcg64" and self.auto_reset:
            self._host_steps += 1
            fin = np.nonzero(self._host_steps >= self.episode_length)[0]
            if fin.size:      # those envs just consumed next_path: draw the one after, still without a device sync
                self._host_steps[fin] = 0
                nxt = self._next_path.cpu().numpy()
                nxt[fin] = self._draw_paths(fin)
                self._next_path.copy_(torch.from_numpy(nxt))
        return obs, reward, (self._done_bool if done is self._done else done.view(torch.bool)), VecInfo(self, done, self._terminal_obs)

    def step_many(self, actions, obs_out=None, reward_out=None, done_out=None):
        """``n_steps`` consecutive env-steps on a known action tape ``actions [n_steps, num_envs, 2]`` (float32, on the device)
        in ONE persistent kernel launch (``cantor_env_step_many``): identical results to ``n_steps`` calls of ``step``, but the
        env state stays in registers between steps (81 instead of 137 bytes of memory traffic per env-step).  Returns
        ``(obs [n_steps, n, 13], rewards [n_steps, n], dones [n_steps, n])`` rollout slabs (written in place when given)."""
        if not self._was_reset:
            raise RuntimeError("reset() must be called before step_many()")
        if self._source is not None or not self.auto_reset or self._info is not None:
            raise NotImplementedError("step_many: replay mode with auto_reset and without record_info")
        if self.episode_sampler == "pcg64":
            raise NotImplementedError("step_many draws episodes on the device: use episode_sampler 'philox' or 'same_path'")
        a = actions
        if not (isinstance(a, torch.Tensor) and a.dtype is torch.float32 and a.device == self.device and a.is_contiguous()
                and a.dim() == 3 and a.shape[1:] == (self.num_envs, 2)):
            raise ValueError("actions must be a contiguous float32 device tensor [n_steps, num_envs, 2]")
        k, n, dev = a.shape[0], self.num_envs, self.device
        obs = obs_out if obs_out is not None else torch.empty((k, n, _lib.OBS_DIM), dtype=torch.float32, device=dev)
        reward = reward_out if reward_out is not None else torch.empty((k, n), dtype=self._reward.dtype, device=dev)
        done = done_out if done_out is not None else torch.empty((k, n), dtype=torch.uint8, device=dev)
        self._rule.episode_counter = self._global_step
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cantor_env_step_many(
                C.byref(self._params), C.byref(self._book), C.byref(self._state), n, self._prec, k, a.data_ptr(), obs.data_ptr(),
                reward.data_ptr(), done.data_ptr(), self._terminal_obs.data_ptr(), C.byref(self._rule),
                _lib.current_stream_ptr(dev)), "cantor_env_step_many")
        self._global_step += k
        return obs, reward, done.view(torch.bool)

    def step_async(self, actions):
        self._pending_actions = actions

    def step_wait(self):
        return self.step(self._pending_actions)

    # ----------------------------------------------------------------------------- VecEnv odds and ends
    def get_attr(self, name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, name)] * n

    def set_attr(self, attr_name, value, indices=None):
        raise AttributeError(f"{attr_name}: the envs of a HedgingVecEnv share one configuration; construct a new one")

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        raise AttributeError(f"{method_name}: there are no per-env Python objects behind a HedgingVecEnv")

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def close(self):
        pass

    def render(self, mode=None):
        pass

    # per-env views of the state the reference exposes as attributes (delta_and_nothing.py:71-78)
    @property
    def current_step(self):
        return self._core[:, 1]

    @property
    def current_episode_idx(self):
        return self._core[:, 2]

    @property
    def call_contracts_held(self):
        return (self._core[:, 0] << 16) >> 16

    @property
    def put_contracts_held(self):
        return self._core[:, 0] >> 16

    @property
    def cash_balance(self):
        return self._cash

    @property
    def initial_S0_for_episode(self):
        return self._core[:, 3].view(torch.float32)

    def _gather_current(self, arr: torch.Tensor, stale_at_end: bool):
        step = self._core[:, 1].long()
        if stale_at_end:
            step = torch.minimum(step, torch.full_like(step, self.episode_length - 1))
        return arr[step, self._core[:, 2].long()]

    @property
    def current_stock_price(self):
        if self._sv is not None:
            return self._sv[:, 0]
        return self._gather_current(self.data.S, False)

    @property
    def current_volatility(self):
        if self._sv is not None:
            return torch.clamp(self._sv[:, 1], min=0.0)
        return self._gather_current(self.data.v, False)

    @property
    def current_call_price(self):
        return self._gather_current(self.data.C, True)

    @property
    def current_put_price(self):
        return self._gather_current(self.data.P, True)


class HedgingEnv:
    """Single-environment adapter with the reference's exact gym signatures (NumPy in, NumPy out).

    ``reset(seed=None, options=None) -> (obs float32[13], {})`` and
    ``step(action) -> (obs, reward np.float64, terminated bool, False, info dict)`` as
    ``src/env/hedging_env_v2.py:145,175,294``; stepping a terminated episode raises ``IndexError`` like the
    reference does.  Runs the fp64 ledger by default so scalar callers see reference numerics.
    """

    metadata = HedgingVecEnv.metadata

    def __init__(self, data_file_path=None, *args, precision="fp64", version="v2", device="cuda", data=None, **kwargs):
        self._vec = HedgingVecEnv(data_file_path, *args, num_envs=1, precision=precision, version=version,
                                  device=device, data=data, episode_sampler="pcg64", record_info=True,
                                  auto_reset=False, **kwargs)
        self._np_random = None
        self._terminated = True

    def __getattr__(self, name):            # reference attributes (max_contracts_held, episode_length, ...)
        if name == "_vec":
            raise AttributeError(name)
        v = getattr(self._vec, name)
        if isinstance(v, torch.Tensor) and v.numel() == 1:
            return v.cpu().numpy().reshape(())[()]
        return v

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    @np_random.setter
    def np_random(self, rng):
        self._np_random = rng

    def reset(self, seed=None, options=None):
        if seed is not None:                # hedging_env_v2.py:146-148
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        idx = int(self.np_random.integers(self._vec.num_episodes))        # :150
        obs = self._vec.reset(path_idx=[idx])
        self._terminated = False
        return obs[0].cpu().numpy(), {}

    def step(self, action):
        if self._terminated:
            raise IndexError("index out of bounds: step() on a terminated episode (call reset())")
        obs, reward, done, infos = self._vec.step(np.asarray(action, np.float32).reshape(1, 2))
        self._terminated = bool(done[0])
        info = infos[0]
        info.pop("terminal_observation", None)
        return obs[0].cpu().numpy(), np.float64(reward[0].item()), self._terminated, False, info

    def render(self):
        pass

    def close(self):
        pass
