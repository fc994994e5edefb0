"""Episode-fused rollouts: policy + env step + (optionally) path simulation in one kernel, with episode statistics.

Host-side mirror of the loops the reference runs around its env:

  ``evaluate_baseline_policy``   src/agents/baselines.py:32-72          (no-hedge / delta-every-step policies)
  ``run_benchmark_strategy``     src/benchmark/delta_and_nothing.py:34-114
  ``run_evaluation``             src/agents/train_ppo_v2.py:465-530     (policy -> env.step -> per-episode statistics)
  random policy                  src/agents/test_inf.py:27-39

``HedgingRollout`` takes the reference env's constructor keywords; ``run`` forwards to ``cantor_rollout``
(``include/cantor_hedge.h``), which keeps the whole env state in registers for ``n_steps`` consecutive env-steps,
so nothing is written per step unless rollout storage is asked for.  Envs are identified by a GLOBAL index
(``env_offset + i`` of ``total_envs``) that drives every Philox counter: the statistics of a population do not
depend on how it is sharded over GPUs, and shards combine with ``EpisodeStats.all_reduce``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Union

import numpy as np
import torch

from . import _lib
from .data import ReplayData
from .stats import EpisodeStats

POLICIES = {"no_hedge": _lib.POLICY_NO_HEDGE, "random": _lib.POLICY_RANDOM,
            "delta_every_step": _lib.POLICY_DELTA_BASELINES, "delta_benchmark": _lib.POLICY_DELTA_BENCHMARK}


def pack_mlp(W1, b1, W2, b2, W3, b3, obs_mean=None, obs_var=None, epsilon=1e-8, device="cuda") -> torch.Tensor:
    """Actor weights in ``torch.nn.Linear`` layout (``[out, in]``) -> the flat float32 block ``cantor_policy.mlp`` wants.

    Layout: ``W1[13][64] b1[64] W2[64][64] b2[64] W3[64][2] b3[2] obs_mean[13] obs_inv_std[13]`` (input-major weights).
    ``obs_mean`` / ``obs_var`` are VecNormalize's running statistics; the kernel applies
    ``clip((obs - mean) / sqrt(var + 1e-8), -10, 10)`` (quantconnect/model_wrapper.py:131) before the first layer.
    """
    def t(x, shape):
        x = torch.as_tensor(np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x, np.float32))
        if tuple(x.shape) != shape:
            raise ValueError(f"expected shape {shape}, got {tuple(x.shape)}")
        return x
    W1, b1, W2, b2, W3, b3 = t(W1, (64, 13)), t(b1, (64,)), t(W2, (64, 64)), t(b2, (64,)), t(W3, (2, 64)), t(b3, (2,))
    mean = t(obs_mean, (13,)) if obs_mean is not None else torch.zeros(13)
    var = t(obs_var, (13,)) if obs_var is not None else torch.ones(13) - epsilon
    inv_std = (1.0 / torch.sqrt(var.double() + epsilon)).float()
    flat = torch.cat([W1.T.contiguous().flatten(), b1, W2.T.contiguous().flatten(), b2, W3.T.contiguous().flatten(), b3,
                      mean, inv_std])
    assert flat.numel() == _lib.MLP_FLOATS
    return flat.to(device)


def _bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit patterns (round to nearest even) as uint16."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def _umma_tile(mat: np.ndarray) -> np.ndarray:
    """[rows, K] float32 (rows % 8 == 0, K % 8 == 0) -> the bf16 byte image of the canonical K-major no-swizzle UMMA layout:
    element (r, k) at (r // 8) * (K // 8) * 128 + (k // 8) * 128 + (r % 8) * 16 + (k % 8) * 2 (csrc/mlp_tc.cuh)."""
    rows, K = mat.shape
    bits = _bf16_bits(mat).reshape(rows // 8, 8, K // 8, 8)          # [row group, row in group, k chunk, k in chunk]
    return np.ascontiguousarray(bits.transpose(0, 2, 1, 3)).reshape(-1).view(np.uint8)


def pack_lstm(w_ih, w_hh, b_ih, b_hh, W1, b1, W2, b2, W3, b3, obs_mean=None, obs_var=None, epsilon=1e-8, device="cuda") -> torch.Tensor:
    """Weights of the recurrent actor in ``torch.nn`` layout -> the byte image ``cantor_policy.mlp`` wants for
    ``CANTOR_POLICY_LSTM`` (csrc/lstm_tc.cuh).

    ``w_ih [512, 13]``, ``w_hh [512, 128]``, ``b_ih``, ``b_hh [512]`` as ``torch.nn.LSTM`` stores them (gate order i, f, g, o);
    head ``W1 [64, 128]``, ``W2 [64, 64]``, ``W3 [2, 64]`` + biases (``mlp_extractor.policy_net.*`` / ``action_net.*`` of the
    shipped ``policy_weights.pth``); ``obs_mean`` / ``obs_var``: VecNormalize statistics.  Biases are folded onto a ones column.
    """
    f = lambda x, shape: _check(np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x, np.float32), shape)  # noqa: E731

    def _check(a, shape):
        if a.shape != shape:
            raise ValueError(f"expected shape {shape}, got {a.shape}")
        return a
    w_ih, w_hh = f(w_ih, (512, 13)), f(w_hh, (512, 128))
    bias = f(b_ih, (512,)) + f(b_hh, (512,))
    W1, b1, W2, b2, W3, b3 = f(W1, (64, 128)), f(b1, (64,)), f(W2, (64, 64)), f(b2, (64,)), f(W3, (2, 64)), f(b3, (2,))
    parts = []
    for q in range(8):                                                # half-pass q = hidden units 16 q .. 16 q + 15, rows {i, f, g, o} x 16
        tile = np.zeros((64, 144), np.float32)
        for g in range(4):
            rows = slice(g * 128 + 16 * q, g * 128 + 16 * q + 16)
            tile[g * 16:(g + 1) * 16, 0:13] = w_ih[rows]
            tile[g * 16:(g + 1) * 16, 13] = bias[rows]
            tile[g * 16:(g + 1) * 16, 16:144] = w_hh[rows]
            if g != 2:                                                # sigmoid gates: the kernel evaluates 0.5 + 0.5 tanh(x / 2), and the
                tile[g * 16:(g + 1) * 16] *= 0.5                      # halving is done here (a power of two: exact in bf16)
        parts.append(_umma_tile(tile))
    t1 = np.zeros((64, 144), np.float32)
    t1[:, 13], t1[:, 16:144] = b1, W1
    t2 = np.zeros((64, 80), np.float32)
    t2[:, :64], t2[:, 64] = W2, b2
    t3 = np.zeros((16, 80), np.float32)
    t3[:2, :64], t3[:2, 64] = W3, b3
    parts += [_umma_tile(t1), _umma_tile(t2), _umma_tile(t3)]
    norm = np.zeros(32, np.float32)
    norm[16:29] = 1.0
    if obs_mean is not None:
        norm[:13] = np.asarray(obs_mean, np.float32)
        norm[16:29] = (1.0 / np.sqrt(np.asarray(obs_var, np.float64) + epsilon)).astype(np.float32)
    img = np.concatenate(parts + [norm.view(np.uint8)])
    assert img.size == _lib.LSTM_IMAGE_BYTES
    return torch.from_numpy(img).to(device)


def load_reference_policy(policy_weights_path, normalization_stats_path=None, device="cuda") -> torch.Tensor:
    """The policy files the reference ships -> the ``pack_lstm`` image for ``run(..., "lstm_bf16", mlp=...)``.

    ``policy_weights_path``: ``quantconnect/model_files/policy_weights.pth`` (keys ``lstm_actor.*``, ``mlp_extractor.policy_net.*``,
    ``action_net.*``, mapped like ``ModelWrapper.LoadModel``, quantconnect/model_wrapper.py:84-103); ``normalization_stats_path``:
    ``normalization_stats.pkl`` (``obs_mean`` / ``obs_var``, applied as :131).  The critic tensors are ignored."""
    import pickle
    sd = torch.load(policy_weights_path, map_location="cpu", weights_only=True)
    need = ["lstm_actor.weight_ih_l0", "lstm_actor.weight_hh_l0", "lstm_actor.bias_ih_l0", "lstm_actor.bias_hh_l0",
            "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
            "mlp_extractor.policy_net.2.bias", "action_net.weight", "action_net.bias"]
    missing = [k for k in need if k not in sd]
    if missing:
        raise KeyError(f"{policy_weights_path}: missing {missing}")
    mean = var = None
    if normalization_stats_path is not None:
        with open(normalization_stats_path, "rb") as f:
            st = pickle.load(f)
        mean, var = np.asarray(st["obs_mean"], np.float32), np.asarray(st["obs_var"], np.float32)
    return pack_lstm(*[sd[k] for k in need], obs_mean=mean, obs_var=var, device=device)


@dataclass
class RolloutResult:
    stats: EpisodeStats
    n_steps: int
    obs: Optional[torch.Tensor] = None        # [n_steps, n_envs, 13] observation the policy acted on
    actions: Optional[torch.Tensor] = None    # [n_steps, n_envs, 2]
    reward: Optional[torch.Tensor] = None     # [n_steps, n_envs]
    done: Optional[torch.Tensor] = None       # [n_steps, n_envs] bool


class HedgingRollout:
    """``num_envs`` hedging environments advanced for many steps by one kernel launch.

    Reference keywords as in ``HedgingVecEnv`` (hedging_env_v2.py:10-22).  Data source: ``data`` (a ``ReplayData``
    book or a dict of the npz arrays, replayed) or ``simulate=dict(model="gbm"|"heston", seed=..., s0=..., v0=...,
    kappa=..., theta=..., sigma_v=..., rho=..., n_steps=252)`` (paths and ATM marks generated on the fly from the
    same Philox counters ``sim.generate_paths_and_options`` uses).  Episode ``e`` of global env ``g`` runs on
    global path ``e * total_envs + g`` (modulo the book size when replaying).
    """

    def __init__(self, data_file_path=None, transaction_cost_per_contract=0.65, lambda_cost=1.0, pnl_penalty_weight=0.01,
                 theta_weight=0.0, slippage_bps=0.0, loss_type="abs", initial_cash=0.0, shares_to_hedge=10000,
                 max_contracts_held_per_type=200, max_trade_per_step=15, profile_print_interval=0, record_metrics=True,
                 *, num_envs, data=None, simulate: Optional[dict] = None, device="cuda", env_offset=0, total_envs=None,
                 one_call_only=False):
        _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CantorError("HedgingRollout runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.env_offset = int(env_offset)
        self.total_envs = int(total_envs) if total_envs is not None else self.env_offset + self.num_envs
        self.one_call_only = bool(one_call_only)
        self.loss_type = loss_type
        self._params = _lib.EnvParams(
            float(transaction_cost_per_contract), float(lambda_cost), float(pnl_penalty_weight), float(theta_weight),
            float(slippage_bps), float(initial_cash), 0.04, 30 / 252, _lib.LOSS_MSE if loss_type == "mse" else _lib.LOSS_ABS,
            int(shares_to_hedge), int(max_contracts_held_per_type), int(max_trade_per_step), 100, int(bool(record_metrics)))
        self.data = None
        self._sim = None
        if simulate is not None:
            if data is not None or data_file_path is not None:
                raise ValueError("give either data or simulate, not both")
            kw = dict(model="gbm", seed=42, s0=100.0, v0=0.04, r=0.04, dt=1 / 252, kappa=2.0, theta=0.04, sigma_v=0.5,
                      rho=-0.7, tenor=30 / 252, n_steps=252)
            unknown = set(simulate) - set(kw)
            if unknown:
                raise TypeError(f"unknown simulate keys: {sorted(unknown)}")
            kw.update(simulate)
            if kw["model"] not in ("gbm", "heston"):
                raise ValueError("model must be 'gbm' or 'heston'")
            self.episode_length = int(kw["n_steps"])
            self._sim = _lib.SimParams(_lib.MODEL_GBM if kw["model"] == "gbm" else _lib.MODEL_HESTON, 1, kw["s0"], kw["v0"],
                                       kw["r"], kw["dt"], kw["kappa"], kw["theta"], kw["sigma_v"], kw["rho"], kw["tenor"],
                                       int(kw["seed"]) & (2 ** 64 - 1), 0)
        else:
            if isinstance(data, ReplayData):
                self.data = data
            elif isinstance(data, dict):
                self.data = ReplayData.from_arrays(*(data[k] for k in ("paths", "volatilities", "call_prices_atm",
                                                                        "put_prices_atm")), device=self.device)
            elif data_file_path is not None:
                self.data = ReplayData.from_npz(data_file_path, device=self.device)
            else:
                raise FileNotFoundError("Could not load or parse data from None. Error: no data_file_path / data / simulate given")
            self.episode_length = self.data.episode_length
            self._book = self.data.book()

    def new_stats(self, hist_bins=4096, hist_max=4.0, keep_episodes=0) -> EpisodeStats:
        return EpisodeStats(self.device, hist_bins, hist_max, keep_episodes, self.num_envs)

    def run(self, n_steps: int, policy: Union[str, torch.Tensor] = "delta_every_step", *, mlp: Optional[torch.Tensor] = None,
            actions: Optional[torch.Tensor] = None, seed: int = 0, stats: Optional[EpisodeStats] = None,
            store: bool = False, squash: str = "clip", obs_clip: Optional[float] = None, first_episode: int = 0) -> RolloutResult:
        """``n_steps`` env-steps of every env (auto-reset at episode ends), statistics accumulated into ``stats``.

        policy   "no_hedge" | "random" | "delta_every_step" | "delta_benchmark" | "actions" (open loop, ``actions``
                 float32 ``[n_steps, num_envs, 2]`` on the device) | "mlp" (``mlp=pack_mlp(...)``; float32 FFMA actor,
                 the parity form) | "mlp_bf16" (same weights on the tensor cores: bf16 operands, float32 accumulation)
                 | "lstm_bf16" (``mlp=pack_lstm(...)``: the recurrent LSTM + MLP actor the reference trained, tensor cores)
        store    also write the rollout (obs the policy saw, actions, reward, done), time-major
        squash   network policies: "clip" the action means to [-1, 1] (SB3 stepping the env) or "tanh" (the reference's
                 deployment wrapper, quantconnect/model_wrapper.py:202)
        obs_clip network policies: clip of the normalised observation.  Default: 10 with squash="clip" (SB3 VecNormalize's
                 clip_obs, train_ppo_v2.py:204-208) and none with squash="tanh" (quantconnect/model_wrapper.py:131 does not clip)
        first_episode  episode number every env starts with: the launch plays episodes first_episode, first_episode + 1, ... (global
                 path ``e * total_envs + g``; the random policy's stream continues likewise).  0 replays the same rollout on
                 every call; pass the number of episodes already played (``n_steps // T`` per call) to collect fresh data.
                 Statistics count finished episodes only: a trailing partial episode (``n_steps % T``) runs but is not reported.
        """
        n, dev = self.num_envs, self.device
        pol = _lib.Policy()
        pol.put_leg_disabled = int(self.one_call_only)
        if squash not in ("clip", "tanh"):
            raise ValueError("squash must be 'clip' or 'tanh'")
        pol.action_squash = _lib.SQUASH_TANH if squash == "tanh" else _lib.SQUASH_CLIP
        if obs_clip is None:
            obs_clip = float("inf") if squash == "tanh" else 10.0
        pol.obs_clip = float(obs_clip) if obs_clip > 0 else float("inf")
        if first_episode < 0:
            raise ValueError("first_episode must be >= 0")
        pol.seed = int(seed) & (2 ** 64 - 1)
        if policy in ("mlp", "mlp_bf16"):
            pol.mlp_tensor_cores = int(policy == "mlp_bf16")
            if mlp is None or mlp.dtype != torch.float32 or mlp.numel() != _lib.MLP_FLOATS or mlp.device != dev:
                raise ValueError("policy='mlp' needs mlp=pack_mlp(...) on the rollout's device")
            pol.kind, pol.mlp = _lib.POLICY_MLP, mlp.data_ptr()
        elif policy == "lstm_bf16":
            if mlp is None or mlp.dtype != torch.uint8 or mlp.numel() != _lib.LSTM_IMAGE_BYTES or mlp.device != dev:
                raise ValueError("policy='lstm_bf16' needs mlp=pack_lstm(...) on the rollout's device")
            pol.kind, pol.mlp = _lib.POLICY_LSTM, mlp.data_ptr()
        elif policy == "actions":
            if actions is None or actions.dtype != torch.float32 or tuple(actions.shape) != (n_steps, n, 2) \
                    or not actions.is_contiguous() or actions.device != dev:
                raise ValueError("policy='actions' needs a contiguous float32 [n_steps, num_envs, 2] device tensor")
            pol.kind, pol.actions = _lib.POLICY_ACTIONS, actions.data_ptr()
        elif policy in POLICIES:
            pol.kind = POLICIES[policy]
        else:
            raise ValueError(f"unknown policy {policy!r}")
        stats = stats if stats is not None else self.new_stats()
        st = stats.c_struct()
        res = RolloutResult(stats, int(n_steps))
        out = None
        if store:
            res.obs = torch.empty((n_steps, n, _lib.OBS_DIM), dtype=torch.float32, device=dev)
            res.actions = torch.empty((n_steps, n, 2), dtype=torch.float32, device=dev)
            res.reward = torch.empty((n_steps, n), dtype=torch.float32, device=dev)
            done = torch.empty((n_steps, n), dtype=torch.uint8, device=dev)
            res.done = done.view(torch.bool)
            out = _lib.RolloutOut(res.obs.data_ptr(), res.actions.data_ptr(), res.reward.data_ptr(), done.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cantor_rollout(
                C.byref(self._params), C.byref(self._book) if self._sim is None else None,
                C.byref(self._sim) if self._sim is not None else None, self.episode_length, C.byref(pol), n,
                self.env_offset, self.total_envs, int(first_episode), int(n_steps), C.byref(st), C.byref(out) if out is not None else None,
                _lib.current_stream_ptr(dev)), "cantor_rollout")
        return res
