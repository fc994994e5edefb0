"""CPU ORACLE for the episode-fused rollout  --  TEST INFRASTRUCTURE ONLY (never imported by the product package).

Restates, in NumPy, the loop the reference runs around its env

  * ``evaluate_baseline_policy``  src/agents/baselines.py:32-72      (reset; while not done: policy -> env.step)
  * ``run_evaluation``            src/agents/train_ppo_v2.py:465-530 (per-episode sums, statistics)

on top of ``hedge_oracle.OracleVecEnv`` (the env step) and ``policy_oracle`` (the hand-written policies), plus the
two policy sources the CUDA rollout adds: counter-based uniform actions (the stand-in for
``action_space.sample()``, src/agents/test_inf.py:29) and the ReLU MLP actor 13-64-64-2 on the normalised
observation (quantconnect/model_wrapper.py:131,177-185; output clipped like SB3 clips a Box action).

Episode ``e`` of global env ``g`` runs on path ``(e * total_envs + g) % n_paths`` -- the rule of
``cantorrl_b200/csrc/rollout.cu`` -- so a population sharded over ranks visits the same paths.

Parity status: the env step and the two delta policies are PINNED (see hedge_oracle.py / policy_oracle.py); the
uniform-action stream is pinned by the Random123 known-answer vectors of ``sim_oracle.philox4x32_10``; the recurrent
actor is PINNED on ``tests/golden/lstm_golden.npz`` and the plain MLP on ``tests/golden/mlp_golden.npz`` -- both produced by
the reference's own network modules (quantconnect/model_wrapper.py:167-204) with the shipped ``policy_weights.pth``
(``tests/golden/make_golden.py --lstm-only / --mlp-only``; the MLP is the shipped head behind a fixed 13 -> 128 projection).
"""
from __future__ import annotations

import numpy as np

from . import policy_oracle
from .hedge_oracle import EnvParams, OracleVecEnv
from .sim_oracle import philox4x32_10

F32 = np.float32
STREAM_ACTIONS = 0x4143544E      # "ACTN"


def uniform_actions(seed, global_env, step):
    """uniform[-1, 1) float32 pairs of Philox(seed; global env, rollout step, "ACTN") (rollout.cu: policy_random)."""
    g = np.asarray(global_env, np.uint64).ravel()
    ctr = np.zeros((g.size, 4), np.uint32)
    ctr[:, 0] = (g & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (g >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(step)
    ctr[:, 3] = STREAM_ACTIONS
    x = philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32))
    u = (x[:, :2] >> np.uint32(8)).astype(F32)                    # 24 bits: exact in float32
    return (u * F32(2.0 ** -23) + F32(-1.0)).astype(F32)            # fmaf is exact here (24-bit integer * 2^-23 - 1)


def _squash(out, squash):
    """SB3 clips the action means to the Box when it steps the env; the deployment wrapper applies tanh (model_wrapper.py:202)."""
    return np.tanh(out).astype(F32) if squash == "tanh" else np.clip(out, -1, 1).astype(F32)


def mlp_actor(obs, W1, b1, W2, b2, W3, b3, mean=None, var=None, epsilon=1e-8, squash="clip", obs_clip=10.0):
    """float32 actor: clip((obs - mean) / sqrt(var + eps), +-obs_clip) -> ReLU(64) -> ReLU(64) -> 2, clipped to [-1, 1] (or tanh).

    Weights in ``torch.nn.Linear`` layout (``[out, in]``).  Accumulation in float64 then rounded: the CUDA kernel's
    float32 FMA chain agrees to ~1e-6 relative.
    """
    x = np.asarray(obs, F32)
    if mean is not None:
        inv = (1.0 / np.sqrt(np.asarray(var, np.float64) + epsilon)).astype(F32)
        x = np.clip((x - np.asarray(mean, F32)) * inv, F32(-obs_clip), F32(obs_clip))
    h = np.maximum(x.astype(np.float64) @ np.asarray(W1, np.float64).T + b1, 0)
    h = np.maximum(h @ np.asarray(W2, np.float64).T + b2, 0)
    out = h @ np.asarray(W3, np.float64).T + b3
    return _squash(out, squash)


def _bf16(x):
    """Round float32 to bfloat16 (round to nearest even), returned as float32."""
    u = np.asarray(x, F32).view(np.uint32).astype(np.uint64)
    r = ((u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)) << np.uint64(16)
    return r.astype(np.uint32).view(F32)


def mlp_actor_bf16(obs, W1, b1, W2, b2, W3, b3, mean=None, var=None, epsilon=1e-8, squash="clip", obs_clip=10.0):
    """The tensor-core form of ``mlp_actor`` (cantorrl_b200/csrc/mlp_tc.cuh): inputs, weights, biases and hidden
    activations rounded to bfloat16, products accumulated in float32 (float64 here: the order is the hardware's)."""
    x = np.asarray(obs, F32)
    if mean is not None:
        inv = (1.0 / np.sqrt(np.asarray(var, np.float64) + epsilon)).astype(F32)
        x = np.clip((x - np.asarray(mean, F32)) * inv, F32(-obs_clip), F32(obs_clip))
    h = _bf16(x).astype(np.float64) @ _bf16(W1).astype(np.float64).T + _bf16(b1)
    h = _bf16(np.maximum(h, 0).astype(F32)).astype(np.float64) @ _bf16(W2).astype(np.float64).T + _bf16(b2)
    out = _bf16(np.maximum(h, 0).astype(F32)).astype(np.float64) @ _bf16(W3).astype(np.float64).T + _bf16(b3)
    return _squash(out, squash)


def lstm_actor_sequence(obs_seq, done_seq, w_ih, w_hh, b_ih, b_hh, W1, b1, W2, b2, W3, b3, mean=None, var=None, epsilon=1e-8,
                        bf16=True, squash="clip", obs_clip=10.0):
    """The recurrent actor LSTM(13 -> 128) -> ReLU MLP(128 -> 64 -> 64) -> 2 (quantconnect/model_wrapper.py:167-204) over an
    observation sequence ``[n_steps, n_envs, 13]``; the state (h, c) of an env is zeroed after a step on which it finished
    (``done_seq [n_steps, n_envs]``), as SB3 does at episode starts.  ``torch.nn.LSTM`` weight layout, gate order i, f, g, o.

    ``squash``: "clip" (SB3 clips the action means to the Box when it steps the env) or "tanh" (the reference's deployment
    wrapper, quantconnect/model_wrapper.py:202).  ``obs_clip``: SB3 VecNormalize's clip of the normalised observation (10); the
    deployment wrapper does not clip (:131) -- pass ``np.inf``.  Pinned: with ``bf16=False, squash="tanh"`` this reproduces the reference's
    own ``RecurrentPPOModel`` run on the shipped ``policy_weights.pth`` (tests/golden/lstm_golden.npz, tests/test_oracle_policy.py).
    ``bf16=True`` follows the rounding points of cantorrl_b200/csrc/lstm_tc.cuh (inputs, weights, biases, h and the head's
    activations in bfloat16; products accumulated wide; c in float32); ``bf16=False`` is the plain float64 network.
    """
    q = _bf16 if bf16 else (lambda a: np.asarray(a, np.float64))
    n_steps, n, _ = obs_seq.shape
    wi, wh, bb = q(w_ih).astype(np.float64), q(w_hh).astype(np.float64), q(np.asarray(b_ih, F32) + np.asarray(b_hh, F32)).astype(np.float64)
    A1, c1 = q(W1).astype(np.float64), q(b1).astype(np.float64)
    A2, c2 = q(W2).astype(np.float64), q(b2).astype(np.float64)
    A3, c3 = q(W3).astype(np.float64), q(b3).astype(np.float64)
    h = np.zeros((n, 128))
    c = np.zeros((n, 128))
    out = np.zeros((n_steps, n, 2), F32)
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))  # noqa: E731
    for t in range(n_steps):
        x = np.asarray(obs_seq[t], F32)
        if mean is not None:
            inv = (1.0 / np.sqrt(np.asarray(var, np.float64) + epsilon)).astype(F32)
            x = np.clip((x - np.asarray(mean, F32)) * inv, F32(-obs_clip), F32(obs_clip))
        g = q(x).astype(np.float64) @ wi.T + q(h.astype(F32)).astype(np.float64) @ wh.T + bb
        gi, gf, gg, go = g[:, 0:128], g[:, 128:256], g[:, 256:384], g[:, 384:512]
        c = (sig(gf) * c + sig(gi) * np.tanh(gg)).astype(F32).astype(np.float64) if bf16 else sig(gf) * c + sig(gi) * np.tanh(gg)
        h = sig(go) * np.tanh(c)
        hq = q(h.astype(F32)).astype(np.float64)
        a1 = np.maximum(hq @ A1.T + c1, 0)
        a2 = np.maximum(q(a1.astype(F32)).astype(np.float64) @ A2.T + c2, 0)
        o = q(a2.astype(F32)).astype(np.float64) @ A3.T + c3
        out[t] = np.tanh(o) if squash == "tanh" else np.clip(o, -1, 1)
        fin = np.asarray(done_seq[t], bool)
        h[fin] = 0
        c[fin] = 0
    return out


def run_rollout(paths, vols, calls, puts, params: EnvParams, policy, n_envs, n_steps, env_offset=0, total_envs=None,
                forced_actions=None, one_call_only=False, seed=0, mlp=None):
    """Free-running (or, with ``forced_actions`` [n_steps, n_envs, 2], teacher-forced) rollout of the oracle env.

    Returns a dict: ``obs`` (what the policy saw) [n_steps, n_envs, 13], ``actions``, ``policy_actions`` (what the
    oracle policy would have played on that observation), ``reward`` / ``pps`` / ``cost`` [n_steps, n_envs],
    ``done``, and the finished episodes' per-step series ``ep_pps`` / ``ep_cost`` / ``ep_reward`` [(episodes), T].
    """
    total = n_envs + env_offset if total_envs is None else total_envs
    env = OracleVecEnv(paths, vols, calls, puts, params, n_envs)
    T, n_paths = env.episode_length, env.num_episodes
    genv = env_offset + np.arange(n_envs, dtype=np.int64)
    episode = np.zeros(n_envs, np.int64)
    obs = env.reset((episode * total + genv) % n_paths)
    rec = {k: [] for k in ("obs", "actions", "policy_actions", "reward", "pps", "cost", "done")}
    cur = {k: np.zeros((n_envs, T)) for k in ("pps", "cost", "reward")}
    fin = {k: [] for k in ("pps", "cost", "reward")}
    for g in range(n_steps):
        if policy == "no_hedge":
            a = policy_oracle.no_hedge(obs)
        elif policy == "random":
            a = uniform_actions(seed, genv, g)
        elif policy == "delta_every_step":
            a = policy_oracle.delta_every_step(obs, params.max_contracts_held_per_type, 100, params.shares_to_hedge,
                                               params.max_trade_per_step)
        elif policy == "delta_benchmark":
            a = policy_oracle.delta_benchmark(obs, env.pos_c, env.pos_p, 100, params.shares_to_hedge, params.max_trade_per_step)
        elif policy == "mlp":
            a = mlp_actor(obs, *mlp)
        elif policy == "mlp_bf16":
            a = mlp_actor_bf16(obs, *mlp)
        elif policy == "actions":
            a = np.asarray(forced_actions[g], F32)
        else:
            raise ValueError(policy)
        a = np.array(a, F32)
        if one_call_only:
            a[:, 1] = 0
        played = a if forced_actions is None else np.asarray(forced_actions[g], F32)
        t = env.step_count.copy()
        episode_next = episode + 1
        nobs, r, d, _, info = env.step_autoreset(played, (episode_next * total + genv) % n_paths)
        rows = np.arange(n_envs)
        cur["pps"][rows, t] = info["per_share_step_pnl"]
        cur["cost"][rows, t] = info["transaction_costs_total"]
        cur["reward"][rows, t] = r
        for k, v in (("obs", obs), ("actions", played), ("policy_actions", a), ("reward", r),
                     ("pps", info["per_share_step_pnl"]), ("cost", info["transaction_costs_total"]), ("done", d)):
            rec[k].append(np.array(v))
        if d.any():
            for k in fin:
                fin[k].append(cur[k][d].copy())
            episode = np.where(d, episode_next, episode)
        obs = nobs
    out = {k: np.stack(v) for k, v in rec.items()}
    for k in fin:
        out["ep_" + k] = np.concatenate(fin[k]) if fin[k] else np.zeros((0, T))
    out["episode_length"] = T
    return out


def stats_vector(ep_pps, ep_cost, ep_reward, T):
    """The ``sums[0:11]`` vector of include/cantor_hedge.h from finished episodes' per-step series."""
    a = np.abs(ep_pps).sum(1) / T
    b = np.abs(ep_pps.sum(1)) / T
    c = ep_cost.sum(1) / T
    R = ep_reward.sum(1)
    s = ep_pps.sum(1)
    v = [float(len(a))]
    for x in (a, b, c, R, s):
        v += [x.sum(), (x * x).sum()]
    return np.array(v), b
