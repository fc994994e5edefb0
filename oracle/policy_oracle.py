"""CPU ORACLE for the hand-written policies and the evaluation statistics  --  TEST INFRASTRUCTURE ONLY.

Reference followed:

  * ``policy_no_hedge``               src/agents/baselines.py:74-75
  * ``policy_delta_every_step``       src/agents/baselines.py:77-103          -> ``delta_every_step``
  * ``delta_hedging_action_selector`` src/benchmark/delta_and_nothing.py:122-163 -> ``delta_benchmark``
  * ``evaluate_baseline_policy``      src/agents/baselines.py:32-72           -> ``episode_statistics`` (a, c)
  * ``run_evaluation`` statistics     src/agents/train_ppo_v2.py:482-530      -> ``episode_statistics`` (b, CVaR95)

Both policies read ``env.max_trade_per_step``, which the reference env never defines (AttributeError as shipped,
SURVEY appendix A.13); here it is the constructor's ``max_trade_per_step``.  Both return CONTRACT COUNTS where the
env expects fractions in [-1, 1]; the env multiplies by max_trade again and clips -- reproduced, not "fixed".

Parity status: PINNED against the unmodified reference functions in the build container
(tests/test_oracle_policy.py, skipped where /root/reference is absent).  The functions are vectorised over envs
but keep NumPy's scalar promotion (float32 observation entries, weak python scalars).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def no_hedge(obs):
    return np.zeros((obs.shape[0], 2), F32)


def delta_every_step(obs, max_contracts_held=200, option_contract_multiplier=100, shares_held_fixed=10000,
                     max_trade_per_step=15):
    """baselines.py:77-103 for a batch of float32 observations (n, 13) -> float32 actions (n, 2)."""
    obs = np.asarray(obs, F32)
    cd, pd = obs[:, 7], obs[:, 9]
    with np.errstate(all="ignore"):
        cpos = obs[:, 3] * F32(max_contracts_held)
        ppos = obs[:, 4] * F32(max_contracts_held)
        opt = (cpos * cd + ppos * pd) * F32(option_contract_multiplier)
        total = F32(shares_held_fixed) + opt
        target = -total
        cdm = cd * F32(option_contract_multiplier)
        pdm = pd * F32(option_contract_multiplier)
        use_c = np.abs(cdm) > F32(1e-1)
        use_p = (~use_c) & (np.abs(pdm) > F32(1e-1))
        tc = np.where(use_c, target / np.where(use_c, cdm, F32(1)), F32(0)).astype(F32)
        tp = np.where(use_p, target / np.where(use_p, pdm, F32(1)), F32(0)).astype(F32)
    mt = F32(max_trade_per_step)
    return np.stack([np.clip(tc, -mt, mt), np.clip(tp, -mt, mt)], axis=1).astype(F32)


def delta_benchmark(obs, pos_c, pos_p, option_contract_multiplier=100, shares_to_hedge=10000, max_trade_per_step=15):
    """delta_and_nothing.py:122-163 for a batch: int64 positions times float32 deltas promote to float64."""
    obs = np.asarray(obs, F32)
    cd, pd = obs[:, 7], obs[:, 9]
    n = obs.shape[0]
    mult = option_contract_multiplier
    cur = (np.asarray(pos_c, np.int64) * cd.astype(np.float64) + np.asarray(pos_p, np.int64) * pd.astype(np.float64)) * mult
    change = -shares_to_hedge - cur
    thr = (F32(0.5) * np.abs(cd) * F32(mult)).astype(np.float64)
    out = np.zeros((n, 2), np.float64)
    with np.errstate(all="ignore"):
        act = ~(np.abs(change) < thr)
        pos = act & (change > 0) & (np.abs(cd) > 1e-6)
        neg = act & (change < 0) & (np.abs(pd) > 1e-6)
        # `delta * multiplier` is a float32 product (python int is weak); the quotient is float64
        out[pos, 0] = np.clip(change[pos] / (cd[pos] * F32(mult)).astype(np.float64), -max_trade_per_step, max_trade_per_step)
        out[neg, 1] = np.clip(change[neg] / (pd[neg] * F32(mult)).astype(np.float64), -max_trade_per_step, max_trade_per_step)
    return out.astype(F32)


def episode_statistics(pps, costs, rewards, episode_length):
    """Per-episode and aggregate statistics from per-step arrays shaped (n_episodes, T).

    a = mean_t |pps| (baselines.py:49,54), b = |sum_t pps| / T (train_ppo_v2.py:482,520), c = sum_t cost / T
    (:483,521), R = sum_t reward (:484); mean / population std over episodes (np.mean / np.std), CVaR95 = mean of the
    sorted b from index int(0.95 n) on (:527-530).
    """
    pps, costs, rewards = (np.asarray(x, np.float64) for x in (pps, costs, rewards))
    T = episode_length
    a = np.abs(pps).sum(1) / T
    b = np.abs(pps.sum(1)) / T
    c = costs.sum(1) / T
    R = rewards.sum(1)
    sb = np.sort(b)
    return dict(n_episodes=len(a), mean_abs_pnl_baseline=a.mean(), std_abs_pnl_baseline=a.std(),
                mean_abs_pnl=b.mean(), std_abs_pnl=b.std(), mean_cost=c.mean(), std_cost=c.std(),
                mean_reward=R.mean(), std_reward=R.std(), cvar95_abs_pnl=sb[int(0.95 * len(sb)):].mean(),
                mean_signed_pnl=pps.sum(1).mean(), a=a, b=b, c=c, R=R)
