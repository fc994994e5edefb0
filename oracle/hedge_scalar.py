"""Scalar, one-environment port of the reference env step  --  TEST / BASELINE INFRASTRUCTURE ONLY.

``bench.py``'s CPU-baseline legs time THIS: it has the cost profile of the reference
(``src/env/hedging_env_v2.py:175-294``): one Python object per environment, NumPy scalars, two
``scipy.stats.norm`` calls per step in the greeks (:100-106), a 24-key info dict per step (:268-293).
``oracle/hedge_oracle.py`` is the vectorised restatement used as the parity checker; the two are compared
in ``tests/test_oracle_env.py`` so the baseline being timed is known to compute the same thing.

The product package never imports this module.
"""
from __future__ import annotations

import numpy as np
from scipy.stats import norm

from .hedge_oracle import EnvParams

F32 = np.float32


class ScalarEnv:
    """One reference-shaped environment over float32 copies of the env-schema arrays."""

    def __init__(self, paths, vols, calls, puts, params: EnvParams = EnvParams(), seed=None):
        self.p = params
        self.S, self.V = np.asarray(paths).astype(F32), np.asarray(vols).astype(F32)
        self.C, self.P = np.asarray(calls).astype(F32), np.asarray(puts).astype(F32)
        self.num_episodes, self.T = self.S.shape[0], self.S.shape[1] - 1
        self.rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    # hedging_env_v2.py:79-107
    def _greeks(self, S, K, v):
        p = self.p
        if not p.record_metrics:
            return 0.0, 0.0, 0.0
        T, r = p.option_tenor_years, p.risk_free_rate
        sigma = np.sqrt(np.maximum(v, 1e-8))
        cd = pd = gm = 0.0
        if S <= 1e-6:
            cd = 0.5 if K == 0 else (0.0 if K > 0 else 1.0)
            pd = -0.5 if K == 0 else (0.0 if K < 0 else -1.0)
        elif T <= 1e-6 or sigma <= 1e-6:
            cd = 1.0 if S > K else (0.5 if S == K else 0.0)
            pd = -1.0 if S < K else (-0.5 if S == K else 0.0)
        else:
            Kc = np.maximum(K, 1e-6)
            sst = sigma * np.sqrt(T)
            num = np.log(S / Kc) + (r + 0.5 * sigma ** 2) * T
            d1 = np.sign(num) * 10.0 if sst < 1e-9 else num / sst
            cd = norm.cdf(d1)
            pd = cd - 1.0
            den = S * sst
            gm = 0.0 if abs(den) < 1e-9 else norm.pdf(d1) / den
        return cd, gm, pd

    # hedging_env_v2.py:109-143
    def _obs(self):
        p = self.p
        s0 = np.maximum(self.S0, 25.0)
        cd, gm, pd = self._greeks(self.cur_S, np.round(self.cur_S), self.cur_v)
        if self.t == 0 or self.S_prev == 0:
            ret = dv = 0.0
        else:
            ret, dv = (self.cur_S - self.S_prev) / self.S_prev, self.cur_v - self.v_prev
        mc = p.max_contracts_held_per_type
        return np.array([self.cur_S / s0, self.cur_C / s0, self.cur_P / s0,
                         self.pos_c / mc if mc else 0.0, self.pos_p / mc if mc else 0.0, self.cur_v,
                         (self.T - self.t) / self.T if self.T else 0.0, cd, gm, pd, gm,
                         np.clip(ret, -1.0, 1.0), np.clip(dv, -1.0, 1.0)], dtype=F32)

    # hedging_env_v2.py:145-173
    def reset(self, idx=None):
        p = self.p
        self.idx = int(self.rng.integers(self.num_episodes)) if idx is None else int(idx)
        self.row_S, self.row_v = self.S[self.idx], self.V[self.idx]
        self.row_C, self.row_P = self.C[self.idx], self.P[self.idx]
        self.t = 0
        self.S0 = self.row_S[0]
        if self.S0 < 1e-6:
            self.S0 = 1.0
        self.cur_S, self.cur_v, self.cur_C, self.cur_P = self.row_S[0], self.row_v[0], self.row_C[0], self.row_P[0]
        self.pos_c = self.pos_p = 0
        self.cash = p.initial_cash
        self.pv_prev = (p.shares_to_hedge * self.cur_S) + 0 + self.cash
        self.S_prev, self.v_prev = self.cur_S, self.cur_v
        return self._obs(), {}

    # hedging_env_v2.py:175-294
    def step(self, action):
        p = self.p
        mt, mc, mult = p.max_trade_per_step, p.max_contracts_held_per_type, p.option_contract_multiplier
        cf_c, cf_p = action[0] * mt, action[1] * mt
        req_c = np.clip(np.rint(cf_c).astype(int), -mt, mt)
        req_p = np.clip(np.rint(cf_p).astype(int), -mt, mt)
        prev_c, prev_p = self.pos_c, self.pos_p
        self.pos_c = np.clip(prev_c + req_c, -mc, mc).astype(int)
        self.pos_p = np.clip(prev_p + req_p, -mc, mc).astype(int)
        tc, tp = self.pos_c - prev_c, self.pos_p - prev_p
        commission = (abs(tc) + abs(tp)) * p.transaction_cost_per_contract
        slip = abs(tc) * self.cur_C * mult * (p.slippage_bps / 10000.0) + \
            abs(tp) * self.cur_P * mult * (p.slippage_bps / 10000.0)
        costs = commission + slip
        self.cash -= costs
        self.S_prev, self.v_prev = self.cur_S, self.cur_v
        self.t += 1
        terminated = self.t >= self.T
        self.cur_S, self.cur_v = self.row_S[self.t], self.row_v[self.t]
        k = self.t - 1 if terminated else self.t
        self.cur_C, self.cur_P = self.row_C[k], self.row_P[k]
        opt_val = (self.pos_c * self.cur_C * mult) + (self.pos_p * self.cur_P * mult)
        pv = (p.shares_to_hedge * self.cur_S) + opt_val + self.cash
        step_pnl = pv - self.pv_prev
        pps = step_pnl / p.shares_to_hedge if p.shares_to_hedge != 0 else step_pnl
        s0_floor = np.maximum(self.S0, 25.0)
        if p.loss_type == "mse":
            term = (pps ** 2) / (s0_floor ** 2 + 1e-9)
        else:
            term = np.abs(pps) / (s0_floor + 1e-9)
        rpc = -p.pnl_penalty_weight * term
        tcp = p.lambda_cost * costs
        theta_pen = p.theta_weight * ((self.T - self.t) / 252.0)
        reward = rpc - tcp - theta_pen
        self.pv_prev = pv
        obs = self._obs()
        info = {
            "step_pnl_total": step_pnl, "per_share_step_pnl": pps, "raw_pnl_deviation_abs": np.abs(pps),
            "transaction_costs_total": costs, "commission_cost": commission, "slippage_cost": slip,
            "reward_pnl_component": rpc, "transaction_cost_penalty": tcp, "theta_penalty": theta_pen,
            "reward_step": reward, "portfolio_value": pv, "call_contracts": self.pos_c, "put_contracts": self.pos_p,
            "cash": self.cash, "raw_action_call": action[0], "raw_action_put": action[1],
            "scaled_float_call": cf_c, "scaled_float_put": cf_p, "requested_calls_rounded_clipped": req_c,
            "requested_puts_rounded_clipped": req_p, "actual_calls_traded": tc, "actual_puts_traded": tp,
            "loss_type_used": p.loss_type, "initial_S0_for_episode": self.S0,
        }
        return obs, reward, terminated, False, info


def time_scalar_env(paths, vols, calls, puts, params: EnvParams, seconds: float, seed: int = 0):
    """Step one ScalarEnv with pre-generated uniform float32 actions for ~``seconds``; returns (env_steps, elapsed)."""
    import time
    env = ScalarEnv(paths, vols, calls, puts, params, seed=seed)
    acts = np.random.default_rng(seed).uniform(-1, 1, (4096, 2)).astype(F32)
    env.reset()
    n, t0 = 0, time.perf_counter()
    while True:
        for a in acts:
            _, _, term, _, _ = env.step(a)
            n += 1
            if term:
                env.reset()
        el = time.perf_counter() - t0
        if el >= seconds:
            return n, el
