"""CPU ORACLE for the hedging environment step  --  TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy, the arithmetic of the reference's gym
environment so the CUDA path can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it;
the product package ``cantorrl_b200`` never does.

Reference followed (paths relative to the reference repo root):

  * ``src/env/hedging_env_v2.py:10-77``   constructor / constants      -> ``EnvParams``
  * ``src/env/hedging_env_v2.py:79-107``  ``_calculate_greeks``        -> ``greeks``
  * ``src/env/hedging_env_v2.py:109-143`` ``_get_observation``         -> ``OracleVecEnv.observation``
  * ``src/env/hedging_env_v2.py:145-173`` ``reset``                    -> ``OracleVecEnv.reset``
  * ``src/env/hedging_env_v2.py:175-294`` ``step``                     -> ``OracleVecEnv.step``
  * ``src/env/hedging_env.py`` (v1) is the same code with
    ``slippage_bps = 0``, ``theta_weight = 0`` and commission default 0.05
    (``hedging_env.py:10-20, 198-200, 240-242``)                       -> ``EnvParams.v1``

Parity status: PINNED.  ``tests/test_oracle_env.py`` checks this restatement
(a) against the unmodified reference class run in the build container
(skipped where ``/root/reference`` is absent) and (b) against the golden
vectors in ``tests/golden/env_*.npz`` that ``tests/golden/make_golden.py``
produced by running the unmodified reference class.

The reference is a *single* environment holding NumPy scalars; NumPy-2
promotion (NEP 50) makes it a float32/float64 mixture.  The ledger below is
reproduced operation by operation:

  =====================================  =========================================
  quantity                               dtype / rounding in the reference
  =====================================  =========================================
  paths, variances, option prices        float32 (``astype`` in ``__init__``)
  ``action * max_trade``                 float32 multiply, then ``rint`` (half-even)
  requested / traded / held contracts    int64
  commission, slippage, cash             float64
  ``shares * S`` (stock leg)             **float32** multiply (python int is weak)
  options value, portfolio value, P&L    float64
  first ``portfolio_value_t_minus_1``    float32 (reset computes it in float32)
  ``max(S0, 25)``                        float32; ``+ 1e-9`` is absorbed in float32
  ``s0_floor**2`` (mse loss)             float32 *scalar* power (libm ``powf``)
  reward                                 float64
  observation                            float32 after a final cast; d1 has a
                                         float32 numerator and float64 denominator
  =====================================  =========================================
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace

import numpy as np
from scipy.special import ndtr

F32 = np.float32
F64 = np.float64
I64 = np.int64

OBS_DIM = 13
_SQRT_2PI = np.sqrt(2 * np.pi)  # scipy.stats._continuous_distns._norm_pdf_C

# info keys of hedging_env_v2.py:268-293, numeric ones only, in reference order
INFO_FLOAT_KEYS = (
    "step_pnl_total", "per_share_step_pnl", "raw_pnl_deviation_abs",
    "transaction_costs_total", "commission_cost", "slippage_cost",
    "reward_pnl_component", "transaction_cost_penalty", "theta_penalty",
    "reward_step", "portfolio_value", "cash",
    "raw_action_call", "raw_action_put", "scaled_float_call", "scaled_float_put",
    "initial_S0_for_episode",
)
INFO_INT_KEYS = (
    "call_contracts", "put_contracts",
    "requested_calls_rounded_clipped", "requested_puts_rounded_clipped",
    "actual_calls_traded", "actual_puts_traded",
)


@dataclass(frozen=True)
class EnvParams:
    """Constructor arguments of the reference env (hedging_env_v2.py:10-22), same names and defaults."""
    transaction_cost_per_contract: float = 0.65
    lambda_cost: float = 1.0
    pnl_penalty_weight: float = 0.01
    theta_weight: float = 0.0
    slippage_bps: float = 0.0
    loss_type: str = "abs"
    initial_cash: float = 0.0
    shares_to_hedge: int = 10000
    max_contracts_held_per_type: int = 200
    max_trade_per_step: int = 15
    record_metrics: bool = True
    # constants hedging_env_v2.py:56-58
    option_contract_multiplier: int = 100
    risk_free_rate: float = 0.04
    option_tenor_years: float = 30 / 252

    @staticmethod
    def v1(**kw) -> "EnvParams":
        """hedging_env.py:10-20: commission default 0.05, no slippage / theta arguments."""
        kw.setdefault("transaction_cost_per_contract", 0.05)
        assert "theta_weight" not in kw and "slippage_bps" not in kw
        return EnvParams(**kw)


def _scalar_pow2_f32(x: np.ndarray) -> np.ndarray:
    """``np.float32(x) ** 2`` element by element.

    The reference squares NumPy *scalars* (hedging_env_v2.py:97,99,247), which
    goes through libm ``powf`` and is not always equal to ``x*x`` (≈0.07 % of
    inputs differ by one float32 ulp).  Iterating keeps the scalar code path.
    """
    x = np.asarray(x, dtype=F32)
    return np.fromiter((v ** 2 for v in x.ravel()), dtype=F32, count=x.size).reshape(x.shape)


def _scalar_pow2_f64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=F64)
    return np.fromiter((v ** 2 for v in x.ravel()), dtype=F64, count=x.size).reshape(x.shape)


def greeks(S, K, T, r, v_spot, record_metrics=True):
    """hedging_env_v2.py:79-107 on arrays.  S, K, v_spot float32 arrays; T, r python floats.

    Returns (call_delta, gamma, put_delta) as float64 arrays (the reference
    returns gamma twice).
    """
    S = np.asarray(S, dtype=F32)
    K = np.asarray(K, dtype=F32)
    v_spot = np.asarray(v_spot, dtype=F32)
    n = S.shape
    cd = np.zeros(n, F64)
    pd = np.zeros(n, F64)
    gm = np.zeros(n, F64)
    if not record_metrics:
        return cd, gm, pd
    with np.errstate(all="ignore"):
        sigma = np.sqrt(np.maximum(v_spot, F32(1e-8)))                    # :84  float32
        tiny_s = S <= F32(1e-6)                                            # :87
        degenerate = (~tiny_s) & ((T <= 1e-6) | (sigma <= F32(1e-6)))      # :90
        regular = ~(tiny_s | degenerate)
        # :88-89
        cd[tiny_s] = np.where(K[tiny_s] == 0, 0.5, np.where(K[tiny_s] > 0, 0.0, 1.0))
        pd[tiny_s] = np.where(K[tiny_s] == 0, -0.5, np.where(K[tiny_s] < 0, 0.0, -1.0))
        # :91-92
        Sd, Kd = S[degenerate], K[degenerate]
        cd[degenerate] = np.where(Sd > Kd, 1.0, np.where(Sd == Kd, 0.5, 0.0))
        pd[degenerate] = np.where(Sd < Kd, -1.0, np.where(Sd == Kd, -0.5, 0.0))
        # :94-106
        Sr, Kr, sg = S[regular], K[regular], sigma[regular]
        K_checked = np.maximum(Kr, F32(1e-6))                              # float32
        sigma_sqrt_T = sg.astype(F64) * np.sqrt(F64(T))                    # float32 * float64 -> float64
        num = np.log(Sr / K_checked) + (F32(r) + F32(0.5) * _scalar_pow2_f32(sg)) * F32(T)   # float32
        d1 = np.where(sigma_sqrt_T < 1e-9,
                      np.sign(num).astype(F64) * 10.0,
                      num.astype(F64) / sigma_sqrt_T)
        cdf = ndtr(d1)                                                     # == scipy.stats.norm.cdf
        cd[regular] = cdf
        pd[regular] = cdf - 1.0
        den = Sr.astype(F64) * sigma_sqrt_T
        pdf = np.exp(-_scalar_pow2_f64(d1) / 2.0) / _SQRT_2PI              # == scipy.stats.norm.pdf
        gm[regular] = np.where(np.abs(den) < 1e-9, 0.0, pdf / den)
    return cd, gm, pd


class OracleVecEnv:
    """N independent copies of the reference ``HedgingEnv`` advanced in lock-step.

    ``paths``/``vols`` are (n_paths, T+1), ``calls``/``puts`` (n_paths, T) as in
    the env-schema npz (hedging_env_v2.py:36-48).  Like the reference, ``step``
    does **not** auto-reset; ``step_autoreset`` adds the VecEnv convention used
    by the CUDA path (post-reset observation returned, terminal observation
    kept aside).
    """

    def __init__(self, paths, vols, calls, puts, params: EnvParams = EnvParams(), n_envs: int = 1):
        self.p = params
        self.S = np.asarray(paths).astype(F32)
        self.V = np.asarray(vols).astype(F32)
        self.C = np.asarray(calls).astype(F32)
        self.P = np.asarray(puts).astype(F32)
        if not (self.S.shape == self.V.shape
                and self.S.shape[0] == self.C.shape[0] == self.P.shape[0]
                and self.S.shape[1] == self.C.shape[1] + 1 == self.P.shape[1] + 1):
            raise ValueError("Data shapes are inconsistent.")        # :45-48
        self.num_episodes = self.S.shape[0]
        self.episode_length = self.S.shape[1] - 1
        self.n = int(n_envs)
        n = self.n
        self.idx = np.full(n, -1, I64)
        self.step_count = np.zeros(n, I64)
        self.pos_c = np.zeros(n, I64)
        self.pos_p = np.zeros(n, I64)
        self.cash = np.zeros(n, F64)
        self.pv_prev = np.zeros(n, F64)
        self.S0 = np.ones(n, F32)
        self.cur_S = np.zeros(n, F32)
        self.cur_v = np.zeros(n, F32)
        self.cur_C = np.zeros(n, F32)
        self.cur_P = np.zeros(n, F32)
        self.S_prev = np.zeros(n, F32)
        self.v_prev = np.zeros(n, F32)

    # ------------------------------------------------------------------ reset
    def reset(self, idx, mask=None):
        """hedging_env_v2.py:145-173 for the envs selected by ``mask`` (default all).

        ``idx`` are the episode indices the reference would have drawn with
        ``np_random.integers(num_episodes)`` (PCG64); the caller supplies them.
        """
        p = self.p
        if mask is None:
            mask = np.ones(self.n, bool)
        m = np.asarray(mask, bool)
        idx = np.broadcast_to(np.asarray(idx, I64), (self.n,))[m]
        self.idx[m] = idx
        self.step_count[m] = 0
        s0 = self.S[idx, 0].copy()
        cur = s0.copy()
        s0[s0 < F32(1e-6)] = F32(1.0)                                     # :157
        self.S0[m] = s0
        self.cur_S[m] = cur
        self.cur_v[m] = self.V[idx, 0]
        self.cur_C[m] = self.C[idx, 0]
        self.cur_P[m] = self.P[idx, 0]
        self.pos_c[m] = 0
        self.pos_p[m] = 0
        self.cash[m] = p.initial_cash
        # :167-168  float32: (shares * S) + 0 + cash, python scalars are weak
        pv0 = (F32(p.shares_to_hedge) * cur + F32(0)) + F32(p.initial_cash)
        self.pv_prev[m] = pv0.astype(F64)
        self.S_prev[m] = cur
        self.v_prev[m] = self.V[idx, 0]
        return self.observation()

    # ------------------------------------------------------------ observation
    def observation(self):
        """hedging_env_v2.py:109-143 -> float32 (n, 13)."""
        p = self.p
        T = self.episode_length
        with np.errstate(all="ignore"):
            s0 = np.maximum(self.S0, F32(25.0))                            # :116 float32
            o = np.zeros((self.n, OBS_DIM), F32)
            o[:, 0] = self.cur_S / s0
            o[:, 1] = self.cur_C / s0
            o[:, 2] = self.cur_P / s0
            if p.max_contracts_held_per_type != 0:
                o[:, 3] = (self.pos_c / p.max_contracts_held_per_type).astype(F32)   # int64 / int -> float64
                o[:, 4] = (self.pos_p / p.max_contracts_held_per_type).astype(F32)
            o[:, 5] = self.cur_v
            if T != 0:
                o[:, 6] = ((T - self.step_count) / T).astype(F32)
            K = np.round(self.cur_S)                                        # :124 float32 half-even
            cd, gm, pd = greeks(self.cur_S, K, p.option_tenor_years, p.risk_free_rate,
                                self.cur_v, p.record_metrics)
            o[:, 7] = cd.astype(F32)
            o[:, 8] = gm.astype(F32)
            o[:, 9] = pd.astype(F32)
            o[:, 10] = gm.astype(F32)
            lag_off = (self.step_count == 0) | (self.S_prev == 0)           # :129
            ret = np.where(lag_off, F32(0), (self.cur_S - self.S_prev) / self.S_prev).astype(F32)
            dv = np.where(lag_off, F32(0), self.cur_v - self.v_prev).astype(F32)
            o[:, 11] = np.clip(ret, F32(-1.0), F32(1.0))
            o[:, 12] = np.clip(dv, F32(-1.0), F32(1.0))
        return o

    # ------------------------------------------------------------------- step
    def step(self, actions):
        """hedging_env_v2.py:175-294 for every env.  ``actions`` float32 (n, 2).

        Returns obs float32 (n,13), reward float64 (n,), terminated bool (n,), info dict of arrays.
        Stepping an env whose previous step terminated raises IndexError, as the reference does.
        """
        p = self.p
        T = self.episode_length
        a = np.asarray(actions, dtype=F32).reshape(self.n, 2)
        if np.any(self.step_count >= T):
            raise IndexError("step() called on a terminated episode")
        mt = p.max_trade_per_step
        with np.errstate(all="ignore"):
            cf = a * F32(mt)                                               # :181-182 float32
            req = np.rint(cf).astype(I64)                                  # :184-185 (NaN/inf -> INT64_MIN on x86)
            req = np.clip(req, -mt, mt)                                    # :187-188
            prev_c, prev_p = self.pos_c.copy(), self.pos_p.copy()
            mc = p.max_contracts_held_per_type
            self.pos_c = np.clip(prev_c + req[:, 0], -mc, mc).astype(I64)  # :193-197
            self.pos_p = np.clip(prev_p + req[:, 1], -mc, mc).astype(I64)
            tc = self.pos_c - prev_c                                       # :199-200
            tp = self.pos_p - prev_p

            commission = (np.abs(tc) + np.abs(tp)) * F64(p.transaction_cost_per_contract)   # :203-204
            bps_frac = p.slippage_bps / 10000.0
            slip_c = np.abs(tc) * self.cur_C.astype(F64) * F64(p.option_contract_multiplier) * F64(bps_frac)  # :206-207
            slip_p = np.abs(tp) * self.cur_P.astype(F64) * F64(p.option_contract_multiplier) * F64(bps_frac)  # :208-209
            slippage = slip_c + slip_p
            costs = commission + slippage                                   # :212
            self.cash = self.cash - costs                                   # :213

            self.S_prev = self.cur_S.copy()                                 # :216-217
            self.v_prev = self.cur_v.copy()
            self.step_count = self.step_count + 1                           # :219
            terminated = self.step_count >= T                               # :220
            rows = self.idx
            self.cur_S = self.S[rows, self.step_count]                      # :223-224
            self.cur_v = self.V[rows, self.step_count]
            opt_col = np.where(terminated, self.step_count - 1, self.step_count)   # :226-231 (stale at the end)
            self.cur_C = self.C[rows, opt_col]
            self.cur_P = self.P[rows, opt_col]

            mult = F64(p.option_contract_multiplier)
            opt_val = (self.pos_c * self.cur_C.astype(F64) * mult) + (self.pos_p * self.cur_P.astype(F64) * mult)  # :233-234
            stock_leg = F32(p.shares_to_hedge) * self.cur_S                 # :235 float32 !
            pv = stock_leg.astype(F64) + opt_val + self.cash                # :235-236
            step_pnl = pv - self.pv_prev                                    # :237
            pps = step_pnl / F64(p.shares_to_hedge) if p.shares_to_hedge != 0 else step_pnl   # :238
            raw_abs = np.abs(pps)                                           # :240

            s0_floor = np.maximum(self.S0, F32(25.0))                       # :244 float32
            if p.loss_type == "mse":
                den = _scalar_pow2_f32(s0_floor) + F32(1e-9)                # :247 float32
                term = _scalar_pow2_f64(pps) / den.astype(F64)
            else:                                                           # abs / cvar / anything else :248-253
                den = s0_floor + F32(1e-9)
                term = np.abs(pps) / den.astype(F64)
            rpc = -p.pnl_penalty_weight * term                              # :255
            tcp = p.lambda_cost * costs                                     # :257
            tte = (T - self.step_count) / 252.0                             # :259
            theta_pen = p.theta_weight * tte                                # :260
            reward = rpc - tcp - theta_pen                                  # :262
            self.pv_prev = pv                                               # :265
        obs = self.observation()                                            # :266
        info = {
            "step_pnl_total": step_pnl, "per_share_step_pnl": pps, "raw_pnl_deviation_abs": raw_abs,
            "transaction_costs_total": costs, "commission_cost": commission, "slippage_cost": slippage,
            "reward_pnl_component": rpc, "transaction_cost_penalty": tcp,
            "theta_penalty": np.asarray(theta_pen, F64), "reward_step": reward,
            "portfolio_value": pv, "cash": self.cash.copy(),
            "call_contracts": self.pos_c.copy(), "put_contracts": self.pos_p.copy(),
            "raw_action_call": a[:, 0].copy(), "raw_action_put": a[:, 1].copy(),
            "scaled_float_call": cf[:, 0].copy(), "scaled_float_put": cf[:, 1].copy(),
            "requested_calls_rounded_clipped": req[:, 0].copy(), "requested_puts_rounded_clipped": req[:, 1].copy(),
            "actual_calls_traded": tc, "actual_puts_traded": tp,
            "initial_S0_for_episode": self.S0.copy(),
        }
        return obs, reward, terminated, info

    # ------------------------------------------------- VecEnv-style auto reset
    def step_autoreset(self, actions, next_idx):
        """``step`` followed by ``reset`` of the envs that terminated (SB3 VecEnv convention).

        Returns obs (post-reset rows for finished envs), reward, done, terminal_obs (the
        pre-reset observation, meaningful where done), info.
        """
        obs, reward, done, info = self.step(actions)
        terminal_obs = obs.copy()
        if done.any():
            obs = self.reset(next_idx, mask=done)
        return obs, reward, done, terminal_obs, info


def with_params(p: EnvParams, **kw) -> EnvParams:
    return replace(p, **kw)
