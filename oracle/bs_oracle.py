"""CPU ORACLE for Black-Scholes repricing along simulated paths  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this; the product package never does.

Reference followed (paths relative to the reference repo root):

  * ``src/sim/option_price_assignment.py:10-21``  ``black_scholes_vectorized``        -> ``black_scholes``
  * ``src/sim/option_price_assignment.py:23-31``  ``calculate_annualized_vol_matrix`` -> ``realised_vol_matrix``
  * ``src/sim/option_price_assignment.py:33-52``  ``process_price_paths``             -> ``schema_b_book``
  * ``src/tools/bs_delta.py:11-55``               single-call delta hedge             -> ``bs_delta_hedge``
  * ``src/env/hedging_env_v2.py:56-58,124``       ATM strike / tenor convention       -> ``atm_book``
    (closed-form stand-in for the nested-MC pricer ``src/sim/rbergomi_sim.py:246-306``,
    as BASELINE.json's north_star prescribes)

Parity status: PINNED by the reference's own known-answer pair
``data/paths.npy -> data/paths_options.npz`` (a 48-path slice is committed as
``tests/golden/schema_b_golden.npz``), reproduced to <= 2e-13 absolute with the
NaN pattern (column 1) identical, and by ``tests/golden/bs_delta_golden.npz``
generated from the unmodified ``bs_delta.py``.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import ndtr

RISK_FREE_RATE = 0.04          # option_price_assignment.py:8, bs_delta.py:8, hedging_env_v2.py:57
TRADING_DAYS = 252


def black_scholes(S, K, T, r, sigma, epsilon=1e-8):
    """option_price_assignment.py:10-21.  Returns (call, put), float64."""
    S = np.asarray(S, np.float64)
    K = np.asarray(K, np.float64)
    T = np.asarray(T, np.float64)
    sigma = np.asarray(sigma, np.float64)
    with np.errstate(all="ignore"):
        T_safe = np.where(T <= 0, 1e-8, T)
        sigma_safe = np.where(sigma < epsilon, epsilon, sigma)      # NaN < eps is False -> NaN propagates
        sq = np.sqrt(T_safe)
        d1 = (np.log(S / K) + (r + 0.5 * sigma_safe ** 2) * T_safe) / (sigma_safe * sq)
        d2 = d1 - sigma_safe * sq
        disc = np.exp(-r * T_safe)
        call = S * ndtr(d1) - K * disc * ndtr(d2)
        put = K * disc * ndtr(-d2) - S * ndtr(-d1)
        disc0 = np.exp(-r * T)
        call = np.where(T <= 0, np.maximum(S - K * disc0, 0), call)
        put = np.where(T <= 0, np.maximum(K * disc0 - S, 0), put)
    return call, put


def realised_vol_matrix(paths):
    """option_price_assignment.py:23-31: sigma[:, t] = std(log-returns of paths[:, :t+1], ddof=1) * sqrt(252).

    Column 0 is 0, column 1 is NaN (one sample with ddof=1).  O(N T^2) like the reference.
    """
    paths = np.asarray(paths, np.float64)
    n, t1 = paths.shape
    vols = np.zeros((n, t1))
    with np.errstate(all="ignore"), np.testing.suppress_warnings() as sup:
        sup.filter(RuntimeWarning)
        for t in range(1, t1):
            sl = paths[:, :t + 1]
            lr = np.log(sl[:, 1:] / sl[:, :-1])
            vols[:, t] = np.std(lr, axis=1, ddof=1) * math.sqrt(252)
    return vols


def schema_b_book(paths, r=RISK_FREE_RATE):
    """option_price_assignment.py:33-52 -> (calls, puts, vols), each (n, T+1) float64 ("schema B")."""
    paths = np.asarray(paths, np.float64)
    n, t1 = paths.shape
    strikes = np.round(paths[:, 0])                                   # :36
    T = np.clip(1 - np.arange(t1) / 252, 0, None)                     # :38
    vols = realised_vol_matrix(paths)
    calls = np.zeros((n, t1))
    puts = np.zeros((n, t1))
    for t in range(t1):
        calls[:, t], puts[:, t] = black_scholes(paths[:, t], strikes, T[t], r, vols[:, t])
    return calls, puts, vols


def atm_book(paths, variances, r=RISK_FREE_RATE, tenor=30 / 252):
    """Env-schema ATM option columns: for t < T, K = round(S_t), maturity = tenor, sigma = sqrt(v_t).

    Shapes: paths, variances (n, T+1) -> calls, puts (n, T).  This is the closed-form replacement
    for rbergomi_sim.py:418,437-446 named by the north star; strike rounding follows
    rbergomi_sim.py:418 (``cp.round``, half-to-even) and the tenor rbergomi_sim.py:19.
    """
    S = np.asarray(paths, np.float64)[:, :-1]
    v = np.asarray(variances, np.float64)[:, :-1]
    K = np.round(S)
    return black_scholes(S, K, tenor, r, np.sqrt(np.maximum(v, 0.0)))


def call_delta_gamma(S, K, T, r, sigma, epsilon=1e-8):
    """Closed-form call delta Phi(d1) and gamma phi(d1) / (S sigma sqrt(T)) in float64, the formulas of
    ``HedgingEnv._calculate_greeks`` (src/env/hedging_env_v2.py:94-106) for a general strike / maturity; at T <= 0 the
    step delta of :90-92 and gamma 0.  sigma floored like ``black_scholes`` (option_price_assignment.py:12)."""
    S, K, T, sigma = (np.asarray(x, np.float64) for x in (S, K, T, sigma))
    with np.errstate(all="ignore"):
        T_safe = np.where(T <= 0, 1e-8, T)
        sig = np.where(sigma < epsilon, epsilon, sigma)
        sst = sig * np.sqrt(T_safe)
        d1 = (np.log(S / K) + (r + 0.5 * sig ** 2) * T_safe) / sst
        delta = ndtr(d1)
        gamma = np.exp(-0.5 * d1 * d1) / math.sqrt(2 * math.pi) / (S * sst)
        step = np.where(S > K, 1.0, np.where(S == K, 0.5, 0.0))
        delta = np.where(T <= 0, step, delta)
        gamma = np.where(T <= 0, 0.0, gamma)
    return delta, gamma


def _scalar_call_and_delta(S, K, T, r, sigma, epsilon=1e-8):
    """bs_delta.py:11-24 (scalar, math-module arithmetic)."""
    if sigma < epsilon or T <= 0:
        return max(S - K * math.exp(-r * T), 0), (1.0 if S > K else 0.0)
    d1 = (math.log(S / K) + (r + 0.5 * sigma ** 2) * T) / (sigma * math.sqrt(T))
    d2 = d1 - sigma * math.sqrt(T)
    return S * ndtr(d1) - K * math.exp(-r * T) * ndtr(d2), float(ndtr(d1))


def bs_delta_hedge(paths, r=RISK_FREE_RATE, dt=1 / 252):
    """bs_delta.py:36-55.  paths (n, T+1) -> pnl (n, T+1) float64.

    K = S_0 (unrounded), T_total = (T+1) * dt, sigma = realised vol of the prefix
    (0 with fewer than two returns), no premium, no transaction costs.
    """
    paths = np.asarray(paths, np.float64)
    n, t1 = paths.shape
    T_total = t1 * dt
    pnl = np.zeros((n, t1))
    for i in range(n):
        prices = paths[i]
        K = prices[0]
        cash = 0.0
        prev_delta = 0.0
        lr = np.log(prices[1:] / prices[:-1])
        for t in range(t1):
            S = prices[t]
            T_rem = max(T_total - t * dt, 0.0)
            sigma = 0.0 if t < 2 else np.std(lr[:t], ddof=1) * math.sqrt(252)   # bs_delta.py:26-34
            price, delta = _scalar_call_and_delta(S, K, T_rem, r, sigma)
            cash -= (delta - prev_delta) * S
            prev_delta = delta
            pnl[i, t] = cash + prev_delta * S - price
    return pnl
