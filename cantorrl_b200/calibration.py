"""Base rough-Bergomi parameters from a price history -- host-side scalar calibration, not part of the GPU path.

Mirror of ``estimate_base_params`` (``src/sim/rbergomi_sim.py:171-193`` and its helpers ``:56-168``): five numbers
``(S0, xi, H, eta, rho)`` from ~1000 closing prices, computed once before ``generate_rbergomi_paths_and_options``.
It is O(n) NumPy on a kilobyte of data, so it stays on the host; it is here only so
that the generator accepts the same first argument as the reference's (a price history).

  xi   annualised variance of the log-returns (sample variance / dt)                                  (:62-64)
  H    detrended-fluctuation-analysis slope of log F(w) on log w, windows 10, 20, 40, ..., n // 4, clipped to [.01, .49]  (:82-134)
  eta  sqrt(252) x sample std of the day-to-day change of log(20-day mean squared return)             (:139-157)
  rho  correlation of returns with squared returns; a positive estimate becomes -0.3; clipped to [-.99, -.01]            (:159-172)
"""
from __future__ import annotations

import numpy as np

XI_DEFAULT, H_DEFAULT, ETA_DEFAULT, RHO_DEFAULT, S0_DEFAULT = 0.04, 0.1, 1.0, -0.7, 100.0      # rbergomi_sim.py:23-27
CLIP_H = (0.01, 0.49)                                                                          # :37-38
CLIP_RHO = (-0.99, -0.01)                                                                      # :39-40


def _window_sizes(n_points: int, smallest: int = 10):
    """10, 20, 40, ... and finally n // 4 itself (the reference's doubling schedule, :93-118)."""
    largest = n_points // 4
    out, w = [], smallest
    while w <= largest:
        out.append(w)
        if w == largest:
            break
        w = largest if 2 * w > largest else 2 * w
    return out


def hurst_dfa(x) -> float:
    """Detrended fluctuation analysis (:82-134): slope of log mean-RMS-residual against log window size."""
    x = np.asarray(x, np.float64)
    if x.size < 20:
        return H_DEFAULT
    profile = np.cumsum(x - x.mean())
    log_w, log_f = [], []
    for w in _window_sizes(profile.size):
        m = profile.size // w                                   # non-overlapping windows from the start
        seg = profile[: m * w].reshape(m, w)
        t = np.arange(1, w + 1, dtype=np.float64)
        tc = t - t.mean()
        slope = (seg - seg.mean(axis=1, keepdims=True)) @ tc / (tc @ tc)          # least-squares line per window
        resid = seg - seg.mean(axis=1, keepdims=True) - slope[:, None] * tc[None, :]
        rms = np.sqrt((resid ** 2).mean(axis=1))
        rms = rms[rms > 1e-8]
        if rms.size and rms.mean() > 1e-8:
            log_w.append(np.log(w))
            log_f.append(np.log(rms.mean()))
    if len(log_w) < 2:
        return H_DEFAULT
    lw, lf = np.array(log_w), np.array(log_f)
    n = lw.size
    den = n * (lw ** 2).sum() - lw.sum() ** 2
    if abs(den) < 1e-14:
        return H_DEFAULT
    return float(np.clip((n * (lw * lf).sum() - lw.sum() * lf.sum()) / den, *CLIP_H))


def vol_of_vol(logrets, window: int = 20) -> float:
    """:139-157 -- annualised std of the daily change of the log rolling mean-square return."""
    r = np.asarray(logrets, np.float64)
    if r.size < window + 1:
        return ETA_DEFAULT
    c = np.concatenate([[0.0], np.cumsum(r * r)])
    rv = (c[window:] - c[:-window]) / window                     # trailing means over `window` returns
    d = np.diff(np.log(rv))
    if d.size < 2:
        return ETA_DEFAULT
    return float(np.std(d, ddof=1) * np.sqrt(252.0))


def leverage_correlation(logrets) -> float:
    """:159-172 -- corr(r, r^2), forced negative."""
    r = np.asarray(logrets, np.float64)
    if r.size < 2:
        return RHO_DEFAULT
    q = r * r
    vr, vq = np.var(r, ddof=1), np.var(q, ddof=1)
    if vr == 0 or vq == 0:
        return RHO_DEFAULT
    den = np.sqrt(vr * vq)
    rho = np.cov(r, q, ddof=1)[0, 1] / den if den != 0.0 else 0.0
    if rho > 0.0:
        rho = -0.3
    return float(np.clip(rho, *CLIP_RHO))


def estimate_base_params(historical_prices, dt: float = 1 / 252):
    """``(S0, xi, H, eta, rho)`` exactly as the reference derives them (:171-193), defaults and guards included."""
    p = np.asarray(historical_prices, np.float64)
    if p.size < 21:
        return (float(p[-1]) if p.size else S0_DEFAULT), XI_DEFAULT, H_DEFAULT, ETA_DEFAULT, RHO_DEFAULT
    r = np.log(p[1:] / p[:-1])
    xi = (np.var(r, ddof=1) if r.size >= 2 else 0.0) / dt
    H = hurst_dfa(r)
    eta = vol_of_vol(r)
    rho = leverage_correlation(r)
    xi = XI_DEFAULT if (not np.isfinite(xi) or xi <= 1e-6) else float(xi)
    H = H_DEFAULT if not np.isfinite(H) else H
    eta = ETA_DEFAULT if (not np.isfinite(eta) or eta <= 1e-6) else eta
    rho = RHO_DEFAULT if not np.isfinite(rho) else rho
    return float(p[-1]), xi, H, eta, rho
