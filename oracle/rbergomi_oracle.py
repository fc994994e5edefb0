"""CPU ORACLE for the rough-Bergomi generator and its nested-Monte-Carlo ATM pricer  --  TEST INFRASTRUCTURE ONLY.

Restates, in NumPy float64, the reference's simulator ``src/sim/rbergomi_sim.py``:

  * per-path parameter perturbation and clips            :363-367 (constants :29-40)     -> ``perturb_params``
  * rbergomi_lambda_gpu / rbergomi_phi_gpu                :206-215                        -> ``lam``, ``phi``
  * fractional_gaussian_gpu                               :217-229                        -> ``fgn_fft`` (the reference's FFT form)
  * forward_variance_gpu                                  :231-243                        -> ``forward_variance``
  * the Brownian increments dW1, dW2 = Re / Im ifft(Z) sqrt(M)  :380-382, :278-280        -> ``brownian_from_Z``
  * price_rbergomi_option_gpu                             :246-306                        -> ``price_option``
  * outer variance path + log-Euler step                  :400-406, :454-464              -> ``outer_paths``

and the identity the CUDA kernels use instead of FFTs (exact, since lambda is real and Z = fft(dW1 + i dW2) / sqrt(M)):

      X_k = sqrt(2H) eta / sqrt(M) * sum_n lambda_n dW1[(k - n) mod M]                    -> ``fgn_conv``

i.e. the reference's "fractional Gaussian" driver is a circular FIR filter of the first Brownian increment stream.

Parity status: PINNED.  ``tests/golden/rbergomi_golden.npz`` (tests/golden/make_golden.py --rbergomi-only) holds inputs,
draws and outputs of the UNMODIFIED reference functions run on the CPU under ``oracle/_cupy_stub``;
tests/test_oracle_rbergomi.py checks every function here against it (and against the reference itself where present).
The reference's own random stream (cuRAND via CuPy) is not reproducible: draws are imported, never re-generated.
"""
from __future__ import annotations

import numpy as np

# rbergomi_sim.py:13-40
R, DT, N_STEPS = 0.04, 1 / 252, 252
T_OPTION_TENOR, N_PATHS_OPTION_MC = 30 / 252, 5000
PERTURB = dict(S0=0.01, xi=0.20, H=0.20, eta=0.20, rho=0.10)
MIN_XI_FACTOR, MIN_ETA_FACTOR = 0.5, 0.5
CLIP_H, CLIP_RHO = (0.01, 0.49), (-0.99, -0.01)


def next_power_of_two(n):
    p = 1
    while p < n:
        p <<= 1
    return p


def perturb_params(base, z):
    """:363-367.  base = (S0, xi, H, eta, rho); z (5, n) standard normals -> per-path arrays (S0, xi, H, eta, rho)."""
    S0b, xib, Hb, etab, rhob = base
    z = np.asarray(z, np.float64)
    S0 = S0b * (1 + PERTURB["S0"] * z[0])
    xi = xib * np.maximum(MIN_XI_FACTOR, 1 + PERTURB["xi"] * z[1])
    H = np.clip(Hb * (1 + PERTURB["H"] * z[2]), *CLIP_H)
    eta = etab * np.maximum(MIN_ETA_FACTOR, 1 + PERTURB["eta"] * z[3])
    rho = np.clip(rhob * (1 + PERTURB["rho"] * z[4]), *CLIP_RHO)
    return S0, xi, H, eta, rho


def lam(t, H):
    """:206-207  lambda[p, k] = 0.5 t_k^(2 H_p)."""
    return 0.5 * (np.asarray(t, np.float64)[None, :] ** (2 * np.asarray(H, np.float64)[:, None]))


def phi(lam_arr):
    """:209-215  FFT of lambda zero-padded to the next power of two."""
    n, nt = lam_arr.shape
    M = next_power_of_two(nt)
    pad = np.zeros((n, M))
    pad[:, :nt] = lam_arr
    return np.fft.fft(pad, axis=1)


def fgn_fft(phi_arr, Z, H, eta, out_len):
    """:217-229 (Z 2-D: (n, M); Z 3-D: (n, paths, M))."""
    H, eta = np.asarray(H, np.float64), np.asarray(eta, np.float64)
    if Z.ndim == 3:
        A = np.fft.ifft(phi_arr[:, None, :] * Z, axis=2).real
        return (np.sqrt(2 * H) * eta)[:, None, None] * A[..., :out_len]
    A = np.fft.ifft(phi_arr * Z, axis=1).real
    return (np.sqrt(2 * H) * eta)[:, None] * A[..., :out_len]


def brownian_from_Z(Z):
    """:380-382 / :278-280  unscaled N(0,1) increments dW1, dW2 from the complex Gaussian array Z (last axis = M)."""
    M = Z.shape[-1]
    w = np.fft.ifft(Z, axis=-1)
    return w.real * np.sqrt(float(M)), w.imag * np.sqrt(float(M))


def fgn_conv(lam_arr, dW1, H, eta, out_len):
    """The FIR form: X_k = sqrt(2H) eta / sqrt(M) sum_n lambda_n dW1[(k - n) mod M]; dW1 (..., M), lam_arr (n, nt)."""
    M = dW1.shape[-1]
    n, nt = lam_arr.shape
    c = np.sqrt(2 * np.asarray(H, np.float64)) * np.asarray(eta, np.float64) / np.sqrt(float(M))
    X = np.zeros(dW1.shape[:-1] + (out_len,))
    for k in range(out_len):
        idx = (k - np.arange(nt)) % M
        if dW1.ndim == 3:
            X[..., k] = np.einsum("pn,pqn->pq", lam_arr, dW1[..., idx])
        else:
            X[..., k] = (lam_arr * dW1[..., idx]).sum(-1)
    return X * (c[:, None, None] if dW1.ndim == 3 else c[:, None])


def forward_variance(X, t, xi, H, eta):
    """:231-243  v = xi exp(X - 0.5 eta^2 t^(2H))."""
    xi, H, eta = (np.asarray(a, np.float64) for a in (xi, H, eta))
    ma = -0.5 * (eta * eta)[:, None] * (np.asarray(t, np.float64)[None, :] ** (2 * H[:, None]))     # (n, nt)
    if X.ndim == 3:
        return xi[:, None, None] * np.exp(X + ma[:, None, :])
    return xi[:, None] * np.exp(X + ma)


def euler_terminal(S0, v, dW1, dW2, rho, r, dt, n_steps):
    """:285-295  inner log-Euler loop; S0, rho (n,), v / dW (n, paths, >= n_steps) -> terminal prices (n, paths)."""
    S = np.repeat(np.asarray(S0, np.float64)[:, None], v.shape[1], axis=1)
    rho = np.asarray(rho, np.float64)[:, None]
    sq = np.sqrt(dt)
    for j in range(1, n_steps + 1):
        dW = rho * (sq * dW1[..., j - 1]) + np.sqrt(np.maximum(0.0, 1.0 - rho * rho)) * (sq * dW2[..., j - 1])
        vt = v[..., j - 1]
        S = np.maximum(S * np.exp((r - 0.5 * vt) * dt + np.sqrt(np.maximum(0.0, vt)) * dW), 1e-8)
    return S


def price_option(S0, K, T_opt, r, xi, H, eta, rho, option_type, Z, dt):
    """:246-306 with the complex draws ``Z`` (n, n_mc, M) supplied by the caller instead of ``cp.random.normal``."""
    S0, K = np.asarray(S0, np.float64), np.asarray(K, np.float64)
    n_steps = int(T_opt / dt)
    if n_steps <= 0:
        pay = np.maximum(S0 - K, 0.0) if option_type == "call" else np.maximum(K - S0, 0.0)
        return pay * np.exp(-r * T_opt)
    t = np.linspace(0, n_steps * dt, n_steps + 1)
    la = lam(t, H)
    X = fgn_fft(phi(la), Z, H, eta, n_steps + 1)
    v = forward_variance(X, t, xi, H, eta)
    dW1, dW2 = brownian_from_Z(Z)
    ST = euler_terminal(S0, v, dW1, dW2, rho, r, dt, n_steps)
    pay = np.maximum(ST - K[:, None], 0.0) if option_type == "call" else np.maximum(K[:, None] - ST, 0.0)
    return pay.mean(axis=1) * np.exp(-r * T_opt)


def price_option_from_increments(S0, K, T_opt, r, xi, H, eta, rho, option_type, dW1, dW2, dt):
    """The same price from the Brownian increments (n, n_mc, M) through the FIR identity: what the CUDA kernel computes."""
    n_steps = int(T_opt / dt)
    t = np.linspace(0, n_steps * dt, n_steps + 1)
    la = lam(t, H)
    X = fgn_conv(la, dW1, H, eta, n_steps + 1)
    v = forward_variance(X, t, xi, H, eta)
    ST = euler_terminal(S0, v, dW1, dW2, rho, r, dt, n_steps)
    K = np.asarray(K, np.float64)
    pay = np.maximum(ST - K[:, None], 0.0) if option_type == "call" else np.maximum(K[:, None] - ST, 0.0)
    return pay.mean(axis=1) * np.exp(-r * T_opt)


def outer_paths(S0, xi, H, eta, rho, dW1, dW2, r=R, dt=DT, n_steps=N_STEPS):
    """:400-406 + :454-464 from the main Brownian increments (n, M >= n_steps + 1) -> paths, v, each (n, n_steps + 1)."""
    t = np.linspace(0, n_steps * dt, n_steps + 1)
    la = lam(t, H)
    X = fgn_conv(la, dW1, H, eta, n_steps + 1)
    v = forward_variance(X, t, xi, H, eta)
    paths = np.zeros_like(v)
    paths[:, 0] = S0
    rho = np.asarray(rho, np.float64)
    sq = np.sqrt(dt)
    for j in range(1, n_steps + 1):
        dW = rho * (sq * dW1[:, j - 1]) + np.sqrt(np.maximum(0.0, 1.0 - rho * rho)) * (sq * dW2[:, j - 1])
        vt = v[:, j - 1]
        paths[:, j] = np.maximum(paths[:, j - 1] * np.exp((r - 0.5 * vt) * dt + np.sqrt(np.maximum(0.0, vt)) * dW), 1e-8)
    return paths, v
