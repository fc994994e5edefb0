"""CPU ORACLE for path simulation  --  TEST INFRASTRUCTURE ONLY (never imported by the product package).

Restates, in NumPy:

  * the outer log-Euler stock step of the reference simulator, ``src/sim/rbergomi_sim.py:454-464``
    (``euler_from_normals``; float64 like the reference, fed with exported normals and variances);
  * the counter-based generator the CUDA simulator uses, Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
    "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 1.09 reference implementation) and the
    Box-Muller transform on top of it (``philox4x32_10``, ``philox_normals``);
  * the GBM and Heston (full-truncation Euler) dynamics named by BASELINE.json's north_star, with the
    reference's step form (``S *= exp(drift + diffusion)``, floor 1e-8) (``gbm_paths``, ``heston_paths``).

Parity status:
  * ``philox4x32_10`` is PINNED by the Random123 known-answer vectors (tests/test_oracle_sim.py).
  * ``euler_from_normals`` is PINNED against the unmodified ``generate_paths_and_options`` run on CPU under a
    cupy stand-in in the build container (tests/test_oracle_sim.py, skipped where /root/reference is absent)
    and by tests/golden/outer_euler_golden.npz produced from that run.
  * The reference's own random stream (cuRAND XORWOW through CuPy, seed 42) is NOT reproducible here:
    "parity unpinned" for the draws themselves -- parity runs import the reference's normals instead.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
U64 = np.uint64
M0, M1 = U64(0xD2511F53), U64(0xCD9E8D57)
W0, W1 = U32(0x9E3779B9), U32(0xBB67AE85)
MASK = U64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32.  Ten rounds, key bumped by the Weyl constants."""
    c = np.array(np.broadcast_arrays(np.asarray(counter, U32))[0], dtype=U32, copy=True)
    k = np.array(np.asarray(key, U32), dtype=U32, copy=True)
    k = np.broadcast_to(k, c.shape[:-1] + (2,)).copy()
    c0, c1, c2, c3 = (c[..., i].astype(U64) for i in range(4))
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            hi0, lo0 = p0 >> U64(32), p0 & MASK
            hi1, lo1 = p1 >> U64(32), p1 & MASK
            n0 = hi1 ^ c1 ^ k0.astype(U64)
            n2 = hi0 ^ c3 ^ k1.astype(U64)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = (k0 + W0).astype(U32)
            k1 = (k1 + W1).astype(U32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(U32)


STREAM_PATHS = 0x50415448      # "PATH": 4th counter word of the path-simulation stream


def philox_normals(seed, path_index, n_steps, normals_per_step, dtype=np.float32):
    """The normals the CUDA simulator draws for the given global path indices (exact libm here; the kernel evaluates
    log / sqrt / sin / cos with single MUFU instructions, ~1e-6 absolute on a draw -- inside the 1e-4 path tolerance).

    Layout (cantorrl_b200/csrc/path_sim.cu): call ``c`` of path ``p`` has counter
    ``(p_lo, p_hi, c, STREAM_PATHS)`` and key ``(seed_lo, seed_hi)``; its four words give two Box-Muller pairs
    ``(n0, n1) = f(x0, x1)``, ``(n2, n3) = f(x2, x3)``; step ``t`` uses normals ``t*nps .. t*nps + nps - 1`` of
    the concatenated sequence.  Returns (n_paths, n_steps, normals_per_step).
    """
    p = np.asarray(path_index, dtype=np.uint64).ravel()
    total = n_steps * normals_per_step
    n_calls = (total + 3) // 4
    ctr = np.zeros((p.size, n_calls, 4), U32)
    ctr[..., 0] = (p & MASK).astype(U32)[:, None]
    ctr[..., 1] = (p >> U64(32)).astype(U32)[:, None]
    ctr[..., 2] = np.arange(n_calls, dtype=U32)[None, :]
    ctr[..., 3] = STREAM_PATHS
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], U32)
    x = philox4x32_10(ctr, key)                                   # (P, n_calls, 4)
    ft = np.dtype(dtype).type
    two_m32 = ft(2.0 ** -32)
    xf = x.astype(dtype)                                            # round-to-nearest like cvt.rn
    u1 = (xf[..., 0::2] + ft(1.0)) * two_m32                        # (0, 1]
    u2 = xf[..., 1::2] * two_m32                                    # [0, 1]
    rad = np.sqrt(ft(-2.0) * np.log(u1))
    ang = ft(2.0 * np.pi) * u2
    n_even, n_odd = rad * np.cos(ang), rad * np.sin(ang)
    z = np.stack([n_even, n_odd], axis=-1).reshape(p.size, n_calls * 4)[:, :total]
    return z.reshape(p.size, n_steps, normals_per_step).astype(dtype)


def euler_from_normals(S0, v, dW1, dW2, rho, r=0.04, dt=1 / 252):
    """rbergomi_sim.py:454-464 in float64.

    S0 (n,), v (n, T+1) variances, dW1/dW2 (n, >=T) *unscaled* N(0,1) draws, rho (n,) or scalar.
    Returns paths (n, T+1).
    """
    S0 = np.asarray(S0, np.float64)
    v = np.asarray(v, np.float64)
    n, t1 = v.shape
    rho = np.broadcast_to(np.asarray(rho, np.float64), (n,))
    paths = np.zeros((n, t1))
    paths[:, 0] = S0
    sqrt_dt = np.sqrt(dt)
    for j in range(1, t1):
        dw1 = sqrt_dt * dW1[:, j - 1]
        dw2 = sqrt_dt * dW2[:, j - 1]
        dW = rho * dw1 + np.sqrt(np.maximum(0.0, 1.0 - rho * rho)) * dw2          # :457
        vt = v[:, j - 1]
        drift = (r - 0.5 * vt) * dt                                               # :460
        diff = np.sqrt(np.maximum(0.0, vt)) * dW                                  # :461
        paths[:, j] = np.maximum(paths[:, j - 1] * np.exp(drift + diff), 1e-8)    # :463-464
    return paths


def gbm_paths(seed, path_index, n_steps, s0=100.0, sigma2=0.04, r=0.04, dt=1 / 252, dtype=np.float32):
    """Constant-variance special case of the reference step, Philox normals (one per step)."""
    ft = np.dtype(dtype).type
    z = philox_normals(seed, path_index, n_steps, 1, dtype)[..., 0]
    n = z.shape[0]
    S = np.empty((n, n_steps + 1), dtype)
    S[:, 0] = ft(s0)
    drift = ft((r - 0.5 * sigma2) * dt)
    vol = ft(np.sqrt(sigma2 * dt))
    for t in range(n_steps):
        S[:, t + 1] = np.maximum(S[:, t] * np.exp(drift + vol * z[:, t]), ft(1e-8))
    V = np.full_like(S, ft(sigma2))
    return S, V


def heston_paths(seed, path_index, n_steps, s0=100.0, v0=0.04, kappa=2.0, theta=0.04, sigma_v=0.5, rho=-0.7,
                 r=0.04, dt=1 / 252, dtype=np.float32):
    """Heston, full-truncation Euler; correlation built as rbergomi_sim.py:457."""
    ft = np.dtype(dtype).type
    z = philox_normals(seed, path_index, n_steps, 2, dtype)
    n = z.shape[0]
    S = np.empty((n, n_steps + 1), dtype)
    V = np.empty((n, n_steps + 1), dtype)
    S[:, 0] = ft(s0)
    V[:, 0] = ft(v0)
    v = np.full(n, ft(v0), dtype)
    rho_c = ft(np.sqrt(max(0.0, 1.0 - rho * rho)))
    for t in range(n_steps):
        vp = np.maximum(v, ft(0.0))
        sq = np.sqrt(vp * ft(dt))
        z1 = z[:, t, 0]
        zv = ft(rho) * z1 + rho_c * z[:, t, 1]
        S[:, t + 1] = np.maximum(S[:, t] * np.exp((ft(r) - ft(0.5) * vp) * ft(dt) + sq * z1), ft(1e-8))
        v = v + ft(kappa) * (ft(theta) - vp) * ft(dt) + ft(sigma_v) * sq * zv
        V[:, t + 1] = np.maximum(v, ft(0.0))
    return S, V
