"""Runs the UNMODIFIED reference files staged in ``oracle/_ref/`` on the host CPU (TEST / BENCH INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs import this.  The reference classes are loaded
from ``oracle/_ref/*.py`` (byte-for-byte copies made by ``oracle/stage_ref.py``; their sha256 is re-checked against the
manifest) with ``oracle/_gym_stub`` standing in for gymnasium -- so what is timed is the reference's own
``HedgingEnv.step`` (src/env/hedging_env_v2.py:175-294), its own ``black_scholes_vectorized``
(src/sim/option_price_assignment.py:10-21) and its own ``bs_delta_hedge`` (src/tools/bs_delta.py:36-55).

Workload of the env baseline = BASELINE.md section 3: env-schema npz of 4096 GBM paths x 253 (r = 0.04, dt = 1/252, S0 = 100,
variance 0.04, ATM Black-Scholes marks K = round(S_t), tenor 30/252, seed 42), ``record_metrics=True``, the v2 training
keywords (src/agents/train_ppo_v2.py:74-80), pre-generated uniform(-1, 1) float32 actions, ``reset()`` on ``terminated``.
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
GYM_STUB = os.path.join(HERE, "_gym_stub")
R, DT, S0, XI = 0.04, 1 / 252, 100.0, 0.04                                      # rbergomi_sim.py:13,14,27,23
TRAIN_KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4,
                transaction_cost_per_contract=0.65, loss_type="abs")            # train_ppo_v2.py:74-80


def staged():
    """True when every staged file is present and matches the manifest's sha256 (i.e. is still the unmodified reference)."""
    try:
        manifest = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))
        for name, m in manifest.items():
            with open(os.path.join(REF_DIR, name), "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != m["sha256"]:
                    return False
        return len(manifest) >= 4
    except (OSError, ValueError, KeyError):
        return False


def _load(name):
    if GYM_STUB not in sys.path:
        sys.path.insert(0, GYM_STUB)
    path = os.path.join(REF_DIR, name)
    spec = importlib.util.spec_from_file_location("cantor_ref_" + name[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_env_class(version="v2"):
    """The reference's own ``HedgingEnv`` class (v2: hedging_env_v2.py, v1: hedging_env.py)."""
    return _load("hedging_env_v2.py" if version == "v2" else "hedging_env.py").HedgingEnv


def reference_module(name):
    """'option_price_assignment' or 'bs_delta' as staged."""
    return _load(name + ".py")


def gbm_env_schema(n_paths=4096, T=252, seed=42):
    """Env-schema arrays with the GPU config's dynamics: exact log-Euler GBM + ATM Black-Scholes."""
    from oracle import bs_oracle
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n_paths, T))
    logS = np.cumsum((R - 0.5 * XI) * DT + np.sqrt(XI * DT) * z, axis=1)
    S = S0 * np.exp(np.concatenate([np.zeros((n_paths, 1)), logS], axis=1))
    V = np.full_like(S, XI)
    Cc, Pp = bs_oracle.atm_book(S, V)
    return S, V, Cc, Pp


def write_env_schema_npz(path, n_paths=4096, T=252, seed=42):
    S, V, Cc, Pp = gbm_env_schema(n_paths, T, seed)
    np.savez(path, paths=S, volatilities=V, call_prices_atm=Cc, put_prices_atm=Pp)
    return path


def step_reference_env(npz_path, seed, *, seconds=None, env_steps=None, version="v2", one_call_only=True, kw=None):
    """Step ONE reference env with pre-generated uniform float32 actions, resetting on ``terminated``.

    Runs for ``seconds`` of wall time or exactly ``env_steps`` steps.  Returns (env_steps done, elapsed seconds, checksum).
    """
    kw = dict(TRAIN_KW if kw is None else kw)
    if version == "v1":
        kw = {k: v for k, v in kw.items() if k not in ("slippage_bps", "theta_weight")}
    env = reference_env_class(version)(npz_path, record_metrics=True, **kw)
    acts = np.random.default_rng(seed).uniform(-1, 1, (4096, 2)).astype(np.float32)
    if one_call_only:
        acts[:, 1] = 0.0                                 # configs[1]: one European call, the put leg is never traded
    env.reset(seed=seed)
    n, acc, t0 = 0, 0.0, time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        while True:
            for a in acts:
                _, r, term, _, _ = env.step(a)
                acc += float(r)
                n += 1
                if term:
                    env.reset()
                if env_steps is not None and n >= env_steps:
                    return n, time.perf_counter() - t0, acc
            el = time.perf_counter() - t0
            if seconds is not None and el >= seconds:
                return n, el, acc


# ---- side baselines --------------------------------------------------------------------
def time_black_scholes_vectorized(n=1 << 20, repeats=3, seed=0):
    """The reference's black_scholes_vectorized on n elements; returns (repricings per second, seconds per call)."""
    bs = reference_module("option_price_assignment").black_scholes_vectorized
    rng = np.random.default_rng(seed)
    S = 100 * np.exp(rng.normal(0, 0.2, n))
    K = np.round(S * np.exp(rng.normal(0, 0.05, n)))
    sigma = np.abs(rng.normal(0.2, 0.05, n))
    bs(S[:1024], K[:1024], 0.5, R, sigma[:1024])
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        bs(S, K, 0.5, R, sigma)
        best = min(best, time.perf_counter() - t0)
    return n / best, best


def time_bs_delta_hedge(n_paths=16, T=252, seed=0):
    """The reference's bs_delta_hedge on n_paths GBM paths x (T + 1); returns (path-steps per second, seconds)."""
    fn = reference_module("bs_delta").bs_delta_hedge
    S = gbm_env_schema(n_paths, T, seed)[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        fn(S)
        el = time.perf_counter() - t0
    return n_paths * (T + 1) / el, el


def time_numpy_outer_step(n=1 << 20, steps=8, seed=0):
    """NumPy restatement of the simulator's outer log-Euler step (rbergomi_sim.py:454-464; the file itself needs CuPy and a GPU,
    so this one is a port): float64, n paths, `steps` consecutive days.  Returns (path-steps per second, seconds)."""
    rng = np.random.default_rng(seed)
    S = np.full(n, S0)
    v = np.full(n, XI)
    rho = np.full(n, -0.7)
    sqrt_dt = np.sqrt(DT)
    z1 = rng.standard_normal((steps, n))
    z2 = rng.standard_normal((steps, n))
    t0 = time.perf_counter()
    for j in range(steps):
        dw1, dw2 = sqrt_dt * z1[j], sqrt_dt * z2[j]
        dW = rho * dw1 + np.sqrt(np.maximum(0.0, 1.0 - rho * rho)) * dw2
        drift = (R - 0.5 * v) * DT
        diff = np.sqrt(np.maximum(0.0, v)) * dW
        S = np.maximum(S * np.exp(drift + diff), 1e-8)
    el = time.perf_counter() - t0
    return n * steps / el, el


def tmp_npz(n_paths=4096, T=252, seed=42):
    """Write the env-schema file to a fresh temporary directory; returns (directory object to keep alive, path)."""
    d = tempfile.TemporaryDirectory(prefix="cantor_ref_")
    return d, write_env_schema_npz(os.path.join(d.name, "paths_gbm_env_schema.npz"), n_paths, T, seed)
