"""Pins oracle/hedge_oracle.py: golden vectors made by the unmodified reference, and the live reference when present."""
import importlib.util
import os
import sys
import tempfile
import warnings

import numpy as np
import pytest

from conftest import REFERENCE, ROOT
from oracle.hedge_oracle import INFO_FLOAT_KEYS, INFO_INT_KEYS, EnvParams, OracleVecEnv

# observations go through float32 log / float64 erf; NumPy's SIMD float32 log may differ by an ulp between CPUs
OBS_RTOL, OBS_ATOL = 2e-6, 1e-7


def replay_case(z, kwargs):
    """Drive the oracle with the golden actions / episode indices; yield per-step outputs."""
    p = EnvParams(**kwargs)
    env = OracleVecEnv(z["paths"], z["volatilities"], z["call_prices_atm"], z["put_prices_atm"], p, n_envs=z["actions"].shape[1])
    ep = np.zeros(env.n, int)
    obs0 = env.reset(z["episode_idx"][0])
    yield "reset", obs0
    for t in range(z["actions"].shape[0]):
        obs, reward, done, info = env.step(z["actions"][t])
        yield t, obs, reward, done, info
        if done.any():
            ep[done] += 1
            if ep.max() >= z["episode_idx"].shape[0]:
                return
            nxt = z["episode_idx"][ep, np.arange(env.n)]
            robs = env.reset(nxt, mask=done)
            np.testing.assert_allclose(robs[done], z["reset_obs"][ep[done], np.nonzero(done)[0]], rtol=OBS_RTOL, atol=OBS_ATOL)


def test_oracle_matches_golden(env_case):
    name, z, kwargs, is_v1 = env_case
    it = replay_case(z, kwargs)
    _, obs0 = next(it)
    np.testing.assert_allclose(obs0, z["reset_obs"][0], rtol=OBS_RTOL, atol=OBS_ATOL)
    n_checked = 0
    for t, obs, reward, done, info in it:
        np.testing.assert_array_equal(done, z["terminated"][t])                      # bit-exact flags
        for j, k in enumerate(INFO_INT_KEYS):
            np.testing.assert_array_equal(info[k], z["info_i"][t, :, j], err_msg=f"{k} step {t}")
        np.testing.assert_allclose(reward, z["reward"][t], rtol=1e-12, atol=0, err_msg=f"reward step {t}")
        for j, k in enumerate(INFO_FLOAT_KEYS):
            ref = z["info_f"][t, :, j]
            if np.isnan(ref).all() and is_v1:
                continue                                                            # key absent in v1
            np.testing.assert_allclose(np.asarray(info[k], np.float64), ref, rtol=1e-12, atol=1e-12, err_msg=f"{k} step {t}")
        np.testing.assert_allclose(obs, z["obs"][t], rtol=OBS_RTOL, atol=OBS_ATOL, err_msg=f"obs step {t}")
        n_checked += 1
    assert n_checked >= z["actions"].shape[0] - 60


def test_oracle_reward_bit_exact_on_golden():
    """On the training case the float64 reward path has no transcendental: the oracle must equal the reference bit for bit."""
    from conftest import load_env_case
    z, kwargs, _ = load_env_case("v2_train")
    for out in replay_case(z, kwargs):
        if out[0] == "reset":
            continue
        t, obs, reward, done, info = out
        assert np.array_equal(reward, z["reward"][t])
        assert np.array_equal(info["step_pnl_total"], z["info_f"][t, :, 0])


def test_step_after_termination_raises(env_case):
    name, z, kwargs, _ = env_case
    env = OracleVecEnv(z["paths"], z["volatilities"], z["call_prices_atm"], z["put_prices_atm"], EnvParams(**kwargs), 2)
    env.reset([0, 1])
    a = np.zeros((2, 2), np.float32)
    for _ in range(env.episode_length):
        env.step(a)
    with pytest.raises(IndexError):        # the reference raises IndexError too
        env.step(a)


def test_shape_validation():
    with pytest.raises(ValueError, match="inconsistent"):
        OracleVecEnv(np.ones((3, 5)), np.ones((3, 5)), np.ones((3, 5)), np.ones((3, 4)))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("version", ["v1", "v2"])
def test_oracle_matches_live_reference(version):
    """Run the unmodified reference class next to the oracle on fresh random data (build container only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_gym_stub"))
    fn = "hedging_env.py" if version == "v1" else "hedging_env_v2.py"
    spec = importlib.util.spec_from_file_location(f"ref_{version}", f"{REFERENCE}/src/env/{fn}")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(7)
    n_paths, T = 12, 30
    S = 80 * np.exp(np.cumsum(rng.normal(0, 0.02, (n_paths, T + 1)), axis=1))
    V = np.abs(rng.normal(0.04, 0.02, (n_paths, T + 1)))
    C = np.abs(rng.normal(3, 1, (n_paths, T)))
    P = np.abs(rng.normal(3, 1, (n_paths, T)))
    kwargs = dict(pnl_penalty_weight=0.3, lambda_cost=0.7, loss_type="mse" if version == "v2" else "abs")
    if version == "v2":
        kwargs.update(slippage_bps=3.0, theta_weight=5e-4)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "a.npz")
        np.savez(f, paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P)
        ref = mod.HedgingEnv(f, **kwargs)
    if version == "v1":
        kwargs["transaction_cost_per_contract"] = 0.05
    orc = OracleVecEnv(S, V, C, P, EnvParams(**kwargs), 1)
    o_ref, _ = ref.reset(seed=3)
    o = orc.reset([ref.current_episode_idx])
    np.testing.assert_array_equal(o[0], o_ref)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(4 * T):
            a = rng.uniform(-1.3, 1.3, 2).astype(np.float32)
            o_ref, r_ref, te, _, inf = ref.step(a)
            o, r, d, info = orc.step(a[None])
            assert r[0] == r_ref and d[0] == te
            np.testing.assert_array_equal(o[0], o_ref)
            assert info["cash"][0] == inf["cash"] and info["call_contracts"][0] == inf["call_contracts"]
            if te:
                o_ref, _ = ref.reset()
                o = orc.reset([ref.current_episode_idx])
                np.testing.assert_array_equal(o[0], o_ref)


def test_scalar_port_matches_golden():
    """The scalar, reference-shaped port that bench.py times as the CPU baseline computes the golden numbers too."""
    from conftest import load_env_case
    from oracle.hedge_scalar import ScalarEnv
    z, kwargs, _ = load_env_case("v2_mse_saturating")
    env = ScalarEnv(z["paths"], z["volatilities"], z["call_prices_atm"], z["put_prices_atm"], EnvParams(**kwargs),
                    seed=int(z["seeds"][1]))
    obs, _ = env.reset()
    assert env.idx == z["episode_idx"][0, 1]
    np.testing.assert_array_equal(obs, z["reset_obs"][0, 1])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(z["actions"].shape[0]):
            obs, r, term, trunc, info = env.step(z["actions"][t, 1])
            assert r == z["reward"][t, 1] and term == z["terminated"][t, 1] and trunc is False
            np.testing.assert_array_equal(obs, z["obs"][t, 1])
            assert info["call_contracts"] == z["info_i"][t, 1, 0]
            if term:
                env.reset()
