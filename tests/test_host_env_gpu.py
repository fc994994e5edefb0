"""GPU parity of the host-buffer C ABI (cantor_vecenv_*): NumPy in, NumPy out, against the oracle."""
import os
import tempfile

import numpy as np
import pytest

from conftest import load_env_case
from oracle.hedge_oracle import EnvParams, OracleVecEnv

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,rtol", [("fp64", 1e-6), ("fp32", 1e-4)])
@pytest.mark.parametrize("n_envs,n_chunks", [(5, 0), (1000, 3), (70000, 0)])
def test_host_step_matches_oracle(prec, rtol, n_envs, n_chunks):
    from cantorrl_b200.host_env import HostVecEnv
    rng = np.random.default_rng(n_envs)
    T, n_paths = 9, 211
    S = 100 * np.exp(np.cumsum(rng.normal(0, 0.02, (n_paths, T + 1)), axis=1))
    V = np.abs(rng.normal(0.04, 0.02, (n_paths, T + 1)))
    Cc = np.abs(rng.normal(3, 1, (n_paths, T)))
    Pp = np.abs(rng.normal(3, 1, (n_paths, T)))
    kw = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)
    env = HostVecEnv(data=dict(paths=S, volatilities=V, call_prices_atm=Cc, put_prices_atm=Pp), num_envs=n_envs,
                     precision=prec, episode_sampler="array", n_chunks=n_chunks, **kw)
    orc = OracleVecEnv(S, V, Cc, Pp, EnvParams(**kw), n_envs)
    idx = rng.integers(n_paths, size=n_envs)
    np.testing.assert_allclose(env.reset(idx), orc.reset(idx), rtol=rtol, atol=2e-6)
    assert env.episode_length == T and env.num_episodes == n_paths
    for t in range(T + 3):
        a = rng.uniform(-1, 1, (n_envs, 2)).astype(np.float32)
        nxt = rng.integers(n_paths, size=n_envs)
        o_ref, r_ref, d_ref, _, _ = orc.step_autoreset(a, nxt)
        o, r, d, _ = env.step(a, next_path=nxt)
        assert np.array_equal(d, d_ref)
        np.testing.assert_allclose(r, r_ref, rtol=rtol, atol=1e-7)
        np.testing.assert_allclose(o, o_ref, rtol=rtol, atol=2e-6)
    env.close()


def test_host_env_loads_reference_npz_and_rejects_bad_files():
    from cantorrl_b200.host_env import HostVecEnv
    z, kwargs, _ = load_env_case("v2_train")
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "a.npz")
        np.savez(f, **{k: z[k] for k in ("paths", "volatilities", "call_prices_atm", "put_prices_atm")})
        env = HostVecEnv(f, num_envs=4, precision="fp64", episode_sampler="array", **kwargs)
        with pytest.raises(FileNotFoundError):
            HostVecEnv(os.path.join(d, "nope.npz"), num_envs=4)
        with pytest.raises(ValueError, match="inconsistent"):
            HostVecEnv(data=dict(paths=z["paths"], volatilities=z["volatilities"][:, :-1],
                                 call_prices_atm=z["call_prices_atm"], put_prices_atm=z["put_prices_atm"]), num_envs=4)
    obs = env.reset(z["episode_idx"][0])
    np.testing.assert_allclose(obs, z["reset_obs"][0], rtol=1e-6, atol=2e-7)
    for t in range(100):
        obs, r, d, _ = env.step(z["actions"][t])
        assert np.array_equal(r, z["reward"][t])           # float64 ledger: bit-exact rewards through the host ABI too
        np.testing.assert_allclose(obs, z["obs"][t], rtol=1e-6, atol=2e-7)


def test_simulated_book_on_the_host_path():
    from cantorrl_b200.host_env import HostVecEnv
    env = HostVecEnv(num_envs=4096, simulate=dict(num_paths=4096, n_steps=30, model="heston", seed=7), slippage_bps=1.0)
    obs = env.reset()
    assert obs.shape == (4096, 13) and np.isfinite(obs).all() and np.allclose(obs[:, 0], 1.0)
    tot = 0
    for t in range(30):
        obs, r, d, _ = env.step(np.full((4096, 2), 0.5, np.float32))
        tot += int(d.sum())
        assert np.isfinite(r).all() and np.isfinite(obs).all()
    assert tot == 4096 and d.all()


def test_host_env_speaks_the_sb3_vecenv_protocol():
    """The duck-typed surface SB3's algorithms use on a VecEnv (base_vec_env.VecEnv): spaces, step_async / step_wait,
    get_attr("episode_length") (train_ppo_v2.py:463), env_is_wrapped, seed, indexable infos."""
    from cantorrl_b200.host_env import HostVecEnv
    env = HostVecEnv(num_envs=6, simulate=dict(num_paths=6, n_steps=5), slippage_bps=1.0, episode_sampler="philox", seed=3)
    assert env.observation_space.shape == (13,) and env.action_space.shape == (2,) and env.action_space.dtype == np.float32
    assert env.get_attr("episode_length") == [5] * 6 and env.get_attr("episode_length", [0, 2]) == [5, 5]
    assert env.env_is_wrapped(object) == [False] * 6 and env.seed(11) == [11] * 6
    obs = env.reset()
    assert obs.shape == (6, 13) and obs.dtype == np.float32
    for t in range(5):
        env.step_async(np.stack([env.action_space.sample() for _ in range(6)]))
        obs, rew, done, infos = env.step_wait()
        assert len(infos) == 6 and infos[3].get("terminal_observation") is None and infos[-1] == {}
    assert done.all() and done.dtype == np.bool_
    with pytest.raises(AttributeError):
        env.set_attr("lambda_cost", 2.0)
    env.close()


def test_sb3_vecenv_subclass_against_a_stub_base_class():
    """cantorrl_b200.sb3.cantor_vec_env_class over a stand-in for SB3's VecEnv ABC: instantiable (every abstract method
    implemented), base-class constructor fed, step = step_async + step_wait, arrays owned by the caller, infos a list."""
    from cantorrl_b200.host_env import HostVecEnv
    from cantorrl_b200.sb3 import cantor_vec_env_class
    from sb3_stub import VecEnv, VecNormalizeLike
    cls = cantor_vec_env_class(VecEnv)
    assert issubclass(cls, VecEnv) and issubclass(cls, HostVecEnv)
    kw = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)
    sim = dict(num_paths=300, n_steps=6, model="gbm", seed=3)
    venv = cls(num_envs=300, simulate=sim, **kw)
    plain = HostVecEnv(num_envs=300, simulate=sim, **kw)
    assert venv.num_envs == 300 and venv.observation_space.shape == (13,) and venv.action_space.shape == (2,)
    assert venv.get_attr("episode_length") == [6] * 300
    wrapped = VecNormalizeLike(venv)
    o0 = venv.reset()
    np.testing.assert_array_equal(o0, plain.reset())
    rng = np.random.default_rng(0)
    kept = []
    for t in range(14):
        a = rng.uniform(-1, 1, (300, 2)).astype(np.float32)
        o, r, d, infos = wrapped.step(a)
        po, pr, pd, _ = plain.step(a)
        np.testing.assert_array_equal(o, po)
        np.testing.assert_array_equal(r, pr)
        np.testing.assert_array_equal(d, pd)
        assert bool(d.all()) == ((t + 1) % 6 == 0)
        kept.append((o, po.copy()))
    for o, po in kept:                      # returned arrays are the caller's: later steps did not overwrite them
        np.testing.assert_array_equal(o, po)
    venv.close()
    plain.close()


def test_host_copy_probe_reports_plausible_bandwidths():
    """cantor_host_copy_probe: the raw copies of one host-buffer step, no kernel; argument validation on the host."""
    from cantorrl_b200 import _lib
    from cantorrl_b200.host_env import host_copy_probe
    r = host_copy_probe(device=0, d2h_bytes=57 << 18, h2d_bytes=8 << 18, n_chunks=4, sync_each_round=True, seconds=0.1)
    assert 1.0 < r["d2h_gbs"] < 200.0 and 0.1 < r["h2d_gbs"] < 200.0 and r["rounds_per_s"] > 10
    assert abs(r["d2h_gbs"] / r["h2d_gbs"] - 57 / 8) < 1e-6
    one_way = host_copy_probe(device=0, d2h_bytes=1 << 24, h2d_bytes=0, n_chunks=1, sync_each_round=False, seconds=0.05)
    assert one_way["h2d_gbs"] == 0.0 and one_way["d2h_gbs"] > 1.0
    with pytest.raises(_lib.CantorError):
        host_copy_probe(device=0, d2h_bytes=0, h2d_bytes=0)
    with pytest.raises(_lib.CantorError):
        host_copy_probe(device=99, d2h_bytes=1 << 20, h2d_bytes=0)
