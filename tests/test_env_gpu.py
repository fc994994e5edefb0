"""GPU parity tests of the fused hedge step (K3) against the golden vectors and the CPU oracle.

Everything goes through the C ABI (ctypes -> libcantor_hedge.so).  Tolerances (BASELINE.json north_star):
integers / done flags bit-exact; floats <= 1e-6 relative in the fp64 ledger, <= 1e-4 in fp32.
"""
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_env_case
from oracle.hedge_oracle import INFO_FLOAT_KEYS, INFO_INT_KEYS, EnvParams, OracleVecEnv

pytestmark = pytest.mark.gpu

RTOL = {"fp64": 1e-6, "fp32": 1e-4}
# observation entries are O(1) float32 numbers produced through log / erfc: absolute floor of a few float32 ulps
OBS_ATOL = {"fp64": 2e-7, "fp32": 2e-6}


def _vec_kwargs(kwargs, is_v1):
    kw = dict(kwargs)
    if is_v1:
        kw.pop("transaction_cost_per_contract", None)
    return kw


def _check_floats(got, want, prec, what, atol=0.0):
    np.testing.assert_allclose(np.asarray(got, np.float64), want, rtol=RTOL[prec], atol=atol, err_msg=what)


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
def test_step_matches_reference_golden(env_case, prec):
    """Same paths, same actions, same seeds as N unmodified reference envs -> same everything, step by step."""
    from cantorrl_b200 import HedgingVecEnv
    from cantorrl_b200._lib import INFO_F64_KEYS, INFO_I32_KEYS
    name, z, kwargs, is_v1 = env_case
    assert tuple(INFO_F64_KEYS) == tuple(INFO_FLOAT_KEYS) and tuple(INFO_I32_KEYS) == tuple(INFO_INT_KEYS)
    n = z["actions"].shape[1]
    data = {k: z[k] for k in ("paths", "volatilities", "call_prices_atm", "put_prices_atm")}
    env = HedgingVecEnv(data=data, num_envs=n, precision=prec, version="v1" if is_v1 else "v2",
                        episode_sampler="pcg64", record_info=True, **_vec_kwargs(kwargs, is_v1))
    obs = env.reset(seed=int(z["seeds"][0]))                 # env i is seeded seeds[0] + i like the golden run
    np.testing.assert_array_equal(env.current_episode_idx.cpu().numpy(), z["episode_idx"][0])   # PCG64 draws
    np.testing.assert_allclose(obs.cpu().numpy(), z["reset_obs"][0], rtol=RTOL[prec], atol=OBS_ATOL[prec])
    ep = np.zeros(n, int)
    scale = float(np.nanmax(np.abs(z["info_f"][:, :, 10])))  # portfolio value scale: absolute floor for P&L-like keys
    for t in range(z["actions"].shape[0]):
        obs, reward, done, info = env.step(z["actions"][t])
        d = done.cpu().numpy()
        np.testing.assert_array_equal(d, z["terminated"][t], err_msg=f"done, step {t}")
        for j, k in enumerate(INFO_INT_KEYS):
            np.testing.assert_array_equal(info[k].cpu().numpy(), z["info_i"][t, :, j], err_msg=f"{k}, step {t}")
        _check_floats(reward.cpu().numpy(), z["reward"][t], prec, f"reward, step {t}",
                      atol=0.0 if prec == "fp64" else 1e-7)
        for j, k in enumerate(INFO_FLOAT_KEYS):
            ref = z["info_f"][t, :, j]
            if is_v1 and np.isnan(ref).all():
                continue
            # fp32 keeps cash in float32, so cash / portfolio value carry its rounding (<= 1e-7 * |value|)
            atol = 0.0 if prec == "fp64" else 2e-7 * scale
            if k in ("step_pnl_total", "per_share_step_pnl", "raw_pnl_deviation_abs", "reward_pnl_component") and prec == "fp32":
                atol = 1e-8 * scale
            _check_floats(info[k].cpu().numpy(), ref, prec, f"{k}, step {t}", atol=atol)
        o = obs.cpu().numpy()
        term_obs = info["terminal_observation"].cpu().numpy()
        live = ~d
        np.testing.assert_allclose(o[live], z["obs"][t][live], rtol=RTOL[prec], atol=OBS_ATOL[prec], err_msg=f"obs, step {t}")
        if d.any():
            np.testing.assert_allclose(term_obs[d], z["obs"][t][d], rtol=RTOL[prec], atol=OBS_ATOL[prec],
                                       err_msg=f"terminal obs, step {t}")
            ep[d] += 1
            if ep.max() >= z["episode_idx"].shape[0]:
                break
            want_idx = z["episode_idx"][ep, np.arange(n)]
            np.testing.assert_array_equal(env.current_episode_idx.cpu().numpy()[d], want_idx[d])
            np.testing.assert_allclose(o[d], z["reset_obs"][ep[d], np.nonzero(d)[0]], rtol=RTOL[prec], atol=OBS_ATOL[prec],
                                       err_msg=f"auto-reset obs, step {t}")


def test_fp64_reward_and_pnl_bit_exact_on_training_case():
    """The float64 reward path has no transcendental: the kernel reproduces the reference bit for bit."""
    from cantorrl_b200 import HedgingVecEnv
    z, kwargs, _ = load_env_case("v2_train")
    n = z["actions"].shape[1]
    data = {k: z[k] for k in ("paths", "volatilities", "call_prices_atm", "put_prices_atm")}
    env = HedgingVecEnv(data=data, num_envs=n, precision="fp64", record_info=True, **kwargs)
    env.reset(seed=int(z["seeds"][0]))
    for t in range(300):
        _, reward, _, info = env.step(z["actions"][t])
        assert np.array_equal(reward.cpu().numpy(), z["reward"][t]), t
        assert np.array_equal(info["step_pnl_total"].cpu().numpy(), z["info_f"][t, :, 0]), t
        assert np.array_equal(info["cash"].cpu().numpy(), z["info_f"][t, :, 11]), t


def _random_book(rng, n_paths, T, s0=100.0):
    S = s0 * np.exp(np.cumsum(rng.normal(0, 0.02, (n_paths, T + 1)), axis=1))
    V = np.abs(rng.normal(0.04, 0.02, (n_paths, T + 1)))
    Cc = np.abs(rng.normal(3, 1, (n_paths, T)))
    Pp = np.abs(rng.normal(3, 1, (n_paths, T)))
    return S, V, Cc, Pp


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("n_envs", [1, 3, 255, 256, 1000, 1001, 4099])
def test_step_matches_oracle_ragged_sizes(prec, n_envs):
    """Sizes around the 256-env CTA tile and the 4-row TMA granule; envs gather arbitrary paths."""
    from cantorrl_b200 import HedgingVecEnv
    rng = np.random.default_rng(n_envs)
    T, n_paths = 12, 97
    S, V, Cc, Pp = _random_book(rng, n_paths, T)
    kw = dict(slippage_bps=1.5, theta_weight=3e-4, pnl_penalty_weight=0.02, lambda_cost=0.5, initial_cash=250.0)
    orc = OracleVecEnv(S, V, Cc, Pp, EnvParams(**kw), n_envs)
    env = HedgingVecEnv(data=dict(paths=S, volatilities=V, call_prices_atm=Cc, put_prices_atm=Pp), num_envs=n_envs,
                        precision=prec, episode_sampler="pcg64", **kw)
    idx = rng.integers(n_paths, size=n_envs)
    o_ref = orc.reset(idx)
    o = env.reset(path_idx=idx)
    np.testing.assert_allclose(o.cpu().numpy(), o_ref, rtol=RTOL[prec], atol=OBS_ATOL[prec])
    for t in range(2 * T + 3):
        a = rng.uniform(-1, 1, (n_envs, 2)).astype(np.float32)
        nxt = rng.integers(n_paths, size=n_envs)
        env.set_next_paths(nxt)
        o_ref, r_ref, d_ref, term_ref, _ = orc.step_autoreset(a, nxt)
        o, r, d, info = env.step(a)
        np.testing.assert_array_equal(d.cpu().numpy(), d_ref)
        np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=RTOL[prec], atol=0 if prec == "fp64" else 1e-7)
        np.testing.assert_allclose(o.cpu().numpy(), o_ref, rtol=RTOL[prec], atol=OBS_ATOL[prec])
        np.testing.assert_array_equal(env.call_contracts_held.cpu().numpy(), orc.pos_c)
        np.testing.assert_array_equal(env.put_contracts_held.cpu().numpy(), orc.pos_p)
        np.testing.assert_array_equal(env.current_step.cpu().numpy(), orc.step_count)
        if d_ref.any():
            np.testing.assert_allclose(info["terminal_observation"].cpu().numpy()[d_ref], term_ref[d_ref],
                                       rtol=RTOL[prec], atol=OBS_ATOL[prec])


def test_single_env_adapter_reads_like_the_reference():
    """The gym-signature adapter: construct from an npz path, reset(seed), step, exceptions (hedging_env_v2.py:42-48)."""
    from cantorrl_b200 import HedgingEnv
    z, kwargs, _ = load_env_case("v2_lowprice")
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "schema_a.npz")
        np.savez(f, **{k: z[k] for k in ("paths", "volatilities", "call_prices_atm", "put_prices_atm")})
        env = HedgingEnv(f, **kwargs)
        bad = os.path.join(d, "bad.npz")
        np.savez(bad, paths=z["paths"], volatilities=z["volatilities"], call_prices_atm=z["call_prices_atm"],
                 put_prices_atm=z["put_prices_atm"][:, :-1])
        with pytest.raises(ValueError, match="inconsistent"):
            HedgingEnv(bad)
        with pytest.raises(FileNotFoundError):
            HedgingEnv(os.path.join(d, "missing.npz"))
        schema_b = os.path.join(d, "b.npz")
        np.savez(schema_b, calls=z["paths"], puts=z["paths"])
        with pytest.raises(FileNotFoundError):          # the reference's own schema mismatch
            HedgingEnv(schema_b)
    assert env.episode_length == z["paths"].shape[1] - 1 and env.max_contracts_held == 200
    assert env.action_space.shape == (2,) and env.observation_space.shape == (13,)
    obs, info = env.reset(seed=int(z["seeds"][0]))
    assert info == {} and obs.dtype == np.float32 and obs.shape == (13,)
    assert env.current_episode_idx == z["episode_idx"][0, 0]
    np.testing.assert_allclose(obs, z["reset_obs"][0, 0], rtol=1e-6, atol=2e-7)
    for t in range(env.episode_length):
        obs, reward, terminated, truncated, info = env.step(z["actions"][t, 0])
        assert isinstance(reward, np.float64) and truncated is False and terminated == z["terminated"][t, 0]
        assert reward == z["reward"][t, 0]
        assert info["actual_calls_traded"] == z["info_i"][t, 0, 4]
        np.testing.assert_allclose(obs, z["obs"][t, 0], rtol=1e-6, atol=2e-7)
    assert terminated
    with pytest.raises(IndexError):
        env.step(np.zeros(2, np.float32))


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
def test_full_size_invariants(prec):
    """BASELINE config 2 size (2^20 envs), properties that need no oracle: a no-trade episode's P&L telescopes to the
    float32 stock leg, positions/trades respect their clamps, done fires exactly every T steps."""
    from cantorrl_b200 import HedgingVecEnv, ReplayData
    n, T = 1 << 20, 6
    g = torch.Generator(device="cuda").manual_seed(5)
    S = 100 * torch.exp(torch.cumsum(0.02 * torch.randn((T + 1, n), device="cuda", generator=g), 0)).float()
    v = torch.full((T + 1, n), 0.04, device="cuda")
    Cc = torch.rand((T, n), device="cuda", generator=g) * 5 + 1
    Pp = torch.rand((T, n), device="cuda", generator=g) * 5 + 1
    data = ReplayData.from_time_major(S, v, Cc, Pp)
    env = HedgingVecEnv(data=data, num_envs=n, precision=prec, episode_sampler="same_path", slippage_bps=1.0,
                        pnl_penalty_weight=1.0, lambda_cost=0.0)
    env.reset()
    zero = torch.zeros((n, 2), device="cuda")
    for t in range(T):
        obs, r, d, _ = env.step(zero)
        stock_new = (torch.tensor(10000.0, device="cuda") * S[t + 1]).double()      # float32 product, then widened
        stock_old = (torch.tensor(10000.0, device="cuda") * S[t]).double()
        shares = torch.full_like(stock_new, 10000.0)          # tensor divisor: torch turns `x / python_float` into x * (1/f)
        want = -(torch.abs((stock_new - stock_old) / shares) / torch.clamp(S[0], min=25.0).double())
        if prec == "fp64":
            assert torch.equal(r, want)
        else:
            torch.testing.assert_close(r.double(), want, rtol=1e-6, atol=1e-9)
        assert bool(d.all()) == (t == T - 1) and bool(d.any()) == (t == T - 1)
    assert torch.equal(env.current_step, torch.zeros_like(env.current_step))          # auto-reset happened
    a = torch.rand((n, 2), device="cuda", generator=g) * 4 - 2
    prev_c = env.call_contracts_held.clone()
    cash0 = env.cash_balance.clone()
    for t in range(T + 2):
        env.step(a)
        c = env.call_contracts_held
        fresh = env.current_step == 0                                                # all envs move in lock-step here
        if not bool(fresh.any()):
            assert int((c - prev_c).abs().max()) <= 15
            assert bool((env.cash_balance <= cash0).all())                           # costs are non-negative
        assert int(c.abs().max()) <= 200
        prev_c, cash0 = c.clone(), env.cash_balance.clone()


def test_config4_shard_size_full_episode():
    """BASELINE configs[3] per-GPU shard: 2^23 envs x 252 steps replayed from a 34 GB book simulated in HBM.  Size-independent
    checks: done flags only at the episode end, the unhedged reward in closed form on a sample, auto-reset to step 0."""
    free, _ = torch.cuda.mem_get_info()
    if free < 70 * 2 ** 30:
        pytest.skip("needs ~50 GB of free HBM")
    from cantorrl_b200 import HedgingVecEnv, sim
    n, T = 1 << 23, 252
    book = sim.generate_paths_and_options(n, n_steps=T, model="gbm")
    env = HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", slippage_bps=1.0, theta_weight=2e-4,
                        pnl_penalty_weight=1e-3, lambda_cost=1e-4)
    env.reset()
    zero = torch.zeros((n, 2), device="cuda")
    sample = torch.arange(0, n, 4099, device="cuda")
    S = book.S
    for t in range(T):
        obs, r, d, _ = env.step(zero)
        if t in (0, 100, T - 1):
            s_new = (torch.tensor(10000.0, device="cuda") * S[t + 1, sample]).double()
            s_old = (torch.tensor(10000.0, device="cuda") * S[t, sample]).double()
            want = -(1e-3 * torch.abs((s_new - s_old) / 10000.0) / torch.clamp(S[0, sample], min=25.0).double()) - 2e-4 * (T - t - 1) / 252
            torch.testing.assert_close(r[sample].double(), want, rtol=1e-4, atol=1e-7)
        assert bool(d.all()) == (t == T - 1) and bool(d.any()) == (t == T - 1)
    assert int(env.current_step.max()) == 0 and bool(torch.isfinite(obs).all())


def test_empty_and_zero_length_calls_are_no_ops():
    """n_envs = 0, n_steps = 0, empty day ranges: accepted, nothing launched, nothing written."""
    import ctypes as C
    from cantorrl_b200 import HedgingVecEnv, _lib, sim
    from cantorrl_b200.rollout import HedgingRollout
    book = sim.generate_paths_and_options(8, n_steps=4)
    env = HedgingVecEnv(data=book, num_envs=8, episode_sampler="same_path")
    env.reset()
    L = _lib.lib()
    sentinel = torch.full((8, 13), 7.0, device="cuda")
    rew = torch.zeros(8, device="cuda")
    done = torch.zeros(8, dtype=torch.uint8, device="cuda")
    act = torch.zeros((8, 2), device="cuda")
    rc = L.cantor_env_step(C.byref(env._params), C.byref(env._book), C.byref(env._state), 0, _lib.F32, act.data_ptr(),
                           sentinel.data_ptr(), rew.data_ptr(), done.data_ptr(), None, 1, None, None, None)
    assert rc == 0
    rc = L.cantor_env_step_many(C.byref(env._params), C.byref(env._book), C.byref(env._state), 8, _lib.F32, 0, act.data_ptr(),
                                sentinel.data_ptr(), rew.data_ptr(), done.data_ptr(), None, None, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert bool((sentinel == 7.0).all()) and int(env.current_step.max()) == 0
    res = HedgingRollout(data=book, num_envs=8).run(0, "random")
    assert float(res.stats.sums.abs().sum()) == 0.0
    rb = sim.generate_rbergomi_paths_and_options(4, n_steps=3, n_mc=16, price=False)
    rb.price_days(2, 2)
    assert float(rb.book.C.abs().sum()) == 0.0
