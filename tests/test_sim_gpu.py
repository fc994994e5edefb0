"""GPU parity tests of K1 (path simulation) and K2 (Black-Scholes repricing) against the oracle and golden vectors."""
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import bs_oracle, sim_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["gbm", "heston"])
def test_simulated_paths_match_philox_oracle(model):
    """float32 kernel vs the NumPy float32 restatement on the same Philox counters (fp32 tolerance 1e-4)."""
    from cantorrl_b200 import sim
    n, T, off = 1003, 50, 123456789012
    book = sim.generate_paths_and_options(n, seed=42, n_steps=T, model=model, path_offset=off, reprice=True)
    pm = {k: v.cpu().numpy() for k, v in book.to_path_major().items()}
    idx = off + np.arange(n)
    if model == "gbm":
        S, V = sim_oracle.gbm_paths(42, idx, T)
    else:
        S, V = sim_oracle.heston_paths(42, idx, T)
    np.testing.assert_allclose(pm["paths"], S, rtol=1e-4)
    np.testing.assert_allclose(pm["volatilities"], V, rtol=1e-4, atol=1e-6)
    # option columns: closed-form ATM prices of the kernel's own (S, v), K = round(S), tenor 30/252
    C, P = bs_oracle.atm_book(pm["paths"].astype(np.float64), pm["volatilities"].astype(np.float64))
    np.testing.assert_allclose(pm["call_prices_atm"], C, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pm["put_prices_atm"], P, rtol=1e-4, atol=2e-5)
    # row T of the packed book repeats the marks of row T-1 (stale terminal marks, hedging_env_v2.py:226-231)
    assert torch.equal(book.C[T, :n], book.C[T - 1, :n]) and torch.equal(book.P[T, :n], book.P[T - 1, :n])


def test_paths_do_not_depend_on_sharding():
    """Global path index drives the counter: 1 shard of 4096 == 4 shards of 1024 (bit-exact), any launch geometry."""
    from cantorrl_b200 import sim
    whole = sim.generate_paths_and_options(4096, n_steps=20, model="heston", path_offset=1000).tensor[:, :4096]
    for s in range(4):
        part = sim.generate_paths_and_options(1024, n_steps=20, model="heston", path_offset=1000 + 1024 * s).tensor[:, :1024]
        assert torch.equal(part, whole[:, 1024 * s:1024 * (s + 1)])
    other = sim.generate_paths_and_options(1024, n_steps=20, model="heston", path_offset=1000, seed=43).tensor[:, :1024]
    assert not torch.equal(other, whole[:, :1024])


def test_gbm_distribution_at_full_size():
    """2^20 paths x 252 steps: terminal log-return moments of the exact log-Euler GBM."""
    from cantorrl_b200 import sim
    book = sim.generate_paths_and_options(1 << 20, n_steps=252, model="gbm", reprice=False)
    lr = torch.log(book.S[252, : 1 << 20].double() / 100.0)
    n = float(1 << 20)
    assert abs(float(lr.mean()) - (0.04 - 0.02)) < 4 * 0.2 / n ** 0.5
    assert abs(float(lr.var()) - 0.04) < 0.04 * 4 * (2 / n) ** 0.5
    skew = float(((lr - lr.mean()) ** 3).mean() / lr.std() ** 3)
    assert abs(skew) < 0.02
    assert bool((book.v[:, : 1 << 20] == 0.04).all()) and bool((book.C == 0).all())


def test_outer_euler_step_on_exported_normals():
    """The reference's own float64 step (rbergomi_sim.py:454-464) on the draws of the unmodified simulator."""
    from cantorrl_b200 import sim
    z = np.load(os.path.join(GOLDEN, "outer_euler_golden.npz"))
    paths = sim.euler_from_normals(z["S0"], z["v"], z["dW1"], z["dW2"], z["rho"], float(z["r"]), float(z["dt"]))
    np.testing.assert_allclose(paths.cpu().numpy(), z["paths"], rtol=1e-6)          # north_star fp64 tolerance
    np.testing.assert_allclose(paths.cpu().numpy(), z["paths"], rtol=1e-12)         # what it actually achieves


def test_black_scholes_vectorized_matches_oracle_incl_edge_cases():
    from cantorrl_b200 import sim
    rng = np.random.default_rng(3)
    n = 5000
    S = rng.uniform(20, 600, n)
    K = np.round(S * rng.uniform(0.7, 1.3, n))
    sig = rng.uniform(0.01, 1.2, n)
    sig[:50] = 0.0                  # below epsilon -> floored
    sig[50:60] = np.nan             # propagates (column 1 of the reference's vol matrix)
    for T in (1.0, 30 / 252, 1 / 252, 0.0, -0.5):
        c, p = sim.black_scholes_vectorized(S, K, T, 0.04, sig)
        c0, p0 = bs_oracle.black_scholes(S, K, T, 0.04, sig)
        for got, want in ((c.cpu().numpy(), c0), (p.cpu().numpy(), p0)):
            assert np.array_equal(np.isnan(got), np.isnan(want))
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9 * 600)
    # array maturities broadcast against scalars
    Tt = np.clip(1 - np.arange(n) / 252, 0, None)
    c, p = sim.black_scholes_vectorized(S, 100.0, Tt, 0.04, 0.3)
    c0, p0 = bs_oracle.black_scholes(S, 100.0, Tt, 0.04, 0.3)
    np.testing.assert_allclose(c.cpu().numpy(), c0, rtol=1e-6, atol=1e-9)


def test_schema_b_known_answer_pair_on_gpu():
    """data/paths.npy -> data/paths_options.npz, the reference's shipped golden pair (48-path slice)."""
    from cantorrl_b200 import sim
    z = np.load(os.path.join(GOLDEN, "schema_b_golden.npz"))
    calls, puts, vols = sim.process_price_paths(z["paths"], return_vols=True)
    for got, want in ((calls.cpu().numpy(), z["calls_shipped"]), (puts.cpu().numpy(), z["puts_shipped"])):
        assert got.shape == want.shape
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got[:, 1]).all()
        # 1e-6 relative (north_star, fp64) with an absolute floor for far out-of-the-money puts near 1e-300
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-10)                     # what it actually achieves
    np.testing.assert_allclose(vols.cpu().numpy(), z["vols"], rtol=1e-9, equal_nan=True)
    assert abs(float(calls[0, 0]) - 19.928449166775806) < 1e-9
    v2 = sim.calculate_annualized_vol_matrix(z["paths"][:5])
    np.testing.assert_allclose(v2.cpu().numpy(), z["vols"][:5], rtol=1e-9, equal_nan=True)


def test_multi_strike_book_matches_oracle():
    """BASELINE config 3 layout: M strikes K_m = round(S0) * {0.90 .. 1.10}, maturities to the episode end."""
    from cantorrl_b200 import sim
    z = np.load(os.path.join(GOLDEN, "schema_b_golden.npz"))
    paths = z["paths"][:16]
    mult = np.linspace(0.9, 1.1, 8)
    calls, puts = sim.process_price_paths(paths, strike_multipliers=mult)
    assert calls.shape == (8, 16, paths.shape[1])
    vols = bs_oracle.realised_vol_matrix(paths)
    Tt = np.clip(1 - np.arange(paths.shape[1]) / 252, 0, None)
    for m in range(8):
        K = np.round(paths[:, :1]) * mult[m]
        c0, p0 = bs_oracle.black_scholes(paths, K, Tt[None, :], 0.04, vols)
        np.testing.assert_allclose(calls[m].cpu().numpy(), c0, rtol=1e-6, atol=1e-9, equal_nan=True)
        np.testing.assert_allclose(puts[m].cpu().numpy(), p0, rtol=1e-6, atol=1e-9, equal_nan=True)


def test_bs_delta_hedge_golden():
    from cantorrl_b200 import sim
    z = np.load(os.path.join(GOLDEN, "bs_delta_golden.npz"))
    pnl = sim.bs_delta_hedge(z["paths"])
    np.testing.assert_allclose(pnl.cpu().numpy(), z["pnl"], rtol=1e-6, atol=1e-8)


def test_formats_roundtrip_and_npz_drop_in():
    """Schema A file written from a simulated book loads back into the env, and pack/unpack are inverse."""
    from cantorrl_b200 import HedgingVecEnv, ReplayData, sim
    book = sim.generate_paths_and_options(300, n_steps=17, model="heston")
    pm = book.to_path_major(torch.float64)
    again = ReplayData.from_arrays(pm["paths"], pm["volatilities"], pm["call_prices_atm"], pm["put_prices_atm"])
    assert torch.equal(again.tensor[:, :300], book.tensor[:, :300])
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "paths_rbergomi_options_100k.npz")
        book.save_npz(f)
        with np.load(f) as z:
            assert sorted(z.files) == ["call_prices_atm", "paths", "put_prices_atm", "volatilities"]
            assert z["paths"].shape == (300, 18) and z["call_prices_atm"].shape == (300, 17) and z["paths"].dtype == np.float64
        env = HedgingVecEnv(f, num_envs=8, episode_sampler="same_path")
    assert env.episode_length == 17 and env.num_episodes == 300
    assert torch.equal(env.data.tensor[:, :300], book.tensor[:, :300])


def test_env_schema_from_bare_paths_and_schema_b_files():
    """data/paths.npy-style input -> env-schema book on the GPU, and the schema-B file round trip."""
    from cantorrl_b200 import HedgingVecEnv, ReplayData, load_schema_b, save_schema_b, sim
    z = np.load(os.path.join(GOLDEN, "schema_b_golden.npz"))
    paths = z["paths"][:20]
    book = ReplayData.from_paths(paths)
    pm = {k: v.cpu().numpy() for k, v in book.to_path_major(torch.float64).items()}
    sig = bs_oracle.realised_vol_matrix(paths)
    sig[:, 0] = sig[:, 2]
    sig[:, 1] = sig[:, 2]
    np.testing.assert_allclose(pm["paths"], paths.astype(np.float32), rtol=0)
    np.testing.assert_allclose(pm["volatilities"], (sig ** 2).astype(np.float32), rtol=1e-6)
    C, P = bs_oracle.atm_book(pm["paths"], pm["volatilities"])
    np.testing.assert_allclose(pm["call_prices_atm"], C, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(pm["put_prices_atm"], P, rtol=1e-4, atol=1e-4)
    const = ReplayData.from_paths(paths, variances=0.02903)
    assert bool((const.v[:, :20] == np.float32(0.02903)).all())
    env = HedgingVecEnv(data=book, num_envs=20, episode_sampler="same_path")
    assert env.episode_length == 252 and bool(torch.isfinite(env.reset()).all())
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "paths.npy")
        np.save(f, paths)
        again = ReplayData.from_paths_npy(f)
        assert torch.equal(again.tensor, book.tensor)
        with pytest.raises(FileNotFoundError):
            ReplayData.from_paths_npy(os.path.join(d, "missing.npy"))
        calls, puts = sim.process_price_paths(paths)
        g = os.path.join(d, "paths_options.npz")
        save_schema_b(g, calls, puts)
        with np.load(g) as w:
            assert sorted(w.files) == ["calls", "puts"] and w["calls"].dtype == np.float64 and w["calls"].shape == (20, 253)
        c2, p2 = load_schema_b(g)
        assert torch.equal(c2.nan_to_num(-1), calls.nan_to_num(-1)) and torch.equal(p2.nan_to_num(-1), puts.nan_to_num(-1))
        with pytest.raises(FileNotFoundError):          # the reference env cannot load a schema-B file either (missing keys)
            HedgingVecEnv(g, num_envs=4)
