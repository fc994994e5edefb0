"""GPU parity tests of the float32 multi-strike Black-Scholes book kernel (K2, throughput form; BASELINE configs[2])."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import bs_oracle, sim_oracle

pytestmark = pytest.mark.gpu


def _book_from_paths(paths, variances=None):
    from cantorrl_b200 import ReplayData
    paths = np.asarray(paths, np.float64)
    v = np.full_like(paths, 0.04) if variances is None else variances
    z = np.zeros((paths.shape[0], paths.shape[1] - 1))
    return ReplayData.from_arrays(paths, v, z, z)


def test_float32_book_reproduces_the_reference_known_answer_pair():
    """M = 1, realised volatility, maturity to the episode end == process_price_paths: data/paths.npy ->
    data/paths_options.npz as shipped by the reference (48-path slice), at the fp32 tolerance of the north star."""
    from cantorrl_b200 import sim
    z = np.load(os.path.join(GOLDEN, "schema_b_golden.npz"))
    book = _book_from_paths(z["paths"])
    calls, puts = sim.reprice_book(book)
    for got, want in ((calls[0].T.cpu().numpy(), z["calls_shipped"]), (puts[0].T.cpu().numpy(), z["puts_shipped"])):
        assert got.shape == want.shape
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got[:, 1]).all()      # column 1 is NaN (ddof = 1)
        # prices are O(1..100) on S ~ 500: 1e-4 relative with an absolute floor of 1e-4 * (a 1-vol-point move) for tiny puts
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=5e-4, equal_nan=True)
    assert abs(float(calls[0, 0, 0]) - 19.928449166775806) < 1e-4 * 19.93                          # sigma-floor branch at t = 0


@pytest.mark.parametrize("sigma", ["realised", "book"])
@pytest.mark.parametrize("tenor", [None, 30 / 252])
def test_multi_strike_prices_deltas_gammas_match_oracle(sigma, tenor):
    from cantorrl_b200 import sim
    n, T = 257, 60
    S, V = sim_oracle.heston_paths(8, np.arange(n), T)
    book = _book_from_paths(S, V.astype(np.float64))
    mult = np.array([0.9, 0.95, 1.0, 1.02, 1.1], np.float32)
    calls, puts, deltas, gammas = sim.reprice_book(book, mult, sigma=sigma, tenor=tenor, greeks=True)
    assert calls.shape == (5, T + 1, n)
    S64 = S.astype(np.float64)
    sig = bs_oracle.realised_vol_matrix(S64) if sigma == "realised" else np.sqrt(np.maximum(V.astype(np.float64), 0))
    Tt = np.full(T + 1, tenor) if tenor else np.clip(1 - np.arange(T + 1) / 252, 0, None)
    for m in range(5):
        K = np.round(S64[:, :1]) * np.float64(mult[m])
        c0, p0 = bs_oracle.black_scholes(S64, K, Tt[None, :], 0.04, sig)
        d0, g0 = bs_oracle.call_delta_gamma(S64, K, Tt[None, :], 0.04, sig)
        np.testing.assert_allclose(calls[m].T.cpu().numpy(), c0, rtol=1e-4, atol=2e-4, equal_nan=True)
        np.testing.assert_allclose(puts[m].T.cpu().numpy(), p0, rtol=1e-4, atol=2e-4, equal_nan=True)
        ok = ~np.isnan(d0)
        if sigma == "realised":
            ok[:, 0] = False            # sigma floor 1e-8: delta is a 0/1 step, gamma a spike that float32 cannot represent
        np.testing.assert_allclose(deltas[m].T.cpu().numpy()[ok], d0[ok], rtol=1e-4, atol=2e-5)
        # gamma = phi(d1) / (S sigma sqrt(T)): relative error |d1| * delta(d1); a small realised sigma amplifies delta(d1)
        np.testing.assert_allclose(gammas[m].T.cpu().numpy()[ok], g0[ok], rtol=1e-3 if sigma == "realised" else 2e-4, atol=1e-5)
    # put-call parity of the kernel's own outputs: C - P = S - K e^{-rT}
    K = np.round(S64[:, :1]) * np.float64(mult[2])
    lhs = (calls[2] - puts[2]).T.cpu().numpy()
    rhs = S64 - K * np.exp(-0.04 * Tt[None, :])
    fin = ~np.isnan(lhs)
    np.testing.assert_allclose(lhs[fin], rhs[fin], rtol=0, atol=2e-4)


def test_book_on_simulated_heston_paths_full_width_properties():
    """2^20 Heston paths x 32 days x 8 strikes generated and repriced in HBM: monotone in strike, parity, bounds."""
    from cantorrl_b200 import sim
    n, T = 1 << 20, 32
    book = sim.generate_paths_and_options(n, n_steps=T, model="heston", reprice=False)
    mult = np.linspace(0.9, 1.1, 8).astype(np.float32)
    calls, puts = sim.reprice_book(book, mult, sigma="book")
    assert bool(torch.isfinite(calls).all()) and bool(torch.isfinite(puts).all())
    assert bool((calls[:-1] >= calls[1:] - 1e-4).all()) and bool((puts[:-1] <= puts[1:] + 1e-4).all())   # monotone in K
    assert bool((calls >= 0).all()) and bool((puts >= 0).all())
    S = book.S[:, :n]
    Tt = torch.clamp(1 - torch.arange(T + 1, device="cuda") / 252, min=0)[:, None]
    for m in (0, 3, 7):
        rhs = S - float(mult[m]) * 100.0 * torch.exp(-0.04 * Tt)
        assert float(((calls[m] - puts[m]) - rhs).abs().max()) < 5e-4


def test_book_in_path_slices_equals_the_book_in_one_piece():
    """``reprice_book_slices`` / ``cantor_reprice_book_strided``: ragged slices of the paths priced into slice-sized arrays are
    bit-equal to the corresponding columns of the one-piece book (realised sigma: a per-path running statistic; with greeks)."""
    from cantorrl_b200 import sim
    n, T = 1000, 20
    book = sim.generate_paths_and_options(n, n_steps=T, model="heston", reprice=False)
    mult = np.array([0.95, 1.0, 1.07], np.float32)
    for kw in (dict(sigma="realised"), dict(sigma="book", greeks=True, tenor=30 / 252)):
        whole = sim.reprice_book(book, mult, **kw)
        seen = 0
        for first, last, part in sim.reprice_book_slices(book, mult, paths_per_slice=301, **kw):
            assert first == seen and len(part) == len(whole)
            for w, q in zip(whole, part):
                a, b = w[:, :, first:last], q
                assert a.shape == b.shape and bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
            seen = last
        assert seen == n
    with pytest.raises(ValueError):
        sim.reprice_book(book, mult, path_range=(10, 5))


def test_config2_size_multi_strike_book_in_slices():
    """BASELINE configs[2] at full size: 2^24 Heston paths x 253 days x 8 strikes priced in 2^21-path slices (the one-piece
    output would be 259 GB); put-call parity and monotonicity in the strike on every slice.  Needs ~100 GB of HBM."""
    from cantorrl_b200 import sim
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs a 180 GB B200")
    n, T = 1 << 24, 252
    book = sim.generate_paths_and_options(n, n_steps=T, model="heston", reprice=False)
    mult = np.linspace(0.9, 1.1, 8).astype(np.float32)
    Tt = torch.clamp(1 - torch.arange(T + 1, device="cuda") / 252, min=0)[:, None]
    n_slices = 0
    for first, last, (calls, puts) in sim.reprice_book_slices(book, mult, paths_per_slice=1 << 21, sigma="book"):
        n_slices += 1
        assert bool((calls[:-1] >= calls[1:] - 1e-4).all()) and bool((puts[0] >= 0).all()) and bool(torch.isfinite(puts[7]).all())
        rhs = book.S[:, first:last] - float(mult[3]) * 100.0 * torch.exp(-0.04 * Tt)
        assert float(((calls[3] - puts[3]) - rhs).abs().max()) < 1e-3
    assert n_slices == 8


def test_reprice_book_argument_errors():
    from cantorrl_b200 import CantorError, sim
    book = sim.generate_paths_and_options(64, n_steps=8, reprice=False)
    with pytest.raises(ValueError):
        sim.reprice_book(book, sigma="implied")
    with pytest.raises(CantorError):
        sim.reprice_book(book, np.ones(33))
