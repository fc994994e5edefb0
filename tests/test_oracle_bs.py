"""Pins oracle/bs_oracle.py against the reference's shipped known-answer pair and bs_delta golden."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import bs_oracle


def test_schema_b_known_answer_pair():
    """data/paths.npy -> data/paths_options.npz (SURVEY §8(c) golden vector 1), 48-path slice."""
    z = np.load(os.path.join(GOLDEN, "schema_b_golden.npz"))
    calls, puts, vols = bs_oracle.schema_b_book(z["paths"])
    for got, want in ((calls, z["calls_shipped"]), (puts, z["puts_shipped"])):
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.isnan(want[:, 1]).all()                                   # column 1 is all-NaN
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-13)
    np.testing.assert_allclose(vols, z["vols"], rtol=1e-14, atol=0, equal_nan=True)
    # golden vector 2: sigma = eps branch at t = 0
    assert abs(z["calls_shipped"][0, 0] - 19.928449166775806) < 1e-12
    np.testing.assert_allclose(calls[:, 0], z["paths"][:, 0] - np.round(z["paths"][:, 0]) * np.exp(-0.04), atol=1e-10)


def test_bs_delta_golden():
    z = np.load(os.path.join(GOLDEN, "bs_delta_golden.npz"))
    pnl = bs_oracle.bs_delta_hedge(z["paths"])
    np.testing.assert_allclose(pnl, z["pnl"], rtol=1e-12, atol=1e-10)


def test_put_call_parity_and_expiry():
    rng = np.random.default_rng(0)
    S = rng.uniform(50, 150, 1000)
    K = np.round(S * rng.uniform(0.8, 1.2, 1000))
    sig = rng.uniform(0.05, 0.8, 1000)
    for T in (30 / 252, 1.0, 1e-3):
        c, p = bs_oracle.black_scholes(S, K, T, 0.04, sig)
        np.testing.assert_allclose(c - p, S - K * np.exp(-0.04 * T), atol=1e-9)
    c, p = bs_oracle.black_scholes(S, K, 0.0, 0.04, sig)
    np.testing.assert_array_equal(c, np.maximum(S - K, 0))
    np.testing.assert_array_equal(p, np.maximum(K - S, 0))
