"""Pins oracle/sim_oracle.py: Random123 known-answer vectors for Philox4x32-10 and the reference's outer Euler step."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import sim_oracle


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors, philox4x32 with 10 rounds."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        got = sim_oracle.philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))
        assert [int(x) for x in got] == want
    # vectorised call == scalar calls
    ctrs = np.array([k[0] for k in kat], np.uint32)
    assert np.array_equal(sim_oracle.philox4x32_10(ctrs, np.array([0, 0], np.uint32))[0],
                          sim_oracle.philox4x32_10(ctrs[0], np.array([0, 0], np.uint32)))


def test_outer_euler_step_matches_unmodified_simulator():
    """tests/golden/outer_euler_golden.npz was produced by the reference's generate_paths_and_options (CPU, cupy stand-in)."""
    z = np.load(os.path.join(GOLDEN, "outer_euler_golden.npz"))
    paths = sim_oracle.euler_from_normals(z["S0"], z["v"], z["dW1"], z["dW2"], z["rho"], float(z["r"]), float(z["dt"]))
    np.testing.assert_allclose(paths, z["paths"], rtol=1e-14, atol=0)
    assert paths.shape == z["paths"].shape and (paths >= 1e-8).all()


def test_philox_normals_are_standard_normal_and_index_stable():
    z = sim_oracle.philox_normals(42, np.arange(4096), 64, 2)
    assert z.shape == (4096, 64, 2) and z.dtype == np.float32
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert abs(np.corrcoef(z[..., 0].ravel(), z[..., 1].ravel())[0, 1]) < 5e-3
    # a path's draws depend on its global index only
    sub = sim_oracle.philox_normals(42, np.arange(100, 110), 64, 2)
    assert np.array_equal(sub, z[100:110])
    # GBM (1 normal per step) and Heston (2 per step) read the same underlying sequence
    z1 = sim_oracle.philox_normals(42, np.arange(8), 128, 1)
    assert np.array_equal(z1.reshape(8, 64, 2), z[:8])


def test_gbm_and_heston_oracle_moments():
    S, V = sim_oracle.gbm_paths(7, np.arange(20000), 60, dtype=np.float64)
    lr = np.log(S[:, -1] / S[:, 0])
    T = 60 / 252
    assert abs(lr.mean() - (0.04 - 0.02) * T) < 3 * 0.2 * np.sqrt(T) / np.sqrt(20000)
    assert abs(lr.var() - 0.04 * T) < 0.04 * T * 0.05
    assert (V == 0.04).all()
    S, V = sim_oracle.heston_paths(7, np.arange(20000), 60, dtype=np.float64)
    assert (V >= 0).all() and abs(V[:, -1].mean() - 0.04) < 2e-3
    assert abs(np.log(S[:, -1] / 100).mean() - (0.04 - 0.02) * T) < 6e-3
