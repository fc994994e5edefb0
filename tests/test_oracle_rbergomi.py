"""Pins oracle/rbergomi_oracle.py against golden vectors produced by the UNMODIFIED reference simulator functions."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import rbergomi_oracle as ro

Z = dict(np.load(os.path.join(GOLDEN, "rbergomi_golden.npz")))


def test_nested_mc_price_restatement_matches_reference_on_its_own_draws():
    for kind in ("call", "put"):
        got = ro.price_option(Z["S0"], Z["K"], float(Z["tenor"]), float(Z["r"]), Z["xi"], Z["H"], Z["eta"], Z["rho"], kind,
                              Z[f"Z_{kind}"], float(Z["dt"]))
        np.testing.assert_allclose(got, Z[f"price_{kind}"], rtol=1e-12)


def test_fir_identity_equals_the_reference_fft_form():
    """X = sqrt(2H) eta / sqrt(M) * (lambda circularly convolved with dW1): what the CUDA kernels compute."""
    n_steps = int(float(Z["tenor"]) / float(Z["dt"]))
    t = np.linspace(0, n_steps * float(Z["dt"]), n_steps + 1)
    la = ro.lam(t, Z["H"])
    Zc = Z["Z_call"]
    X_fft = ro.fgn_fft(ro.phi(la), Zc, Z["H"], Z["eta"], n_steps + 1)
    dW1, dW2 = ro.brownian_from_Z(Zc)
    np.testing.assert_allclose(ro.fgn_conv(la, dW1, Z["H"], Z["eta"], n_steps + 1), X_fft, rtol=0, atol=1e-14)
    for kind in ("call", "put"):
        d1, d2 = ro.brownian_from_Z(Z[f"Z_{kind}"])
        got = ro.price_option_from_increments(Z["S0"], Z["K"], float(Z["tenor"]), float(Z["r"]), Z["xi"], Z["H"], Z["eta"],
                                              Z["rho"], kind, d1, d2, float(Z["dt"]))
        np.testing.assert_allclose(got, Z[f"price_{kind}"], rtol=1e-12)
    assert abs(dW1.std() - 1) < 0.02 and abs(dW2.std() - 1) < 0.02 and abs(np.mean(dW1 * dW2)) < 0.02


def test_generator_variance_and_paths_match_reference_run():
    n_steps = int(Z["gen_n_steps"])
    d1, d2 = ro.brownian_from_Z(Z["gen_Z"])
    np.testing.assert_allclose(d1, Z["gen_dW1"], rtol=0, atol=1e-13)
    paths, v = ro.outer_paths(Z["gen_S0"], Z["gen_xi"], Z["gen_H"], Z["gen_eta"], Z["gen_rho"], Z["gen_dW1"], Z["gen_dW2"],
                              n_steps=n_steps)
    np.testing.assert_allclose(v, Z["gen_v"], rtol=1e-12)
    np.testing.assert_allclose(paths, Z["gen_paths"], rtol=1e-12)


def test_parameter_perturbation_clips():
    base = tuple(Z["gen_base"])
    z = np.array([[0, 3, -3], [0, 5, -5], [0, 10, -10], [0, 5, -5], [0, 20, -20]], float)
    S0, xi, H, eta, rho = ro.perturb_params(base, z)
    assert S0[0] == base[0] and xi[0] == base[1] and eta[0] == base[3]
    assert xi[2] == 0.5 * base[1] and eta[2] == 0.5 * base[3]                  # MIN_*_FACTOR (:35-36)
    assert H.max() <= 0.49 and H.min() >= 0.01 and rho.max() <= -0.01 and rho.min() >= -0.99    # :37-40
    # the reference run's own parameters respect the same clips
    assert Z["gen_H"].max() <= 0.49 and Z["gen_rho"].min() >= -0.99 and (Z["gen_xi"] >= 0.5 * base[1]).all()


def test_host_calibration_matches_the_reference_estimate_base_params():
    """cantorrl_b200.calibration (host-side scalar routine) against the unmodified reference's outputs."""
    from cantorrl_b200.calibration import estimate_base_params
    g = np.load(os.path.join(GOLDEN, "calibration_golden.npz"))
    names = [k[len("prices_"):] for k in g.files if k.startswith("prices_")]
    assert "shipped" in names and len(names) >= 8
    for k in names:
        with np.errstate(all="ignore"):
            got = np.array(estimate_base_params(g[f"prices_{k}"]), np.float64)
        np.testing.assert_allclose(got, g[f"params_{k}"], rtol=1e-10, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(g["params_shipped"], [496.48, 0.02903, 0.4656, 1.985, -0.2022], rtol=2e-4)   # SURVEY [probe]
