"""GPU parity tests of the rough-Bergomi generator and nested-MC pricer against golden vectors of the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import rbergomi_oracle as ro

pytestmark = pytest.mark.gpu
Z = dict(np.load(os.path.join(GOLDEN, "rbergomi_golden.npz")))


def test_outer_generator_on_the_reference_run_draws():
    """Variance path (the FIR form of the reference's FFT pipeline) and log-Euler paths, float64, from the parameters and
    increments the unmodified generate_paths_and_options produced (rbergomi_sim.py:357-406, 454-464)."""
    from cantorrl_b200 import sim
    n_steps = int(Z["gen_n_steps"])
    n = Z["gen_S0"].shape[0]
    prm = np.stack([Z["gen_S0"], Z["gen_xi"], Z["gen_H"], Z["gen_eta"], Z["gen_rho"]])
    M = ro.next_power_of_two(n_steps + 1)
    paths, v = sim.rbergomi_outer_paths(n, n_steps, prm, Z["gen_dW1"][:, :M], Z["gen_dW2"][:, :M])
    np.testing.assert_allclose(v.cpu().numpy(), Z["gen_v"], rtol=1e-6)            # north star, fp64
    np.testing.assert_allclose(paths.cpu().numpy(), Z["gen_paths"], rtol=1e-6)
    np.testing.assert_allclose(v.cpu().numpy(), Z["gen_v"], rtol=1e-11)           # what it achieves
    np.testing.assert_allclose(paths.cpu().numpy(), Z["gen_paths"], rtol=1e-11)
    # the packed float32 book of the same run
    rb = sim.generate_rbergomi_paths_and_options(n, n_steps=n_steps, price=False,
                                                 exported=dict(params=prm, dW1=Z["gen_dW1"][:, :M], dW2=Z["gen_dW2"][:, :M]))
    np.testing.assert_allclose(rb.book.S[:, :n].T.cpu().numpy(), Z["gen_paths"], rtol=1e-6)
    np.testing.assert_allclose(rb.book.v[:, :n].T.cpu().numpy(), Z["gen_v"], rtol=1e-6)
    np.testing.assert_allclose(rb.H.cpu().numpy(), Z["gen_H"], rtol=0)


@pytest.mark.parametrize("tc", [False, True], ids=["ffma", "tcgen05"])
@pytest.mark.parametrize("kind", ["call", "put"])
def test_nested_mc_price_on_the_reference_draws(kind, tc):
    """price_rbergomi_option_gpu (:246-306) on its own complex draws, float32 inner arithmetic: 1e-4 relative."""
    from cantorrl_b200 import sim
    dW1, dW2 = ro.brownian_from_Z(Z[f"Z_{kind}"])
    got = sim.price_rbergomi_option(Z["S0"], Z["K"], float(Z["tenor"]), float(Z["r"]), Z["xi"], Z["H"], Z["eta"], Z["rho"], kind,
                                    dW1, dW2, float(Z["dt"]), tensor_cores=tc)
    np.testing.assert_allclose(got.cpu().numpy(), Z[f"price_{kind}"], rtol=1e-4)


@pytest.mark.parametrize("tc", [False, True], ids=["ffma", "tcgen05"])
def test_generated_book_statistics_and_resumability(tc):
    """Philox mode: parameters within the reference's clips, unbiased increments, prices consistent with an independent
    oracle Monte Carlo within sampling error, and pricing any day range in any order gives the same book."""
    from cantorrl_b200 import sim
    base = (496.48, 0.02903, 0.4656, 1.985, -0.2022)              # estimate_base_params on the shipped CSV (SURVEY [probe])
    n, T, n_mc = 512, 8, 4000
    rb = sim.generate_rbergomi_paths_and_options(n, base_params=base, n_steps=T, n_mc=n_mc, seed=7, days_per_launch=3, tensor_cores=tc)
    H, rho, xi, eta = (getattr(rb, a).cpu().numpy() for a in ("H", "rho", "xi", "eta"))
    assert H.min() >= 0.01 and H.max() <= 0.49 and rho.min() >= -0.99 and rho.max() <= -0.01
    assert (xi >= 0.5 * base[1] - 1e-15).all() and (eta >= 0.5 * base[3] - 1e-15).all()
    assert abs(rb.S0.mean().item() / base[0] - 1) < 4 * 0.01 / np.sqrt(n)
    book = rb.book
    assert torch.equal(book.C[T, :n], book.C[T - 1, :n]) and bool((book.C[:T, :n] > 0).all()) and bool((book.P[:T, :n] > 0).all())
    # resumability / order independence
    again = sim.generate_rbergomi_paths_and_options(n, base_params=base, n_steps=T, n_mc=n_mc, seed=7, price=False, tensor_cores=tc)
    again.price_days(5, 8).price_days(0, 5)
    assert torch.equal(again.book.tensor, book.tensor)
    # independent Monte Carlo of the same (S, K, v, H, eta, rho) states by the oracle
    idx = np.arange(0, n, 64)
    t = 3
    S = book.S[t, idx].double().cpu().numpy()
    v = book.v[t, idx].double().cpu().numpy()
    rng = np.random.default_rng(0)
    Zo = rng.normal(size=(len(idx), 20000, 32)) + 1j * rng.normal(size=(len(idx), 20000, 32))
    for kind, col in (("call", book.C), ("put", book.P)):
        want = ro.price_option(S, np.round(S), 30 / 252, 0.04, v, H[idx], eta[idx], rho[idx], kind, Zo, 1 / 252)
        got = col[t, idx].double().cpu().numpy()
        # standard error of the difference of two MC means: payoff std ~ 1.5 x price
        se = 1.5 * want * np.sqrt(1 / n_mc + 1 / 20000)
        assert (np.abs(got - want) < 5 * se).all(), (kind, got, want)
    # calls and puts are priced on independent draws (rbergomi_sim.py:437-446): put-call parity holds only statistically
    par = (book.C[t, :n] - book.P[t, :n]).double().cpu().numpy() - (book.S[t, :n].double().cpu().numpy()
                                                                      - np.round(book.S[t, :n].cpu().numpy()) * np.exp(-0.04 * 30 / 252))
    assert abs(par.mean()) < 0.5 and par.std() > 1e-3


def test_rbergomi_book_drives_the_env():
    from cantorrl_b200 import HedgingVecEnv, sim
    rb = sim.generate_rbergomi_paths_and_options(256, n_steps=6, n_mc=500)
    env = HedgingVecEnv(data=rb.book, num_envs=256, episode_sampler="same_path")
    obs = env.reset()
    for _ in range(7):
        obs, r, d, _ = env.step(torch.zeros((256, 2), device="cuda"))
    assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(r).all())
