"""GPU parity tests of the episode-fused rollout kernel (policy + env step + statistics) against the oracle.

Parity is checked TEACHER-FORCED: the kernel's stored actions drive the oracle env, so a policy output that differs
in the last float32 bit cannot fork the two trajectories; the policy itself is compared on the kernel's own
observations.  fp32 tolerance of the north star: 1e-4 relative; done flags and episode counts exact.
"""
import os
import numpy as np
import pytest
import torch

from oracle import bs_oracle, policy_oracle, rollout_oracle, sim_oracle
from oracle.hedge_oracle import EnvParams

pytestmark = pytest.mark.gpu

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)


def _book(n_paths, T, seed=5, heston=False):
    idx = np.arange(n_paths)
    S, V = (sim_oracle.heston_paths if heston else sim_oracle.gbm_paths)(seed, idx, T)
    C, P = bs_oracle.atm_book(S.astype(np.float64), V.astype(np.float64))
    return S, V, C.astype(np.float32), P.astype(np.float32)


def _mlp_weights(seed=0):
    g = np.random.default_rng(seed)
    W1, b1 = g.normal(0, 0.5, (64, 13)).astype(np.float32), g.normal(0, 0.1, 64).astype(np.float32)
    W2, b2 = g.normal(0, 0.2, (64, 64)).astype(np.float32), g.normal(0, 0.1, 64).astype(np.float32)
    W3, b3 = g.normal(0, 0.3, (2, 64)).astype(np.float32), g.normal(0, 0.1, 2).astype(np.float32)
    mean = g.normal(0, 0.2, 13).astype(np.float32)
    var = g.uniform(0.05, 2.0, 13).astype(np.float32)
    return W1, b1, W2, b2, W3, b3, mean, var


@pytest.mark.parametrize("policy", ["no_hedge", "random", "delta_every_step", "delta_benchmark", "mlp", "mlp_bf16", "actions"])
@pytest.mark.parametrize("loss", ["abs", "mse"])
def test_rollout_matches_oracle_teacher_forced(policy, loss):
    from cantorrl_b200.rollout import HedgingRollout, pack_mlp
    n_paths, T, n_envs, n_steps, off, total = 61, 12, 203, 41, 1000, 5000
    S, V, C, P = _book(n_paths, T, heston=True)
    kw = dict(KW, loss_type=loss)
    ro = HedgingRollout(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=n_envs,
                        env_offset=off, total_envs=total, **kw)
    w = _mlp_weights()
    forced = None
    if policy == "actions":
        forced = np.random.default_rng(9).uniform(-1.3, 1.3, (n_steps, n_envs, 2)).astype(np.float32)
    stats = ro.new_stats(hist_bins=512, hist_max=2.0, keep_episodes=4)
    res = ro.run(n_steps, policy, mlp=pack_mlp(*w) if policy.startswith("mlp") else None, seed=77, stats=stats, store=True,
                 actions=torch.from_numpy(forced).cuda() if forced is not None else None)
    torch.cuda.synchronize()
    got = {k: getattr(res, k).cpu().numpy() for k in ("obs", "actions", "reward", "done")}
    ref = rollout_oracle.run_rollout(S, V, C, P, EnvParams(**kw), policy, n_envs, n_steps, off, total,
                                     forced_actions=got["actions"], seed=77, mlp=w)
    assert np.array_equal(got["done"], ref["done"])
    np.testing.assert_allclose(got["obs"], ref["obs"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(got["reward"], ref["reward"], rtol=1e-4, atol=1e-7)
    # the policy, evaluated by the oracle on the observations the kernel itself produced (same float32 inputs)
    flat = got["obs"].reshape(-1, 13)
    if policy in ("random", "actions", "no_hedge"):
        assert np.array_equal(got["actions"], ref["policy_actions"])           # same Philox words / same inputs: bit-exact
    else:
        if policy == "delta_every_step":
            want_a = policy_oracle.delta_every_step(flat)
        elif policy == "delta_benchmark":
            pos = np.rint(flat[:, 3:5].astype(np.float64) * 200).astype(np.int64)
            want_a = policy_oracle.delta_benchmark(flat, pos[:, 0], pos[:, 1])
        elif policy == "mlp":
            want_a = rollout_oracle.mlp_actor(flat, *w)
        else:
            # tensor-core form: bf16 operands, float32 accumulation.  An activation that lands on a bf16 rounding boundary
            # may round the other way under a different summation order (one bf16 ulp = 2^-8 relative of ONE hidden unit),
            # so the comparison with the bf16-emulating oracle is at 1e-2 absolute on actions in [-1, 1] -- and well
            # inside it on average -- while the float32 oracle bounds the quantisation error itself.
            want_a = rollout_oracle.mlp_actor_bf16(flat, *w)
            assert float(stats.sums[15]) == 0.0, "a tcgen05 MMA timed out"
            err = np.abs(got["actions"].reshape(-1, 2) - want_a)
            assert err.max() < 1e-2 and err.mean() < 5e-4
            f32 = rollout_oracle.mlp_actor(flat, *w)
            assert np.abs(got["actions"].reshape(-1, 2) - f32).max() < 8e-2
            want_a = None
        if want_a is not None:
            np.testing.assert_allclose(got["actions"].reshape(-1, 2), want_a, rtol=1e-4, atol=1e-4)
    # statistics of the finished episodes
    want, b = rollout_oracle.stats_vector(ref["ep_pps"], ref["ep_cost"], ref["ep_reward"], T)
    sums = stats.sums.cpu().numpy()
    assert sums[0] == want[0] == n_envs * (n_steps // T)
    scale = np.abs(ref["ep_pps"]).sum(1).mean() / T
    np.testing.assert_allclose(sums[1:11], want[1:], rtol=2e-4, atol=2e-4 * max(scale, 1e-6) * want[0])
    assert sums[11] == n_envs * n_steps
    r = stats.result()
    sb = np.sort(b)
    np.testing.assert_allclose(r["cvar95_abs_pnl_exact"], sb[int(0.95 * len(sb)):].mean(), rtol=2e-4, atol=1e-6)
    assert int(stats.hist.sum()) == int(want[0])
    np.testing.assert_allclose(r["cvar95_abs_pnl"], r["cvar95_abs_pnl_exact"], rtol=0, atol=2.0 / 512)


@pytest.mark.parametrize("model", ["gbm", "heston"])
def test_on_the_fly_rollout_equals_replay_of_simulated_book(model):
    """Same Philox counters, same device functions: simulating inside the rollout == replaying K1's book, bit for bit."""
    from cantorrl_b200 import sim
    from cantorrl_b200.rollout import HedgingRollout
    n_envs, T, episodes = 1000, 21, 3
    n_steps = T * episodes - 5
    book = sim.generate_paths_and_options(n_envs * episodes, seed=11, n_steps=T, model=model)
    a = HedgingRollout(data=book, num_envs=n_envs, **KW)
    b = HedgingRollout(simulate=dict(model=model, seed=11, n_steps=T), num_envs=n_envs, **KW)
    for policy in ("random", "delta_every_step"):
        ra, rb = a.run(n_steps, policy, seed=3, store=True), b.run(n_steps, policy, seed=3, store=True)
        for k in ("obs", "actions", "reward", "done"):
            assert torch.equal(getattr(ra, k), getattr(rb, k)), (policy, k)
        np.testing.assert_allclose(ra.stats.sums.cpu().numpy(), rb.stats.sums.cpu().numpy(), rtol=1e-12)
        assert torch.equal(ra.stats.hist, rb.stats.hist)


def test_statistics_do_not_depend_on_sharding():
    """2 shards of the global env index == 1 shard: histogram counts exact, float64 sums to summation-order noise."""
    from cantorrl_b200.rollout import HedgingRollout
    total, T = 6000, 10
    sim_kw = dict(model="heston", seed=4, n_steps=T)
    whole = HedgingRollout(simulate=sim_kw, num_envs=total, **KW).run(35, "delta_benchmark")
    parts = [HedgingRollout(simulate=sim_kw, num_envs=cnt, env_offset=off, total_envs=total, **KW).run(35, "delta_benchmark")
             for off, cnt in ((0, 2500), (2500, 3500))]
    sums = sum(p.stats.sums for p in parts)
    hist = sum(p.stats.hist for p in parts)
    assert torch.equal(hist, whole.stats.hist) and int(hist.sum()) == total * 3
    np.testing.assert_allclose(sums.cpu().numpy(), whole.stats.sums.cpu().numpy(), rtol=1e-11)
    np.testing.assert_allclose(sum(p.stats.hist_sum for p in parts).cpu().numpy(), whole.stats.hist_sum.cpu().numpy(), rtol=1e-11)


def test_full_size_rollout_invariants():
    """BASELINE configs[1] shape (2^20 envs x 252 steps, GBM on the fly): size-independent properties."""
    from cantorrl_b200.rollout import HedgingRollout
    n, T = 1 << 20, 252
    ro = HedgingRollout(simulate=dict(model="gbm", seed=42, n_steps=T), num_envs=n, one_call_only=True, **KW)
    r0 = ro.run(T, "no_hedge").stats.result()
    assert r0["n_episodes"] == n and r0["env_steps"] == n * T
    assert r0["mean_cost"] == 0.0 and r0["std_cost"] == 0.0                      # no trades, no costs
    # unhedged P&L per share = S_T - S_0: mean of |S_T - S_0| / T for GBM(mu = .04, sigma = .2), S0 = 100
    assert abs(r0["mean_signed_pnl"] - 100 * (np.exp(0.04) - 1)) < 0.1
    # theta penalty alone is deterministic: sum_t 2e-4 (T - t) / 252, t = 1..T
    theta = 2e-4 * sum(T - t for t in range(1, T + 1)) / 252
    assert r0["mean_reward"] < -theta
    r1 = ro.run(T, "random", seed=1).stats.result()
    assert r1["n_episodes"] == n and r1["mean_cost"] > 0
    # idempotence: the same launch twice gives the same statistics (counter-based randomness, no hidden state)
    r2 = ro.run(T, "random", seed=1).stats.result()
    assert r1["n_episodes"] == r2["n_episodes"] and abs(r1["mean_abs_pnl"] - r2["mean_abs_pnl"]) < 1e-12
    assert r1["cvar95_abs_pnl"] >= r1["mean_abs_pnl"]


def test_rollout_argument_errors():
    from cantorrl_b200 import CantorError
    from cantorrl_b200.rollout import HedgingRollout
    ro = HedgingRollout(simulate=dict(n_steps=8), num_envs=16)
    with pytest.raises(ValueError):
        ro.run(4, "mlp")
    with pytest.raises(ValueError):
        ro.run(4, "nonsense")
    with pytest.raises(ValueError):
        ro.run(4, "actions", actions=torch.zeros((3, 16, 2), device="cuda"))
    with pytest.raises(TypeError):
        HedgingRollout(simulate=dict(nsteps=8), num_envs=16)
    with pytest.raises(FileNotFoundError):
        HedgingRollout(num_envs=4)
    assert issubclass(CantorError, RuntimeError)


def _lstm_weights(seed=3):
    g = np.random.default_rng(seed)
    k = 1 / np.sqrt(128)
    return dict(w_ih=g.uniform(-k, k, (512, 13)).astype(np.float32) * 3, w_hh=g.uniform(-k, k, (512, 128)).astype(np.float32) * 2,
                b_ih=g.uniform(-k, k, 512).astype(np.float32), b_hh=g.uniform(-k, k, 512).astype(np.float32),
                W1=g.normal(0, 0.15, (64, 128)).astype(np.float32), b1=g.normal(0, 0.1, 64).astype(np.float32),
                W2=g.normal(0, 0.2, (64, 64)).astype(np.float32), b2=g.normal(0, 0.1, 64).astype(np.float32),
                W3=g.normal(0, 0.3, (2, 64)).astype(np.float32), b3=g.normal(0, 0.1, 2).astype(np.float32))


@pytest.mark.parametrize("n_envs", [100, 203, 900])
def test_recurrent_lstm_actor_matches_oracle_over_episodes(n_envs):
    """The tensor-core LSTM + MLP actor inside the rollout: env parity teacher-forced on its own actions, and the actions
    themselves against the oracle network run over the kernel's observation sequence (state reset at episode ends).
    100 envs = one CTA whose second group is EMPTY (the groups store their observation pieces independently: nothing to store for
    it); 203 = one ragged CTA (second group partly empty); 900 = three full CTAs of 256 + one of 132."""
    from cantorrl_b200.rollout import HedgingRollout, pack_lstm
    n_paths, T, n_steps = 61, 12, 41
    S, V, C, P = _book(n_paths, T, heston=True)
    w = _lstm_weights()
    g = np.random.default_rng(1)
    mean, var = g.normal(0, 0.2, 13).astype(np.float32), g.uniform(0.05, 2.0, 13).astype(np.float32)
    ro = HedgingRollout(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=n_envs, **KW)
    stats = ro.new_stats()
    res = ro.run(n_steps, "lstm_bf16", mlp=pack_lstm(**w, obs_mean=mean, obs_var=var), stats=stats, store=True)
    torch.cuda.synchronize()
    assert float(stats.sums[15]) == 0.0, "a tcgen05 MMA timed out"
    got = {k: getattr(res, k).cpu().numpy() for k in ("obs", "actions", "reward", "done")}
    ref = rollout_oracle.run_rollout(S, V, C, P, EnvParams(**KW), "actions", n_envs, n_steps, forced_actions=got["actions"])
    assert np.array_equal(got["done"], ref["done"])
    np.testing.assert_allclose(got["obs"], ref["obs"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(got["reward"], ref["reward"], rtol=1e-4, atol=1e-7)
    want = rollout_oracle.lstm_actor_sequence(got["obs"], got["done"], **w, mean=mean, var=var, bf16=True)
    err = np.abs(got["actions"] - want)
    # bf16 operands + tanh.approx (2^-11 relative) through a 12-step recurrence: measured max 3e-3, mean 1.5e-4 (a wrong
    # hidden unit out of 128 already shows as max 3e-2, mean 1.5e-2)
    assert err.max() < 8e-3 and err.mean() < 5e-4, (err.max(), err.mean())
    full = rollout_oracle.lstm_actor_sequence(got["obs"], got["done"], **w, mean=mean, var=var, bf16=False)
    assert np.abs(got["actions"] - full).max() < 0.15 and np.abs(got["actions"] - full).mean() < 1e-2
    assert np.abs(want).mean() > 0.05                     # the network is not saturated / trivially zero
    assert stats.sums[0] == n_envs * (n_steps // T)


def test_recurrent_policy_with_the_shipped_weights_and_tanh_squash():
    """The policy the reference ships (quantconnect/model_files/policy_weights.pth + normalization_stats.pkl, carried in
    tests/golden/lstm_golden.npz) inside the rollout kernel, squashed with tanh like the reference's deployment wrapper
    (quantconnect/model_wrapper.py:202), against the oracle network -- which tests/test_oracle_policy.py pins on the
    reference's own RecurrentPPOModel -- run over the kernel's observation sequence."""
    from cantorrl_b200.rollout import HedgingRollout, pack_lstm
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lstm_golden.npz"))
    w = {k: g[k] for k in ("w_ih", "w_hh", "b_ih", "b_hh", "W1", "b1", "W2", "b2", "W3", "b3")}
    n_paths, T, n_envs, n_steps = 61, 16, 300, 40
    S, V, C, P = _book(n_paths, T, heston=True)
    ro = HedgingRollout(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=n_envs, **KW)
    stats = ro.new_stats()
    res = ro.run(n_steps, "lstm_bf16", mlp=pack_lstm(**w, obs_mean=g["obs_mean"], obs_var=g["obs_var"]), stats=stats, store=True,
                 squash="tanh")
    torch.cuda.synchronize()
    assert float(stats.sums[15]) == 0.0, "a tcgen05 MMA timed out"
    got = {k: getattr(res, k).cpu().numpy() for k in ("obs", "actions", "reward", "done")}
    ref = rollout_oracle.run_rollout(S, V, C, P, EnvParams(**KW), "actions", n_envs, n_steps, forced_actions=got["actions"])
    assert np.array_equal(got["done"], ref["done"])
    np.testing.assert_allclose(got["obs"], ref["obs"], rtol=1e-4, atol=2e-6)
    # squash="tanh" is the deployment wrapper's convention, which does not clip the normalised observation (model_wrapper.py:131)
    want = rollout_oracle.lstm_actor_sequence(got["obs"], got["done"], **w, mean=g["obs_mean"], var=g["obs_var"], bf16=True, squash="tanh",
                                              obs_clip=np.inf)
    err = np.abs(got["actions"] - want)
    assert err.max() < 2e-2 and err.mean() < 1e-3, (err.max(), err.mean())
    full = rollout_oracle.lstm_actor_sequence(got["obs"], got["done"], **w, mean=g["obs_mean"], var=g["obs_var"], bf16=False, squash="tanh",
                                              obs_clip=np.inf)
    assert np.abs(got["actions"] - full).mean() < 2e-2
    assert np.abs(got["actions"]).max() <= 1.0 and got["actions"].std() > 0.05


def test_config2_size_heston_16m_paths_replay_equals_on_the_fly():
    """BASELINE configs[2] at its full size: 2^24 Heston paths x 252 days generated into a 68 GB packed book (K1), a whole
    episode of every env replayed from it, and the same episode simulated on the fly inside the rollout kernel: the
    statistics of the two are identical (same device functions, same Philox counters), every env finishes exactly one
    episode, the martingale holds.  Needs ~75 GB of HBM."""
    from cantorrl_b200 import sim
    from cantorrl_b200.rollout import HedgingRollout
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    n, T = 1 << 24, 252
    book = sim.generate_paths_and_options(n, model="heston", n_steps=T)
    assert book.tensor.shape[0] == T + 1 and bool(torch.isfinite(book.tensor[-1]).all())
    assert abs(float(book.tensor[T, :n, 0].double().mean()) / (100 * np.exp(0.04)) - 1) < 2e-3      # E[S_T] = S_0 e^{rT}
    ro = HedgingRollout(data=book, num_envs=n, **KW)
    st = ro.new_stats()
    ro.run(T, "delta_every_step", stats=st)
    torch.cuda.synchronize()
    replayed = st.sums[:12].clone()
    del ro, book, st
    torch.cuda.empty_cache()
    ro = HedgingRollout(simulate=dict(model="heston", n_steps=T), num_envs=n, **KW)
    st = ro.new_stats()
    ro.run(T, "delta_every_step", stats=st)
    torch.cuda.synchronize()
    assert float(st.sums[0]) == n and float(st.sums[11]) == float(n) * T
    assert torch.allclose(st.sums[:12], replayed, rtol=1e-9, atol=0)


@pytest.mark.parametrize("policy,tol_max,tol_mean", [("mlp", 2e-5, 2e-6), ("mlp_bf16", 3e-2, 3e-3)])
def test_plain_mlp_policy_with_the_shipped_head_weights(policy, tol_max, tol_mean):
    """BASELINE configs[4]'s MLP 13-64-64-2 pinned on the reference: the shipped mlp_extractor.policy_net / action_net tensors
    behind the fixed 13 -> 128 projection of tests/golden/mlp_golden.npz (the oracle network reproduces the reference's own
    modules on it to 3e-6, tests/test_oracle_policy.py), tanh squash and no observation clip like the deployment wrapper
    (quantconnect/model_wrapper.py:131, 202), run inside the rollout kernel and compared on the kernel's own observations."""
    from cantorrl_b200.rollout import HedgingRollout, pack_mlp
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g, l = np.load(os.path.join(here, "mlp_golden.npz")), np.load(os.path.join(here, "lstm_golden.npz"))
    W1 = (l["W1"].astype(np.float64) @ g["projection"].astype(np.float64)).astype(np.float32)
    w = dict(W1=W1, b1=l["b1"], W2=l["W2"], b2=l["b2"], W3=l["W3"], b3=l["b3"])
    n_paths, T, n_envs, n_steps = 61, 16, 300, 40
    S, V, C, P = _book(n_paths, T, heston=True)
    ro = HedgingRollout(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=n_envs, **KW)
    stats = ro.new_stats()
    res = ro.run(n_steps, policy, mlp=pack_mlp(**w, obs_mean=l["obs_mean"], obs_var=l["obs_var"]), stats=stats, store=True, squash="tanh")
    torch.cuda.synchronize()
    assert float(stats.sums[15]) == 0.0, "a tcgen05 MMA timed out"
    obs, act = res.obs.cpu().numpy(), res.actions.cpu().numpy()
    fn = rollout_oracle.mlp_actor_bf16 if policy == "mlp_bf16" else rollout_oracle.mlp_actor
    want = fn(obs.reshape(-1, 13), **w, mean=l["obs_mean"], var=l["obs_var"], squash="tanh", obs_clip=np.inf).reshape(act.shape)
    err = np.abs(act - want)
    assert err.max() < tol_max and err.mean() < tol_mean, (err.max(), err.mean())
    full = rollout_oracle.mlp_actor(obs.reshape(-1, 13), **w, mean=l["obs_mean"], var=l["obs_var"], squash="tanh", obs_clip=np.inf).reshape(act.shape)
    assert np.abs(act - full).mean() < 1e-2 and act.std() > 0.05
    # with SB3's conventions (default squash="clip": Box clip + observation clip at 10) the same weights give clipped means
    res2 = ro.run(n_steps, policy, mlp=pack_mlp(**w, obs_mean=l["obs_mean"], obs_var=l["obs_var"]), store=True)
    want2 = fn(res2.obs.cpu().numpy().reshape(-1, 13), **w, mean=l["obs_mean"], var=l["obs_var"]).reshape(act.shape)
    assert np.abs(res2.actions.cpu().numpy() - want2).max() < max(tol_max, 5e-5) * 3


@pytest.mark.gpu
def test_umma_probe_reports_the_math_floor_at_n_256():
    """cantor_umma_probe (the measurement behind DESIGN section 4's tensor-core bounds): a 128 x 256 x 16 bf16 MMA cannot beat the pipe's
    math floor of 128 clk, and narrow MMAs cost about the same per instruction whatever N."""
    import ctypes as C
    from cantorrl_b200 import _lib
    L = _lib.lib()
    out = {}
    for n in (16, 64, 256):
        i, t = C.c_double(), C.c_double()
        _lib.check(L.cantor_umma_probe(n, 9, 16, 0, 4, C.byref(i), C.byref(t)), "cantor_umma_probe")
        out[n] = t.value
    assert 120.0 <= out[256] <= 200.0, out
    assert 20.0 <= out[16] <= 110.0 and abs(out[16] - out[64]) <= 15.0, out
    with pytest.raises(_lib.CantorError):
        _lib.check(L.cantor_umma_probe(24, 9, 16, 0, 4, C.byref(i), C.byref(t)), "cantor_umma_probe")
