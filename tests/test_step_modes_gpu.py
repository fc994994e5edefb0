"""GPU tests of the two other forms of the fused hedge step, both pinned on the per-step replay kernel (which the golden
vectors of the unmodified reference pin, tests/test_env_gpu.py):

  * cantor_env_step_many as ONE persistent launch (state in registers)  == the same steps as chained per-step launches, bit for bit;
  * cantor_env_step_sim, the on-the-fly mode (path generated inside the step kernel) == cantor_env_step replaying the book that
    cantor_sim_paths writes with the same parameters, bit for bit (hedging_env_v2.py:206-231 semantics: pre-advance marks for the
    slippage, stale marks at the terminal step).

Plus the episode sampler fixes (distinct first episodes per reset and per rank, reproducible re-seeding).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4, initial_cash=125.0)


def _tape(n_steps, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.rand((n_steps, n, 2), device="cuda", generator=g) * 2.6 - 1.3).contiguous()


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
@pytest.mark.parametrize("n_envs,sampler", [(1, "same_path"), (130, "philox"), (1000, "philox"), (1003, "same_path"), (4096, "philox")])
def test_step_many_persistent_equals_per_step_launches(prec, n_envs, sampler):
    """One persistent launch over an action tape (3.4 episodes, so auto-resets and Philox episode draws happen inside it)
    against the same tape stepped launch by launch: observations, rewards, dones, terminal observations and the final state."""
    from cantorrl_b200 import HedgingVecEnv, sim
    T, n_paths, k = 7, 61, 24
    book = sim.generate_paths_and_options(n_paths, n_steps=T, model="heston", seed=11)
    envs = [HedgingVecEnv(data=book, num_envs=n_envs, precision=prec, episode_sampler=sampler, seed=5, env_offset=17, **KW)
            for _ in range(2)]
    tape = _tape(k, n_envs, 3)
    for e in envs:
        e.reset()
    assert torch.equal(envs[0].current_episode_idx, envs[1].current_episode_idx)
    obs_a, rew_a, done_a = envs[0].step_many(tape[:10])
    obs_a2, rew_a2, done_a2 = envs[0].step_many(tape[10:])          # a second call continues from the stored state
    obs_a, rew_a, done_a = torch.cat([obs_a, obs_a2]), torch.cat([rew_a, rew_a2]), torch.cat([done_a, done_a2])
    term_seen = torch.zeros_like(envs[1]._terminal_obs)
    for t in range(k):
        o, r, d, info = envs[1].step(tape[t])
        assert torch.equal(o, obs_a[t]), f"obs, step {t}"
        assert torch.equal(r, rew_a[t]), f"reward, step {t}"
        assert torch.equal(d, done_a[t]), f"done, step {t}"
        term_seen = torch.where(d[:, None], info["terminal_observation"], term_seen)
    assert bool(done_a.any()) and bool(done_a[T - 1].all())
    assert torch.equal(envs[0]._core, envs[1]._core) and torch.equal(envs[0]._cash, envs[1]._cash)
    assert torch.equal(envs[0]._terminal_obs, term_seen)
    if prec == "fp64":
        assert torch.equal(envs[0]._pv_prev, envs[1]._pv_prev)


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
def test_step_many_with_monitor_and_statistics(prec):
    """The persistent kernel keeps the Monitor's running sums in registers for the whole launch and forms the statistics of finished
    episodes on the spot: after 3 1/3 episodes the statistics, the per-env episode returns and the running sums of the unfinished
    episode must equal those of 20 per-step launches."""
    from cantorrl_b200 import HedgingVecEnv, sim
    from cantorrl_b200.stats import EpisodeStats
    T, n, k = 6, 700, 20
    book = sim.generate_paths_and_options(n, n_steps=T, model="gbm", seed=2)
    tape = _tape(k, n, 9)
    res = []
    for many in (True, False):
        st = EpisodeStats("cuda", hist_bins=128, hist_max=4.0)
        env = HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", monitor=True, stats=st, precision=prec, **KW)
        env.reset()
        if many:
            env.step_many(tape)
        else:
            for t in range(k):
                env.step(tape[t])
        res.append((st.sums.cpu().numpy().copy(), st.hist.cpu().numpy().copy(), env._ep_return.clone(), env._ep_acc.clone()))
    assert res[0][0][0] == n * (k // T) and res[0][0][11] == n * k
    np.testing.assert_array_equal(res[0][1], res[1][1])                             # histogram counts
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-12)                    # float64 sums: summation order only
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
@pytest.mark.parametrize("model", ["gbm", "heston"])
def test_on_the_fly_step_equals_replay_of_the_simulated_book(prec, model):
    """Three episodes of 1003 envs: the on-the-fly kernel (no book) against the replay kernel over the book K1 writes for the
    same seed -- episode e of env g is path e * total_envs + g.  Bit-equal observations, rewards, dones, info and positions."""
    from cantorrl_b200 import HedgingVecEnv, sim
    T, n, E = 9, 1003, 3
    simkw = dict(model=model, seed=77, s0=100.0, v0=0.04, sigma_v=0.9 if model == "heston" else 0.5)      # sigma_v 0.9: v hits 0
    book = sim.generate_paths_and_options(E * n, n_steps=T, **simkw)
    replay = HedgingVecEnv(data=book, num_envs=n, precision=prec, episode_sampler="pcg64", record_info=True, **KW)
    fly = HedgingVecEnv(simulate=dict(simkw, n_steps=T), num_envs=n, total_envs=n, precision=prec, record_info=True, **KW)
    o_r = replay.reset(path_idx=np.arange(n))
    o_f = fly.reset()
    assert torch.equal(o_r, o_f)
    tape = _tape(E * T, n, 21)
    for t in range(E * T):
        replay.set_next_paths(np.arange(n) + n * min(t // T + 1, E - 1))
        o_r, r_r, d_r, i_r = replay.step(tape[t])
        o_f, r_f, d_f, i_f = fly.step(tape[t])
        assert torch.equal(d_r, d_f) and bool(d_f.all()) == ((t + 1) % T == 0)
        assert torch.equal(r_r, r_f), f"reward, step {t}"
        assert torch.equal(o_r, o_f), f"obs, step {t}"
        if bool(d_r.any()):
            assert torch.equal(i_r["terminal_observation"], i_f["terminal_observation"])
        for key in ("step_pnl_total", "slippage_cost", "portfolio_value", "actual_calls_traded", "put_contracts"):
            assert torch.equal(i_r[key], i_f[key]), f"{key}, step {t}"
        if t < E * T - 1:
            assert torch.equal(fly.current_stock_price, replay.current_stock_price)
    assert int(fly.current_episode_idx.min()) == E and int(fly.current_step.max()) == 0


def test_on_the_fly_sharded_and_without_auto_reset():
    """Two shards of a 600-env population reproduce the unsharded run (global path indices); with auto_reset=False a finished env
    repeats its terminal observation with reward 0 when stepped again, like the replay kernel."""
    from cantorrl_b200 import HedgingVecEnv
    T, n = 5, 600
    simkw = dict(model="gbm", seed=5, n_steps=T)
    full = HedgingVecEnv(simulate=simkw, num_envs=n, total_envs=n, **KW)
    lo = HedgingVecEnv(simulate=simkw, num_envs=256, total_envs=n, env_offset=0, **KW)
    hi = HedgingVecEnv(simulate=simkw, num_envs=n - 256, total_envs=n, env_offset=256, **KW)
    for e in (full, lo, hi):
        e.reset()
    tape = _tape(2 * T + 2, n, 8)
    for t in range(2 * T + 2):
        o, r, d, _ = full.step(tape[t])
        o1, r1, d1, _ = lo.step(tape[t, :256].contiguous())
        o2, r2, d2, _ = hi.step(tape[t, 256:].contiguous())
        assert torch.equal(o, torch.cat([o1, o2])) and torch.equal(r, torch.cat([r1, r2])) and torch.equal(d, torch.cat([d1, d2]))
    stay = HedgingVecEnv(simulate=simkw, num_envs=n, total_envs=n, auto_reset=False, **KW)
    stay.reset()
    for t in range(T):
        o, r, d, info = stay.step(tape[t])
    assert bool(d.all())
    term = o.clone()
    o2, r2, d2, _ = stay.step(tape[T])
    assert torch.equal(o2, term) and bool(d2.all()) and float(r2.abs().max()) == 0.0
    assert int(stay.current_step.min()) == T


def test_philox_sampler_draws_fresh_first_episodes_per_reset_and_rank_and_reseeds_reproducibly():
    from cantorrl_b200 import HedgingVecEnv, sim
    from cantorrl_b200.env import philox_episode_draw
    T, n, n_paths = 4, 5000, 100000
    book = sim.generate_paths_and_options(n_paths, n_steps=T, seed=1)
    a = HedgingVecEnv(data=book, num_envs=n, episode_sampler="philox", seed=9, env_offset=0, **KW)
    b = HedgingVecEnv(data=book, num_envs=n, episode_sampler="philox", seed=9, env_offset=n, **KW)
    a.reset()
    b.reset()
    first = a.current_episode_idx.clone()
    assert float((first == b.current_episode_idx).float().mean()) < 0.01        # two ranks: different episodes
    a.reset()
    assert float((first == a.current_episode_idx).float().mean()) < 0.01        # two resets: different episodes
    zero = torch.zeros((n, 2), device="cuda")
    for _ in range(T):
        a.step(zero)
    # the auto-reset at global step T - 1 drew with the device's counter scheme; the host mirror computes the same indices
    want = philox_episode_draw(9, np.arange(n), T - 1, n_paths)
    np.testing.assert_array_equal(a.current_episode_idx.cpu().numpy(), want)
    after = a.current_episode_idx.clone()
    a.reset(seed=9)                                                              # re-seeding restarts every counter
    assert torch.equal(a.current_episode_idx, first)
    for _ in range(T):
        a.step(zero)
    assert torch.equal(a.current_episode_idx, after)


def test_float32_info_arrays_in_fp32_mode():
    from cantorrl_b200 import HedgingVecEnv, sim
    book = sim.generate_paths_and_options(64, n_steps=5, seed=3)
    e32 = HedgingVecEnv(data=book, num_envs=64, precision="fp32", episode_sampler="same_path", record_info=True, **KW)
    e64 = HedgingVecEnv(data=book, num_envs=64, precision="fp64", episode_sampler="same_path", record_info=True, **KW)
    assert e32._info_f64.dtype == torch.float32 and e64._info_f64.dtype == torch.float64
    e32.reset()
    e64.reset()
    a = _tape(1, 64, 1)[0]
    _, _, _, i32 = e32.step(a)
    _, _, _, i64 = e64.step(a)
    for key in ("step_pnl_total", "transaction_costs_total", "portfolio_value", "cash", "reward_step", "scaled_float_call"):
        torch.testing.assert_close(i32[key].double(), i64[key], rtol=1e-4, atol=2e-2 if key in ("portfolio_value", "step_pnl_total") else 1e-6)
    assert torch.equal(i32["actual_calls_traded"], i64["actual_calls_traded"])
    d = i32[3]
    assert isinstance(d["cash"], np.float64) and d["loss_type_used"] == "abs"


def test_step_many_edge_cases_single_step_ragged_tail_and_unaligned_slabs():
    """n_steps = 1 (falls back to the per-step kernel), a 5-env population (no TMA: rows % 4 != 0), slabs whose per-step offset is
    not 16-byte aligned (n_envs % 4 != 0), and an episode length of 1 (every step ends an episode)."""
    from cantorrl_b200 import HedgingVecEnv, sim
    for n, T in ((5, 3), (1001, 1), (130, 2)):
        book = sim.generate_paths_and_options(17, n_steps=T, model="gbm", seed=8)
        a, b = (HedgingVecEnv(data=book, num_envs=n, episode_sampler="philox", seed=2, **KW) for _ in range(2))
        a.reset()
        b.reset()
        tape = _tape(7, n, n)
        o1, r1, d1 = a.step_many(tape[:1])
        o2, r2, d2 = a.step_many(tape[1:])
        for t in range(7):
            o, r, d, _ = b.step(tape[t])
            want = (o1[0], r1[0], d1[0]) if t == 0 else (o2[t - 1], r2[t - 1], d2[t - 1])
            assert torch.equal(o, want[0]) and torch.equal(r, want[1]) and torch.equal(d, want[2]), (n, T, t)
        assert torch.equal(a._core, b._core)


def test_v1_env_positional_call_of_the_reference_and_record_metrics_off_on_the_fly():
    """src/agents/test_rand_ppo.py:26-27 calls the v1 class positionally: HedgingEnv(DATA_FILE, 0.05, 1.0, 0.0, 10000, 200) binds
    loss_type=10000 (-> the abs formula, hedging_env.py:240-242) and initial_cash=200.  Same call here, same numbers as keywords."""
    import os
    import tempfile
    from cantorrl_b200 import HedgingVecEnv, sim
    book = sim.generate_paths_and_options(32, n_steps=6, model="gbm", seed=5)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "book.npz")
        book.save_npz(f)
        pos = HedgingVecEnv(f, 0.05, 1.0, 0.0, 10000, 200, version="v1", num_envs=32, episode_sampler="same_path", precision="fp64")
        kw = HedgingVecEnv(f, transaction_cost_per_contract=0.05, lambda_cost=1.0, pnl_penalty_weight=0.0, loss_type=10000,
                           initial_cash=200, version="v1", num_envs=32, episode_sampler="same_path", precision="fp64")
    assert pos.initial_cash == 200 and pos.loss_type == 10000 and pos.shares_held_fixed == 10000 and pos.slippage_bps == 0.0
    pos.reset()
    kw.reset()
    a = _tape(1, 32, 4)[0]
    o1, r1, d1, _ = pos.step(a)
    o2, r2, d2, _ = kw.step(a)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and float(pos.cash_balance.max()) <= 200.0
    with pytest.raises(TypeError):
        HedgingVecEnv(data=book, version="v1", theta_weight=1e-4)
    # record_metrics=False zeroes obs[7:11] (hedging_env_v2.py:80-81) in the on-the-fly kernel as in the replay kernel
    simkw = dict(model="gbm", seed=5, n_steps=6)
    fly = HedgingVecEnv(simulate=simkw, num_envs=32, record_metrics=False, **KW)
    rep = HedgingVecEnv(data=sim.generate_paths_and_options(32, n_steps=6, model="gbm", seed=5), num_envs=32, record_metrics=False,
                        episode_sampler="same_path", **KW)
    of, orp = fly.reset(), rep.reset()
    assert torch.equal(of, orp) and float(of[:, 7:11].abs().max()) == 0.0
    of, rf, _, _ = fly.step(a)
    orp, rr, _, _ = rep.step(a)
    assert torch.equal(of, orp) and torch.equal(rf, rr) and float(of[:, 7:11].abs().max()) == 0.0


def test_full_shard_size_the_three_step_forms_agree():
    """BASELINE configs[3] per-GPU shard (2^23 envs, 34 GB simulated book): a whole 252-step episode stepped by the on-the-fly kernel
    (no book) and by the replay kernel gives identical rewards / dones / final observations, and a 6-step action tape through the
    persistent kernel equals the same six steps launched one by one."""
    free, _ = torch.cuda.mem_get_info()
    if free < 80 * 2 ** 30:
        pytest.skip("needs ~60 GB of free HBM")
    from cantorrl_b200 import HedgingVecEnv, sim
    n, T = 1 << 23, 252
    kw = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)
    book = sim.generate_paths_and_options(n, n_steps=T, model="gbm", seed=42)
    replay = HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", **kw)
    fly = HedgingVecEnv(simulate=dict(model="gbm", seed=42, n_steps=T), num_envs=n, total_envs=n, **kw)
    assert torch.equal(replay.reset(), fly.reset())
    g = torch.Generator(device="cuda").manual_seed(3)
    acts = [torch.rand((n, 2), device="cuda", generator=g) * 2 - 1 for _ in range(4)]
    for t in range(T):
        o_r, r_r, d_r, _ = replay.step(acts[t & 3])
        o_f, r_f, d_f, _ = fly.step(acts[t & 3])
        if t in (0, 1, 100, T - 2, T - 1):
            assert torch.equal(r_r, r_f) and torch.equal(d_r, d_f), t
            if t < T - 1:                       # after the last step the two envs start different second episodes
                assert torch.equal(o_r, o_f), t
        assert bool(d_f.any()) == (t == T - 1)
    assert int(fly.current_episode_idx.min()) == 1 and int(replay.current_step.max()) == 0
    del fly
    torch.cuda.empty_cache()
    other = HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", **kw)
    other.reset()
    replay.reset()
    tape = torch.stack([acts[j & 3] for j in range(6)]).contiguous()
    o_m, r_m, d_m = other.step_many(tape)
    for t in range(6):
        o, r, d, _ = replay.step(tape[t])
        assert torch.equal(r, r_m[t]) and torch.equal(d, d_m[t]) and torch.equal(o, o_m[t]), t
    assert torch.equal(other._core, replay._core) and torch.equal(other._cash, replay._cash)
