"""GPU tests of the fused Monitor (episode return / length / statistics in the step kernel) and the device VecNormalize."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import bs_oracle, rollout_oracle, sim_oracle
from oracle.hedge_oracle import EnvParams, OracleVecEnv
from oracle.vecnorm_oracle import VecNormalizeOracle

pytestmark = pytest.mark.gpu
KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)


def _data(n_paths=40, T=9, seed=2):
    S, V = sim_oracle.heston_paths(seed, np.arange(n_paths), T)
    C, P = bs_oracle.atm_book(S.astype(np.float64), V.astype(np.float64))
    return S, V, C.astype(np.float32), P.astype(np.float32)


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
def test_monitor_returns_lengths_and_statistics_match_oracle(prec):
    from cantorrl_b200 import HedgingVecEnv
    from cantorrl_b200.stats import EpisodeStats
    S, V, C, P = _data()
    n, T, steps = 300, S.shape[1] - 1, 31
    stats = EpisodeStats("cuda", hist_bins=256, hist_max=4.0)
    env = HedgingVecEnv(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=n, precision=prec,
                        episode_sampler="same_path", monitor=True, stats=stats, **KW)
    orc = OracleVecEnv(S, V, C, P, EnvParams(**KW), n)
    idx = np.arange(n) % S.shape[0]
    env.reset(path_idx=idx)
    orc.reset(idx)
    rng = np.random.default_rng(4)
    ep_r = np.zeros(n)
    cur = {k: np.zeros((n, T)) for k in ("pps", "cost", "reward")}
    fin = {k: [] for k in cur}
    rtol = 1e-6 if prec == "fp64" else 1e-4
    for g in range(steps):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        t = orc.step_count.copy()
        _, r_ref, d_ref, _, info = orc.step_autoreset(a, orc.idx.copy())
        _, r, d, infos = env.step(a)
        ep_r += r_ref
        rows = np.arange(n)
        cur["pps"][rows, t], cur["cost"][rows, t], cur["reward"][rows, t] = info["per_share_step_pnl"], info["transaction_costs_total"], r_ref
        assert np.array_equal(d.cpu().numpy(), d_ref)
        if d_ref.any():
            ep = infos["episode"]
            np.testing.assert_allclose(ep["r"].cpu().numpy()[d_ref], ep_r[d_ref], rtol=rtol, atol=1e-6)
            assert (ep["l"].cpu().numpy()[d_ref] == T).all()
            one = infos[int(np.nonzero(d_ref)[0][0])]
            assert set(one["episode"]) == {"r", "l", "t"} and one["episode"]["l"] == T
            for k in fin:
                fin[k].append(cur[k][d_ref].copy())
            ep_r[d_ref] = 0
    want, _ = rollout_oracle.stats_vector(*(np.concatenate(fin[k]) for k in ("pps", "cost", "reward")), T)
    sums = stats.sums.cpu().numpy()
    assert sums[0] == want[0] == n * (steps // T) and sums[11] == n * steps
    np.testing.assert_allclose(sums[1:11], want[1:], rtol=2e-4, atol=1e-4 * want[0])
    assert int(stats.hist.sum()) == int(want[0])


def test_vecnormalize_matches_oracle_over_many_steps_and_checkpoints():
    from cantorrl_b200 import HedgingVecEnv
    from cantorrl_b200.vecnorm import VecNormalize
    S, V, C, P = _data(n_paths=64, T=7)
    n, T = 1000, 7
    env = HedgingVecEnv(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=n, precision="fp64",
                        episode_sampler="same_path", **KW)
    vn = VecNormalize(env, gamma=0.95, clip_obs=5.0, clip_reward=3.0)
    orc_env = OracleVecEnv(S, V, C, P, EnvParams(**KW), n)
    orc = VecNormalizeOracle(n, gamma=0.95, clip_obs=5.0, clip_reward=3.0)
    idx = np.arange(n) % 64
    obs = vn.reset(path_idx=idx)
    want = orc.reset(orc_env.reset(idx))
    np.testing.assert_allclose(obs.cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    rng = np.random.default_rng(8)
    for g in range(20):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        o_ref, r_ref, d_ref, t_ref, _ = orc_env.step_autoreset(a, orc_env.idx.copy())
        wo, wr, wt = orc.step(o_ref, r_ref, d_ref, t_ref)
        o, r, d, infos = vn.step(a)
        np.testing.assert_allclose(o.cpu().numpy(), wo, rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(r.cpu().numpy(), wr, rtol=1e-6, atol=1e-9)
        if d_ref.any():
            np.testing.assert_allclose(infos["terminal_observation"].cpu().numpy()[d_ref], wt[d_ref], rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(vn.returns.cpu().numpy(), orc.returns, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(vn.obs_rms.mean.cpu().numpy(), orc.obs_rms.mean, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(vn.obs_rms.var.cpu().numpy(), orc.obs_rms.var, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(float(vn.ret_rms.var), orc.ret_rms.var, rtol=1e-9)
    assert abs(vn.obs_rms.count - orc.obs_rms.count) < 1e-6 and abs(vn.ret_rms.count - orc.ret_rms.count) < 1e-6
    # checkpoint / resume: statistics survive a save / load; evaluation mode stops updating them
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "vecnormalize.pkl")
        vn.save(f)
        vn2 = VecNormalize.load(f, env)
    assert torch.equal(vn2._rms[:30], vn._rms[:30]) and vn2.gamma == 0.95
    vn2.training = False
    vn2.keep_original = True
    before = vn2._rms[:30].clone()
    o_eval, r_eval, _, _ = vn2.step(torch.zeros((n, 2), device="cuda"))
    assert torch.equal(vn2._rms[:30], before)
    # evaluation mode = the apply kernel alone, deriving 1 / sqrt(var + eps) itself (quantconnect/model_wrapper.py:131)
    torch.testing.assert_close(o_eval, vn2.normalize_obs(vn2.get_original_obs()), rtol=1e-5, atol=5e-6)
    torch.testing.assert_close(r_eval, vn2.normalize_reward(vn2.get_original_reward()), rtol=1e-6, atol=1e-9)
    st = vn.export_stats()
    assert st["obs_mean"].shape == (13,) and st["obs_var"].dtype == np.float32


def test_vecnormalize_fp32_rewards_and_full_size():
    """2^20 envs, float32 rewards: normalised observations have ~zero mean / unit variance after a few steps."""
    from cantorrl_b200 import HedgingVecEnv, sim
    from cantorrl_b200.vecnorm import VecNormalize
    n = 1 << 20
    book = sim.generate_paths_and_options(n, n_steps=16, model="heston")
    vn = VecNormalize(HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", **KW))
    obs = vn.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(6):
        obs, r, d, _ = vn.step(torch.rand((n, 2), device="cuda", generator=g) * 2 - 1)
    assert bool(torch.isfinite(obs).all()) and float(obs.abs().max()) <= 10.0 and float(r.abs().max()) <= 10.0
    raw_like = obs[:, [0, 1, 2, 5, 7]]                  # columns with real cross-sectional spread
    assert float(raw_like.mean(0).abs().max()) < 0.5 and 0.3 < float(raw_like.std(0).mean()) < 3.0
    assert abs(vn.obs_rms.count - (7 * n + 1e-4)) < 1e-3


def test_vecnormalize_statistics_are_bitwise_reproducible():
    """The moments are reduced without atomics on the data and folded in a fixed order by whichever CTA finishes last:
    two identical runs give bit-identical running statistics (70 000 envs = 547 chunks over one wave of CTAs, ragged tail)."""
    from cantorrl_b200 import HedgingVecEnv, sim
    from cantorrl_b200.vecnorm import VecNormalize
    n = 70000
    book = sim.generate_paths_and_options(n, n_steps=8, model="gbm")
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = [torch.rand((n, 2), device="cuda", generator=g) * 2 - 1 for _ in range(10)]
    runs = []
    for _ in range(2):
        vn = VecNormalize(HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", **KW))
        vn.reset()
        for a in acts:
            obs, r, _, _ = vn.step(a)
        runs.append((vn._rms[:30].clone(), obs.clone(), r.clone(), vn.returns.clone()))
    for x, y in zip(*runs):
        assert torch.equal(x, y)


def test_vecnormalize_load_accepts_an_sb3_pickle():
    """``VecNormalize.load(path, venv)`` on a file in the format SB3's ``VecNormalize.save`` writes (the reference:
    train_ppo_v2.py:315-317, 449-455).  SB3 is not installed, so the pickle is produced from stand-in classes registered under
    SB3's module paths, which are removed again before loading."""
    import pickle
    import sys
    import types
    from cantorrl_b200 import HedgingVecEnv
    from cantorrl_b200.vecnorm import VecNormalize
    names = ["stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.running_mean_std",
             "stable_baselines3.common.vec_env", "stable_baselines3.common.vec_env.vec_normalize"]
    mods = {n: types.ModuleType(n) for n in names}
    rms_cls = type("RunningMeanStd", (), {"__module__": names[2]})
    vn_cls = type("VecNormalize", (), {"__module__": names[4]})
    mods[names[2]].RunningMeanStd, mods[names[4]].VecNormalize = rms_cls, vn_cls
    sys.modules.update(mods)
    try:
        rng = np.random.default_rng(4)
        o, r, v = rms_cls(), rms_cls(), vn_cls()
        o.mean, o.var, o.count = rng.normal(0, 1, 13), rng.uniform(0.1, 2, 13), 5000.0001
        r.mean, r.var, r.count = np.float64(-0.05), np.float64(0.0025), 4996.0001
        v.obs_rms, v.ret_rms, v.clip_obs, v.clip_reward, v.gamma, v.epsilon = o, r, 10.0, 10.0, 0.98, 1e-8
        v.norm_obs, v.norm_reward, v.training, v.venv = True, True, True, None
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "final_vecnormalize.pkl")
            with open(path, "wb") as f:
                pickle.dump(v, f)
            for n in names:
                del sys.modules[n]
            S, V, C, P = _data(n_paths=16, T=5)
            env = HedgingVecEnv(data=dict(paths=S, volatilities=V, call_prices_atm=C, put_prices_atm=P), num_envs=64,
                                episode_sampler="same_path", **KW)
            vn = VecNormalize.load(path, env)
    finally:
        for n in names:
            sys.modules.pop(n, None)
    np.testing.assert_array_equal(vn.obs_rms.mean.cpu().numpy(), o.mean)
    np.testing.assert_array_equal(vn.obs_rms.var.cpu().numpy(), o.var)
    assert vn.obs_rms.count == 5000.0001 and vn.ret_rms.count == 4996.0001 and vn.gamma == 0.98 and float(vn.ret_rms.var) == 0.0025
    vn.training, vn.keep_original = False, True
    vn.reset()
    obs, rew, _, _ = vn.step(torch.zeros((64, 2), device="cuda"))
    raw = vn.get_original_obs().double().cpu().numpy()
    want = np.clip((raw - o.mean) / np.sqrt(o.var + 1e-8), -10, 10)
    np.testing.assert_allclose(obs.cpu().numpy(), want, rtol=1e-5, atol=5e-6)


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
@pytest.mark.parametrize("n", [1003, 70000])
def test_vecnormalize_fused_with_the_step_kernel_equals_the_two_kernel_form(prec, n):
    """The batch moments produced inside the step kernel (per-CTA partials while the observation tile is in shared memory, then a
    two-level fold) against the stand-alone moments kernel: same running statistics up to float64 summation order, same normalised
    outputs; replay and on-the-fly envs; record_info / evaluation mode fall back to the unfused path."""
    from cantorrl_b200 import HedgingVecEnv, sim
    from cantorrl_b200.vecnorm import VecNormalize
    T = 7
    book = sim.generate_paths_and_options(n, n_steps=T, model="heston", seed=4)
    g = torch.Generator(device="cuda").manual_seed(6)
    acts = [torch.rand((n, 2), device="cuda", generator=g) * 2 - 1 for _ in range(2 * T + 3)]
    for make in (lambda **kw: HedgingVecEnv(data=book, num_envs=n, precision=prec, episode_sampler="same_path", **KW, **kw),
                 lambda **kw: HedgingVecEnv(simulate=dict(model="heston", seed=4, n_steps=T), num_envs=n, precision=prec, **KW, **kw)):
        fused, plain = VecNormalize(make(), gamma=0.97), VecNormalize(make(), gamma=0.97, fuse=False)
        assert fused._fuse is not None and plain._fuse is None
        assert VecNormalize(make(record_info=True))._fuse is None            # info output keeps the two-kernel form
        of, op = fused.reset(), plain.reset()
        assert torch.equal(of, op)
        for i, a in enumerate(acts):
            if i == 2 * T:                                                    # evaluation mode: statistics frozen, no partials written
                fused.training = plain.training = False
            of, rf, df, inf_f = fused.step(a)
            op, rp, dp, inf_p = plain.step(a)
            assert torch.equal(df, dp)
            torch.testing.assert_close(of, op, rtol=2e-6, atol=2e-6)
            torch.testing.assert_close(rf, rp, rtol=1e-6 if prec == "fp32" else 1e-12, atol=1e-9)
            torch.testing.assert_close(fused.returns, plain.returns, rtol=1e-12, atol=1e-15)
            if bool(df.any()):
                torch.testing.assert_close(inf_f["terminal_observation"][df], inf_p["terminal_observation"][dp], rtol=2e-6, atol=2e-6)
        torch.testing.assert_close(fused._rms[:30], plain._rms[:30], rtol=1e-11, atol=1e-13)
        assert abs(fused.obs_rms.count - (1e-4 + (2 * T + 1) * n)) < 1e-3
