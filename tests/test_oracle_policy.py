"""Pins oracle/policy_oracle.py against the unmodified reference policy functions (build container only)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from conftest import REFERENCE, ROOT
from oracle import policy_oracle

needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")


def _load(name, path):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_gym_stub"))
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _random_obs(rng, n):
    obs = rng.uniform(-1, 1, (n, 13)).astype(np.float32)
    obs[:, 7] = rng.uniform(0, 1, n)                 # call delta
    obs[:, 9] = obs[:, 7] - 1                        # put delta
    obs[::7, 7] = 0.0                                # falls through to the put branch
    obs[::11, 9] = 0.0
    obs[:, 3] = rng.integers(-200, 201, n) / 200.0
    obs[:, 4] = rng.integers(-200, 201, n) / 200.0
    return obs


@needs_reference
def test_delta_every_step_matches_reference():
    mod = _load("ref_baselines", f"{REFERENCE}/src/agents/baselines.py")
    env = types.SimpleNamespace(max_contracts_held=200, option_contract_multiplier=100, shares_held_fixed=10000,
                                max_trade_per_step=15, action_space=types.SimpleNamespace(dtype=np.float32))
    rng = np.random.default_rng(0)
    obs = _random_obs(rng, 2000)
    got = policy_oracle.delta_every_step(obs)
    want = np.stack([mod.policy_delta_every_step(o, env) for o in obs])
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(policy_oracle.no_hedge(obs[:3]), np.stack([mod.policy_no_hedge(o, env) for o in obs[:3]]))


@needs_reference
def test_delta_benchmark_matches_reference():
    mod = _load("ref_delta_and_nothing", f"{REFERENCE}/src/benchmark/delta_and_nothing.py")
    rng = np.random.default_rng(1)
    obs = _random_obs(rng, 2000)
    pos_c = rng.integers(-200, 201, 2000)
    pos_p = rng.integers(-200, 201, 2000)
    pos_c[:100] = 0
    pos_p[:100] = 0
    got = policy_oracle.delta_benchmark(obs, pos_c, pos_p)
    want = np.stack([mod.delta_hedging_action_selector(dict(
        S_t=100.0, v_t=0.04, call_delta_atm=o[7], put_delta_atm=o[9], current_call_contracts=np.int64(c),
        current_put_contracts=np.int64(p), shares_to_hedge=10000, option_contract_multiplier=100, max_trade_per_step=15))
        for o, c, p in zip(obs, pos_c, pos_p)])
    np.testing.assert_array_equal(got, want)


def test_policies_match_golden_vectors_from_the_unmodified_reference():
    """tests/golden/policy_golden.npz (make_golden.py --policy-only): travels to the GPU box, unlike /root/reference."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "policy_golden.npz"))
    np.testing.assert_array_equal(policy_oracle.delta_every_step(z["obs"]), z["delta_every_step"])
    np.testing.assert_array_equal(policy_oracle.no_hedge(z["obs"]), z["no_hedge"])
    np.testing.assert_array_equal(policy_oracle.delta_benchmark(z["obs"], z["pos_c"], z["pos_p"]), z["delta_benchmark"])


def test_episode_statistics_definitions():
    rng = np.random.default_rng(2)
    pps = rng.normal(0, 1, (200, 30))
    cost = np.abs(rng.normal(5, 1, (200, 30)))
    rew = -np.abs(pps) * 1e-3 - cost * 1e-4
    s = policy_oracle.episode_statistics(pps, cost, rew, 30)
    # train_ppo_v2.py:520-530 written out literally
    ep = [np.abs(x) / 30 for x in pps.sum(1)]
    srt = sorted(ep)
    assert np.isclose(s["mean_abs_pnl"], np.mean(ep)) and np.isclose(s["std_abs_pnl"], np.std(ep))
    assert np.isclose(s["cvar95_abs_pnl"], np.mean(srt[int(0.95 * len(srt)):]))
    # baselines.py:49-65
    assert np.isclose(s["mean_abs_pnl_baseline"], np.mean([np.abs(r).sum() / 30 for r in pps]))
    assert np.isclose(s["mean_cost"], np.mean([r.sum() / 30 for r in cost]))


def test_recurrent_policy_oracle_reproduces_the_reference_network_on_the_shipped_weights():
    """oracle/rollout_oracle.lstm_actor_sequence against tests/golden/lstm_golden.npz: the reference's own RecurrentPPOModel
    (quantconnect/model_wrapper.py:167-204) with the shipped policy_weights.pth and normalization_stats.pkl, hidden state
    carried over 80 steps of 4 envs with an episode boundary inside (generator: tests/golden/make_golden.py --lstm-only)."""
    from oracle import rollout_oracle
    g = np.load(os.path.join(ROOT, "tests", "golden", "lstm_golden.npz"))
    w = {k: g[k] for k in ("w_ih", "w_hh", "b_ih", "b_hh", "W1", "b1", "W2", "b2", "W3", "b3")}
    assert g["done"].any() and float(g["max_abs_normalised_obs"]) < 10.0
    got = rollout_oracle.lstm_actor_sequence(g["obs"], g["done"], **w, mean=g["obs_mean"], var=g["obs_var"], bf16=False, squash="tanh")
    np.testing.assert_allclose(got, g["actions"], rtol=0, atol=2e-6)           # float32 torch vs float64 NumPy
    assert np.abs(g["actions"]).max() <= 1.0 and np.abs(g["actions"]).std() > 0.05
    # the bf16 rounding points of the tensor-core kernel stay close to the float32 network on the real weights
    q = rollout_oracle.lstm_actor_sequence(g["obs"], g["done"], **w, mean=g["obs_mean"], var=g["obs_var"], bf16=True, squash="tanh")
    assert np.abs(q - g["actions"]).max() < 0.1 and np.abs(q - g["actions"]).mean() < 1e-2


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs /root/reference (build container only)")
def test_load_reference_policy_packs_the_shipped_files_like_the_golden_weights():
    """Host-side packing only (no GPU): the shipped files and the weights carried in the golden fixture give the same image."""
    from cantorrl_b200.rollout import load_reference_policy, pack_lstm
    mf = os.path.join(REFERENCE, "quantconnect", "model_files")
    img = load_reference_policy(os.path.join(mf, "policy_weights.pth"), os.path.join(mf, "normalization_stats.pkl"), device="cpu")
    g = np.load(os.path.join(ROOT, "tests", "golden", "lstm_golden.npz"))
    w = {k: g[k] for k in ("w_ih", "w_hh", "b_ih", "b_hh", "W1", "b1", "W2", "b2", "W3", "b3")}
    want = pack_lstm(**w, obs_mean=g["obs_mean"], obs_var=g["obs_var"], device="cpu")
    assert img.dtype == want.dtype and img.shape == want.shape and bool((img == want).all())


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs /root/reference (build container only)")
def test_shipped_sb3_vecnormalize_pickle_is_readable_without_sb3_and_confirms_the_restated_conventions():
    """quantconnect/model_files/final_vecnormalize.pkl (SB3 2.6.0's VecNormalize.save) read with SB3 absent.  It is the one
    artefact of SB3's VecNormalize the reference ships, and it pins the conventions oracle/vecnorm_oracle.py restates from
    documentation: count starts at 1e-4, the observation statistics see one extra batch (the reset) that the return
    statistics do not, clip_obs = clip_reward = 10, epsilon = 1e-8; its mean / var are the shipped normalization_stats.pkl."""
    import pickle
    from cantorrl_b200.vecnorm import read_sb3_vecnormalize
    mf = os.path.join(REFERENCE, "quantconnect", "model_files")
    st = read_sb3_vecnormalize(os.path.join(mf, "final_vecnormalize.pkl"))
    with open(os.path.join(mf, "normalization_stats.pkl"), "rb") as f:
        shipped = pickle.load(f)
    assert np.array_equal(st["obs_mean"], np.asarray(shipped["obs_mean"])) and np.array_equal(st["obs_var"], np.asarray(shipped["obs_var"]))
    assert st["clip_obs"] == 10.0 and st["clip_reward"] == 10.0 and st["epsilon"] == 1e-8 and st["norm_obs"] and st["norm_reward"]
    steps = st["ret_count"] - 1e-4
    assert abs(steps - round(steps)) < 1e-6 and steps > 1e5                                   # RunningMeanStd(epsilon=1e-4).count
    n_envs = st["obs_count"] - st["ret_count"]
    assert abs(n_envs - round(n_envs)) < 1e-6 and 1 <= round(n_envs) <= 8                     # the reset batch: obs only
    with pytest.raises(Exception):
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".pkl") as f:                                 # anything but numpy / builtins is refused
            pickle.dump(os.path.join, f)
            f.flush()
            read_sb3_vecnormalize(f.name)


def _shipped_mlp():
    """The plain 13-64-64-2 MLP built from the reference's shipped head behind the fixture's fixed 13 -> 128 projection."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "mlp_golden.npz"))
    l = np.load(os.path.join(ROOT, "tests", "golden", "lstm_golden.npz"))
    W1 = (l["W1"].astype(np.float64) @ g["projection"].astype(np.float64)).astype(np.float32)          # Linear(128, 64) o P
    return g, dict(W1=W1, b1=l["b1"], W2=l["W2"], b2=l["b2"], W3=l["W3"], b3=l["b3"]), l["obs_mean"], l["obs_var"]


def test_plain_mlp_oracle_reproduces_the_reference_head_on_the_shipped_weights():
    """oracle/rollout_oracle.mlp_actor against tests/golden/mlp_golden.npz: the reference's own mlp_extractor.policy_net +
    action_net + tanh (quantconnect/model_wrapper.py:177-185, 202) with policy_weights.pth, fed from a fixed 13 -> 128 projection
    of the normalised observation (:131) -- BASELINE configs[4]'s MLP, no longer 'parity unpinned'
    (generator: tests/golden/make_golden.py --mlp-only)."""
    from oracle import rollout_oracle
    g, w, mean, var = _shipped_mlp()
    got = rollout_oracle.mlp_actor(g["obs"], **w, mean=mean, var=var, squash="tanh", obs_clip=np.inf)
    np.testing.assert_allclose(got, g["actions"], rtol=0, atol=3e-6)           # float32 torch (two matmuls) vs float64 NumPy (one)
    assert (np.abs(g["actions"]) > 0.99).mean() < 0.05 and g["actions"].std() > 0.3
    q = rollout_oracle.mlp_actor_bf16(g["obs"], **w, mean=mean, var=var, squash="tanh", obs_clip=np.inf)
    assert np.abs(q - g["actions"]).max() < 2e-2 and np.abs(q - g["actions"]).mean() < 3e-3
    # SB3's own conventions (clip the means to the Box, clip the normalised observation at 10) differ where they should
    c = rollout_oracle.mlp_actor(g["obs"], **w, mean=mean, var=var)
    big = np.abs(g["action_means"]) < 0.3
    np.testing.assert_allclose(c[big], g["action_means"][big], atol=3e-6)
