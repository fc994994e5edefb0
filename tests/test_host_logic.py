"""CPU tests of host-side logic that needs no GPU: reference-style argument binding (both env versions), the host mirror of
the device's episode draw, the staged-reference runner that bench.py times as the CPU arm."""
import os

import numpy as np
import pytest

from conftest import REFERENCE


def test_v2_and_v1_positional_order_follow_the_reference_signatures():
    """src/env/hedging_env_v2.py:10-22 vs src/env/hedging_env.py:10-20: v1 has no theta_weight / slippage_bps, so the 5th
    positional argument is loss_type there.  src/agents/test_rand_ppo.py:26-27 calls the v1 class positionally."""
    from cantorrl_b200.env import bind_reference_arguments
    v2 = bind_reference_arguments("v2", ("f.npz", 0.65, 1.0, 0.01, 2e-4, 1.0, "mse", 5.0), {})
    assert (v2["theta_weight"], v2["slippage_bps"], v2["loss_type"], v2["initial_cash"]) == (2e-4, 1.0, "mse", 5.0)
    v1 = bind_reference_arguments("v1", ("f.npz", 0.05, 1.0, 0.0, 10000, 200), {})          # test_rand_ppo.py:26-27
    assert (v1["transaction_cost_per_contract"], v1["lambda_cost"], v1["pnl_penalty_weight"]) == (0.05, 1.0, 0.0)
    assert (v1["loss_type"], v1["initial_cash"], v1["shares_to_hedge"], v1["max_contracts_held_per_type"]) == (10000, 200, 10000, 200)
    assert v1["theta_weight"] == 0.0 and v1["slippage_bps"] == 0.0
    assert bind_reference_arguments("v1", (), {})["transaction_cost_per_contract"] == 0.05
    assert bind_reference_arguments("v2", (), {})["transaction_cost_per_contract"] == 0.65
    with pytest.raises(TypeError, match="theta_weight"):
        bind_reference_arguments("v1", ("f.npz",), {"theta_weight": 1e-4})
    with pytest.raises(TypeError):
        bind_reference_arguments("v2", tuple(range(14)), {})


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("version,fn", [("v2", "hedging_env_v2.py"), ("v1", "hedging_env.py")])
def test_bound_signatures_equal_the_reference_signatures(version, fn):
    """Names, order and defaults read off the unmodified reference class with inspect."""
    import inspect
    from cantorrl_b200 import env as E
    from oracle import ref_runner
    if not ref_runner.staged():
        from oracle import stage_ref
        stage_ref.stage(REFERENCE)
    ref = inspect.signature(ref_runner.reference_env_class(version).__init__)
    ours = inspect.signature(E._ref_signature_v1 if version == "v1" else E._ref_signature_v2)
    ref_params = [(p.name, p.default) for p in list(ref.parameters.values())[1:]]          # drop self
    our_params = [(p.name, p.default) for p in ours.parameters.values()]
    assert [n for n, _ in ref_params] == [n for n, _ in our_params]
    assert ref_params[1:] == our_params[1:]              # data_file_path is required in the reference, optional here (data=)


def test_host_mirror_of_the_device_episode_draw():
    """cantorrl_b200.env.philox_episode_draw against the oracle's Philox4x32-10 (pinned on Random123 known answers)."""
    from cantorrl_b200.env import philox_episode_draw
    from oracle import sim_oracle
    g = np.arange(1000, dtype=np.int64) + (1 << 33) + 5
    for counter in (0, 3, 251, -1, -2):
        c = counter & (2 ** 64 - 1)
        ctr = np.stack([g & 0xFFFFFFFF, g >> 32, np.full_like(g, c & 0xFFFFFFFF), np.full_like(g, ((c >> 32) ^ 0x52455345) & 0xFFFFFFFF)], 1)
        x = sim_oracle.philox4x32_10(ctr.astype(np.uint32), np.array([42, 7], np.uint32))
        want = (x[:, 0].astype(np.uint64) * np.uint64(12345) >> np.uint64(32)).astype(np.int32)
        np.testing.assert_array_equal(philox_episode_draw(42 + (7 << 32), g, counter, 12345), want)
    a, b = philox_episode_draw(1, g, -1, 10 ** 6), philox_episode_draw(1, g, -2, 10 ** 6)
    assert (a == b).mean() < 0.01 and a.min() >= 0 and a.max() < 10 ** 6


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_staged_reference_is_the_unmodified_reference_and_runs():
    """oracle/stage_ref.py copies byte for byte; oracle/ref_runner.py steps the staged class; the scalar port agrees with it."""
    from oracle import ref_runner, stage_ref
    manifest = stage_ref.stage(REFERENCE)
    assert ref_runner.staged()
    for name, m in manifest.items():
        assert stage_ref.sha256(os.path.join(REFERENCE, m["source"])) == m["sha256"]
    tmp, npz = ref_runner.tmp_npz(32, 12, 1)
    n, el, acc = ref_runner.step_reference_env(npz, 3, env_steps=40)
    assert n == 40 and np.isfinite(acc) and acc < 0
    # the same 40 steps through the port that bench.py reports next to it
    from oracle.hedge_oracle import EnvParams
    from oracle.hedge_scalar import ScalarEnv
    z = np.load(npz)
    env = ScalarEnv(z["paths"], z["volatilities"], z["call_prices_atm"], z["put_prices_atm"], EnvParams(**ref_runner.TRAIN_KW), seed=3)
    acts = np.random.default_rng(3).uniform(-1, 1, (4096, 2)).astype(np.float32)
    acts[:, 1] = 0.0
    env.reset()
    tot = 0.0
    for a in acts[:40]:
        _, r, term, _, _ = env.step(a)
        tot += float(r)
        if term:
            env.reset()
    assert tot == acc
    tmp.cleanup()
