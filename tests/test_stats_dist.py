"""Host-side logic of the multi-GPU path on CPU: sharding by global env index, the statistics all-reduce (gloo,
world_size 2) and the reference's statistics recovered from the reduced accumulators."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bs_oracle, policy_oracle, rollout_oracle, sim_oracle
from oracle.hedge_oracle import EnvParams

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)
TOTAL, T, STEPS, BINS, HMAX = 24, 6, 15, 1024, 4.0


def test_shard_covers_the_population_exactly_once():
    from cantorrl_b200.distributed import shard
    for total in (0, 1, 7, 8, 1 << 20, (1 << 26) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard(10, 2, 2)


def _oracle_shard(offset, count):
    """What one rank's rollout kernel would accumulate, computed by the oracle."""
    S, V = sim_oracle.heston_paths(3, np.arange(40), T)
    C, P = bs_oracle.atm_book(S.astype(np.float64), V.astype(np.float64))
    out = rollout_oracle.run_rollout(S, V, C, P, EnvParams(**KW), "random", count, STEPS, offset, TOTAL, seed=5)
    vec, b = rollout_oracle.stats_vector(out["ep_pps"], out["ep_cost"], out["ep_reward"], T)
    return out, vec, b


def _fill(stats, vec, b, n_env_steps):
    stats.sums[:11] = torch.from_numpy(vec)
    stats.sums[11] = n_env_steps
    bins = np.minimum((b.astype(np.float32) * np.float32(BINS / HMAX)).astype(np.int64), BINS - 1)
    stats.hist.add_(torch.from_numpy(np.bincount(bins, minlength=BINS)))
    stats.hist_sum.add_(torch.from_numpy(np.bincount(bins, weights=b, minlength=BINS)))


def _worker(rank, world, init_file, q):
    try:
        from cantorrl_b200.distributed import shard
        from cantorrl_b200.stats import EpisodeStats
        import datetime
        dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=60))
        off, cnt = shard(TOTAL, rank, world)
        _, vec, b = _oracle_shard(off, cnt)
        st = EpisodeStats("cpu", hist_bins=BINS, hist_max=HMAX)
        _fill(st, vec, b, cnt * STEPS)
        st.all_reduce()
        if rank == 0:
            q.put((st.sums.numpy().copy(), st.hist.numpy().copy(), st.result()))
        dist.barrier()
        dist.destroy_process_group()
    except BaseException as e:          # never leave the parent waiting on the queue
        if rank == 0:
            q.put(("error", repr(e), None))
        raise


def test_all_reduced_statistics_equal_the_single_shard_statistics():
    from cantorrl_b200.stats import EpisodeStats
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    with tempfile.TemporaryDirectory() as d:
        procs = [ctx.Process(target=_worker, args=(r, 2, os.path.join(d, "rdzv"), q)) for r in range(2)]
        for p in procs:
            p.start()
        import time
        t0 = time.time()
        while q.empty() and time.time() - t0 < 120 and any(p.is_alive() for p in procs):
            time.sleep(0.2)
        assert not q.empty(), "workers died or timed out without a result"
        sums, hist, res = q.get()
        assert not isinstance(sums, str), hist
        for p in procs:
            p.join(60)
            if p.is_alive():
                p.kill()
            assert p.exitcode == 0
    out, vec, b = _oracle_shard(0, TOTAL)
    np.testing.assert_allclose(sums[:11], vec, rtol=1e-12)
    assert sums[11] == TOTAL * STEPS and hist.sum() == vec[0] == TOTAL * (STEPS // T)
    # the statistics the reference computes from Python lists (train_ppo_v2.py:520-530, baselines.py:63-65)
    want = policy_oracle.episode_statistics(out["ep_pps"], out["ep_cost"], out["ep_reward"], T)
    for k in ("mean_abs_pnl", "std_abs_pnl", "mean_cost", "std_cost", "mean_reward", "std_reward",
              "mean_abs_pnl_baseline", "std_abs_pnl_baseline", "mean_signed_pnl"):
        np.testing.assert_allclose(res[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)
    assert res["n_episodes"] == want["n_episodes"]
    # histogram CVaR: exact for whole bins, one bin width at worst for the straddling bin
    assert abs(res["cvar95_abs_pnl"] - want["cvar95_abs_pnl"]) <= HMAX / BINS
    single = EpisodeStats("cpu", hist_bins=BINS, hist_max=HMAX)
    _fill(single, vec, b, TOTAL * STEPS)
    assert single.all_reduce() == []                       # no process group: no-op
    assert np.array_equal(single.hist.numpy(), hist)


def test_cvar_from_histogram_matches_sorted_definition():
    from cantorrl_b200.stats import EpisodeStats
    rng = np.random.default_rng(0)
    for n in (1, 19, 20, 21, 1000, 4567):
        b = np.abs(rng.normal(0.2, 0.3, n))
        b[:3] = 5.0                                        # beyond hist_max: clamps to the last bin, sum kept exact
        st = EpisodeStats("cpu", hist_bins=4096, hist_max=2.0)
        bins = np.minimum((b * (4096 / 2.0)).astype(np.int64), 4095)
        st.hist.add_(torch.from_numpy(np.bincount(bins, minlength=4096)))
        st.hist_sum.add_(torch.from_numpy(np.bincount(bins, weights=b, minlength=4096)))
        sb = np.sort(b)
        want = sb[int(0.95 * n):].mean()
        assert abs(st.cvar95() - want) <= 2.0 / 4096 + 1e-12, n
    empty = EpisodeStats("cpu", hist_bins=16, hist_max=1.0)
    assert np.isnan(empty.cvar95()) and np.isnan(empty.result()["mean_abs_pnl"])


def test_rollout_oracle_free_running_is_self_consistent():
    """Teacher-forcing the oracle with its own actions reproduces the free-running rollout; stats vector matches."""
    S, V = sim_oracle.gbm_paths(1, np.arange(9), 5)
    C, P = bs_oracle.atm_book(S.astype(np.float64), V.astype(np.float64))
    p = EnvParams(**KW)
    free = rollout_oracle.run_rollout(S, V, C, P, p, "delta_every_step", 14, 12)
    forced = rollout_oracle.run_rollout(S, V, C, P, p, "delta_every_step", 14, 12, forced_actions=free["actions"])
    for k in ("obs", "reward", "done", "policy_actions"):
        assert np.array_equal(free[k], forced[k])
    assert free["done"].sum() == 14 * 2 and free["ep_pps"].shape == (28, 5)
    vec, b = rollout_oracle.stats_vector(free["ep_pps"], free["ep_cost"], free["ep_reward"], 5)
    s = policy_oracle.episode_statistics(free["ep_pps"], free["ep_cost"], free["ep_reward"], 5)
    assert np.isclose(vec[3] / vec[0], s["mean_abs_pnl"]) and np.isclose(vec[5] / vec[0], s["mean_cost"])
    u = rollout_oracle.uniform_actions(7, np.arange(1000), 3)
    assert u.dtype == np.float32 and u.min() >= -1 and u.max() < 1 and abs(u.mean()) < 0.05


def test_gpu_local_cpu_list_from_sysfs(tmp_path):
    """cantorrl_b200.distributed: the CPU set a rank binds to comes from the GPU's sysfs ``local_cpulist``."""
    from cantorrl_b200.distributed import _parse_cpulist, gpu_local_cpus
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("\n") == set()
    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    (d / "local_cpulist").write_text("0-55,112-167\n")
    cpus = gpu_local_cpus("0000:1B:00.0", sysfs=str(tmp_path))
    assert len(cpus) == 112 and 55 in cpus and 56 not in cpus and 167 in cpus
    assert gpu_local_cpus("0000:9a:00.0", sysfs=str(tmp_path)) == set()       # unknown device: no binding, no error


def test_numa_bind_is_a_no_op_without_a_gpu_or_when_disabled(monkeypatch):
    import os
    from cantorrl_b200.distributed import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None                                   # no CUDA device here: must not raise
    monkeypatch.setenv("CANTOR_NO_NUMA_BIND", "1")
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
