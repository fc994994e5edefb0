#!/usr/bin/env python
"""Env-steps/s of the per-step API in its variants (one GPU, CUDA events): fast path, Monitor + statistics, VecNormalize,
record_info, fp64 ledger.  2^20 envs, 252-step episode sweeps through HedgingVecEnv.step (one launch per step)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import HedgingVecEnv, sim  # noqa: E402
from cantorrl_b200.stats import EpisodeStats  # noqa: E402
from cantorrl_b200.vecnorm import VecNormalize  # noqa: E402

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)


def main():
    n, T = 1 << 20, 252
    book = sim.generate_paths_and_options(n, n_steps=T, model="gbm")
    g = torch.Generator(device="cuda").manual_seed(0)
    actions = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
    out = {}
    variants = {
        "fp32_fast": dict(), "fp32_monitor_stats": dict(monitor=True, stats=EpisodeStats("cuda")), "fp32_record_info": dict(record_info=True),
        "fp64": dict(precision="fp64"), "fp64_monitor_stats": dict(precision="fp64", monitor=True, stats=EpisodeStats("cuda")),
        "fp32_vecnormalize": dict(_vn=True),
    }
    only = os.environ.get("VARIANTS")
    for name, kw in variants.items():
        if only and name not in only.split(","):
            continue
        vn = kw.pop("_vn", False)
        env = HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", **KW, **kw)
        if vn:
            env = VecNormalize(env)
        env.reset()
        for _ in range(20):
            env.step(actions)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(T):
            env.step(actions)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = dict(us_per_step=ms / T * 1e3, env_steps_per_s=n * T / ms * 1e3)
        print(name, json.dumps(out[name]), flush=True)
        del env
    print(json.dumps(out))


if __name__ == "__main__":
    main()
