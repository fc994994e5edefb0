#!/usr/bin/env python
"""GPU-side cost of a VecNormalize step: env.step + cantor_vecnorm_step captured in a CUDA graph (no Python / ctypes time)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import HedgingVecEnv, sim  # noqa: E402
from cantorrl_b200.vecnorm import VecNormalize  # noqa: E402

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)


def timed_graph(fn, steps_per_graph=21, reps=12):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(steps_per_graph):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * steps_per_graph) * 1e3


def main():
    n = int(os.environ.get("ENVS", 1 << 20))
    book = sim.generate_paths_and_options(n, model="gbm", n_steps=252)
    actions = (torch.rand(n, 2, device="cuda") * 2 - 1).float()
    out = {}
    env = HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", **KW)
    env.reset()
    out["step_us"] = timed_graph(lambda: env.step(actions))
    vn = VecNormalize(HedgingVecEnv(data=book, num_envs=n, episode_sampler="same_path", **KW))
    vn.reset()
    out["step_keep_l2_us"] = timed_graph(lambda: vn.venv.step(actions))
    out["step_vecnorm_us"] = timed_graph(lambda: vn.step(actions))
    out["vecnorm_only_us"] = out["step_vecnorm_us"] - out["step_keep_l2_us"]
    out["env_steps_per_s_vecnorm"] = n / out["step_vecnorm_us"] * 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
