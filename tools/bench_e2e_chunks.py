#!/usr/bin/env python
"""e2e (host buffers) env-steps/s of HostVecEnv.step for several chunk counts (one GPU)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200.host_env import HostVecEnv  # noqa: E402

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)
n, T = 1 << 20, 252
rng = np.random.default_rng(0)
for chunks in [int(x) for x in (sys.argv[1:] or ["2", "4", "8", "16", "32", "64"])]:
    env = HostVecEnv(num_envs=n, simulate=dict(num_paths=n, n_steps=T, model="gbm", seed=42), n_chunks=chunks, **KW)
    a = env.pin(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
    env.reset()
    for _ in range(10):
        env.step(a)
    t0 = time.perf_counter()
    for _ in range(100):
        env.step(a)
    dt = time.perf_counter() - t0
    print(f"chunks={chunks:3d}  {dt / 100 * 1e3:.3f} ms/step  {n * 100 / dt:.3e} env-steps/s  D2H {57 * n * 100 / dt / 1e9:.1f} GB/s", flush=True)
    env.close()
