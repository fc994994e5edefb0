#!/usr/bin/env python
"""The fused hedge-step kernel timed where nothing it touches survives in L2 between launches.

BASELINE configs[3] shard size: 2^23 envs per GPU replaying a 34 GB simulated book (state 168 MB + one path slab
134 MB >> 126 MB of L2).  Observations / rewards / dones / actions cycle through a ring of R slabs (each slab of
observations is 436 MB), so every launch reads and writes lines that left L2 long ago.  One sweep = 252 launches of
``hedge_step_kernel`` chained from C (``cantor_env_step_many`` over the ring, R steps per call).

    python tools/bench_step_l2free.py [--envs N] [--precision fp32|fp64] [--sweeps K] [--ring R] [--steps-per-sweep S]

Prints one JSON line: launch time, env-steps/s, algorithmic GB/s (137 / 157 B per env-step) and its fraction of the
measured HBM copy peak.  CANTOR_HEDGE_LIB selects a variant build of the library (tools/build_variants.sh).
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cantorrl_b200 import HedgingVecEnv, _lib, sim  # noqa: E402

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)
B_ALG = {"fp32": 137, "fp64": 157}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 23)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--sweeps", type=int, default=3)
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--steps-per-sweep", type=int, default=252)
    ap.add_argument("--episode-length", type=int, default=252)
    ap.add_argument("--tag", default=os.path.basename(os.path.dirname(os.environ.get("CANTOR_HEDGE_LIB", ""))) or "shipped")
    a = ap.parse_args()
    n, T, R, S = a.envs, a.episode_length, a.ring, a.steps_per_sweep
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    book = sim.generate_paths_and_options(n, n_steps=T, model="gbm", device=dev)
    env = HedgingVecEnv(data=book, num_envs=n, device=dev, precision=a.precision, episode_sampler="same_path", **KW)
    g = torch.Generator(device=dev).manual_seed(7)
    actions = torch.rand((R, n, 2), device=dev, generator=g) * 2 - 1
    actions[:, :, 1] = 0.0
    rdt = torch.float64 if a.precision == "fp64" else torch.float32
    obs = torch.empty((R, n, 13), dtype=torch.float32, device=dev)
    reward = torch.empty((R, n), dtype=rdt, device=dev)
    done = torch.empty((R, n), dtype=torch.uint8, device=dev)
    env.reset()
    stream = torch.cuda.current_stream(dev)

    def sweep():
        left = S
        while left > 0:
            k = min(R, left)
            _lib.check(L.cantor_env_step_many(C.byref(env._params), C.byref(env._book), C.byref(env._state), n, env._prec, k,
                                              actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), done.data_ptr(), None,
                                              C.byref(env._rule), stream.cuda_stream), "cantor_env_step_many")
            left -= k

    sweep()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.sweeps):
        sweep()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    launch_us = ms * 1e3 / (a.sweeps * S)
    peak = 6565.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        pass
    gbs = B_ALG[a.precision] * n / (launch_us * 1e-6) / 1e9
    assert bool(torch.isfinite(reward).all())
    print(json.dumps(dict(tag=a.tag, envs=n, precision=a.precision, launches=a.sweeps * S, launch_us=launch_us,
                          env_steps_per_s=n / (launch_us * 1e-6), algorithmic_gbs=gbs, frac=gbs / peak, peak=peak)), flush=True)


if __name__ == "__main__":
    main()
