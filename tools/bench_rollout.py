#!/usr/bin/env python
"""Throughput of the episode-fused rollout kernel for each policy / data source (one GPU, CUDA events)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import sim  # noqa: E402
from cantorrl_b200.rollout import HedgingRollout, pack_lstm, pack_mlp  # noqa: E402

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=252)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--policies", default="no_hedge,random,delta_every_step,mlp,mlp_bf16,lstm_bf16")
    ap.add_argument("--sources", default="gbm,heston,replay")
    ap.add_argument("--store", action="store_true")
    a = ap.parse_args()
    g = np.random.default_rng(0)
    w = pack_mlp(g.normal(0, .5, (64, 13)), g.normal(0, .1, 64), g.normal(0, .2, (64, 64)), g.normal(0, .1, 64),
                 g.normal(0, .3, (2, 64)), g.normal(0, .1, 2), g.normal(0, .2, 13), g.uniform(.05, 2, 13))
    kk = 1 / np.sqrt(128)
    wl = pack_lstm(g.uniform(-kk, kk, (512, 13)) * 3, g.uniform(-kk, kk, (512, 128)) * 2, g.uniform(-kk, kk, 512), g.uniform(-kk, kk, 512),
                   g.normal(0, .15, (64, 128)), g.normal(0, .1, 64), g.normal(0, .2, (64, 64)), g.normal(0, .1, 64),
                   g.normal(0, .3, (2, 64)), g.normal(0, .1, 2), g.normal(0, .2, 13), g.uniform(.05, 2, 13))
    out = {}
    for src in a.sources.split(","):
        if src == "replay":
            book = sim.generate_paths_and_options(a.envs, n_steps=252, model="gbm")
            ro = HedgingRollout(data=book, num_envs=a.envs, **KW)
        else:
            ro = HedgingRollout(simulate=dict(model=src, n_steps=252), num_envs=a.envs, **KW)
        for pol in a.policies.split(","):
            st = ro.new_stats()
            ro.run(a.steps, pol, mlp=wl if pol == "lstm_bf16" else w, stats=st, store=a.store)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                ro.run(a.steps, pol, mlp=wl if pol == "lstm_bf16" else w, stats=st, store=a.store)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            out[f"{src}/{pol}"] = dict(ms=round(ms, 3), env_steps_per_s=a.envs * a.steps / ms * 1e3, err_flag=float(st.sums[15]))
            print(f"{src:8s} {pol:18s} {ms:9.3f} ms  {a.envs * a.steps / ms * 1e3:.3e} env-steps/s  err={float(st.sums[15])}", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
