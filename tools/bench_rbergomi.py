#!/usr/bin/env python
"""Throughput of the rough-Bergomi generator + nested-MC pricer (one GPU, CUDA events)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import sim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=252)
    ap.add_argument("--n-mc", type=int, default=5000)
    ap.add_argument("--days-per-launch", type=int, default=32)
    ap.add_argument("--no-tc", action="store_true", help="float32 FFMA filter instead of the tcgen05 one")
    a = ap.parse_args()
    base = (496.48, 0.02903, 0.4656, 1.985, -0.2022)
    sim.generate_rbergomi_paths_and_options(64, base_params=base, n_steps=8, n_mc=64)          # warm-up
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    rb = sim.generate_rbergomi_paths_and_options(a.paths, base_params=base, n_steps=a.steps, n_mc=a.n_mc, price=False, tensor_cores=not a.no_tc)
    e[1].record()
    for t0 in range(0, a.steps, a.days_per_launch):
        rb.price_days(t0, min(a.steps, t0 + a.days_per_launch))
    e[2].record()
    torch.cuda.synchronize()
    outer_ms, price_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    inner = a.paths * a.steps * 2 * a.n_mc
    out = dict(tensor_cores=not a.no_tc, paths=a.paths, steps=a.steps, n_mc=a.n_mc, outer_ms=outer_ms, price_ms=price_ms,
               pricings_per_s=a.paths * a.steps * 2 / price_ms * 1e3, inner_paths_per_s=inner / price_ms * 1e3,
               inner_path_steps_per_s=inner * 30 / price_ms * 1e3,
               reference_workload_seconds=(100000 * 252 * 2 * 5000) / (inner / price_ms * 1e3),
               mean_call=float(rb.book.C[: a.steps, : a.paths].mean()), mean_put=float(rb.book.P[: a.steps, : a.paths].mean()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
