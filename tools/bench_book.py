#!/usr/bin/env python
"""Throughput of the float32 multi-strike book kernel (BASELINE configs[2] shape, one GPU, CUDA events)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import sim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=252)
    ap.add_argument("--strikes", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    book = sim.generate_paths_and_options(a.paths, n_steps=a.steps, model="heston", reprice=False)
    mult = np.linspace(0.9, 1.1, a.strikes).astype(np.float32)
    out = {}
    for name, kw in (("prices_book_sigma", dict(sigma="book")), ("prices_realised_sigma", dict(sigma="realised")),
                     ("prices_greeks_book_sigma", dict(sigma="book", greeks=True))):
        r = sim.reprice_book(book, mult, **kw)
        torch.cuda.synchronize()
        del r
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            r = sim.reprice_book(book, mult, **kw)     # includes torch.empty of the outputs (cached allocator after the first)
            del r
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        cells = a.paths * (a.steps + 1)
        n_out = 4 if kw.get("greeks") else 2
        bytes_alg = cells * (16 + a.strikes * n_out * 4)
        out[name] = dict(ms=round(ms, 3), path_steps_per_s=cells / ms * 1e3, reprices_per_s=cells * a.strikes / ms * 1e3,
                         algorithmic_gbs=bytes_alg / ms / 1e6, mufu_per_s=cells * (5 + 3 * a.strikes) / ms * 1e3)
        print(name, json.dumps(out[name]), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
