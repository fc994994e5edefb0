#!/usr/bin/env python
"""One of the step-kernel forms at one size (wrapper over bench.ring_step_bench / bench.side_kernel_bench), for ncu captures.

    python tools/bench_modes.py --mode replay|many|sim [--envs N] [--precision fp32|fp64] [--model gbm|heston] [--sweeps K]
    python tools/bench_modes.py --mode forms
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cantorrl_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="replay")
    ap.add_argument("--envs", type=int, default=1 << 23)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--model", default="gbm")
    ap.add_argument("--sweeps", type=int, default=2)
    ap.add_argument("--steps", type=int, default=252)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream(dev)
    peak = 6565.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        pass
    if a.mode == "forms":
        print(json.dumps(bench.side_kernel_bench(dev, stream, peak)))
        return
    out = bench.ring_step_bench(_lib.lib(), dev, stream, a.envs, a.steps, a.precision, peak, a.mode, sweeps=a.sweeps, model=a.model)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
