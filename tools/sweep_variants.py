#!/usr/bin/env python
"""Run bench.py against every library under build/variants/ (one subprocess each) and print value / launch time."""
import glob
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
extra = sys.argv[1:] or ["--steps", "20"]
for lib in sorted(glob.glob(os.path.join(root, "build", "variants", "*", "libcantor_hedge.so"))):
    env = dict(os.environ, CANTOR_HEDGE_LIB=lib)
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--e2e-steps", "0", "--no-cpu-baseline", "--rollout-steps", "0",
                        "--mlp-rollout-steps", "0", "--lstm-rollout-steps", "0", "--book-strikes", "0", "--rbergomi-paths", "0",
                        "--l2free-envs", "0", "--no-forms", *extra],
                       env=env, capture_output=True, text=True)
    name = os.path.basename(os.path.dirname(lib))
    try:
        j = json.loads(p.stdout.strip().splitlines()[-1])
        print(f"{name:28s} value={j['value']:.4e} sweep_us={j['roofline']['launch_us']:.1f} frac={j['roofline']['frac']:.3f} "
              f"per_step_launch_us={j['roofline']['per_step_kernel']['launch_us']:.2f}", flush=True)
    except Exception:
        print(name, "FAILED", p.stdout[-300:], p.stderr[-600:], flush=True)
