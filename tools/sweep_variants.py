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
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--e2e-steps", "0", "--no-cpu-baseline", *extra],
                       env=env, capture_output=True, text=True)
    name = os.path.basename(os.path.dirname(lib))
    try:
        j = json.loads(p.stdout.strip().splitlines()[-1])
        print(f"{name:28s} value={j['value']:.4e} launch_us={j['roofline']['launch_us']:.2f} frac={j['roofline']['frac']:.3f}", flush=True)
    except Exception:
        print(name, "FAILED", p.stdout[-300:], p.stderr[-600:], flush=True)
