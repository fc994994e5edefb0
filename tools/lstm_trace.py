#!/usr/bin/env python
"""Timeline of the recurrent actor's three actors (group 0, group 1, issuer) in CTA 0, from a -DLSTM_TRACE=<step> build:
CANTOR_HEDGE_LIB=build/variants/trace/libcantor_hedge.so python tools/lstm_trace.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import _lib  # noqa: E402
from cantorrl_b200.rollout import HedgingRollout, pack_lstm  # noqa: E402

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)
g = np.random.default_rng(0)
kk = 1 / np.sqrt(128)
wl = pack_lstm(g.uniform(-kk, kk, (512, 13)) * 3, g.uniform(-kk, kk, (512, 128)) * 2, g.uniform(-kk, kk, 512), g.uniform(-kk, kk, 512),
               g.normal(0, .15, (64, 128)), g.normal(0, .1, 64), g.normal(0, .2, (64, 64)), g.normal(0, .1, 64),
               g.normal(0, .3, (2, 64)), g.normal(0, .1, 2), g.normal(0, .2, 13), g.uniform(.05, 2, 13))
ro = HedgingRollout(simulate=dict(model="gbm", n_steps=252), num_envs=148 * 256, **KW)
ro.run(16, "lstm_bf16", mlp=wl, stats=ro.new_stats())
torch.cuda.synchronize()
lib = C.CDLL(os.environ["CANTOR_HEDGE_LIB"])
out = np.zeros((3, 256, 2), np.int64)
cnt = np.zeros(3, np.int32)
assert lib.cantor_debug_lstm_trace(C.c_void_p(out.ctypes.data), C.c_void_p(cnt.ctypes.data)) == 0
ev = []
for a in range(3):
    for i in range(int(cnt[a])):
        ev.append((int(out[a, i, 1]), a, int(out[a, i, 0])))
ev.sort()
t0 = ev[0][0]
names = {1: "forward in", 2: "x published", 30: "h published", 31: "L1 done", 32: "a2 published", 33: "L2 done", 34: "a2 published", 35: "L3 done"}
for t, a, tag in ev:
    if a < 2:
        n = names.get(tag) or (f"product {tag - 10} seen" if 10 <= tag < 20 else f"epilogue {tag - 20} done")
    else:
        if 100 <= tag < 120:
            q, r = divmod(tag - 100, 2)
            n = f"gate product {q}" + (" issued" if r else " ready")
        else:
            n = f"head layer {tag - 119} ready"
    print(f"{t - t0:8d}  {'  ' * 20 * a}{['G0', 'G1', 'IS'][a]} {n}")
