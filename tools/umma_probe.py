#!/usr/bin/env python
"""Clocks per tcgen05.mma (M = 128, K = 16, bf16) as a function of N and of where A lives (cantor_umma_probe): the figure behind the
tensor-core actors' design.  python tools/umma_probe.py [--ts 1]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ts", type=int, default=0, help="1: A operand in tensor memory")
    ap.add_argument("--ctas", default="1,148")
    ap.add_argument("--reps", type=int, default=64)
    a = ap.parse_args()
    L = _lib.lib()
    rows = []
    for ctas in [int(x) for x in a.ctas.split(",")]:
        for k_steps in (9, 5, 1):
            for n in (16, 32, 64, 128, 256):
                i, t = C.c_double(), C.c_double()
                _lib.check(L.cantor_umma_probe(n, k_steps, a.reps, a.ts, ctas, C.byref(i), C.byref(t)), "cantor_umma_probe")
                floor = 128 * n / 256
                rows.append(dict(a_in_tmem=a.ts, ctas=ctas, k_steps=k_steps, n=n, issue_clk=round(i.value, 1), clk_per_mma=round(t.value, 1),
                                 math_floor_clk=floor, smem_bytes=(0 if a.ts else 128 * 32) + n * 32))
                print(f"A in {'TMEM' if a.ts else 'smem'}  CTAs {ctas:4d}  K-steps {k_steps}  N {n:3d}: {t.value:6.1f} clk per MMA "
                      f"(issue {i.value:6.1f}; math floor {floor:5.1f}; operand bytes from smem {rows[-1]['smem_bytes']})", flush=True)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
