#!/usr/bin/env python
"""Raw page-locked copy ceilings of this box, per GPU and for several GPUs at once (one thread per GPU).

    python tools/host_copy_probe.py [--gpus 1,2,4,8] [--envs 1048576]

For each GPU count: the per-step traffic of HostVecEnv.step at `envs` envs per GPU (57 B out, 8 B in per env), with and without
the per-round stream synchronisation a gym step needs, D2H alone, H2D alone.  JSON lines.
"""
import argparse
import json
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200.host_env import host_copy_probe  # noqa: E402


def run(gpus, **kw):
    res = [None] * gpus
    bar = threading.Barrier(gpus)

    def work(g):
        host_copy_probe(device=g, **dict(kw, seconds=0.05))      # context + allocation warm-up
        bar.wait()
        res[g] = host_copy_probe(device=g, **kw)

    th = [threading.Thread(target=work, args=(g,)) for g in range(gpus)]
    [t.start() for t in th]
    [t.join() for t in th]
    return dict(gpus=gpus, d2h_gbs_total=sum(r["d2h_gbs"] for r in res), h2d_gbs_total=sum(r["h2d_gbs"] for r in res),
                d2h_gbs_per_gpu=[round(r["d2h_gbs"], 2) for r in res], rounds_per_s=[round(r["rounds_per_s"], 1) for r in res])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--seconds", type=float, default=0.5)
    a = ap.parse_args()
    n = a.envs
    for g in [int(x) for x in a.gpus.split(",")]:
        for name, kw in (("step_traffic_sync", dict(d2h_bytes=57 * n, h2d_bytes=8 * n, n_chunks=8, sync_each_round=True)),
                         ("step_traffic_nosync", dict(d2h_bytes=57 * n, h2d_bytes=8 * n, n_chunks=8, sync_each_round=False)),
                         ("step_traffic_sync_16chunks", dict(d2h_bytes=57 * n, h2d_bytes=8 * n, n_chunks=16, sync_each_round=True)),
                         ("step_traffic_sync_3chunks", dict(d2h_bytes=57 * n, h2d_bytes=8 * n, n_chunks=3, sync_each_round=True)),
                         ("d2h_only", dict(d2h_bytes=57 * n, h2d_bytes=0, n_chunks=1, sync_each_round=False)),
                         ("h2d_only", dict(d2h_bytes=0, h2d_bytes=57 * n, n_chunks=1, sync_each_round=False))):
            r = run(g, seconds=a.seconds, **kw)
            r["what"] = name
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
