#!/usr/bin/env python
"""The headline form alone: cantor_env_step_many over a full 252-step action tape (ONE persistent launch per sweep), at a list of
env counts.  python tools/bench_many_full.py [--envs 1048576,947200] [--sweeps 10] [--precision fp32]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cantorrl_b200 import HedgingVecEnv, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="1048576")
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--steps", type=int, default=252)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--monitor", type=int, default=0)
    ap.add_argument("--record-metrics", type=int, default=1, help="0: the observation's greeks are zeros (fewer instructions, same bytes)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream(dev)
    L = _lib.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    T = a.steps
    for n in [int(x) for x in a.envs.split(",")]:
        data, _ = bench.synth_replay_data(n, T, 0, dev)
        env = HedgingVecEnv(data=data, num_envs=n, device=dev, precision=a.precision, episode_sampler="same_path",
                            monitor=bool(a.monitor), record_metrics=bool(a.record_metrics), **bench.ENV_KW)
        g = torch.Generator(device=dev).manual_seed(1234)
        actions = torch.rand((T, n, 2), device=dev, generator=g) * 2 - 1
        actions[:, :, 1] = 0.0
        rdt = torch.float64 if a.precision == "fp64" else torch.float32
        obs = torch.empty((T, n, 13), dtype=torch.float32, device=dev)
        reward = torch.empty((T, n), dtype=rdt, device=dev)
        done = torch.empty((T, n), dtype=torch.uint8, device=dev)
        env.reset()

        def sweep():
            _lib.check(L.cantor_env_step_many(C.byref(env._params), C.byref(env._book), C.byref(env._state), n, env._prec, T,
                                              actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), done.data_ptr(), None,
                                              C.byref(env._rule), stream.cuda_stream), "cantor_env_step_many")

        ms = bench.gpu_ms(sweep, a.sweeps, stream, dev, warm=3)
        b_alg = 81 + (4 if a.precision == "fp64" else 0) + (40 + (24 if a.precision == "fp64" else 0)) / T
        gbs = b_alg * n * T / (ms * 1e-3) / 1e9
        print(json.dumps(dict(envs=n, steps=T, ms_per_sweep=ms, env_steps_per_s=n * T / (ms * 1e-3), gbs=gbs, frac=gbs / peak,
                              checksum=float(reward[T - 1].double().sum()))), flush=True)
        del env, data, actions, obs, reward, done
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
