#!/usr/bin/env python
"""Benchmark of the fused hedge step (BASELINE.json metric: env-steps/sec, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): 2^20 synthetic GBM paths x 252 steps per GPU, one European ATM call
(put leg forced to 0), proportional transaction cost (slippage_bps = 1) + commission, the v2 training
reward weights (src/agents/train_ppo_v2.py:74-80).  One bench "step" = one full episode sweep = 252
launches of the fused hedge-step kernel over all envs (ends with the auto-reset of every env).

  value     env-steps/s with everything resident in HBM, replay mode, fp32 state (137 algorithmic B/env-step);
            actions are read from, and obs/reward/done written to, [T, n_envs, ...] rollout slabs, so every
            launch touches fresh lines (inputs >> L2).
  e2e       the same 252-step sweep through the gym-style API with HOST buffers: every step copies that
            step's actions from pinned host memory and brings obs/reward/done back to pinned host memory.
  roofline  137 B x n_envs per launch / mean launch duration (CUDA events over the timed region), against the
            measured HBM copy peak in MEASURED_PEAKS.json.
  cpu_baseline  the scalar reference-shaped port (oracle/hedge_scalar.py) on the host cores of this box.

Multi-GPU (torchrun, one rank per GPU): envs are sharded by global path index, no data-path collective
(weak scaling: 2^20 envs per GPU); time is the max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = {"fp32": 137, "fp64": 157}          # algorithmic bytes per env-step, replay mode
ENV_KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4,
              transaction_cost_per_contract=0.65, loss_type="abs")            # train_ppo_v2.py:74-80
R, DT, S0, XI = 0.04, 1 / 252, 100.0, 0.04                                      # rbergomi_sim.py:13,14,27,23


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cantor", choices=["cantor", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--episode-length", type=int, default=252)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--e2e-steps", type=int, default=2, help="episode sweeps timed through the host-buffer API")
    ap.add_argument("--rollout-steps", type=int, default=3, help="episode sweeps of the on-the-fly rollout kernel (extra)")
    ap.add_argument("--rollout-envs", type=int, default=1 << 23, help="envs per GPU of that extra (configs[3]: 64 M envs over 8 GPUs)")
    ap.add_argument("--mlp-rollout-steps", type=int, default=1000, help="steps of the MLP-policy rollout (configs[4] shape; 0 = skip)")
    ap.add_argument("--lstm-rollout-steps", type=int, default=252, help="steps of the recurrent (LSTM + MLP) policy rollout (0 = skip)")
    ap.add_argument("--book-strikes", type=int, default=8, help="strikes of the multi-strike book extra (configs[2] shape; 0 = skip)")
    ap.add_argument("--rbergomi-paths", type=int, default=512, help="paths of the rough-Bergomi nested-MC extra (x 32 days; 0 = skip)")
    ap.add_argument("--no-fused-allreduce", action="store_true", help="all-reduce the statistics with NCCL instead of in-kernel")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- CPU legs
# The CPU arm is the UNMODIFIED reference env (src/env/hedging_env_v2.py, staged byte for byte into the git-ignored
# oracle/_ref/ by oracle/stage_ref.py -- run by __graft_entry__.build() -- and loaded under oracle/_gym_stub): P forked
# processes, one env each, which is the reference's SubprocVecEnv model (src/agents/train_ppo_v2.py:131-134) without the
# pipes.  The scalar port (oracle/hedge_scalar.py) is timed next to it as a second, labelled number.
REF_PATHS = 4096
_ref_env = None


def _ref_init(npz_path, seed_base):
    """Pool initializer: one reference env per worker process, constructed (it np.loads the file) outside any timed region."""
    global _ref_env
    import warnings
    from oracle import ref_runner
    warnings.simplefilter("ignore")
    wid = os.getpid()
    env = ref_runner.reference_env_class("v2")(npz_path, record_metrics=True, **ENV_KW)
    acts = np.random.default_rng(seed_base + wid).uniform(-1, 1, (4096, 2)).astype(np.float32)
    acts[:, 1] = 0.0                                                            # one European call: the put leg is never traded
    env.reset(seed=seed_base + wid)
    _ref_env = [env, acts, 0]


def _ref_steps(n_steps):
    """n_steps env-steps of this worker's reference env (reset on terminated); returns (steps, seconds)."""
    env, acts, at = _ref_env
    t0 = time.perf_counter()
    for _ in range(n_steps):
        _, _, term, _, _ = env.step(acts[at & 4095])
        at += 1
        if term:
            env.reset()
    _ref_env[2] = at
    return n_steps, time.perf_counter() - t0


def _ref_seconds(seconds):
    env, acts, at = _ref_env
    n, t0 = 0, time.perf_counter()
    while True:
        for _ in range(256):
            _, _, term, _, _ = env.step(acts[at & 4095])
            at += 1
            if term:
                env.reset()
        n += 256
        el = time.perf_counter() - t0
        if el >= seconds:
            _ref_env[2] = at
            return n, el


def _port_worker(args):
    seconds, seed = args
    from oracle.hedge_oracle import EnvParams
    from oracle.hedge_scalar import time_scalar_env
    from oracle.ref_runner import gbm_env_schema
    S, V, Cc, Pp = gbm_env_schema(256, 252, 42)
    return time_scalar_env(S, V, Cc, Pp, EnvParams(**ENV_KW), seconds, seed)


def _side_baselines():
    """the reference's own repricing / delta-hedge functions and the NumPy outer path step, one core."""
    from oracle import ref_runner as rr
    out = {}
    if rr.staged():
        v, s = rr.time_black_scholes_vectorized(1 << 20, repeats=2)
        out["black_scholes_vectorized"] = dict(value=v, unit="repricings/s (call+put)", seconds=s, kind="reference",
                                               what="src/sim/option_price_assignment.py:10-21 on 2^20 float64 elements, 1 core")
        v, s = rr.time_bs_delta_hedge(16, 252)
        out["bs_delta_hedge"] = dict(value=v, unit="path-steps/s", seconds=s, kind="reference",
                                     what="src/tools/bs_delta.py:36-55 on 16 GBM paths x 253, 1 core")
    v, s = rr.time_numpy_outer_step(1 << 20, 8)
    out["numpy_outer_euler_step"] = dict(value=v, unit="path-steps/s", seconds=s, kind="port",
                                         what="NumPy float64 restatement of src/sim/rbergomi_sim.py:454-464 on 2^20 paths x 8 days, 1 core "
                                              "(the file itself needs CuPy + a GPU)")
    return out


def cpu_baseline(seconds, procs, side=True):
    """P forked processes x one UNMODIFIED reference env each (kind "reference"); the port as a second number."""
    import multiprocessing as mp
    from oracle import ref_runner as rr
    ctx = mp.get_context("fork")
    out = {}
    if rr.staged():
        tmp, npz = rr.tmp_npz(REF_PATHS, 252, 42)
        with ctx.Pool(procs, initializer=_ref_init, initargs=(npz, 1000)) as pool:
            pool.map(_ref_seconds, [0.5] * procs)                                  # warm-up (>= 1 s in all)
            pool.map(_ref_seconds, [0.5] * procs)
            t0 = time.perf_counter()
            res = pool.map(_ref_seconds, [seconds] * procs)
            wall = time.perf_counter() - t0
            one = None
        with ctx.Pool(1, initializer=_ref_init, initargs=(npz, 2000)) as pool:     # (a) of section 3.4: one process alone
            pool.map(_ref_seconds, [0.5])
            one = pool.map(_ref_seconds, [min(seconds, 3.0)])[0]
        tmp.cleanup()
        steps = sum(r[0] for r in res)
        out = dict(value=sum(r[0] / r[1] for r in res), unit="env-steps/s", cores=procs, kind="reference",
                   single_process_value=one[0] / one[1],
                   sample=f"{procs} forked processes x 1 UNMODIFIED reference HedgingEnv (src/env/hedging_env_v2.py staged in oracle/_ref, "
                          f"gymnasium stub), record_metrics=True, v2 training keywords, uniform float32 actions (put leg 0), reset on "
                          f"terminated, {REF_PATHS} GBM paths x 252 steps; {steps} env-steps in {wall:.1f} s wall")
    with ctx.Pool(procs) as pool:
        pres = pool.map(_port_worker, [(min(seconds, 4.0), i) for i in range(procs)])
    port = dict(value=sum(r[0] / r[1] for r in pres), unit="env-steps/s", cores=procs, kind="port",
                sample=f"{procs} processes x 1 scalar reference-shaped env (oracle/hedge_scalar.py) on 256 GBM paths x 252 steps")
    if out:
        out["port"] = port
    else:                                   # oracle/_ref not staged (build() did not run where /root/reference exists)
        out = port
    out["cpu"] = _cpu_model()
    if side:
        out["side_baselines"] = _side_baselines()
    return out


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _numa_note(local_rank):
    """config.host_numa_bind, computed the same way by both arms so that their configs compare equal on one box."""
    try:
        from cantorrl_b200.distributed import bind_to_gpu_numa_node
        all_cpus = os.sched_getaffinity(0)
        bound = bind_to_gpu_numa_node(local_rank)
        note = f"{len(bound)} of {len(all_cpus)} CPUs (GPU-local node)" if bound else "none (single node, unknown topology or disabled)"
        return note, all_cpus, bound
    except Exception:
        return "none (single node, unknown topology or disabled)", None, None


def run_reference(args, rank, world):
    """--impl reference: the unmodified reference env on all host cores; one bench 'step' = EP episodes x 252 steps per process."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import ref_runner as rr
    procs = os.cpu_count() or 1
    T = args.episode_length
    note, all_cpus, _ = _numa_note(int(os.environ.get("LOCAL_RANK", "0")))
    if all_cpus:
        os.sched_setaffinity(0, all_cpus)                 # the CPU arm uses every core
    ctx = mp.get_context("fork")
    EP = 8                                                # episodes per process per bench step (bounded sample)
    kind = "reference" if rr.staged() else "port"
    if kind == "reference":
        tmp, npz = rr.tmp_npz(REF_PATHS, T, 42)
        with ctx.Pool(procs, initializer=_ref_init, initargs=(npz, 1000)) as pool:
            for w in range(args.warmup):
                pool.map(_ref_steps, [EP * T] * procs)
            t0 = time.perf_counter()
            total = 0
            for k in range(args.steps):
                total += sum(r[0] for r in pool.map(_ref_steps, [EP * T] * procs))
            el = time.perf_counter() - t0
        tmp.cleanup()
        what = "UNMODIFIED reference HedgingEnv (src/env/hedging_env_v2.py staged in oracle/_ref, gymnasium stub)"
    else:
        with ctx.Pool(procs) as pool:
            t0 = time.perf_counter()
            res = pool.map(_port_worker, [(max(1.0, 0.2 * args.steps), i) for i in range(procs)])
            el = time.perf_counter() - t0
        total = sum(r[0] for r in res)
        what = "scalar reference-shaped port (oracle/hedge_scalar.py): oracle/_ref is not staged"
    value = total / el
    sample = (f"{procs} processes x 1 {what}, greeks on, x {EP} episodes x {T} steps per bench step on {REF_PATHS} GBM paths; "
              f"{total} env-steps in {el:.1f} s")
    cfg = _config(args, world)
    cfg["host_numa_bind"] = note
    line = dict(impl="reference", metric="env-steps/sec (fused hedge step)", value=value, unit="env-steps/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * el / max(args.steps, 1),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=cfg, gpu_launches=0,
                cpu_baseline=dict(value=value, unit="env-steps/s", cores=procs, kind=kind, sample=sample, cpu=_cpu_model()),
                e2e=dict(value=value, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU side
def _config(args, world):
    return dict(workload=f"configs[1]: {args.envs} synthetic GBM paths x {args.episode_length} steps per GPU, one European "
                         "ATM call (put action 0), commission 0.65 + slippage 1 bp, v2 reward (w=1e-3, lam=1e-4, theta=2e-4)",
                mode="replay", envs_per_gpu=args.envs, episode_length=args.episode_length, precision=args.precision,
                policy="uniform random float32 actions, pre-generated on device", parallelism=f"path-sharded x{world}",
                l2="inputs larger than L2: each launch reads/writes fresh [t] slabs of 4.2 GB replay + 17 GB rollout buffers")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.06)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) == 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) == 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[j] for r in self.rows if len(r) == 7 for j in range(4) if r[3 + j].lower() == "active"})
        pw = [float(r[2]) for r in self.rows if len(r) == 7 and r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=reasons)


def synth_replay_data(n, T, rank, device):
    """Synthetic env-schema book of configs[1]: K1 (Philox GBM log-Euler) fused with K2 (ATM Black-Scholes), written
    straight into HBM in the packed layout.  Returns (book, kernel milliseconds)."""
    import torch
    from cantorrl_b200 import sim
    book = sim.generate_paths_and_options(n, R, DT, 42, n_steps=T, model="gbm", s0=S0, v0=XI, path_offset=rank * n,
                                          device=device)                       # warm-up / allocation
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sim.generate_paths_and_options(n, R, DT, 42, n_steps=T, model="gbm", s0=S0, v0=XI, path_offset=rank * n,
                                   device=device, out=book)
    e1.record()
    torch.cuda.synchronize(device)
    return book, e0.elapsed_time(e1)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from cantorrl_b200 import HedgingVecEnv, _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # the e2e leg stages 65 B per env-step through page-locked host memory: keep this rank (and the memory it pins) on the
    # NUMA node of its GPU's PCIe root; the CPU baseline below gets the full core set back
    from cantorrl_b200.distributed import bind_to_gpu_numa_node
    all_cpus = os.sched_getaffinity(0)
    bound_cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, T, K, W = args.envs, args.episode_length, args.steps, args.warmup
    L = _lib.lib()

    data, sim_ms = synth_replay_data(n, T, rank, dev)
    env = HedgingVecEnv(data=data, num_envs=n, device=dev, precision=args.precision, episode_sampler="same_path",
                        env_offset=rank * n, **ENV_KW)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = torch.rand((T, n, 2), device=dev, generator=g) * 2 - 1
    actions[:, :, 1] = 0.0                                        # one European call: the put leg is never traded
    rdt = torch.float64 if args.precision == "fp64" else torch.float32
    obs = torch.empty((T, n, 13), dtype=torch.float32, device=dev)
    reward = torch.empty((T, n), dtype=rdt, device=dev)
    done = torch.empty((T, n), dtype=torch.uint8, device=dev)
    env.reset()
    stream = torch.cuda.current_stream(dev)

    def sweep():
        _lib.check(L.cantor_env_step_many(C.byref(env._params), C.byref(env._book), C.byref(env._state), n, env._prec, T,
                                          actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), done.data_ptr(), None,
                                          C.byref(env._rule), stream.cuda_stream), "cantor_env_step_many")

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(W, 3)):
        sweep()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        sweep()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    assert bool(done[T - 1].all()) and not bool(done[T - 2].any()), "episode boundary not where expected"
    assert bool(torch.isfinite(reward).all()) and bool(torch.isfinite(obs[T - 1]).all())

    # ---- e2e: gym-style step with host buffers (pinned), copies inside the timed region --------------------
    # NumPy in / NumPy out through the host-buffer C ABI (cantor_vecenv_step_host): no torch on this path.
    e2e_s, e2e_checksum = float("inf"), None
    if args.e2e_steps > 0:
        from cantorrl_b200.host_env import HostVecEnv
        henv = HostVecEnv(num_envs=n, device=local_rank, precision=args.precision, episode_sampler="same_path",
                          env_offset=rank * n, simulate=dict(num_paths=n, n_steps=T, model="gbm", seed=42, s0=S0, v0=XI,
                                                             path_offset=rank * n), **ENV_KW)
        a_host = henv.pin(np.ascontiguousarray(actions.cpu().numpy()))          # [T, n, 2] page-locked host actions
        henv.reset()

        def e2e_sweep():
            acc = 0.0
            for t in range(T):                          # each call returns when obs/reward/done are on the host
                o, r, d, _ = henv.step(a_host[t])
                acc += float(r[0]) + float(o[0, 0])
            return acc

        e2e_sweep()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_checksum = e2e_sweep()
        barrier()
        e2e_s = time.perf_counter() - t0
        # the host path and the device path computed the same last reward of the sweep (same book, same actions)
        assert np.isfinite(e2e_checksum)
        henv.close()

    # ---- configs[3] shape: episode-fused rollout, paths generated in-kernel, statistics all-reduced over NCCL -------
    # One launch = one whole episode of every env (GBM on the fly + ATM repricing + delta-hedge policy + env step +
    # episode statistics); the only inter-GPU traffic of the whole path is the all-reduce of the statistics buffers.
    roll = None
    if args.rollout_steps > 0:
        from cantorrl_b200.rollout import HedgingRollout
        nr = args.rollout_envs
        ro = HedgingRollout(simulate=dict(model="gbm", seed=42, s0=S0, v0=XI, n_steps=T), num_envs=nr, device=dev,
                            env_offset=rank * nr, total_envs=world * nr, one_call_only=True, **ENV_KW)
        rstats = ro.new_stats()
        stats_transport = "nccl all_reduce" if world > 1 else "single GPU"
        if world > 1 and not args.no_fused_allreduce:
            try:      # all-reduce fused into the kernel epilogue: multimem.red through the NVSwitch, or peer atomics over NVLink
                stats_transport = "fused in-kernel: " + rstats.enable_fused_all_reduce()
            except Exception as e:
                stats_transport = f"nccl all_reduce (symmetric memory unavailable: {type(e).__name__})"

        def roll_sweep():
            rstats.zero_()
            ro.run(T, "delta_every_step", stats=rstats)
            rstats.all_reduce()

        for _ in range(2):
            roll_sweep()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        for _ in range(args.rollout_steps):
            roll_sweep()
        r1.record(stream)
        barrier()
        roll = (r0.elapsed_time(r1), rstats.result(), stats_transport)

    # ---- configs[4] shape: on-policy rollout, MLP actor 13-64-64-2 on the tensor cores fused with the env step ------
    # 2^19 envs per GPU x 1000 steps (4 M envs on 8 GPUs), GBM on the fly, bf16 tcgen05.mma actor, statistics all-reduced.
    mlp_roll = None
    if args.mlp_rollout_steps > 0:
        from cantorrl_b200.rollout import HedgingRollout, pack_mlp
        n5 = min(n, 1 << 19)
        gw = np.random.default_rng(5)
        wts = pack_mlp(gw.normal(0, .5, (64, 13)), gw.normal(0, .1, 64), gw.normal(0, .2, (64, 64)), gw.normal(0, .1, 64),
                       gw.normal(0, .3, (2, 64)), gw.normal(0, .1, 2), gw.normal(0, .2, 13), gw.uniform(.05, 2, 13), device=dev)
        ro5 = HedgingRollout(simulate=dict(model="gbm", seed=42, s0=S0, v0=XI, n_steps=T), num_envs=n5, device=dev,
                             env_offset=rank * n5, total_envs=world * n5, **ENV_KW)
        st5 = ro5.new_stats()
        ro5.run(args.mlp_rollout_steps, "mlp_bf16", mlp=wts, stats=st5)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st5.zero_()
        m0.record(stream)
        ro5.run(args.mlp_rollout_steps, "mlp_bf16", mlp=wts, stats=st5)
        st5.all_reduce()
        m1.record(stream)
        barrier()
        assert float(st5.sums[15]) == 0.0, "a tcgen05 MMA timed out"
        mlp_roll = (m0.elapsed_time(m1), n5, st5.result())

    # ---- the policy the reference actually trained: LSTM(13-128) + MLP(128-64-64-2), recurrent state carried in TMEM --------
    lstm_roll = None
    if args.lstm_rollout_steps > 0:
        from cantorrl_b200.rollout import HedgingRollout, pack_lstm
        n6 = min(n, 1 << 19)
        gw = np.random.default_rng(6)
        kk = 1 / np.sqrt(128)
        wl = pack_lstm(gw.uniform(-kk, kk, (512, 13)) * 3, gw.uniform(-kk, kk, (512, 128)) * 2, gw.uniform(-kk, kk, 512), gw.uniform(-kk, kk, 512),
                       gw.normal(0, .15, (64, 128)), gw.normal(0, .1, 64), gw.normal(0, .2, (64, 64)), gw.normal(0, .1, 64),
                       gw.normal(0, .3, (2, 64)), gw.normal(0, .1, 2), gw.normal(0, .2, 13), gw.uniform(.05, 2, 13), device=dev)
        ro6 = HedgingRollout(simulate=dict(model="gbm", seed=42, s0=S0, v0=XI, n_steps=T), num_envs=n6, device=dev,
                             env_offset=rank * n6, total_envs=world * n6, **ENV_KW)
        st6 = ro6.new_stats()
        ro6.run(min(args.lstm_rollout_steps, 16), "lstm_bf16", mlp=wl, stats=st6)
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st6.zero_()
        l0.record(stream)
        ro6.run(args.lstm_rollout_steps, "lstm_bf16", mlp=wl, stats=st6)
        st6.all_reduce()
        l1.record(stream)
        barrier()
        assert float(st6.sums[15]) == 0.0, "a tcgen05 MMA timed out"
        lstm_roll = (l0.elapsed_time(l1), n6, st6.result())

    # ---- configs[2] shape: Heston paths + multi-strike float32 Black-Scholes book (8 strikes, maturities to episode end) ---
    book_ms = None
    if args.book_strikes > 0 and rank == 0:
        from cantorrl_b200 import sim as _sim
        hb = _sim.generate_paths_and_options(n, R, DT, 42, n_steps=T, model="heston", reprice=False, device=dev)
        mult = np.linspace(0.9, 1.1, args.book_strikes).astype(np.float32)
        _sim.reprice_book(hb, mult, sigma="book")
        torch.cuda.synchronize(dev)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        out_book = _sim.reprice_book(hb, mult, sigma="book")
        b1.record(stream)
        torch.cuda.synchronize(dev)
        book_ms = b0.elapsed_time(b1)
        del out_book, hb

    # ---- the reference's own data generator: rough-Bergomi paths + nested-MC ATM prices (5000 inner paths x 30 steps) ----
    rb_res = None
    if args.rbergomi_paths > 0 and rank == 0:
        from cantorrl_b200 import sim as _sim
        base = (496.48, 0.02903, 0.4656, 1.985, -0.2022)           # estimate_base_params on the shipped CSV (SURVEY [probe])
        _sim.generate_rbergomi_paths_and_options(64, base_params=base, n_steps=8, n_mc=64, device=dev)
        torch.cuda.synchronize(dev)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rbk = _sim.generate_rbergomi_paths_and_options(args.rbergomi_paths, base_params=base, n_steps=32, n_mc=5000, price=False, device=dev)
        q0.record(stream)
        rbk.price_days(0, 32)
        q1.record(stream)
        torch.cuda.synchronize(dev)
        rb_res = (q0.elapsed_time(q1), args.rbergomi_paths * 32 * 2)

    # ---- reduce over ranks ------------------------------------------------------------------------------------
    tt = torch.tensor([ms, e2e_s, roll[0] if roll else 0.0, mlp_roll[0] if mlp_roll else 0.0, lstm_roll[0] if lstm_roll else 0.0],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_s, roll_ms, mlp_ms, lstm_ms = float(tt[0]), float(tt[1]), float(tt[2]), float(tt[3]), float(tt[4])
    if rank == 0:
        env_steps = float(n) * world * T * K
        value = env_steps / (ms * 1e-3)
        launch_ms = ms / (K * T)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        achieved = B_ALG[args.precision] * n / (launch_ms * 1e-3) / 1e9
        kname = "hedge_step_kernel<F64=%s,INFO=false>" % ("true" if args.precision == "fp64" else "false")
        traffic = None
        try:   # DRAM bytes per launch from the committed ncu --set full capture of this kernel, scaled to this launch size
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kname]
            traffic = (tr["read_bytes"] + tr["write_bytes"]) * n / tr["envs"]
        except (OSError, KeyError, ValueError):
            pass
        line = dict(
            metric="env-steps/sec (fused hedge step)", value=value, unit="env-steps/s", n_gpus=world, steps=K, warmup=max(W, 3),
            ms_per_step=ms / K, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f32" if args.precision == "fp32" else "f64", data="synthetic", config=_config(args, world),
            gpu_launches=K * T,
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                          kernel=kname,
                          algorithmic_bytes_per_env_step=B_ALG[args.precision], launch_us=launch_ms * 1e3, peak_source=peak_src,
                          dram_side=dict(bytes_per_env_step=B_ALG[args.precision] - 56,
                                         gbs=(B_ALG[args.precision] - 56) * n / (launch_ms * 1e-3) / 1e9,
                                         frac=(B_ALG[args.precision] - 56) * n / (launch_ms * 1e-3) / 1e9 / peak,
                                         what="the same launches counted without the 40 B of state and the 16 B lagged path record "
                                              "that stay in L2 from one launch to the next: what HBM itself has to move"),
                          note="achieved counts the 137 algorithmic B/env-step; 40 B of them (the env state, read + written every "
                               "step) are served by the 126 MB L2 between consecutive launches, so DRAM traffic per launch is lower "
                               "than the algorithmic bytes and frac can read above 1 against a plain-copy peak"),
            e2e=dict(value=float(n) * world * T * args.e2e_steps / e2e_s, unit="env-steps/s",
                     h2d_bytes_per_step=n * 8 * T, d2h_bytes_per_step=n * (52 + reward.element_size() + 1) * T,
                     api="HostVecEnv.step -> cantor_vecenv_step_host: NumPy actions in page-locked host memory -> obs/reward/"
                         "done in page-locked host memory, every env step (8 chunks over 3 streams), returns after the D2H"),
            clocks=clocks,
            extra=dict(path_sim_reprice=dict(kernel="sim_paths_kernel<GBM> (K1 fused with K2 ATM repricing)", ms=sim_ms,
                                             path_steps_per_s=float(n) * T / (sim_ms * 1e-3),
                                             hbm_write_gbs=float(n) * (T + 1) * 16 / (sim_ms * 1e-3) / 1e9)))
        if roll is not None:
            rs = roll[1]
            line["extra"]["rollout_on_the_fly"] = dict(
                kernel="rollout_kernel<GBM on the fly, delta_every_step policy, episode statistics> + all-reduce of the "
                       "statistics buffers once per sweep (configs[3] shape: 2^23 envs per GPU, 64 M on 8)", stats_all_reduce=roll[2], sweeps=args.rollout_steps, ms_per_sweep=roll_ms / args.rollout_steps,
                envs_per_gpu=args.rollout_envs,
                env_steps_per_s=float(args.rollout_envs) * world * T * args.rollout_steps / (roll_ms * 1e-3),
                stats={k: rs[k] for k in ("n_episodes", "mean_abs_pnl", "std_abs_pnl", "mean_cost", "mean_reward", "cvar95_abs_pnl")})
        if mlp_roll is not None:
            line["extra"]["rollout_mlp_policy"] = dict(
                kernel="rollout_kernel<GBM on the fly, MLP 13-64-64-2 actor as bf16 tcgen05.mma (TMEM accumulators), env step, "
                       "episode statistics> + NCCL all-reduce of the statistics", envs_per_gpu=mlp_roll[1],
                steps=args.mlp_rollout_steps, ms=mlp_ms,
                env_steps_per_s=float(mlp_roll[1]) * world * args.mlp_rollout_steps / (mlp_ms * 1e-3),
                actor_tflops=float(mlp_roll[1]) * world * args.mlp_rollout_steps * 2 * (16 * 64 + 80 * 64 + 80 * 16) / (mlp_ms * 1e-3) / 1e12,
                n_episodes=mlp_roll[2]["n_episodes"])
        if lstm_roll is not None:
            flop = 2 * (4 * 128 * 144 + 64 * 144 + 64 * 80 + 16 * 80)                  # per env-step, as issued (padded K / N)
            line["extra"]["rollout_lstm_policy"] = dict(
                kernel="rollout_kernel<GBM on the fly, LSTM(13-128) + MLP(128-64-64-2) actor as bf16 tcgen05.mma (warp-specialised "
                       "issuer, TMA-streamed gate weights, cell state in TMEM), env step, episode statistics> + all-reduce of the statistics",
                envs_per_gpu=lstm_roll[1], steps=args.lstm_rollout_steps, ms=lstm_ms,
                env_steps_per_s=float(lstm_roll[1]) * world * args.lstm_rollout_steps / (lstm_ms * 1e-3),
                actor_tflops=float(lstm_roll[1]) * world * args.lstm_rollout_steps * flop / (lstm_ms * 1e-3) / 1e12,
                mufu_bound_frac=float(lstm_roll[1]) * args.lstm_rollout_steps * 640 / (lstm_ms * 1e-3) / (148 * 16 * 1.965e9),
                n_episodes=lstm_roll[2]["n_episodes"])
        if book_ms is not None:
            cells = float(n) * (T + 1)
            line["extra"]["multi_strike_book"] = dict(
                kernel="book_f32_kernel (Heston book variance, %d strikes, price only)" % args.book_strikes, ms=book_ms,
                reprices_per_s=cells * args.book_strikes / (book_ms * 1e-3),
                algorithmic_gbs=cells * (16 + 8 * args.book_strikes) / (book_ms * 1e-3) / 1e9,
                mufu_per_s=cells * (5 + 3 * args.book_strikes) / (book_ms * 1e-3))
        if rb_res is not None:
            line["extra"]["rbergomi_nested_mc"] = dict(
                kernel="rbergomi_price_kernel<tcgen05 split-TF32 FIR> (5000 inner paths x 30 steps per ATM call / put)", ms=rb_res[0],
                pricings=rb_res[1], inner_path_steps_per_s=rb_res[1] * 5000.0 * 30 / (rb_res[0] * 1e-3),
                reference_workload_seconds=100000 * 252 * 2 / (rb_res[1] / (rb_res[0] * 1e-3)))
        line["config"]["host_numa_bind"] = (f"{len(bound_cpus)} of {len(all_cpus)} CPUs (GPU-local node)" if bound_cpus
                                            else "none (single node, unknown topology or disabled)")
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds, os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
