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
    ap.add_argument("--l2free-envs", type=int, default=1 << 23, help="envs of the per-step kernels timed beyond L2 (N = 1 only; 0 = skip)")
    ap.add_argument("--no-forms", dest="forms", action="store_false", help="skip the float64 parity forms / format kernels extra")
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
    """SM clock / throttle reasons sampled WHILE the timed region runs (B200_PROFILING.md's clocks line).  Polls NVML in-process
    every few milliseconds (the timed region of the default run is ~75 ms: a freshly spawned `nvidia-smi -lms` often delivers its
    first row after that); falls back to the nvidia-smi subprocess when pynvml is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx, self.nvml, self.thread, self.stop_flag = [], None, gpu_index, None, None, False
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip() != ""]
        if ids and all(x.strip().isdigit() for x in ids) and self.idx < len(ids):
            return int(ids[self.idx])
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                    pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]

            def poll():
                while not self.stop_flag:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        self.rows.append([sm, mx, pw] + [bool(r & b) for b in bits])
                    except pynvml.NVMLError:
                        pass
                    time.sleep(0.004)
            self.nvml, self.source = pynvml, "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:                       # noqa: BLE001 -- any NVML problem: use the command-line tool instead
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) == 7 and r[0].replace(".", "").isdigit() and r[1].replace(".", "").isdigit():
                pw = float(r[2]) if r[2].replace(".", "").isdigit() else float("nan")
                self.rows.append([float(r[0]), float(r[1]), pw] + [r[3 + j].lower() == "active" for j in range(4)])

    def stop(self):
        if self.nvml is None and self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.06)
            self.proc.terminate()
        rows = list(self.rows)
        sm, mx = [r[0] for r in rows], [r[1] for r in rows]
        pw = [r[2] for r in rows if r[2] == r[2]]
        reasons = sorted({self.NAMES[j] for r in rows for j in range(4) if r[3 + j]})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=reasons, source=self.source)


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


def gpu_ms(fn, iters, stream, dev, warm=1):
    """Mean milliseconds of fn() over `iters` calls, CUDA events on `stream`, after `warm` untimed calls."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters


def _traffic(kernel):
    """DRAM bytes (read + write) per env-step of `kernel` from the committed ncu launch list of the same command, or None."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return tr, (tr["read_bytes"] + tr["write_bytes"]) / tr["envs"] / tr.get("steps_per_launch", 1)
    except (OSError, KeyError, ValueError):
        return None, None


def ring_step_bench(L, dev, stream, n, T, precision, peak, mode, sweeps=2, ring=4, model="gbm"):
    """The gym-style step kernel (one launch = one env-step) where nothing survives in L2 between launches: n envs, the
    observation / reward / done / action slabs cycle through a ring.  mode: "replay" (book simulated in HBM), "sim" (on the fly:
    cantor_env_step_sim, no book) or "many" (replay through ONE persistent launch per `ring` steps... measured over T steps)."""
    import torch
    from cantorrl_b200 import HedgingVecEnv, _lib, sim
    rdt = torch.float64 if precision == "fp64" else torch.float32
    if mode == "sim":
        env = HedgingVecEnv(simulate=dict(model=model, seed=42, s0=S0, v0=XI, n_steps=T), num_envs=n, total_envs=n, device=dev,
                            precision=precision, **ENV_KW)
    else:
        book = sim.generate_paths_and_options(n, R, DT, 42, n_steps=T, model=model, s0=S0, v0=XI, device=dev)
        env = HedgingVecEnv(data=book, num_envs=n, device=dev, precision=precision, episode_sampler="same_path", **ENV_KW)
    g = torch.Generator(device=dev).manual_seed(7)
    R_ = T if mode == "many" else ring
    actions = torch.rand((R_, n, 2), device=dev, generator=g) * 2 - 1
    actions[:, :, 1] = 0.0
    obs = torch.empty((ring, n, 13), dtype=torch.float32, device=dev)
    reward = torch.empty((ring, n), dtype=rdt, device=dev)
    done = torch.empty((ring, n), dtype=torch.uint8, device=dev)
    env.reset()

    def sweep():
        if mode == "sim":
            for t in range(T):
                r = t % ring
                env.step(actions[r], obs_out=obs[r], reward_out=reward[r], done_out=done[r])
            return
        left = T
        while left > 0:
            k = min(ring, left)
            _lib.check(L.cantor_env_step_many(C.byref(env._params), C.byref(env._book), C.byref(env._state), n, env._prec, k,
                                              actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), done.data_ptr(), None,
                                              C.byref(env._rule), stream.cuda_stream), "cantor_env_step_many")
            left -= k

    if mode == "replay":
        os.environ["CANTOR_STEP_MANY_LAUNCHES"] = "1"
    try:
        ms = gpu_ms(sweep, sweeps, stream, dev)
    finally:
        os.environ.pop("CANTOR_STEP_MANY_LAUNCHES", None)
    assert bool(torch.isfinite(reward).all()) and bool(torch.isfinite(obs[0]).all())
    per_step_us = ms * 1e3 / T
    if mode == "sim":
        b_alg, kernel = 121 + (8 if precision == "fp64" else 0) + (12 if precision == "fp64" else 0), f"hedge_step_sim_kernel<{model}>"
    elif mode == "many":
        b_alg, kernel = 81 + (4 if precision == "fp64" else 0) + (40 + (24 if precision == "fp64" else 0)) / ring, "hedge_step_many_kernel"
    else:
        b_alg, kernel = B_ALG[precision], "hedge_step_kernel"
    gbs = b_alg * n / (per_step_us * 1e-6) / 1e9
    out = dict(kernel=kernel, mode=mode, envs=n, precision=precision, us_per_env_step_launch=per_step_us,
               env_steps_per_s=n / (per_step_us * 1e-6), algorithmic_bytes_per_env_step=b_alg, achieved_gbs=gbs, frac=gbs / peak)
    del env
    return out


def multi_gpu_stats_check(dev, rank, world, T, barrier):
    """N > 1 only: a small population (2^18 global envs, one episode) three ways -- statistics all-reduced INSIDE the kernels
    (multimem.red through the NVSwitch / peer atomics), all-reduced by NCCL, and computed by rank 0 alone over the same global
    env indices -- through the episode-fused rollout kernel and through the gym-style step kernel's fused Monitor.  The
    histogram counts must be equal, the float64 sums equal up to summation order (train_ppo_v2.py:520-530 statistics)."""
    import torch
    import torch.distributed as dist
    from cantorrl_b200 import HedgingVecEnv
    from cantorrl_b200.rollout import HedgingRollout
    from cantorrl_b200.stats import EpisodeStats
    total = 1 << 18
    per = total // world
    simkw = dict(model="gbm", seed=42, s0=S0, v0=XI, n_steps=T)
    out = {}

    def rel(a, b):
        a, b = a.cpu().double(), b.cpu().double()
        return float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())

    part = HedgingRollout(simulate=simkw, num_envs=per, device=dev, env_offset=rank * per, total_envs=total, **ENV_KW)
    nccl = part.new_stats()
    part.run(T, "delta_every_step", stats=nccl)
    nccl.all_reduce()
    fused = part.new_stats()
    out["transport"] = fused.enable_fused_all_reduce()
    fused.zero_()
    part.run(T, "delta_every_step", stats=fused)
    fused.all_reduce()
    barrier()
    out["rollout_fused_vs_nccl_hist_equal"] = bool(torch.equal(fused.hist, nccl.hist))
    out["rollout_fused_vs_nccl_max_rel_sums"] = rel(fused.sums[:12], nccl.sums[:12])
    # the gym-style step kernel with the fused Monitor, sharded, statistics all-reduced in its epilogue
    g = torch.Generator(device=dev).manual_seed(99)
    tape_all = torch.rand((T, total, 2), device=dev, generator=g) * 2 - 1            # same seed on every rank: same global tape
    envst = EpisodeStats(dev)
    envst.enable_fused_all_reduce()
    envst.zero_()
    env = HedgingVecEnv(simulate=simkw, num_envs=per, total_envs=total, env_offset=rank * per, device=dev, monitor=True, stats=envst, **ENV_KW)
    env.reset()
    for t in range(T):
        env.step(tape_all[t, rank * per:(rank + 1) * per].contiguous())
    envst.all_reduce()
    barrier()
    if rank == 0:
        whole = HedgingRollout(simulate=simkw, num_envs=total, device=dev, env_offset=0, total_envs=total, **ENV_KW)
        one = whole.new_stats()
        whole.run(T, "delta_every_step", stats=one)
        out["rollout_n_vs_1_hist_equal"] = bool(torch.equal(nccl.hist, one.hist))
        out["rollout_n_vs_1_max_rel_sums"] = rel(nccl.sums[:12], one.sums[:12])
        one_env_st = EpisodeStats(dev)
        env1 = HedgingVecEnv(simulate=simkw, num_envs=total, total_envs=total, device=dev, monitor=True, stats=one_env_st, **ENV_KW)
        env1.reset()
        for t in range(T):
            env1.step(tape_all[t])
        torch.cuda.synchronize(dev)
        out["step_kernel_monitor_fused_n_vs_1_hist_equal"] = bool(torch.equal(envst.hist, one_env_st.hist))
        out["step_kernel_monitor_fused_n_vs_1_max_rel_sums"] = rel(envst.sums[:12], one_env_st.sums[:12])
        out["episodes"] = int(one.sums[0])
        out["ok"] = bool(out["rollout_fused_vs_nccl_hist_equal"] and out["rollout_n_vs_1_hist_equal"]
                         and out["step_kernel_monitor_fused_n_vs_1_hist_equal"] and out["rollout_fused_vs_nccl_max_rel_sums"] < 1e-9
                         and out["rollout_n_vs_1_max_rel_sums"] < 1e-9 and out["step_kernel_monitor_fused_n_vs_1_max_rel_sums"] < 1e-9
                         and out["episodes"] == total)
    barrier()
    if dist.is_initialized():
        dist.barrier()
    return out


def graph_us(fn, dev, steps_per_graph=21, reps=8):
    """Microseconds per fn() with the calls captured in a CUDA graph, so that the host side of a call is not what is timed."""
    import torch
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for _ in range(4):
            fn()
        torch.cuda.synchronize(dev)
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=side):
            for _ in range(steps_per_graph):
                fn()
        gph.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gph.replay()
        e1.record()
        torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / (reps * steps_per_graph) * 1e3


def hbm_mix_context(dev, stream):
    """HBM bandwidth of this GPU for pure reads, a copy, `add` and pure writes (plain torch streaming kernels over 1 GiB float32 buffers,
    best of 5): context for fractions quoted against the 50 % / 50 % copy peak -- pure writes are NOT the slower stream on this part."""
    import torch
    n = (1 << 30) // 4
    a = torch.empty(n, dtype=torch.float32, device=dev).fill_(1.0)
    b = torch.empty_like(a).fill_(2.0)
    c = torch.empty_like(a)

    def best(fn, nbytes):
        t = min(gpu_ms(fn, 1, stream, dev, warm=1) for _ in range(5))
        return nbytes / t / 1e6

    out = dict(what="torch kernels over 1 GiB float32 buffers, best of 5, GB/s",
               read_only_sum=best(lambda: a.sum(), 4 * n), copy=best(lambda: c.copy_(a), 8 * n),
               add_two_reads_one_write=best(lambda: torch.add(a, b, out=c), 12 * n), write_only_fill=best(lambda: c.fill_(1.0), 4 * n))
    del a, b, c
    torch.cuda.empty_cache()
    return out


def side_kernel_bench(dev, stream, peak):
    """Time + algorithmic GB/s + fraction of the HBM peak of the float64 parity forms and the format / wrapper kernels, called
    through the C ABI on preallocated buffers: bs_price (A6), schema_b_book (A7 + A8), bs_delta_hedge (A12), pack / unpack
    book ((f)3), VecNormalize ((f)1)."""
    import torch
    from cantorrl_b200 import HedgingVecEnv, _lib, sim
    from cantorrl_b200.vecnorm import VecNormalize
    L = _lib.lib()
    sp = stream.cuda_stream
    out = {}
    g = torch.Generator(device=dev).manual_seed(3)
    f64 = dict(dtype=torch.float64, device=dev)

    def entry(kernel, ms, nbytes, units, unit_name, note=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        d = dict(kernel=kernel, ms=ms, algorithmic_gbs=gbs, frac=gbs / peak)
        d[unit_name] = units / (ms * 1e-3)
        if note:
            d["bound"] = note
        return d

    # A6 black_scholes_vectorized on 2^24 elements, float64: S, K, sigma read (24 B), call + put written (16 B)
    m = 1 << 24
    S = 100 * torch.exp(0.2 * torch.randn(m, generator=g, **f64))
    Kk = torch.round(S * torch.exp(0.05 * torch.randn(m, generator=g, **f64)))
    sg = (0.2 + 0.05 * torch.randn(m, generator=g, **f64)).abs()
    Tm = torch.full((1,), 0.5, **f64)
    call, put = torch.empty(m, **f64), torch.empty(m, **f64)
    ms = gpu_ms(lambda: _lib.check(L.cantor_bs_price(S.data_ptr(), Kk.data_ptr(), Tm.data_ptr(), sg.data_ptr(), m, 1, 1, 0, 1, R, 1e-8,
                                                     call.data_ptr(), put.data_ptr(), sp)), 5, stream, dev)
    out["bs_price"] = entry("bs_price_kernel (black_scholes_vectorized, option_price_assignment.py:10-21, float64)", ms, 40.0 * m, m,
                            "repricings_per_s", "FP64 pipe (2 erfc + exp + log + divisions in float64 per element)")
    del S, Kk, sg, call, put
    # A7 + A8 process_price_paths: 2^18 paths x 253, float64: 8 B read, 24 B written (vols, calls, puts) per cell
    n, T = 1 << 18, 252
    book = sim.generate_paths_and_options(n, R, DT, 42, n_steps=T, model="gbm", device=dev)
    tm = book.tensor[:, :n, 0].double().contiguous()                       # time-major float64 paths [T+1, n]
    mult = torch.ones(1, **f64)
    vols, calls, puts = (torch.empty((T + 1, n), **f64) for _ in range(3))
    ms = gpu_ms(lambda: _lib.check(L.cantor_schema_b_book(tm.data_ptr(), n, T, n, R, mult.data_ptr(), 1, vols.data_ptr(), calls.data_ptr(),
                                                          puts.data_ptr(), sp)), 3, stream, dev)
    cells = float(n) * (T + 1)
    out["schema_b_book"] = entry("schema_b_book_kernel (process_price_paths + running realised volatility, option_price_assignment.py:23-52, "
                                 "float64)", ms, 32.0 * cells, cells, "repricings_per_s", "FP64 pipe")
    # A12 bs_delta_hedge: 8 B read + 8 B written per cell
    pnl = torch.empty((T + 1, n), **f64)
    ms = gpu_ms(lambda: _lib.check(L.cantor_bs_delta_hedge(tm.data_ptr(), n, T, n, R, DT, pnl.data_ptr(), sp)), 3, stream, dev)
    out["bs_delta_hedge"] = entry("bs_delta_hedge_kernel (src/tools/bs_delta.py:36-55, float64)", ms, 16.0 * cells, cells, "path_steps_per_s",
                                  "FP64 pipe")
    del vols, calls, puts, pnl, tm
    # (f)3 pack / unpack: float64 npz-layout arrays <-> packed float32 book: 32 B on the float64 side + 16 B on the book side per cell
    pm = book.to_path_major(torch.float64)
    ptrs = [pm[k].data_ptr() for k in ("paths", "volatilities", "call_prices_atm", "put_prices_atm")]
    ms = gpu_ms(lambda: _lib.check(L.cantor_pack_book(*ptrs, _lib.F64, n, T, book.tensor.data_ptr(), book.ld, sp)), 3, stream, dev)
    out["pack_book"] = entry("pack_book_kernel (float64 path-major npz arrays -> packed float32 time-major book)", ms, 48.0 * cells, cells,
                             "cells_per_s", "HBM (transposing through 32 x 32 shared-memory tiles)")
    ms = gpu_ms(lambda: _lib.check(L.cantor_unpack_book(book.tensor.data_ptr(), book.ld, n, T, _lib.F64, *ptrs, sp)), 3, stream, dev)
    out["unpack_book"] = entry("unpack_book_kernel (packed book -> float64 path-major arrays)", ms, 48.0 * cells, cells, "cells_per_s", "HBM")
    del pm, book
    torch.cuda.empty_cache()
    # (f)1 VecNormalize around the step at 2^20 envs (moments + apply, 174 B of L2 traffic per env-step on top of the step)
    n = 1 << 20
    book = sim.generate_paths_and_options(n, R, DT, 42, n_steps=T, model="gbm", device=dev)
    actions = torch.rand((n, 2), device=dev, generator=g) * 2 - 1
    env = HedgingVecEnv(data=book, num_envs=n, device=dev, episode_sampler="same_path", **ENV_KW)
    env.reset()
    plain = graph_us(lambda: env.step(actions), dev)
    vn = VecNormalize(HedgingVecEnv(data=book, num_envs=n, device=dev, episode_sampler="same_path", **ENV_KW))
    vn.reset()
    keep = graph_us(lambda: vn.venv.step(actions), dev)
    both = graph_us(lambda: vn.step(actions), dev)
    out["vecnormalize"] = dict(kernel="vecnorm moments + apply kernels behind hedge_step_kernel (2^20 envs, CUDA graph)", step_us=plain,
                               step_keeping_obs_in_l2_us=keep, step_plus_vecnormalize_us=both, vecnormalize_only_us=both - keep,
                               env_steps_per_s=n / both * 1e6)
    # record_info (fp32: float32 info arrays) and the Monitor at 2^20 envs
    for name, kw in (("step_record_info_fp32", dict(record_info=True)), ("step_monitor_fp32", dict(monitor=True)),
                     ("step_fp64", dict(precision="fp64"))):
        e2 = HedgingVecEnv(data=book, num_envs=n, device=dev, episode_sampler="same_path", **ENV_KW, **kw)
        e2.reset()
        out[name] = dict(us_per_step=graph_us(lambda: e2.step(actions), dev), envs=n)
        del e2
    return out


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
    numa_note, all_cpus, bound_cpus = _numa_note(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, T, K, W = args.envs, args.episode_length, args.steps, args.warmup
    L = _lib.lib()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

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

    def sweep():            # one 252-step episode of every env: ONE persistent launch (action tape known; state in registers)
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
    checksum_many = float(reward[T - 1].double().sum())

    # ---- the same sweep as 252 chained launches of the gym-style per-step kernel (what a closed-loop caller gets) -----------
    os.environ["CANTOR_STEP_MANY_LAUNCHES"] = "1"
    try:
        for _ in range(2):
            sweep()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(K):
            sweep()
        p1.record(stream)
        barrier()
        per_launch_ms = p0.elapsed_time(p1)
    finally:
        os.environ.pop("CANTOR_STEP_MANY_LAUNCHES", None)
    # same book, same actions, same starting state (every sweep ends with the auto-reset of every env): identical last rewards
    assert float(reward[T - 1].double().sum()) == checksum_many, "persistent and per-launch sweeps disagree"

    # ---- e2e: gym-style step with host buffers (pinned), copies inside the timed region --------------------
    # NumPy in / NumPy out through the host-buffer C ABI (cantor_vecenv_step_host): no torch on this path.
    e2e_s, e2e_checksum, probe = float("inf"), None, None
    if args.e2e_steps > 0:
        from cantorrl_b200.host_env import HostVecEnv, host_copy_probe
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
        assert np.isfinite(e2e_checksum)
        henv.close()
        # the raw ceiling of that path on this box, all ranks at once: the same bytes per step (57 out, 8 in per env) as plain
        # cudaMemcpyAsync between page-locked host memory and HBM, no kernel, with the per-step stream synchronisation
        rb = 8 if args.precision == "fp64" else 4
        barrier()
        probe = host_copy_probe(device=local_rank, d2h_bytes=(52 + rb + 1) * n, h2d_bytes=8 * n, n_chunks=8, sync_each_round=True, seconds=0.6)
        barrier()
        probe_free = host_copy_probe(device=local_rank, d2h_bytes=(52 + rb + 1) * n, h2d_bytes=0, n_chunks=1, sync_each_round=False, seconds=0.4)
        barrier()
        probe["d2h_only_gbs"] = probe_free["d2h_gbs"]

    del obs, reward, done, actions, env, data
    torch.cuda.empty_cache()

    # ---- the per-step kernels where nothing survives in L2 (configs[3] shard size), and the on-the-fly mode ----------------------
    l2free = {}
    if args.l2free_envs > 0 and world == 1:
        big = args.l2free_envs
        for name, kw in (("per_step_replay_fp32", dict(precision="fp32", mode="replay")),
                         ("per_step_replay_fp64", dict(precision="fp64", mode="replay")),
                         ("persistent_step_many_fp32", dict(precision="fp32", mode="many")),
                         ("per_step_on_the_fly_gbm_fp32", dict(precision="fp32", mode="sim", model="gbm")),
                         ("per_step_on_the_fly_heston_fp32_2x", dict(precision="fp32", mode="sim", model="heston", n_mult=2))):
            kw = dict(kw)
            mult = kw.pop("n_mult", 1)
            l2free[name] = ring_step_bench(L, dev, stream, big * mult, T, peak=peak, **kw)
            torch.cuda.empty_cache()
        for name, kernel in (("per_step_replay_fp32", "hedge_step_kernel<F64=false,INFO=false>@2^23"),
                             ("per_step_on_the_fly_gbm_fp32", "hedge_step_sim_kernel<gbm,F64=false>@2^23")):
            tr, per_env = _traffic(kernel)
            if per_env is not None and name in l2free:
                l2free[name]["dram_traffic_bytes_per_env_step"] = per_env
                l2free[name]["traffic_source"] = tr["source"]

    # ---- configs[3] shape: episode-fused rollout, paths generated in-kernel, statistics all-reduced over NCCL -------
    # One launch = one whole episode of every env (GBM on the fly + ATM repricing + delta-hedge policy + env step +
    # episode statistics); the only inter-GPU traffic of the whole path is the all-reduce of the statistics buffers.
    roll = None
    stats_check = None
    if args.rollout_steps > 0:
        from cantorrl_b200.rollout import HedgingRollout
        nr = args.rollout_envs
        ro = HedgingRollout(simulate=dict(model="gbm", seed=42, s0=S0, v0=XI, n_steps=T), num_envs=nr, device=dev,
                            env_offset=rank * nr, total_envs=world * nr, one_call_only=True, **ENV_KW)
        rstats = ro.new_stats()
        stats_transport = "nccl all_reduce" if world > 1 else "single GPU"
        if world > 1 and not args.no_fused_allreduce:
            # all-reduce fused into the kernel epilogue: multimem.red through the NVSwitch, or peer atomics over NVLink.  A box whose
            # driver / torch build cannot set up symmetric memory keeps NCCL -- said in the line, never silently.
            try:
                stats_transport = "fused in-kernel: " + rstats.enable_fused_all_reduce()
            except (RuntimeError, ImportError, AttributeError, NotImplementedError) as exc:
                stats_transport = f"nccl all_reduce (symmetric memory unavailable: {type(exc).__name__}: {str(exc)[:120]})"
                rstats = ro.new_stats()

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
        del ro, rstats
        if world > 1:
            if stats_transport.startswith("fused"):
                stats_check = multi_gpu_stats_check(dev, rank, world, T, barrier)
            else:
                stats_check = dict(ok=None, skipped="the fused all-reduce could not be enabled on this box: " + stats_transport)

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

    # ---- the float64 parity forms and the format / wrapper kernels (rows A6-A8, A12, (f)1, (f)3) ----------------------------------
    forms = None
    if args.forms and rank == 0:
        torch.cuda.empty_cache()
        forms = side_kernel_bench(dev, stream, peak)
        forms["hbm_bandwidth_vs_mix"] = hbm_mix_context(dev, stream)

    # ---- reduce over ranks ------------------------------------------------------------------------------------
    pr = probe or dict(d2h_gbs=0.0, h2d_gbs=0.0, rounds_per_s=0.0, d2h_only_gbs=0.0)
    tt = torch.tensor([ms, e2e_s, roll[0] if roll else 0.0, mlp_roll[0] if mlp_roll else 0.0, lstm_roll[0] if lstm_roll else 0.0,
                       per_launch_ms], dtype=torch.float64, device=dev)
    ps = torch.tensor([pr["d2h_gbs"], pr["h2d_gbs"], pr["rounds_per_s"], pr["d2h_only_gbs"]], dtype=torch.float64, device=dev)
    pmin = torch.tensor([pr["rounds_per_s"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ps, op=dist.ReduceOp.SUM)
        dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    ms, e2e_s, roll_ms, mlp_ms, lstm_ms, per_launch_ms = (float(x) for x in tt)
    if rank == 0:
        env_steps = float(n) * world * T * K
        value = env_steps / (ms * 1e-3)
        sweep_ms = ms / K
        fp64 = args.precision == "fp64"
        # algorithmic bytes of the persistent kernel per env-step: action 8 + new path record 16 read; obs 52 + reward + done written;
        # the state (16 + cash [+ pv_prev]) read and written once per launch of T steps
        b_many = 8 + 16 + 52 + (8 if fp64 else 4) + 1 + 2 * (16 + (16 if fp64 else 4)) / T
        achieved = b_many * n * T / (sweep_ms * 1e-3) / 1e9
        kname = "hedge_step_many_kernel<F64=%s,MON=false>" % ("true" if fp64 else "false")
        tr, tr_per_env = _traffic(kname)
        launch_ms = per_launch_ms / (K * T)
        b_step = B_ALG[args.precision]
        l2_resident = 56 if not fp64 else 80                      # state read + written (20 / 32 B each way) + the lagged path record (16 B)
        ach_step = b_step * n / (launch_ms * 1e-3) / 1e9
        per_step = dict(kernel="hedge_step_kernel<F64=%s,INFO=false>" % ("true" if fp64 else "false"),
                        what="the same sweep as 252 chained launches of the gym-style per-step kernel (closed-loop callers)",
                        launch_us=launch_ms * 1e3, env_steps_per_s=float(n) * world / (launch_ms * 1e-3),
                        algorithmic_bytes_per_env_step=b_step, achieved=ach_step, frac=ach_step / peak,
                        dram_side=dict(bytes_per_env_step=b_step - l2_resident,
                                       gbs=(b_step - l2_resident) * n / (launch_ms * 1e-3) / 1e9,
                                       frac=(b_step - l2_resident) * n / (launch_ms * 1e-3) / 1e9 / peak,
                                       what="at 2^20 envs the env state (read + written every step) and the lagged path record stay in the "
                                            "126 MB L2 between consecutive launches, so frac above counts bytes HBM never moves and can read "
                                            "above 1; this line counts only what HBM itself moves.  The L2-free measurement of the same kernel "
                                            "is roofline.per_step_kernel_l2_free"))
        line = dict(
            metric="env-steps/sec (fused hedge step)", value=value, unit="env-steps/s", n_gpus=world, steps=K, warmup=max(W, 3),
            ms_per_step=sweep_ms, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f64" if fp64 else "f32", data="synthetic", config=_config(args, world),
            gpu_launches=K,
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=(tr_per_env * n * T if tr_per_env is not None else None),
                          traffic_bytes_per_env_step=tr_per_env, traffic_source=(tr or {}).get("source"),
                          kernel=kname, algorithmic_bytes_per_env_step=b_many, launch_us=sweep_ms * 1e3, env_steps_per_launch=n * T,
                          peak_source=peak_src,
                          note="one launch = one 252-step episode of every env (cantor_env_step_many: the action tape is known, so the env "
                               "state stays in registers).  Every byte is touched once -- nothing can be served from L2 -- so achieved is "
                               "HBM traffic.  The gym-style one-launch-per-step kernel is reported next to it",
                          per_step_kernel=per_step,
                          per_step_kernel_l2_free=l2free.get("per_step_replay_fp32")),
            e2e=dict(value=float(n) * world * T * args.e2e_steps / e2e_s, unit="env-steps/s",
                     h2d_bytes_per_step=n * 8 * T, d2h_bytes_per_step=n * (52 + (8 if fp64 else 4) + 1) * T,
                     api="HostVecEnv.step -> cantor_vecenv_step_host: NumPy actions in page-locked host memory -> obs/reward/"
                         "done in page-locked host memory, every env step (8 chunks over 3 streams), returns after the D2H"),
            clocks=clocks,
            extra=dict(path_sim_reprice=dict(kernel="sim_paths_kernel<GBM> (K1 fused with K2 ATM repricing)", ms=sim_ms,
                                             path_steps_per_s=float(n) * T / (sim_ms * 1e-3),
                                             hbm_write_gbs=float(n) * (T + 1) * 16 / (sim_ms * 1e-3) / 1e9)))
        if probe is not None:
            ceil = float(ps[2]) * n                      # env-steps/s the raw copies alone would sustain, summed over ranks
            slowest = float(pmin[0]) * n                 # the slowest rank's share (the box's GPUs do not get equal shares of its host side)
            e2e_v = line["e2e"]["value"]
            line["e2e"]["copy_ceiling"] = dict(
                what="cantor_host_copy_probe on every rank at once: the same bytes per step (57 out + 8 in per env, 8 chunks, 3 streams, "
                     "stream sync after every round) as plain cudaMemcpyAsync between page-locked host memory and HBM, no kernel",
                d2h_gbs_total=float(ps[0]), h2d_gbs_total=float(ps[1]), env_steps_per_s=ceil, e2e_over_ceiling=e2e_v / ceil if ceil else None,
                slowest_rank_env_steps_per_s=slowest, fastest_to_slowest_note="ranks are not in lock-step: a rank that finishes early leaves its share to the others",
                d2h_only_unsynchronised_gbs_total=float(ps[3]), e2e_d2h_gbs=e2e_v * (52 + (8 if fp64 else 4) + 1) / 1e9)
        if l2free:
            line["extra"]["step_kernels_beyond_l2"] = l2free
        if stats_check is not None:
            line["extra"]["stats_check"] = stats_check
        if forms is not None:
            line["extra"]["parity_forms_and_formats"] = forms
        if roll is not None:
            rs = roll[1]
            line["extra"]["rollout_on_the_fly"] = dict(
                kernel="rollout_kernel<GBM on the fly, delta_every_step policy, episode statistics> + all-reduce of the "
                       "statistics buffers once per sweep (configs[3] shape: 2^23 envs per GPU, 64 M on 8)", stats_all_reduce=roll[2], sweeps=args.rollout_steps, ms_per_sweep=roll_ms / args.rollout_steps,
                envs_per_gpu=args.rollout_envs,
                env_steps_per_s=float(args.rollout_envs) * world * T * args.rollout_steps / (roll_ms * 1e-3),
                stats={k: rs[k] for k in ("n_episodes", "mean_abs_pnl", "std_abs_pnl", "mean_cost", "mean_reward", "cvar95_abs_pnl")})
        if mlp_roll is not None:
            steps5 = float(mlp_roll[1]) * world * args.mlp_rollout_steps
            line["extra"]["rollout_mlp_policy"] = dict(
                kernel="rollout_kernel<GBM on the fly, MLP 13-64-64-2 actor as bf16 tcgen05.mma (TMEM accumulators), env step, "
                       "episode statistics> + all-reduce of the statistics", envs_per_gpu=mlp_roll[1],
                steps=args.mlp_rollout_steps, ms=mlp_ms,
                env_steps_per_s=steps5 / (mlp_ms * 1e-3),
                actor_tflops_issued=steps5 * 2 * (16 * 64 + 80 * 64 + 80 * 16) / (mlp_ms * 1e-3) / 1e12,
                actor_tflops_useful=steps5 * 2 * (13 * 64 + 64 * 64 + 64 * 2) / (mlp_ms * 1e-3) / 1e12,
                n_episodes=mlp_roll[2]["n_episodes"])
        if lstm_roll is not None:
            flop = 2 * (4 * 128 * 144 + 64 * 144 + 64 * 80 + 16 * 80)                  # per env-step, as issued (padded K / N)
            flop_useful = 2 * (4 * 128 * (13 + 128) + 64 * 128 + 64 * 64 + 2 * 64)
            steps6 = float(lstm_roll[1]) * world * args.lstm_rollout_steps
            line["extra"]["rollout_lstm_policy"] = dict(
                kernel="rollout_kernel<GBM on the fly, LSTM(13-128) + MLP(128-64-64-2) actor as bf16 tcgen05.mma (warp-specialised "
                       "issuer, TMA-streamed gate weights, cell state in TMEM), env step, episode statistics> + all-reduce of the statistics",
                envs_per_gpu=lstm_roll[1], steps=args.lstm_rollout_steps, ms=lstm_ms,
                env_steps_per_s=steps6 / (lstm_ms * 1e-3),
                actor_tflops_issued=steps6 * flop / (lstm_ms * 1e-3) / 1e12, actor_tflops_useful=steps6 * flop_useful / (lstm_ms * 1e-3) / 1e12,
                mufu_bound_frac=float(lstm_roll[1]) * args.lstm_rollout_steps * 640 / (lstm_ms * 1e-3) / (148 * 16 * 1.965e9),
                n_episodes=lstm_roll[2]["n_episodes"])
        if book_ms is not None:
            cells = float(n) * (T + 1)
            line["extra"]["multi_strike_book"] = dict(
                kernel="book_f32_kernel (Heston book variance, %d strikes, price only)" % args.book_strikes, ms=book_ms,
                reprices_per_s=cells * args.book_strikes / (book_ms * 1e-3),
                algorithmic_gbs=cells * (16 + 8 * args.book_strikes) / (book_ms * 1e-3) / 1e9,
                frac=cells * (16 + 8 * args.book_strikes) / (book_ms * 1e-3) / 1e9 / peak,
                mufu_per_s=cells * (5 + 3 * args.book_strikes) / (book_ms * 1e-3))
        if rb_res is not None:
            line["extra"]["rbergomi_nested_mc"] = dict(
                kernel="rbergomi_price_kernel<tcgen05 split-TF32 FIR> (5000 inner paths x 30 steps per ATM call / put)", ms=rb_res[0],
                pricings=rb_res[1], inner_path_steps_per_s=rb_res[1] * 5000.0 * 30 / (rb_res[0] * 1e-3),
                reference_workload_seconds=100000 * 252 * 2 / (rb_res[1] / (rb_res[0] * 1e-3)))
        line["config"]["host_numa_bind"] = numa_note
        if world == 1 and not args.no_cpu_baseline:
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds, os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
