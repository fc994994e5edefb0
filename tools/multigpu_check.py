#!/usr/bin/env python
"""Launched with torchrun (one rank per GPU): statistics of a population sharded by global env index and all-reduced
over NCCL must equal the statistics of the whole population computed on one GPU (rank 0 checks and prints OK)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200.distributed import init_process_group, rank_world, shard  # noqa: E402
from cantorrl_b200.rollout import HedgingRollout  # noqa: E402

KW = dict(slippage_bps=1.0, theta_weight=2e-4, pnl_penalty_weight=1e-3, lambda_cost=1e-4)


def main():
    rank, local_rank, world = rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    init_process_group("nccl", dev)
    total, T, steps = 300_001, 21, 50
    sim = dict(model="heston", seed=9, n_steps=T)
    off, cnt = shard(total, rank, world)
    part = HedgingRollout(simulate=sim, num_envs=cnt, env_offset=off, total_envs=total, device=dev, **KW)
    st = part.run(steps, "delta_benchmark").stats
    st.all_reduce()
    torch.cuda.synchronize()
    # the same statistics with the all-reduce fused into the kernel epilogue (multimem.red / peer atomics, no NCCL collective)
    fused_note = "fused all-reduce unavailable"
    fused_ok = True
    try:
      for want in ("auto", "p2p"):
        fs = part.new_stats()
        transport = fs.enable_fused_all_reduce(transport=want)
        for rep in range(2):                                  # twice: local accumulators must come back clean, global must re-zero
            fs.zero_()
            part.run(steps, "delta_benchmark", stats=fs)
            fs.all_reduce()
            torch.cuda.synchronize()
            fused_ok &= torch.equal(fs.hist, st.hist)
            fused_ok &= bool(np.allclose(fs.sums.cpu().numpy(), st.sums.cpu().numpy(), rtol=1e-10, atol=0))
            fused_ok &= bool(np.allclose(fs.hist_sum.cpu().numpy(), st.hist_sum.cpu().numpy(), rtol=1e-10, atol=0))
            fused_ok &= float(fs._l_sums.abs().sum()) == 0.0 and int(fs._l_hist.sum()) == 0
        fused_note = (fused_note + "; " if want == "p2p" else "") + f"fused all-reduce via {transport}: {'OK' if fused_ok else 'MISMATCH'}"
    except Exception as e:                                    # symmetric memory not available on this box: NCCL stays the path
        fused_note = f"fused all-reduce unavailable ({type(e).__name__}: {str(e)[:200]})"
    ok = fused_ok
    print(f"[rank {rank}] {fused_note}", flush=True)
    # the gym-style step kernel's fused Monitor with the all-reduce in ITS epilogue (replay and on-the-fly), against NCCL
    from cantorrl_b200 import HedgingVecEnv
    from cantorrl_b200.stats import EpisodeStats
    n_env_total, Te = 40_003, 9
    eo, ec = shard(n_env_total, rank, world)
    g = torch.Generator(device=dev).manual_seed(17)
    tape = torch.rand((2 * Te + 3, n_env_total, 2), device=dev, generator=g) * 2 - 1
    env_ok = True
    for enable_late in (False, True):
        est_f, est_n = EpisodeStats(dev, hist_bins=512), EpisodeStats(dev, hist_bins=512)
        if not enable_late:
            est_f.enable_fused_all_reduce()
        envs = [HedgingVecEnv(simulate=dict(model="gbm", seed=3, n_steps=Te), num_envs=ec, total_envs=n_env_total, env_offset=eo,
                              device=dev, monitor=True, stats=st, **KW) for st in (est_f, est_n)]
        if enable_late:                 # enabling the fused transport AFTER the env was constructed must take effect too
            est_f.enable_fused_all_reduce()
        est_f.zero_()
        for e in envs:
            e.reset()
        for t in range(tape.shape[0]):
            a = tape[t, eo:eo + ec].contiguous()
            for e in envs:
                e.step(a)
        est_f.all_reduce()
        est_n.all_reduce()
        torch.cuda.synchronize()
        env_ok &= torch.equal(est_f.hist, est_n.hist) and int(est_n.sums[0]) == 2 * n_env_total
        env_ok &= bool(np.allclose(est_f.sums.cpu().numpy()[:12], est_n.sums.cpu().numpy()[:12], rtol=1e-10, atol=0))
    print(f"[rank {rank}] step-kernel Monitor statistics, fused all-reduce vs NCCL: {'OK' if env_ok else 'MISMATCH'}", flush=True)
    ok &= env_ok
    if rank == 0:
        whole = HedgingRollout(simulate=sim, num_envs=total, device=dev, **KW).run(steps, "delta_benchmark").stats
        ok &= torch.equal(whole.hist, st.hist)
        ok &= bool(np.allclose(whole.sums.cpu().numpy(), st.sums.cpu().numpy(), rtol=1e-10, atol=0))
        ok &= int(st.sums[0]) == total * (steps // T) and int(st.sums[11]) == total * steps
        r = st.result()
        print(f"world={world} n_episodes={r['n_episodes']} mean_abs_pnl={r['mean_abs_pnl']:.6f} cvar95={r['cvar95_abs_pnl']:.6f}")
        print("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
