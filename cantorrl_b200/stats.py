"""Episode statistics of the evaluation loops, as a small vector that shards combine by addition.

The reference computes, over finished episodes (one list entry per episode, Python lists on the host):

  ``evaluate_baseline_policy``  src/agents/baselines.py:49-65      a = mean_t |per_share_step_pnl|, c = mean_t cost;
                                                                    mean / std over episodes
  ``run_evaluation``            src/agents/train_ppo_v2.py:482-530 b = |sum_t per_share_step_pnl| / T, c = sum_t cost / T,
                                                                    R = sum_t reward; mean, std, CVaR95 = mean of the
                                                                    top 5 % of sorted b (index int(0.95 n) onward)

Here the kernels reduce ``n, sum x, sum x^2`` of each quantity (``include/cantor_hedge.h``: ``cantor_stats_out``)
plus a fixed-bin histogram of ``b`` (counts and per-bin sums), so the statistics of an env population sharded
over GPUs by path index are ONE all-reduce(sum) of ``16 + 2 * bins`` numbers -- the only collective on the path.
``EpisodeStats`` holds those buffers (caller-owned torch tensors), all-reduces them over ``torch.distributed``
(NCCL on GPUs; any backend works, the tests use gloo) and turns them into the reference's statistics.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib

# sums[] layout (cantor_hedge.h)
N_EPISODES, A_SUM, A_SQ, B_SUM, B_SQ, C_SUM, C_SQ, R_SUM, R_SQ, S_SUM, S_SQ, ENV_STEPS = range(12)


def _mean_std(n, s, sq):
    if n <= 0:
        return float("nan"), float("nan")
    mean = s / n
    return mean, math.sqrt(max(sq / n - mean * mean, 0.0))          # np.std: population standard deviation


class EpisodeStats:
    """Accumulators for the per-episode hedging-error statistics (device tensors, zeroed on construction).

    hist_bins / hist_max  fixed-bin histogram of ``b`` over ``[0, hist_max)`` for CVaR95 (values beyond the range
                          land in the last bin; their exact sum is kept, so the tail mean stays exact)
    episode_slots         optionally keep every episode's ``b`` (``[episode_slots, n_envs]`` floats) for an exact,
                          sort-based CVaR on one GPU
    """

    def __init__(self, device="cuda", hist_bins: int = 4096, hist_max: float = 4.0, episode_slots: int = 0,
                 n_envs: int = 0):
        dev = torch.device(device)
        self.device = dev
        # local accumulators (what a kernel launch on this GPU adds into)
        self._l_sums = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=dev)
        self.hist_bins, self.hist_max = int(hist_bins), float(hist_max)
        self._l_hist = torch.zeros(self.hist_bins, dtype=torch.int64, device=dev) if hist_bins > 0 else None   # u64 counts
        self._l_hist_sum = torch.zeros(self.hist_bins, dtype=torch.float64, device=dev) if hist_bins > 0 else None
        # fused all-reduce (enable_fused_all_reduce): symmetric "global" block that every rank's kernels add into
        self._g = self._hdl = self._ticket = None
        self._c = None                              # the cantor_stats_out struct handed to the C ABI (see c_struct)
        self.fused_transport = None
        self.episode_slots = int(episode_slots)
        self.episode_b = (torch.full((self.episode_slots, int(n_envs)), float("nan"), dtype=torch.float32, device=dev)
                          if episode_slots > 0 else None)

    # the buffers that hold the totals: the symmetric global block when the all-reduce is fused into the kernels, else local
    @property
    def sums(self):
        return self._l_sums if self._g is None else self._g[:_lib.STATS_LEN]

    @property
    def hist(self):
        if self._g is None or self._l_hist is None:
            return self._l_hist
        return self._g[_lib.STATS_LEN:_lib.STATS_LEN + self.hist_bins].view(torch.int64)

    @property
    def hist_sum(self):
        if self._g is None or self._l_hist_sum is None:
            return self._l_hist_sum
        return self._g[_lib.STATS_LEN + self.hist_bins:_lib.STATS_LEN + 2 * self.hist_bins]

    def zero_(self):
        self._l_sums.zero_()
        if self._l_hist is not None:
            self._l_hist.zero_()
            self._l_hist_sum.zero_()
        if self.episode_b is not None:
            self.episode_b.fill_(float("nan"))
        if self._g is not None:                 # nobody may still be adding into a block that is being cleared, and vice versa
            self._hdl.barrier(channel=1)
            self._g.zero_()
            self._hdl.barrier(channel=1)
        return self

    def enable_fused_all_reduce(self, group=None, transport: str = "auto") -> str:
        """Fuse the all-reduce into the kernels: the last CTA of every launch adds the launch's statistics into every
        rank's copy of a symmetric memory block -- with ``multimem.red`` through the NVLS multicast mapping (the NVSwitch
        does the reduction) when the fabric offers one, else with peer atomics over NVLink -- so ``all_reduce`` shrinks
        to a barrier and no NCCL collective runs.  Collective call (every rank of ``group``).  Returns the transport
        (``"multimem"`` or ``"p2p"``; ``transport="p2p"`` forces the peer-atomic form); raises if symmetric memory cannot be
        set up (callers fall back to NCCL)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if self.device.type != "cuda":
            raise RuntimeError("fused all-reduce needs CUDA devices")
        group = group if group is not None else dist.group.WORLD
        world = dist.get_world_size(group)
        if world > 8:
            raise RuntimeError("fused all-reduce is built for one NVSwitch domain (<= 8 ranks)")
        words = _lib.STATS_LEN + 2 * self.hist_bins
        g = symm.empty(words, dtype=torch.float64, device=self.device)
        g.zero_()
        hdl = symm.rendezvous(g, group)
        self._g, self._hdl = g, hdl
        self._ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        if transport not in ("auto", "p2p"):
            raise ValueError("transport must be 'auto' or 'p2p'")
        self._mc = int(hdl.multicast_ptr) if transport == "auto" and hdl.multicast_ptr else 0
        self._peers = [int(p) for p in hdl.buffer_ptrs]
        self.fused_transport = "multimem" if self._mc else "p2p"
        hdl.barrier(channel=1)                  # every block is zero before anyone adds into it
        self.c_struct()                         # refresh the struct in place: envs / rollouts holding a pointer to it switch too
        return self.fused_transport

    def c_struct(self) -> _lib.StatsOut:
        """The ``cantor_stats_out`` view of these buffers.  ONE struct object per EpisodeStats, refreshed in place: an env that
        stored a pointer to it at construction (``HedgingVecEnv(stats=...)``) sees a later ``enable_fused_all_reduce``."""
        if self._c is None:
            self._c = _lib.StatsOut()
        st = self._c
        st.sums, st.hist, st.hist_sum = self._l_sums.data_ptr(), _lib.ptr(self._l_hist), _lib.ptr(self._l_hist_sum)
        st.episode_b, st.hist_max, st.hist_bins, st.episode_slots = _lib.ptr(self.episode_b), self.hist_max, self.hist_bins, self.episode_slots
        if self._g is not None:
            st.mc_global = self._mc or None
            for r, p in enumerate(self._peers):
                st.peer_global[r] = p
            st.n_peers = len(self._peers)
            st.ticket = self._ticket.data_ptr()
        return st

    # ------------------------------------------------------------------------------------------ collective
    def all_reduce(self, group=None, async_op: bool = False):
        """Sum the accumulators over the ranks of ``group`` (no-op without an initialised process group).

        One flat float64 buffer would need a dtype change of the counts; two small all-reduces keep the counts
        exact (int64) and the sums in float64.  Both are latency-bound (<= 64 KB) on NVLink / NVSwitch.
        """
        import torch.distributed as dist
        if self._g is not None:      # fused: the kernels already added into every rank's block; wait for all ranks' launches
            self._hdl.barrier(channel=0)
            return []
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return []
        works = [dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=group, async_op=async_op)]
        if self.hist is not None:
            works.append(dist.all_reduce(self.hist, op=dist.ReduceOp.SUM, group=group, async_op=async_op))
            works.append(dist.all_reduce(self.hist_sum, op=dist.ReduceOp.SUM, group=group, async_op=async_op))
        return works

    # --------------------------------------------------------------------------------------------- results
    def cvar95(self, level: float = 0.95) -> Optional[float]:
        """Mean of the top ``n - int(level * n)`` values of ``b`` (train_ppo_v2.py:527-530) from the histogram.

        Whole bins above the cut contribute their exact sums; the bin that straddles the cut contributes its
        mean value for the remaining count (error <= one bin width times that bin's share of the tail).
        """
        if self.hist is None:
            return None
        cnt = self.hist.cpu().numpy()
        sm = self.hist_sum.cpu().numpy()
        n = int(cnt.sum())
        if n == 0:
            return float("nan")
        k = n - int(level * n)                      # len(sorted[int(0.95 n):])
        if k <= 0:
            return float("nan")
        need, total = k, 0.0
        for j in range(self.hist_bins - 1, -1, -1):
            c = int(cnt[j])
            if c == 0:
                continue
            if c <= need:
                total += float(sm[j])
                need -= c
            else:
                total += float(sm[j]) / c * need
                need = 0
            if need == 0:
                break
        return total / k

    def cvar95_exact(self, level: float = 0.95) -> Optional[float]:
        """Sort-based CVaR over the kept per-episode values (single GPU; NaN slots = episodes not finished)."""
        if self.episode_b is None:
            return None
        b = self.episode_b.flatten()
        b = b[~torch.isnan(b)]
        if b.numel() == 0:
            return float("nan")
        sb = torch.sort(b.double()).values
        return float(sb[int(level * sb.numel()):].mean())

    def result(self) -> dict:
        """The reference's statistics (names follow train_ppo_v2.py:520-530 / baselines.py:63-65)."""
        s = self.sums.cpu().numpy()
        n = float(s[N_EPISODES])
        out = {"n_episodes": int(n), "env_steps": int(s[ENV_STEPS])}
        out["mean_abs_pnl_baseline"], out["std_abs_pnl_baseline"] = _mean_std(n, s[A_SUM], s[A_SQ])
        out["mean_abs_pnl"], out["std_abs_pnl"] = _mean_std(n, s[B_SUM], s[B_SQ])
        out["mean_cost"], out["std_cost"] = _mean_std(n, s[C_SUM], s[C_SQ])
        out["mean_reward"], out["std_reward"] = _mean_std(n, s[R_SUM], s[R_SQ])
        out["mean_signed_pnl"], out["std_signed_pnl"] = _mean_std(n, s[S_SUM], s[S_SQ])
        out["cvar95_abs_pnl"] = self.cvar95()
        if self.episode_b is not None:
            out["cvar95_abs_pnl_exact"] = self.cvar95_exact()
        return out
