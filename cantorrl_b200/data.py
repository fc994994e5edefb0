"""Env-schema path data resident in HBM, time-major.

The reference env loads four path-major arrays from an ``.npz`` and casts them to float32
(``src/env/hedging_env_v2.py:36-48``).  Here the same four arrays live on the GPU as
``[T+1, ld]`` / ``[T, ld]`` float32 tensors so that the 32 envs of a warp read one 128-byte line per
array and time slab.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

SCHEMA_A_KEYS = ("paths", "volatilities", "call_prices_atm", "put_prices_atm")   # rbergomi_sim.py:528


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class ReplayData:
    """Time-major float32 device copies of ``paths``, ``volatilities``, ``call_prices_atm``, ``put_prices_atm``."""

    def __init__(self, S: torch.Tensor, v: torch.Tensor, C: torch.Tensor, P: torch.Tensor, n_paths: int):
        for t in (S, v, C, P):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.dim() != 2:
                raise ValueError("ReplayData tensors must be contiguous 2-D float32 CUDA tensors")
        if not (S.shape == v.shape and C.shape == P.shape and S.shape[1] == C.shape[1]
                and S.shape[0] == C.shape[0] + 1 and 0 < n_paths <= S.shape[1]):
            raise ValueError("Data shapes are inconsistent.")
        self.S, self.v, self.C, self.P = S, v, C, P
        self.n_paths = int(n_paths)
        self.ld = int(S.shape[1])
        self.episode_length = int(S.shape[0] - 1)
        self.device = S.device

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, paths, volatilities, call_prices_atm, put_prices_atm, device="cuda") -> "ReplayData":
        """Path-major ``(n, T+1)/(n, T)`` host arrays (any float dtype) -> time-major float32 on ``device``."""
        arrs = [np.asarray(a) for a in (paths, volatilities, call_prices_atm, put_prices_atm)]
        S, V, Cc, Pp = arrs
        # hedging_env_v2.py:45-48
        if not (S.ndim == 2 and S.shape == V.shape and Cc.ndim == 2 and Pp.ndim == 2
                and S.shape[0] == Cc.shape[0] == Pp.shape[0] and S.shape[1] == Cc.shape[1] + 1 == Pp.shape[1] + 1):
            raise ValueError("Data shapes are inconsistent.")
        n = S.shape[0]
        ld = _round_up(n, 32)                 # 128-byte rows
        out = []
        for a in arrs:
            t = torch.zeros((a.shape[1], ld), dtype=torch.float32, device=device)
            t[:, :n] = torch.from_numpy(np.ascontiguousarray(a.astype(np.float32).T)).to(device)
            out.append(t)
        return cls(*out, n_paths=n)

    @classmethod
    def from_npz(cls, data_file_path, device="cuda") -> "ReplayData":
        """``np.load`` of the reference's env-schema file; any load/parse failure is a FileNotFoundError (:42-43)."""
        try:
            with np.load(data_file_path) as data:
                arrs = [data[k] for k in SCHEMA_A_KEYS]
        except Exception as e:  # same catch-all as the reference
            raise FileNotFoundError(f"Could not load or parse data from {data_file_path}. Error: {e}")
        return cls.from_arrays(*arrs, device=device)

    # -- C ABI view -----------------------------------------------------------------------------
    def book(self) -> _lib.ReplayBook:
        return _lib.ReplayBook(self.S.data_ptr(), self.v.data_ptr(), self.C.data_ptr(), self.P.data_ptr(),
                               self.ld, self.n_paths, self.episode_length)

    def to_path_major(self):
        """Back to the npz layout (host float32 arrays), e.g. to feed the CPU oracle."""
        n = self.n_paths
        return tuple(t[:, :n].T.contiguous().cpu().numpy() for t in (self.S, self.v, self.C, self.P))
