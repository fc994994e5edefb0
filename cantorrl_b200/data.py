"""Env-schema path data resident in HBM as one packed, time-major book.

The reference env loads four path-major arrays from an ``.npz`` and casts them to float32
(``src/env/hedging_env_v2.py:36-48``).  Here the same data live on the GPU as ONE float32 tensor
``book[t, path] = (S, v, C, P)`` of shape ``[T+1, ld, 4]`` so an env-step is two aligned 16-byte loads and
the 32 envs of a warp read 512 contiguous bytes.  Row ``T`` repeats the option marks of row ``T-1`` -- the
stale marks the reference uses at the terminal step (``:226-231``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

SCHEMA_A_KEYS = ("paths", "volatilities", "call_prices_atm", "put_prices_atm")   # rbergomi_sim.py:528


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class ReplayData:
    """Packed float32 device copy of ``paths``, ``volatilities``, ``call_prices_atm``, ``put_prices_atm``."""

    def __init__(self, book: torch.Tensor, n_paths: int):
        if (book.dtype != torch.float32 or not book.is_cuda or not book.is_contiguous() or book.dim() != 3
                or book.shape[2] != 4 or book.shape[0] < 2 or not 0 < n_paths <= book.shape[1]):
            raise ValueError("ReplayData book must be a contiguous float32 CUDA tensor [T+1, ld, 4] with 0 < n_paths <= ld")
        self.tensor = book
        self.n_paths = int(n_paths)
        self.ld = int(book.shape[1])
        self.episode_length = int(book.shape[0] - 1)
        self.device = book.device

    # strided views in the reference's vocabulary (time-major)
    @property
    def S(self):
        return self.tensor[:, :, 0]

    @property
    def v(self):
        return self.tensor[:, :, 1]

    @property
    def C(self):
        return self.tensor[:, :, 2]

    @property
    def P(self):
        return self.tensor[:, :, 3]

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def empty(cls, n_paths: int, episode_length: int, device="cuda") -> "ReplayData":
        """Uninitialised book for a simulator kernel to fill."""
        ld = _round_up(int(n_paths), 8)
        return cls(torch.empty((episode_length + 1, ld, 4), dtype=torch.float32, device=device), n_paths)

    @classmethod
    def from_arrays(cls, paths, volatilities, call_prices_atm, put_prices_atm, device="cuda") -> "ReplayData":
        """Path-major ``(n, T+1)/(n, T)`` arrays (NumPy or torch; float32 or float64) -> packed book on ``device``.

        The cast to float32 (hedging_env_v2.py:38-41), the transposition and the interleave run in one kernel.
        """
        arrs = []
        for a in (paths, volatilities, call_prices_atm, put_prices_atm):
            if isinstance(a, torch.Tensor):
                arrs.append(a)
            else:
                a = np.asarray(a)
                if a.dtype not in (np.float32, np.float64):
                    a = a.astype(np.float64)
                arrs.append(torch.from_numpy(np.ascontiguousarray(a)))
        S, V, Cc, Pp = arrs
        # hedging_env_v2.py:45-48
        if not (S.dim() == 2 and S.shape == V.shape and Cc.dim() == 2 and Pp.dim() == 2
                and S.shape[0] == Cc.shape[0] == Pp.shape[0] and S.shape[1] == Cc.shape[1] + 1 == Pp.shape[1] + 1):
            raise ValueError("Data shapes are inconsistent.")
        dt = torch.float64 if any(a.dtype == torch.float64 for a in arrs) else torch.float32
        dev = torch.device(device)
        arrs = [a.to(device=dev, dtype=dt).contiguous() for a in arrs]
        n, T = int(S.shape[0]), int(S.shape[1]) - 1
        if T < 1 or n < 1:
            raise ValueError("Data shapes are inconsistent.")
        out = cls.empty(n, T, dev)
        out.tensor.zero_()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cantor_pack_book(
                arrs[0].data_ptr(), arrs[1].data_ptr(), arrs[2].data_ptr(), arrs[3].data_ptr(),
                _lib.F64 if dt == torch.float64 else _lib.F32, n, T, out.tensor.data_ptr(), out.ld,
                _lib.current_stream_ptr(dev)), "cantor_pack_book")
        return out

    @classmethod
    def from_time_major(cls, S, v, C, P, n_paths=None) -> "ReplayData":
        """Time-major float32 CUDA tensors ``S, v [T+1, n]``, ``C, P [T, n]`` -> packed book (torch copy; tests/tools)."""
        T1, n = S.shape
        out = cls.empty(n if n_paths is None else n_paths, T1 - 1, S.device)
        out.tensor.zero_()
        out.tensor[:, :n, 0] = S
        out.tensor[:, :n, 1] = v
        out.tensor[:T1 - 1, :n, 2] = C
        out.tensor[:T1 - 1, :n, 3] = P
        out.tensor[T1 - 1, :n, 2] = C[T1 - 2]
        out.tensor[T1 - 1, :n, 3] = P[T1 - 2]
        return out

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
        return _lib.ReplayBook(self.tensor.data_ptr(), self.ld, self.n_paths, self.episode_length)

    def to_path_major(self, dtype=torch.float32):
        """Back to the npz layout: dict of path-major device tensors (``dtype`` float32 or float64)."""
        n, T, dev = self.n_paths, self.episode_length, self.device
        out = {"paths": torch.empty((n, T + 1), dtype=dtype, device=dev),
               "volatilities": torch.empty((n, T + 1), dtype=dtype, device=dev),
               "call_prices_atm": torch.empty((n, T), dtype=dtype, device=dev),
               "put_prices_atm": torch.empty((n, T), dtype=dtype, device=dev)}
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cantor_unpack_book(
                self.tensor.data_ptr(), self.ld, n, T, _lib.F64 if dtype == torch.float64 else _lib.F32,
                *(out[k].data_ptr() for k in SCHEMA_A_KEYS), _lib.current_stream_ptr(dev)), "cantor_unpack_book")
        return out

    def save_npz(self, path, compressed=True):
        """Write the reference's env-schema file (float64 arrays like rbergomi_sim.py:528)."""
        arrs = {k: v.cpu().numpy() for k, v in self.to_path_major(torch.float64).items()}
        (np.savez_compressed if compressed else np.savez)(path, **arrs)

    # -- bare price paths (data/paths.npy) -> env schema ---------------------------------------------------------
    @classmethod
    def from_paths(cls, paths, variances="realised", r=0.04, tenor=30 / 252, device="cuda") -> "ReplayData":
        """Env-schema book from a bare ``(n, T+1)`` price array such as the reference's ``data/paths.npy``.

        The reference ships price paths and a schema-B option file, but no env-schema file;
        this builds one on the GPU: ``volatilities`` = per-step VARIANCE -- ``"realised"``: the square of the reference's
        realised annualised volatility of the path prefix (``option_price_assignment.py:23-31``; columns 0 and 1, which
        are 0 / NaN there, take column 2's value), or a constant (e.g. xi = 0.02903 from ``estimate_base_params``) --
        and ``call_prices_atm / put_prices_atm`` = closed-form ATM Black-Scholes (K = round(S_t), tenor 30/252) in the
        kernel that K1 uses.
        """
        from . import sim                                    # local import: sim imports this module
        dev = torch.device(device)
        p = torch.as_tensor(np.asarray(paths, np.float64) if not isinstance(paths, torch.Tensor) else paths).to(dev, torch.float64)
        if p.dim() != 2 or p.shape[1] < 3:
            raise ValueError("Data shapes are inconsistent.")
        n, T = int(p.shape[0]), int(p.shape[1]) - 1
        if isinstance(variances, str):
            if variances != "realised":
                raise ValueError("variances must be 'realised', a number or an array")
            vol = sim.calculate_annualized_vol_matrix(p, device=dev).clone()
            vol[:, 0] = vol[:, 2]
            vol[:, 1] = vol[:, 2]
            v = vol * vol
        elif np.isscalar(variances):
            v = torch.full_like(p, float(variances))
        else:
            v = torch.as_tensor(np.asarray(variances, np.float64)).to(dev)
            if v.shape != p.shape:
                raise ValueError("Data shapes are inconsistent.")
        zeros = torch.zeros((n, T), dtype=torch.float64, device=dev)
        book = cls.from_arrays(p, v, zeros, zeros, device=dev)
        return sim.reprice_atm(book, r=r, tenor=tenor)

    @classmethod
    def from_paths_npy(cls, path, **kw) -> "ReplayData":
        """``np.load`` of a ``paths.npy``-style file -> ``from_paths``; load failures are FileNotFoundError like the env's."""
        try:
            arr = np.load(path)
        except Exception as e:
            raise FileNotFoundError(f"Could not load or parse data from {path}. Error: {e}")
        return cls.from_paths(arr, **kw)


SCHEMA_B_KEYS = ("calls", "puts")                                                  # option_price_assignment.py:51


def save_schema_b(path, calls, puts, compressed=False):
    """``np.savez(OUTPUT_FILE, calls=..., puts=...)`` (option_price_assignment.py:51): path-major ``(n, T+1)`` float64."""
    c = calls.detach().cpu().numpy() if isinstance(calls, torch.Tensor) else np.asarray(calls)
    q = puts.detach().cpu().numpy() if isinstance(puts, torch.Tensor) else np.asarray(puts)
    if c.shape != q.shape or c.ndim not in (2, 3):
        raise ValueError("Data shapes are inconsistent.")
    (np.savez_compressed if compressed else np.savez)(path, calls=c.astype(np.float64), puts=q.astype(np.float64))


def load_schema_b(path, device="cuda"):
    """The reference's ``data/paths_options.npz`` -> ``(calls, puts)`` device tensors, float64, path-major."""
    try:
        with np.load(path) as z:
            c, q = z["calls"], z["puts"]
    except Exception as e:
        raise FileNotFoundError(f"Could not load or parse data from {path}. Error: {e}")
    return torch.as_tensor(c).to(device), torch.as_tensor(q).to(device)
