"""Path simulation and Black-Scholes repricing on the GPU (host-side mirror of ``src/sim`` / ``src/tools``).

Function names follow the reference where one exists:

  ``generate_paths_and_options``   src/sim/rbergomi_sim.py:309-499  (GBM / Heston + closed-form ATM book here)
  ``black_scholes_vectorized``     src/sim/option_price_assignment.py:10-21
  ``calculate_annualized_vol_matrix`` / ``process_price_paths``   option_price_assignment.py:23-52
  ``bs_delta_hedge``               src/tools/bs_delta.py:36-55

Each forwards to one CUDA kernel through the C ABI; tensors are the buffers.  No CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .data import ReplayData

# defaults of the reference simulator (rbergomi_sim.py:13-27) and env (hedging_env_v2.py:57-58)
R, DT, N_STEPS, SEED = 0.04, 1 / 252, 252, 42
T_OPTION_TENOR = 30 / 252
S0_DEFAULT, XI_DEFAULT, RHO_DEFAULT = 100.0, 0.04, -0.7


def _stream(dev):
    return _lib.current_stream_ptr(dev)


def generate_paths_and_options(num_paths, r=R, dt=DT, seed=SEED, *, n_steps=N_STEPS, model="gbm", s0=S0_DEFAULT,
                               v0=XI_DEFAULT, kappa=2.0, theta=XI_DEFAULT, sigma_v=0.5, rho=RHO_DEFAULT,
                               tenor=T_OPTION_TENOR, reprice=True, path_offset=0, device="cuda",
                               out: ReplayData | None = None) -> ReplayData:
    """Simulate ``num_paths`` paths of ``n_steps`` days and (optionally) their ATM option columns, in HBM.

    Returns the packed book (``ReplayData``) that ``HedgingVecEnv`` consumes directly; ``.to_path_major()`` /
    ``.save_npz()`` give the reference's ``paths / volatilities / call_prices_atm / put_prices_atm`` arrays.
    ``path_offset`` is the global index of path 0 (rank * num_paths when sharding over GPUs).
    """
    if model not in ("gbm", "heston"):
        raise ValueError("model must be 'gbm' or 'heston'")
    dev = torch.device(device)
    book = out if out is not None else ReplayData.empty(num_paths, n_steps, dev)
    if book.n_paths != num_paths or book.episode_length != n_steps:
        raise ValueError("out has the wrong shape")
    p = _lib.SimParams(_lib.MODEL_GBM if model == "gbm" else _lib.MODEL_HESTON, int(bool(reprice)), s0, v0, r, dt,
                       kappa, theta, sigma_v, rho, tenor, int(seed) & (2 ** 64 - 1), int(path_offset))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_sim_paths(p, num_paths, n_steps, book.tensor.data_ptr(), book.ld, _stream(dev)),
                   "cantor_sim_paths")
    return book


def reprice_atm(book: ReplayData, r=R, tenor=T_OPTION_TENOR) -> ReplayData:
    """Fill the option columns of a book that holds S and v with Black-Scholes ATM prices (K = round(S_t))."""
    with torch.cuda.device(book.device):
        _lib.check(_lib.lib().cantor_reprice_atm(book.tensor.data_ptr(), book.ld, book.n_paths, book.episode_length,
                                                 r, tenor, _stream(book.device)), "cantor_reprice_atm")
    return book


def euler_from_normals(S0, v, dW1, dW2, rho, r=R, dt=DT, device="cuda"):
    """rbergomi_sim.py:454-464 on exported draws (path-major float64 in, path-major float64 paths out)."""
    dev = torch.device(device)
    v_t = torch.as_tensor(np.asarray(v, np.float64)).to(dev).T.contiguous()            # [T+1, n]
    n, T = v_t.shape[1], v_t.shape[0] - 1
    d1 = torch.as_tensor(np.asarray(dW1, np.float64)[:, :T]).to(dev).T.contiguous()
    d2 = torch.as_tensor(np.asarray(dW2, np.float64)[:, :T]).to(dev).T.contiguous()
    s0 = torch.as_tensor(np.broadcast_to(np.asarray(S0, np.float64), (n,)).copy()).to(dev)
    rh = torch.as_tensor(np.broadcast_to(np.asarray(rho, np.float64), (n,)).copy()).to(dev)
    out = torch.empty((T + 1, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_euler_from_normals(s0.data_ptr(), v_t.data_ptr(), d1.data_ptr(), d2.data_ptr(),
                                                        rh.data_ptr(), n, T, n, r, dt, out.data_ptr(), _stream(dev)),
                   "cantor_euler_from_normals")
    return out.T.contiguous()


def _as_f64(x, dev):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, np.float64))
    return t.to(device=dev, dtype=torch.float64).contiguous()


def black_scholes_vectorized(S, K, T, r, sigma, epsilon=1e-8, device="cuda"):
    """option_price_assignment.py:10-21.  Array or scalar S, K, T, sigma (broadcast like NumPy) -> (call, put) float64."""
    dev = torch.device(device)
    args = [_as_f64(a, dev) for a in (S, K, T, sigma)]
    shape = torch.broadcast_shapes(*(a.shape for a in args))
    n = int(np.prod(shape)) if len(shape) else 1
    flat, strides = [], []
    for a in args:
        if a.numel() == 1:
            flat.append(a.reshape(1))
            strides.append(0)
        else:
            flat.append(a.expand(shape).contiguous().reshape(-1))
            strides.append(1)
    call = torch.empty(n, dtype=torch.float64, device=dev)
    put = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_bs_price(flat[0].data_ptr(), flat[1].data_ptr(), flat[2].data_ptr(), flat[3].data_ptr(),
                                              n, *strides, float(r), float(epsilon), call.data_ptr(), put.data_ptr(),
                                              _stream(dev)), "cantor_bs_price")
    return call.reshape(shape), put.reshape(shape)


def process_price_paths(paths, r=R, strike_multipliers=(1.0,), device="cuda", return_vols=False):
    """option_price_assignment.py:33-52 on path-major ``paths`` (n, T+1) -> ``calls, puts`` (n, T+1) float64 ("schema B").

    With several ``strike_multipliers`` the outputs gain a leading strike axis (M, n, T+1): K_m = round(S_0) * mult_m.
    """
    dev = torch.device(device)
    p = _as_f64(paths, dev)
    n, T1 = p.shape
    tm = p.T.contiguous()
    mult = _as_f64(list(strike_multipliers), dev)
    M = mult.numel()
    vols = torch.empty((T1, n), dtype=torch.float64, device=dev)
    calls = torch.empty((M, T1, n), dtype=torch.float64, device=dev)
    puts = torch.empty((M, T1, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_schema_b_book(tm.data_ptr(), n, T1 - 1, n, float(r), mult.data_ptr(), M,
                                                   vols.data_ptr(), calls.data_ptr(), puts.data_ptr(), _stream(dev)),
                   "cantor_schema_b_book")
    calls, puts = calls.transpose(1, 2), puts.transpose(1, 2)
    if M == 1:
        calls, puts = calls[0], puts[0]
    return (calls, puts, vols.T) if return_vols else (calls, puts)


def reprice_book(book: ReplayData, strike_multipliers=(1.0,), r=R, sigma="realised", tenor=None, greeks=False):
    """Float32 multi-strike Black-Scholes book along every path of a packed book (BASELINE configs[2]).

    ``process_price_paths`` (option_price_assignment.py:33-52) in throughput form: K_m = round(S_0) * mult_m, maturity to
    the episode end (``tenor=None``) or a fixed ``tenor``; ``sigma="realised"`` is the reference's prefix volatility,
    ``sigma="book"`` uses the book's instantaneous variance (Heston).  Returns time-major device tensors
    ``calls, puts`` of shape ``[M, T+1, n_paths]`` (+ ``deltas, gammas`` with ``greeks=True``); ``.permute(0, 2, 1)``
    gives the reference's path-major ``(n, T+1)`` planes.
    """
    if sigma not in ("realised", "book"):
        raise ValueError("sigma must be 'realised' or 'book'")
    dev = book.device
    mult = torch.as_tensor(np.asarray(list(strike_multipliers), np.float32)).to(dev)
    M, T1, ld = int(mult.numel()), book.episode_length + 1, book.ld
    outs = [torch.empty((M, T1, ld), dtype=torch.float32, device=dev) for _ in range(4 if greeks else 2)]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_reprice_book(
            book.tensor.data_ptr(), ld, book.n_paths, book.episode_length, float(r), mult.data_ptr(), M,
            _lib.SIGMA_REALISED if sigma == "realised" else _lib.SIGMA_BOOK_VARIANCE, float(tenor or 0.0),
            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr() if greeks else None,
            outs[3].data_ptr() if greeks else None, _stream(dev)), "cantor_reprice_book")
    return tuple(o[:, :, :book.n_paths] for o in outs)


def calculate_annualized_vol_matrix(paths, device="cuda"):
    """option_price_assignment.py:23-31: realised annualised volatility of each path prefix (n, T+1) float64."""
    return process_price_paths(paths, device=device, return_vols=True)[2]


def bs_delta_hedge(paths, r=R, dt=DT, device="cuda"):
    """src/tools/bs_delta.py:36-55: (n, T+1) float64 paths -> (n, T+1) float64 delta-hedge P&L."""
    dev = torch.device(device)
    p = _as_f64(paths, dev)
    n, T1 = p.shape
    tm = p.T.contiguous()
    pnl = torch.empty((T1, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_bs_delta_hedge(tm.data_ptr(), n, T1 - 1, n, float(r), float(dt), pnl.data_ptr(),
                                                    _stream(dev)), "cantor_bs_delta_hedge")
    return pnl.T
