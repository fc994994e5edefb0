"""Path simulation and Black-Scholes repricing on the GPU (host-side mirror of ``src/sim`` / ``src/tools``).

Function names follow the reference where one exists:

  ``generate_paths_and_options``   src/sim/rbergomi_sim.py:309-499  (GBM / Heston + closed-form ATM book here)
  ``black_scholes_vectorized``     src/sim/option_price_assignment.py:10-21
  ``calculate_annualized_vol_matrix`` / ``process_price_paths``   option_price_assignment.py:23-52
  ``bs_delta_hedge``               src/tools/bs_delta.py:36-55

Each forwards to one CUDA kernel through the C ABI; tensors are the buffers.  No CPU path.
"""
from __future__ import annotations

import ctypes as C

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


def reprice_book(book: ReplayData, strike_multipliers=(1.0,), r=R, sigma="realised", tenor=None, greeks=False, path_range=None,
                 out=None):
    """Float32 multi-strike Black-Scholes book along every path of a packed book (BASELINE configs[2]).

    ``process_price_paths`` (option_price_assignment.py:33-52) in throughput form: K_m = round(S_0) * mult_m, maturity to
    the episode end (``tenor=None``) or a fixed ``tenor``; ``sigma="realised"`` is the reference's prefix volatility,
    ``sigma="book"`` uses the book's instantaneous variance (Heston).  Returns time-major device tensors
    ``calls, puts`` of shape ``[M, T+1, n_paths]`` (+ ``deltas, gammas`` with ``greeks=True``); ``.permute(0, 2, 1)``
    gives the reference's path-major ``(n, T+1)`` planes.

    ``path_range=(first, last)`` prices that slice of the paths only, into arrays of the slice's width (``out``: a tuple of
    such tensors to reuse): the way through a book whose full output does not fit (``reprice_book_slices``).
    """
    if sigma not in ("realised", "book"):
        raise ValueError("sigma must be 'realised' or 'book'")
    dev = book.device
    mult = torch.as_tensor(np.asarray(list(strike_multipliers), np.float32)).to(dev)
    M, T1, ld = int(mult.numel()), book.episode_length + 1, book.ld
    first, last = (0, book.n_paths) if path_range is None else (int(path_range[0]), int(path_range[1]))
    if not 0 <= first < last <= book.n_paths:
        raise ValueError(f"path_range {path_range} outside [0, {book.n_paths}]")
    width = last - first
    out_ld = ld if path_range is None else (width + 3) // 4 * 4
    n_out = 4 if greeks else 2
    if out is None:
        outs = [torch.empty((M, T1, out_ld), dtype=torch.float32, device=dev) for _ in range(n_out)]
    else:
        outs = list(out)
        if len(outs) != n_out or any(o.shape != (M, T1, out_ld) or o.dtype != torch.float32 or not o.is_contiguous() for o in outs):
            raise ValueError(f"out must be {n_out} contiguous float32 tensors of shape {(M, T1, out_ld)}")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_reprice_book_strided(
            book.tensor.data_ptr() + 16 * first, ld, width, book.episode_length, float(r), mult.data_ptr(), M,
            _lib.SIGMA_REALISED if sigma == "realised" else _lib.SIGMA_BOOK_VARIANCE, float(tenor or 0.0), out_ld,
            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr() if greeks else None,
            outs[3].data_ptr() if greeks else None, _stream(dev)), "cantor_reprice_book_strided")
    return tuple(o[:, :, :width] for o in outs)


def reprice_book_slices(book: ReplayData, strike_multipliers=(1.0,), paths_per_slice=1 << 21, **kw):
    """``reprice_book`` over consecutive path slices, yielding ``(first, last, outputs)``; the output tensors are REUSED from
    slice to slice (copy or consume them before advancing).  BASELINE configs[2]: 2^24 Heston paths x 8 strikes."""
    bufs = None
    for first in range(0, book.n_paths, int(paths_per_slice)):
        last = min(first + int(paths_per_slice), book.n_paths)
        if bufs is not None and bufs[0].shape[2] != (last - first + 3) // 4 * 4:
            bufs = None
        res = reprice_book(book, strike_multipliers, path_range=(first, last), out=bufs, **kw)
        if bufs is None:
            bufs = tuple(r._base if r._base is not None else r for r in res)
        yield first, last, res


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


# ---------------------------------------------------------------------------------------------- rough Bergomi
# module constants of the reference simulator (rbergomi_sim.py:13-40)
N_PATHS_OPTION_MC = 5000
H_DEFAULT, ETA_DEFAULT = 0.1, 1.0
RBERGOMI_DEFAULTS = dict(s0=S0_DEFAULT, xi=XI_DEFAULT, H=H_DEFAULT, eta=ETA_DEFAULT, rho=RHO_DEFAULT,
                         perturb_s0=0.01, perturb_xi=0.20, perturb_H=0.20, perturb_eta=0.20, perturb_rho=0.10,
                         min_xi_factor=0.5, min_eta_factor=0.5, clip_H_min=0.01, clip_H_max=0.49, clip_rho_min=-0.99,
                         clip_rho_max=-0.01)


def _rb_params(base_params, r, dt, tenor, n_mc, shared_draws, seed, path_offset, tensor_cores=True):
    kw = dict(RBERGOMI_DEFAULTS)
    if base_params is not None:
        if isinstance(base_params, dict):
            unknown = set(base_params) - set(kw)
            if unknown:
                raise TypeError(f"unknown rBergomi parameters: {sorted(unknown)}")
            kw.update(base_params)
        else:                                         # (S0, xi, H, eta, rho) as estimate_base_params returns them (:171-193)
            kw.update(dict(zip(("s0", "xi", "H", "eta", "rho"), (float(x) for x in base_params))))
    return _lib.RbergomiParams(*(float(kw[n]) for n, _ in _lib.RbergomiParams._fields_[:16]), float(r), float(dt), float(tenor),
                               int(n_mc), int(bool(shared_draws)), int(bool(tensor_cores)), 0, int(seed) & (2 ** 64 - 1),
                               int(path_offset))


class RbergomiBook:
    """What ``generate_rbergomi_paths_and_options`` returns: the env-schema book plus the per-path parameters."""

    def __init__(self, book: ReplayData, path_params: torch.Tensor, params: "_lib.RbergomiParams"):
        self.book, self.path_params, self._params = book, path_params, params

    S0 = property(lambda self: self.path_params[0])
    xi = property(lambda self: self.path_params[1])
    H = property(lambda self: self.path_params[2])
    eta = property(lambda self: self.path_params[3])
    rho = property(lambda self: self.path_params[4])

    def price_days(self, t_begin: int, t_end: int):
        """Nested-MC ATM call / put prices of days ``[t_begin, t_end)`` into the book (resumable: any range, any order)."""
        b = self.book
        with torch.cuda.device(b.device):
            _lib.check(_lib.lib().cantor_rbergomi_price_atm(C.byref(self._params), b.tensor.data_ptr(), b.ld, b.n_paths,
                                                            b.episode_length, self.path_params.data_ptr(), int(t_begin), int(t_end),
                                                            _stream(b.device)), "cantor_rbergomi_price_atm")
        return self


def generate_rbergomi_paths_and_options(num_paths, r=R, dt=DT, seed=SEED, *, base_params=None, historical_prices=None,
                                        n_steps=N_STEPS,
                                        n_mc=N_PATHS_OPTION_MC, tenor=T_OPTION_TENOR, price=True, days_per_launch=32,
                                        shared_draws=False, path_offset=0, device="cuda", exported=None,
                                        tensor_cores=True) -> RbergomiBook:
    """The reference's data generator (``generate_paths_and_options``, rbergomi_sim.py:309-499) on the GPU.

    ``historical_prices`` (the reference's first argument) is calibrated on the host with
    ``calibration.estimate_base_params``; or pass ``base_params`` = ``(S0, xi, H, eta, rho)`` directly (or a dict also
    overriding the perturbation constants); every path gets its own perturbed parameters.  Returns the packed env-schema book
    (``.book.save_npz(path)`` writes the reference's ``paths_rbergomi_options_100k.npz`` schema) with the ATM call / put
    columns priced by ``n_mc`` inner rough-Bergomi paths per (path, day), calls and puts on independent draws like the
    reference unless ``shared_draws``.  ``tensor_cores`` runs the 30-tap variance filter as split-TF32 ``tcgen05.mma``
    (float32-level accuracy) instead of float32 FFMAs.  ``exported=dict(params=[5, n], dW1=[n, M], dW2=[n, M])`` replays exported draws.
    """
    dev = torch.device(device)
    if historical_prices is not None:                 # the reference's first argument: calibrate on the host (rbergomi_sim.py:360)
        if base_params is not None:
            raise ValueError("give historical_prices or base_params, not both")
        from .calibration import estimate_base_params
        base_params = estimate_base_params(historical_prices, dt)
    p = _rb_params(base_params, r, dt, tenor, n_mc, shared_draws, seed, path_offset, tensor_cores)
    book = ReplayData.empty(num_paths, n_steps, dev)
    book.tensor.zero_()
    pp = torch.empty((5, num_paths), dtype=torch.float64, device=dev)
    ex = exported or {}
    prm = _as_f64(ex["params"], dev) if "params" in ex else None
    d1 = _as_f64(ex["dW1"], dev) if "dW1" in ex else None
    d2 = _as_f64(ex["dW2"], dev) if "dW2" in ex else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_rbergomi_paths(C.byref(p), num_paths, n_steps, _lib.ptr(prm), _lib.ptr(d1), _lib.ptr(d2),
                                                    int(d1.shape[1]) if d1 is not None else 0, book.tensor.data_ptr(), book.ld,
                                                    pp.data_ptr(), None, None, _stream(dev)), "cantor_rbergomi_paths")
    out = RbergomiBook(book, pp, p)
    if price:
        for t0 in range(0, n_steps, max(1, int(days_per_launch))):
            out.price_days(t0, min(n_steps, t0 + int(days_per_launch)))
    return out


def rbergomi_outer_paths(num_paths, n_steps, params, dW1, dW2, r=R, dt=DT, device="cuda"):
    """Float64 paths and variances ``(n, n_steps + 1)`` of the outer generator on exported parameters / increments."""
    dev = torch.device(device)
    p = _rb_params(None, r, dt, T_OPTION_TENOR, 1, False, 0, 0)
    prm, d1, d2 = _as_f64(params, dev), _as_f64(dW1, dev), _as_f64(dW2, dev)
    paths = torch.empty((num_paths, n_steps + 1), dtype=torch.float64, device=dev)
    v = torch.empty_like(paths)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_rbergomi_paths(C.byref(p), num_paths, n_steps, prm.data_ptr(), d1.data_ptr(), d2.data_ptr(),
                                                    int(d1.shape[1]), None, 0, None, paths.data_ptr(), v.data_ptr(), _stream(dev)),
                   "cantor_rbergomi_paths")
    return paths, v


def price_rbergomi_option(S0, K, T_opt, r, xi, H, eta, rho, option_type, dW1, dW2, dt=DT, device="cuda", tensor_cores=True):
    """``price_rbergomi_option_gpu`` (rbergomi_sim.py:246-306) on exported increments ``dW1, dW2 [batch, n_mc, 32]``."""
    if option_type not in ("call", "put"):
        raise ValueError("option_type must be 'call' or 'put'")
    dev = torch.device(device)
    p = _rb_params(None, r, dt, T_opt, 1, False, 0, 0, tensor_cores)
    arrs = [_as_f64(a, dev) for a in (S0, K, xi)]
    hep = torch.stack([_as_f64(a, dev) for a in (H, eta, rho)]).contiguous()        # consecutive rows, as the C ABI asks
    arrs += [hep[0], hep[1], hep[2]]
    d1, d2 = _as_f64(dW1, dev), _as_f64(dW2, dev)
    B, n_mc, M = d1.shape
    out = torch.empty(B, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cantor_rbergomi_price_from_increments(C.byref(p), *(a.data_ptr() for a in arrs), d1.data_ptr(),
                                                                    d2.data_ptr(), B, n_mc, M, int(option_type == "put"),
                                                                    out.data_ptr(), _stream(dev)),
                   "cantor_rbergomi_price_from_increments")
    return out
