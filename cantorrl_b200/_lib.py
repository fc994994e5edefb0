"""ctypes binding of ``libcantor_hedge.so`` (the C ABI declared in ``include/cantor_hedge.h``).

The library is built in-tree by ``cantorrl_b200/csrc/Makefile`` (see ``__graft_entry__.build``).
There is NO fallback: if the shared object is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# CANTOR_HEDGE_LIB selects another build of the same library (kernel-variant sweeps, profiling builds)
LIB_PATH = os.environ.get("CANTOR_HEDGE_LIB") or os.path.join(CSRC, "libcantor_hedge.so")

OBS_DIM = 13
F32, F64 = 32, 64
LOSS_ABS, LOSS_MSE = 0, 1
RESET_SAME_PATH, RESET_FROM_ARRAY, RESET_PHILOX = 0, 1, 2
STEP_KEEP_OBS_IN_L2 = 1
STEP_WALK_BACKWARD = 2
INFO_F64_KEYS = (
    "step_pnl_total", "per_share_step_pnl", "raw_pnl_deviation_abs", "transaction_costs_total",
    "commission_cost", "slippage_cost", "reward_pnl_component", "transaction_cost_penalty", "theta_penalty",
    "reward_step", "portfolio_value", "cash", "raw_action_call", "raw_action_put", "scaled_float_call",
    "scaled_float_put", "initial_S0_for_episode",
)
INFO_I32_KEYS = (
    "call_contracts", "put_contracts", "requested_calls_rounded_clipped", "requested_puts_rounded_clipped",
    "actual_calls_traded", "actual_puts_traded",
)


class CantorError(RuntimeError):
    """A libcantor_hedge call returned a non-zero status."""


class EnvParams(C.Structure):
    _fields_ = [
        ("transaction_cost_per_contract", C.c_double), ("lambda_cost", C.c_double),
        ("pnl_penalty_weight", C.c_double), ("theta_weight", C.c_double), ("slippage_bps", C.c_double),
        ("initial_cash", C.c_double), ("risk_free_rate", C.c_double), ("option_tenor_years", C.c_double),
        ("loss_type", C.c_int32), ("shares_to_hedge", C.c_int32), ("max_contracts_held", C.c_int32),
        ("max_trade_per_step", C.c_int32), ("option_contract_multiplier", C.c_int32), ("record_metrics", C.c_int32),
    ]


class ReplayBook(C.Structure):
    _fields_ = [("svcp", C.c_void_p), ("ld", C.c_int64), ("n_paths", C.c_int32), ("episode_length", C.c_int32)]


class VecNormFuse(C.Structure):
    _fields_ = [("partial", C.c_void_p), ("returns", C.c_void_p), ("gamma", C.c_double), ("n_partial_ctas", C.c_int64),
                ("norm_obs", C.c_int32), ("norm_reward", C.c_int32)]


class EnvState(C.Structure):
    _fields_ = [("core", C.c_void_p), ("cash", C.c_void_p), ("pv_prev", C.c_void_p), ("episode_acc", C.c_void_p),
                ("episode_return", C.c_void_p), ("episode_length", C.c_void_p), ("stats", C.c_void_p), ("vecnorm", C.c_void_p)]


class ResetRule(C.Structure):
    _fields_ = [("mode", C.c_int32), ("flags", C.c_int32), ("next_path", C.c_void_p),
                ("seed", C.c_uint64), ("env_offset", C.c_int64), ("episode_counter", C.c_int64)]


MODEL_GBM, MODEL_HESTON = 0, 1
SIGMA_REALISED, SIGMA_BOOK_VARIANCE = 0, 1


class SimParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("reprice", C.c_int32), ("s0", C.c_double), ("v0", C.c_double),
                ("r", C.c_double), ("dt", C.c_double), ("kappa", C.c_double), ("theta", C.c_double),
                ("sigma_v", C.c_double), ("rho", C.c_double), ("tenor", C.c_double), ("seed", C.c_uint64),
                ("path_offset", C.c_int64)]


class RbergomiParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("s0", "xi", "H", "eta", "rho", "perturb_s0", "perturb_xi", "perturb_H", "perturb_eta",
                                          "perturb_rho", "min_xi_factor", "min_eta_factor", "clip_H_min", "clip_H_max",
                                          "clip_rho_min", "clip_rho_max", "r", "dt", "tenor")] + \
               [("n_mc", C.c_int32), ("shared_draws", C.c_int32), ("tensor_cores", C.c_int32), ("reserved", C.c_int32),
                ("seed", C.c_uint64), ("path_offset", C.c_int64)]


POLICY_NO_HEDGE, POLICY_RANDOM, POLICY_DELTA_BASELINES, POLICY_DELTA_BENCHMARK, POLICY_MLP, POLICY_ACTIONS, POLICY_LSTM = range(7)
LSTM_IMAGE_BYTES = 178816
SQUASH_CLIP, SQUASH_TANH = 0, 1
MLP_FLOATS = 5212
STATS_LEN = 16
VECNORM_DOUBLES = 16704


class Policy(C.Structure):
    _fields_ = [("kind", C.c_int32), ("put_leg_disabled", C.c_int32), ("mlp", C.c_void_p), ("actions", C.c_void_p),
                ("seed", C.c_uint64), ("mlp_tensor_cores", C.c_int32), ("action_squash", C.c_int32), ("obs_clip", C.c_float),
                ("reserved", C.c_int32)]


class StatsOut(C.Structure):
    _fields_ = [("sums", C.c_void_p), ("hist", C.c_void_p), ("hist_sum", C.c_void_p), ("episode_b", C.c_void_p),
                ("hist_max", C.c_double), ("hist_bins", C.c_int32), ("reserved", C.c_int32), ("episode_slots", C.c_int64),
                ("mc_global", C.c_void_p), ("peer_global", C.c_void_p * 8), ("n_peers", C.c_int32), ("reserved2", C.c_int32),
                ("ticket", C.c_void_p)]


class RolloutOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("actions", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p)]


class InfoOut(C.Structure):
    _fields_ = [("f64", C.c_void_p), ("i32", C.c_void_p), ("f32", C.c_void_p)]


class EnvSim(C.Structure):
    _fields_ = [("sim", C.POINTER(SimParams)), ("sv", C.c_void_p), ("total_envs", C.c_int64), ("episode_length", C.c_int32),
                ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/cantor_hedge.h declares must appear here
SIGNATURES = {
    "cantor_abi_version": (C.c_int, []),
    "cantor_last_error": (C.c_char_p, []),
    "cantor_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_size_t)]),
    "cantor_pack_book": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "cantor_unpack_book": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_sim_paths": (C.c_int, [C.POINTER(SimParams), C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "cantor_reprice_atm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p]),
    "cantor_euler_from_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "cantor_rbergomi_paths": (C.c_int, [C.POINTER(RbergomiParams), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_rbergomi_price_atm": (C.c_int, [C.POINTER(RbergomiParams), C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_int32, C.c_int32, C.c_void_p]),
    "cantor_rbergomi_price_from_increments": (C.c_int, [C.POINTER(RbergomiParams)] + [C.c_void_p] * 8 +
                                              [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cantor_bs_price": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_schema_b_book": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_double, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_reprice_book": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_reprice_book_strided": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32,
                                              C.c_int32, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "cantor_bs_delta_hedge": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_double, C.c_double,
                                        C.c_void_p, C.c_void_p]),
    "cantor_vecnorm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cantor_vecnorm_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_void_p]),
    "cantor_vecnorm_step_fused": (C.c_int, [C.c_void_p, C.POINTER(VecNormFuse), C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "cantor_rollout": (C.c_int, [C.POINTER(EnvParams), C.POINTER(ReplayBook), C.POINTER(SimParams), C.c_int32,
                                 C.POINTER(Policy), C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.POINTER(StatsOut),
                                 C.POINTER(RolloutOut), C.c_void_p]),
    "cantor_vecenv_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(EnvParams), C.c_int32, C.c_int64, C.c_int32, C.c_int32]),
    "cantor_vecenv_destroy": (C.c_int, [C.c_void_p]),
    "cantor_vecenv_load_book_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_int32, C.c_int32]),
    "cantor_vecenv_simulate_book": (C.c_int, [C.c_void_p, C.POINTER(SimParams), C.c_int32, C.c_int32]),
    "cantor_vecenv_set_reset_rule": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_int64]),
    "cantor_vecenv_reset_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_vecenv_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_vecenv_episode_length": (C.c_int, [C.c_void_p]),
    "cantor_vecenv_num_paths": (C.c_int, [C.c_void_p]),
    "cantor_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "cantor_host_unregister": (C.c_int, [C.c_void_p]),
    "cantor_host_copy_probe": (C.c_int, [C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_double,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cantor_umma_probe": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cantor_env_reset": (C.c_int, [C.POINTER(EnvParams), C.POINTER(ReplayBook), C.POINTER(EnvState), C.c_int64,
                                   C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_env_step": (C.c_int, [C.POINTER(EnvParams), C.POINTER(ReplayBook), C.POINTER(EnvState), C.c_int64,
                                  C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.POINTER(ResetRule), C.POINTER(InfoOut), C.c_void_p]),
    "cantor_env_reset_sim": (C.c_int, [C.POINTER(EnvParams), C.POINTER(EnvSim), C.POINTER(EnvState), C.c_int64, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cantor_env_step_sim": (C.c_int, [C.POINTER(EnvParams), C.POINTER(EnvSim), C.POINTER(EnvState), C.c_int64, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.POINTER(InfoOut), C.c_int32, C.c_void_p]),
    "cantor_env_step_many": (C.c_int, [C.POINTER(EnvParams), C.POINTER(ReplayBook), C.POINTER(EnvState), C.c_int64,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.POINTER(ResetRule), C.c_void_p]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", CSRC, f"-j{min(8, os.cpu_count() or 1)}"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise CantorError(f"building libcantor_hedge.so failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stdout)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded shared library.  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CantorError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). cantorrl_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)            # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().cantor_last_error().decode(errors="replace")
        raise CantorError(f"{what or 'libcantor_hedge'} failed with status {status}: {msg}")


def ptr(t):
    """Device (or host) address of a torch tensor, or None."""
    return None if t is None else t.data_ptr()


def current_stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
