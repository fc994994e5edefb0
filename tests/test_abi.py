"""CPU checks of the drop-in boundary: the library loads and exports every symbol include/cantor_hedge.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "cantor_hedge.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cantor_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("cantor_env_reset", "cantor_env_step", "cantor_env_step_many", "cantor_last_error", "cantor_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cantorrl_b200 import _lib
    handle = _lib.lib()                       # raises if the .so has not been built: no fallback
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in cantor_hedge.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in cantorrl_b200/_lib.py"
    assert handle.cantor_abi_version() == 1


def test_struct_layouts_match_the_header():
    from cantorrl_b200 import _lib
    assert C.sizeof(_lib.EnvParams) == 8 * 8 + 6 * 4
    assert C.sizeof(_lib.ReplayBook) == 8 + 8 + 2 * 4
    assert C.sizeof(_lib.EnvState) == 3 * 8
    assert C.sizeof(_lib.ResetRule) == 2 * 4 + 8 + 8 + 8 + 8
    assert C.sizeof(_lib.InfoOut) == 16


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected on the host before any CUDA call, with a message."""
    from cantorrl_b200 import _lib
    L = _lib.lib()
    p, b, s = _lib.EnvParams(), _lib.ReplayBook(), _lib.EnvState()
    rc = L.cantor_env_step(C.byref(p), C.byref(b), C.byref(s), 4, 32, None, None, None, None, None, 1, None, None, None)
    assert rc == 1 and b"book" in L.cantor_last_error()
    rc = L.cantor_device_info(0, None, None, None, None)
    assert rc in (0, 3)


def test_env_without_cuda_device_raises():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from cantorrl_b200 import CantorError, HedgingVecEnv
    with pytest.raises((CantorError, RuntimeError, AssertionError)):
        HedgingVecEnv(data={}, num_envs=4, device="cpu")
