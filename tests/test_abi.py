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
    assert handle.cantor_abi_version() == 2


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct, as gcc sees include/cantor_hedge.h (plain C), equal the ctypes mirrors."""
    import subprocess
    from cantorrl_b200 import _lib
    pairs = {"cantor_env_params": _lib.EnvParams, "cantor_replay_book": _lib.ReplayBook, "cantor_env_state": _lib.EnvState,
             "cantor_reset_rule": _lib.ResetRule, "cantor_info_out": _lib.InfoOut, "cantor_env_sim": _lib.EnvSim, "cantor_sim_params": _lib.SimParams,
             "cantor_policy": _lib.Policy, "cantor_stats_out": _lib.StatsOut, "cantor_vecnorm_fuse": _lib.VecNormFuse, "cantor_rollout_out": _lib.RolloutOut, "cantor_rbergomi_params": _lib.RbergomiParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cantor_hedge.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['printf("VECNORM %d STATS %d MLP %d OBS %d\\n", CANTOR_VECNORM_DOUBLES, CANTOR_STATS_LEN, CANTOR_MLP_FLOATS, CANTOR_OBS_DIM);',
              'return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()[:-1])
    for cname, ct in pairs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"
    consts = subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines()[-1].split()
    assert [int(consts[1]), int(consts[3]), int(consts[5]), int(consts[7])] == [_lib.VECNORM_DOUBLES, _lib.STATS_LEN, _lib.MLP_FLOATS, _lib.OBS_DIM]


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


def test_missing_library_fails_loudly_instead_of_falling_back():
    """Without the built .so every product entry point raises (there is no CPU path to fall back to)."""
    import subprocess
    import sys
    code = ("import numpy as np\n"
            "from cantorrl_b200 import CantorError, HedgingVecEnv\n"
            "from cantorrl_b200.host_env import HostVecEnv\n"
            "from cantorrl_b200.rollout import HedgingRollout\n"
            "for make in (lambda: HedgingVecEnv(data={}, num_envs=2), lambda: HostVecEnv(num_envs=2, data={}),\n"
            "             lambda: HedgingRollout(num_envs=2, simulate={})):\n"
            "    try:\n"
            "        make()\n"
            "    except CantorError as e:\n"
            "        assert 'no CPU fallback' in str(e), e\n"
            "    else:\n"
            "        raise SystemExit('constructed without the library')\n"
            "print('LOUD')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                         env=dict(os.environ, CANTOR_HEDGE_LIB="/nonexistent/libcantor_hedge.so"))
    assert out.returncode == 0 and "LOUD" in out.stdout, out.stdout + out.stderr
