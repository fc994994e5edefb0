"""Shared pytest plumbing: the ``gpu`` marker, repo-root imports, golden-case loader."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"          # present in the build container only, never on the GPU box

ENV_CASES = ("v2_train", "v1_default", "v2_mse_saturating", "v2_cvar_nometrics", "v2_lowprice", "v2_degenerate")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_env_case(name):
    """Golden env case -> (npz dict, kwargs for the env constructor, is_v1)."""
    z = dict(np.load(os.path.join(GOLDEN, f"env_{name}.npz")))
    kwargs = {}
    for k in list(z):
        if k.startswith("kw_"):
            v = z.pop(k)
            kwargs[k[3:]] = v.item() if v.dtype.kind != "U" else str(v)
    is_v1 = int(z["version"]) == 1
    if is_v1:
        kwargs.setdefault("transaction_cost_per_contract", 0.05)     # hedging_env.py:11
    return z, kwargs, is_v1


@pytest.fixture(params=ENV_CASES)
def env_case(request):
    return (request.param,) + load_env_case(request.param)


@pytest.fixture(autouse=True)
def _release_device_memory_between_tests(request):
    """The full-size tests (34 - 100 GB books) must not find the caching allocator still holding the previous test's blocks."""
    yield
    if "gpu" in request.keywords:
        import gc

        import torch
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
