"""NumPy-backed stand-in for the `cupy` package (TEST INFRASTRUCTURE ONLY).

Lets the UNMODIFIED reference simulator ``src/sim/rbergomi_sim.py`` run on the CPU in the build container so
that ``tests/golden/make_golden.py`` can export its draws and its outer log-Euler step
(``rbergomi_sim.py:454-464``) as golden vectors.  The random stream is NumPy's, not cuRAND's: this is a
structural oracle (same arithmetic on the same exported normals), not a reproduction of the reference's data.
"""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import fft  # noqa: F401

float64 = _np.float64


def asnumpy(a):
    return _np.asarray(a)


class _Random:
    def __init__(self):
        self._rng = _np.random.default_rng(0)

    def seed(self, s):
        self._rng = _np.random.default_rng(s)

    def normal(self, loc=0.0, scale=1.0, size=None, dtype=_np.float64):
        return self._rng.normal(loc, scale, size).astype(dtype)


random = _Random()


class _Stream:
    def synchronize(self):
        pass


class _StreamNS:
    null = _Stream()


class cuda:  # noqa: N801
    Stream = _StreamNS
