"""Minimal stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

The reference environment (`src/env/hedging_env*.py` in bcosm/CantorRL) imports
gymnasium, which is not installed in this image.  This stub provides exactly
the surface those two files touch, so that the UNMODIFIED reference class can
be imported as the parity oracle by `tests/golden/make_golden.py` and by the
CPU tests that pin `oracle/hedge_oracle.py`:

  * ``gymnasium.Env`` with ``reset(seed=...)`` that (re)creates ``np_random``
    and an ``np_random`` property with a setter
    (hedging_env_v2.py:146-148 assigns it),
  * ``gymnasium.spaces.Box``,
  * ``gymnasium.utils.seeding.np_random`` = Generator(PCG64(SeedSequence(seed)))
    which is what gymnasium 1.1.1 builds, so the episode-index stream equals
    ``np.random.default_rng(seed).integers(n)``.

Nothing in the product package imports this.
"""
from . import spaces, utils  # noqa: F401
from .utils import seeding


class Env:
    metadata = {}
    _np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, _ = seeding.np_random(seed)

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def close(self):
        pass
