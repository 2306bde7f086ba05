"""Import shim for the *live* reference (test harness only, never product code).

The reference (read-only at /root/reference, absent on the GPU box) imports pygame,
gymnasium, rvo2, socialforce and matplotlib at module import time
(social_gym/__init__.py:1, src/agent.py:1, src/obstacle.py:1,7, src/motion_model_manager.py:8,
social_nav_sim.py:1,14-15,28).  None of them is on the SFM/HSFM arithmetic path, so this module
installs inert stand-ins in ``sys.modules`` *before* the first ``import social_gym`` and restores
the ``np.NaN`` alias that NumPy 2 removed (motion_model_manager.py:264,271).

Only ``tests/golden/make_golden.py`` and the optional live cross-check tests use it.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("SNP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "social_gym"))


class _Sprite:
    def __init__(self, *a, **k):
        pass


class _Group:
    def __init__(self, *a):
        self._items = list(a)

    def add(self, *items):
        self._items.extend(items)

    def empty(self):
        self._items.clear()

    def sprites(self):
        return list(self._items)

    def __len__(self):
        return len(self._items)

    def __iter__(self):
        return iter(self._items)


def install():
    """Install stubs and put the reference on sys.path.  Idempotent."""
    if "social_gym" in sys.modules:
        return
    import numpy as np

    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/snp_numba_cache")

    gymn = types.ModuleType("gymnasium")

    class Env:
        pass

    gymn.Env = Env
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Discrete = lambda n: n
    gymn.spaces = spaces
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda **kw: None
    envs.registration = reg
    gymn.envs = envs
    sys.modules["gymnasium"] = gymn
    sys.modules["gymnasium.spaces"] = spaces
    sys.modules["gymnasium.envs"] = envs
    sys.modules["gymnasium.envs.registration"] = reg

    pg = MagicMock()
    pg.sprite.Sprite = _Sprite
    pg.sprite.Group = _Group
    pg.time.get_ticks = lambda: 0
    sys.modules["pygame"] = pg
    sys.modules["pygame.sprite"] = pg.sprite

    sys.modules["rvo2"] = MagicMock()
    sys.modules["socialforce"] = MagicMock()
    mpl = MagicMock()
    mpl.colors.TABLEAU_COLORS = {"a": "#000000"}
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    sys.modules["matplotlib.colors"] = mpl.colors

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
