"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imports the UNMODIFIED reference (``/root/reference/gym_ACAS2D``, or its hot-path
files copied byte for byte into the git-ignored ``oracle/_ref`` by
``oracle/make_ref.py``) under in-process stand-ins for ``gym`` and ``pygame``
(neither is installed).  Every arithmetic line of the
reference then runs verbatim.  Used by ``tests/golden/make_golden.py`` (fixture
generation), by the CPU tests that cross-check the C oracle, and by
``oracle/ref_runner.py`` (the CPU arm of ``bench.py``: ``kind: "reference"``).
"""
from __future__ import annotations

import importlib
import io
import contextlib
import os
import sys
import types

_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/make_ref.py (git-ignored)


def _find_root() -> str:
    """The reference checkout when it is there (build container), else the unmodified copy of its
    hot-path files that ``oracle/make_ref.py`` placed in ``oracle/_ref`` (travels to the GPU box)."""
    for root in (os.environ.get("ACAS2D_REFERENCE_ROOT"), "/root/reference", _LOCAL_COPY):
        if root and os.path.isdir(os.path.join(root, "gym_ACAS2D", "envs")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_ACAS2D", "envs"))


def full_tree() -> bool:
    """True when the whole reference checkout (golden CSV, model zips, notebooks) is present, not just
    the hot-path source files of ``oracle/_ref``."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_ACAS2D", "models"))


def _stub_modules():
    gym = types.ModuleType("gym")

    class Env:  # gym.Env stand-in
        pass

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            import numpy as np
            self.dtype = dtype
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

    gym.Env = Env
    spaces = types.ModuleType("gym.spaces"); spaces.Box = Box
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registry = {}
    registration.register = lambda id, entry_point, **kw: registry.__setitem__(id, entry_point)
    registration.registry = registry
    gym.spaces, gym.envs, envs.registration = spaces, envs, registration

    pg = types.ModuleType("pygame")

    class _Anything:
        def __call__(self, *a, **k):
            return self

        def __getattr__(self, name):
            return self

    pg.init = lambda: None
    pg.QUIT = 256
    for sub in ("display", "image", "font", "event", "draw", "time", "transform"):
        setattr(pg, sub, _Anything())
    pg.event.get = lambda: []

    class _Clock:
        def tick(self, fps=0):
            return 0
    pg.time.Clock = _Clock
    return {"gym": gym, "gym.spaces": spaces, "gym.envs": envs, "gym.envs.registration": registration, "pygame": pg}


def load(n_traffic: int = 1, **settings_overrides):
    """Import the reference with MIN/MAX_TRAFFIC = n_traffic (and any other settings.py constant
    overridden the same way, e.g. AIRSPEED_FACTOR_MIN / MAX); returns the ``gym_ACAS2D`` module.

    The reference star-imports its settings at import time, so a different traffic
    count needs a fresh import (Config)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in [m for m in sys.modules if m == "gym_ACAS2D" or m.startswith("gym_ACAS2D.")]:
        del sys.modules[name]
    for name, mod in _stub_modules().items():
        sys.modules[name] = mod
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        settings = importlib.import_module("gym_ACAS2D.settings")
        settings.MIN_TRAFFIC = settings.MAX_TRAFFIC = int(n_traffic)
        for key, value in settings_overrides.items():
            if not hasattr(settings, key):
                raise KeyError(key)
            setattr(settings, key, value)
        pkg = importlib.import_module("gym_ACAS2D")
        importlib.import_module("gym_ACAS2D.envs")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return pkg


def unload():
    """Drop the reference (and the stand-ins) from ``sys.modules``."""
    for name in [m for m in sys.modules
                 if m in ("gym", "pygame") or m.startswith(("gym.", "gym_ACAS2D"))]:
        del sys.modules[name]


@contextlib.contextmanager
def quiet():
    """The reference prints an "Outcome:" line on every episode end (game.py:311-313)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
