"""The C-ABI shared library: loads on a machine without a GPU, exports every symbol that
``include/acas2d_b200.h`` declares, and validates its arguments before touching the device."""
import ctypes
import os
import re

import pytest

from gym_ACAS2D.envs import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _native.build()
    return _native.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "acas2d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acas2d_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert set(names) == set(_native.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_abi_version_and_constants(lib):
    assert lib.acas2d_abi_version() == _native.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "acas2d_b200.h")).read()
    for name, val in [("ACAS2D_STAT_SLOTS", _native.STAT_SLOTS), ("ACAS2D_STAT_FIELDS", _native.STAT_FIELDS),
                      ("ACAS2D_MAX_TRAFFIC", _native.MAX_TRAFFIC), ("ACAS2D_ABI_VERSION", _native.ABI_VERSION)]:
        assert re.search(rf"#define {name}\s+{val}\b", hdr), name


def test_params_default_equals_settings(lib):
    """acas2d_params_default (C) == params_from_settings(settings.py) field by field."""
    for n in (1, 8):
        c = _native.Params()
        assert lib.acas2d_params_default(ctypes.byref(c), n) == 0
        py = _native.params_from_settings(None, n, auto_reset=False)
        for name, _ in _native.Params._fields_:
            assert getattr(c, name) == getattr(py, name), name
    assert py.d_goal_max == 3408.0 and py.d_dev_max == 2000.0
    assert py.d_separation_max == 5886.796226411321 and py.d_cpa_max == 1886.7962264113207
    assert py.acc_lat_limit == 20 * 9.80665


def test_argument_errors_do_not_touch_the_device(lib):
    p = _native.params_from_settings(None, 1)
    s = _native.State()
    assert lib.acas2d_params_default(None, 1) == -1
    assert lib.acas2d_params_default(ctypes.byref(_native.Params()), 0) == -2            # N_TRAFFIC = 0 (Q20)
    assert lib.acas2d_params_default(ctypes.byref(_native.Params()), _native.MAX_TRAFFIC + 1) == -2
    assert lib.acas2d_step(None, None, None, None, None, None, None, None) == -1
    s.num_envs = 4
    assert lib.acas2d_reset(ctypes.byref(p), ctypes.byref(s), None, None, None) == -1    # state pointers NULL
    s.num_envs = 0
    assert lib.acas2d_reset(ctypes.byref(p), ctypes.byref(s), None, None, None) == 0     # empty batch: nothing to do
    s.num_envs = -5
    assert lib.acas2d_reset(ctypes.byref(p), ctypes.byref(s), None, None, None) == -3
    p.n_traffic = 0
    assert lib.acas2d_extract_state(ctypes.byref(p), ctypes.byref(s), None, None, None, None, None) == -2
    # PPO learner entry points: same conventions
    c = _native.PpoConfig.sb3_defaults()
    assert (c.gamma, c.clip_range, c.normalize_advantage) == (ctypes.c_float(0.99).value, ctypes.c_float(0.2).value, 1)
    assert lib.acas2d_ppo_values(None, None, 16, None, None) == -1
    assert lib.acas2d_ppo_values(None, None, 0, None, None) == 0 and lib.acas2d_ppo_values(None, None, -1, None, None) == -3
    assert lib.acas2d_ppo_gae(ctypes.byref(c), None, None, None, 8, 8, None, None, None) == -1
    assert lib.acas2d_ppo_gae(ctypes.byref(c), None, None, None, -1, 8, None, None, None) == -3
    assert lib.acas2d_ppo_grad(ctypes.byref(c), None, None, None, None, None, None, None, 64, None, None, None, None, None) == -1
    assert lib.acas2d_ppo_grad(ctypes.byref(c), None, None, None, None, None, None, None, 0, None, None, None, None, None) == -3
    assert lib.acas2d_ppo_adam(ctypes.byref(c), None, None, 1.0, None, None, None, None, None) == -1
    step_args = [ctypes.byref(c)] + [None] * 7 + [64] + [None] * 6
    assert lib.acas2d_ppo_step(*step_args, 0, 1, None, None) == -1
    assert lib.acas2d_ppo_step(*step_args, 3, 2, None, None) == -3                      # rank outside the world
    assert lib.acas2d_ppo_step(*step_args, 0, _native.PPO_MAX_RANKS + 1, None, None) == -3
    assert lib.acas2d_launch_count() == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_ACAS2D.envs import ACAS2DEnv, ACAS2DVecEnv, BatchedACAS2D
    for ctor in (lambda: BatchedACAS2D(4), lambda: ACAS2DEnv(), lambda: ACAS2DVecEnv(4)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ctor()


def test_product_never_imports_the_oracle():
    """The oracle and the host-check build are test infrastructure: nothing under the product
    package may import, include, link or load them."""
    pkg = os.path.join(ROOT, "gym-acas2d_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|from\s+tests|import\s+tests|#include\s+[\"<][^\n]*oracle|"
                     r"libacas2d_oracle|libacas2d_hostcheck|oracle/|ref_shim)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)
