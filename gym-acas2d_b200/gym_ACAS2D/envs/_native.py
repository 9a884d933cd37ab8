"""ctypes binding of ``libacas2d_b200.so`` (C ABI: ``include/acas2d_b200.h``).

The library is built in-tree by ``build()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``)
and loaded from ``gym-acas2d_b200/csrc``.  There is no fallback: if the library is missing
or the machine has no CUDA device, the environment constructors raise.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # gym-acas2d_b200/
CSRC_DIR = os.path.join(_PKG_ROOT, "csrc")
REPO_ROOT = os.path.dirname(_PKG_ROOT)
LIB_PATH = os.environ.get("ACAS2D_LIB") or os.path.join(CSRC_DIR, "libacas2d_b200.so")   # ACAS2D_LIB: experiment builds
SOURCES = ("acas2d_kernels.cu", "acas2d_env.cuh", "acas2d_math.cuh", "acas2d_policy.cuh", "acas2d_policy_tc.cuh", "acas2d_dev.cuh", "acas2d_ppo.cuh",
           "acas2d_tiled.cuh")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

ABI_VERSION = 2
MAX_TRAFFIC = 1024
STAT_SLOTS, STAT_FIELDS = 128, 16
STAT_NAMES = ("episodes", "goal", "collision", "timeout", "length_sum", "return_fx", "min_sep_fx")
STAT_FX_SCALE = 1048576.0
FLAG_COLLISION, FLAG_GOAL, FLAG_TIMEOUT, FLAG_DONE, FLAG_OOB = 1, 2, 4, 8, 16
STEPS_RESIDUAL_BIT = 0x40000000
STEPS_COMPACT_BIT, STEPS_DOWN_BIT, STEPS_MASK = 0x20000000, 0x10000000, 0x0FFFFFFF
POLICY_FLOATS = 4804
PPO_PARAM_FLOATS = 2 * POLICY_FLOATS + 4
PPO_LOG_STD = 2 * POLICY_FLOATS
PPO_PARTIAL_FLOATS, PPO_MAX_CTAS = 4816, 148
PPO_WORKSPACE_FLOATS = 64 + 2 * PPO_MAX_CTAS * PPO_PARTIAL_FLOATS
PPO_MAX_RANKS = 16
PSTAGE_BYTES = 112
PPO_EXCHANGE_FLOATS = 2 * PPO_PARAM_FLOATS + PPO_MAX_RANKS
PPO_LOSS_STATS = 8

EXPORTS = ("acas2d_abi_version", "acas2d_params_default", "acas2d_reset", "acas2d_step", "acas2d_step_host",
           "acas2d_inject_state", "acas2d_extract_state", "acas2d_rollout_random", "acas2d_random_actions",
           "acas2d_launch_count", "acas2d_set_tuning", "acas2d_set_n1_kernel", "acas2d_policy_step", "acas2d_observe", "acas2d_render",
           "acas2d_ppo_values", "acas2d_ppo_gae", "acas2d_ppo_grad", "acas2d_ppo_adam", "acas2d_ppo_step",
           "acas2d_ppo_prepare", "acas2d_policy_step_dyn", "acas2d_step_k", "acas2d_set_tiled_tuning", "acas2d_step_host_packed", "acas2d_trace_step", "acas2d_step_mapped")

ERRORS = {-1: "required pointer is NULL", -2: "unsupported n_traffic", -3: "bad size", -4: "no CUDA device"}


class Params(ctypes.Structure):
    """``acas2d_params`` (include/acas2d_b200.h)."""
    _fields_ = [(n, ctypes.c_double) for n in (
        "width", "height", "fps", "max_steps", "aircraft_size", "collision_radius", "goal_radius",
        "safe_distance", "airspeed", "airspeed_factor_min", "airspeed_factor_max", "acc_lat_limit",
        "player_heading_lim", "traffic_heading_lim", "reward_goal", "reward_collision",
        "goal_x", "goal_y", "player_x0", "player_y0", "player_psi_base",
        "d_goal_max", "d_dev_max", "d_separation_max", "d_cpa_max", "v_closing_max")] + [
        ("n_traffic", ctypes.c_int32), ("auto_reset", ctypes.c_int32)]


class State(ctypes.Structure):
    """``acas2d_state``: device pointers owned by the caller."""
    _fields_ = [("num_envs", ctypes.c_int64),
                ("ppos", ctypes.c_void_p), ("paux", ctypes.c_void_p),
                ("thot", ctypes.c_void_p), ("tres", ctypes.c_void_p),
                ("episode_idx", ctypes.c_void_p), ("min_sep", ctypes.c_void_p),
                ("stats", ctypes.c_void_p),
                ("seed", ctypes.c_uint64), ("env_id_offset", ctypes.c_uint64),
                ("tkin", ctypes.c_void_p), ("tpsi0", ctypes.c_void_p), ("pstage", ctypes.c_void_p),
                ("spawn_sep", ctypes.c_void_p)]


class PpoConfig(ctypes.Structure):
    """``acas2d_ppo_config``: SB3 1.1.0 PPO defaults (the reference's ``best_model.zip/data``)."""
    _fields_ = [(n, ctypes.c_float) for n in (
        "gamma", "gae_lambda", "clip_range", "vf_coef", "ent_coef", "max_grad_norm",
        "lr", "beta1", "beta2", "adam_eps")] + [("normalize_advantage", ctypes.c_int32), ("reserved", ctypes.c_int32)]

    @classmethod
    def sb3_defaults(cls, **over) -> "PpoConfig":
        d = dict(gamma=0.99, gae_lambda=0.95, clip_range=0.2, vf_coef=0.5, ent_coef=0.0, max_grad_norm=0.5,
                 lr=3e-4, beta1=0.9, beta2=0.999, adam_eps=1e-5, normalize_advantage=1)
        d.update(over)
        return cls(**d)


TRACE_DOUBLES, TRACE_MAX_TRAFFIC = 18, 16
TRACE_FIELDS = ("x", "y", "psi", "a_lat", "d_sep", "d_goal", "delta_heading", "v_closing", "d_cpa", "d_dev", "r_d_goal",
                "r_h_goal", "r_d_cpa", "r_d_dev", "r_step", "steps", "reward", "flags")


class Trace(ctypes.Structure):
    """``acas2d_trace``: ring buffers of per-step episode records for a window of envs."""
    _fields_ = [("first_env", ctypes.c_int64), ("num_envs", ctypes.c_int64), ("capacity", ctypes.c_int32),
                ("n_traffic_rec", ctypes.c_int32), ("cursor", ctypes.c_void_p), ("rows", ctypes.c_void_p)]


class StepAux(ctypes.Structure):
    """``acas2d_step_aux``: optional per-step outputs."""
    _fields_ = [("flags", ctypes.c_void_p), ("outcome", ctypes.c_void_p), ("term_obs", ctypes.c_void_p),
                ("ep_return", ctypes.c_void_p), ("ep_length", ctypes.c_void_p)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC_DIR, s) for s in SOURCES] + [os.path.join(REPO_ROOT, "include", "acas2d_b200.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB_PATH, os.path.join(CSRC_DIR, "acas2d_kernels.cu")]
        subprocess.run(cmd, check=True)
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None


def declare(lib: ctypes.CDLL) -> ctypes.CDLL:
    """Attach argument / result types to every exported entry point."""
    vp, PP, SP, AP = ctypes.c_void_p, ctypes.POINTER(Params), ctypes.POINTER(State), ctypes.POINTER(StepAux)
    lib.acas2d_abi_version.argtypes = []
    lib.acas2d_params_default.argtypes = [PP, ctypes.c_int32]
    lib.acas2d_reset.argtypes = [PP, SP, vp, vp, vp]
    lib.acas2d_step.argtypes = [PP, SP, vp, vp, vp, vp, AP, vp]
    lib.acas2d_step_mapped.argtypes = [PP, SP, vp, vp, vp, vp, AP, vp]
    lib.acas2d_step_k.argtypes = [PP, SP, ctypes.c_int32, vp, vp, vp, vp, AP, vp]
    lib.acas2d_step_host.argtypes = [PP, SP, vp, vp, vp, vp, vp, vp, vp, vp, AP, vp]
    lib.acas2d_step_host_packed.argtypes = [PP, SP, vp, vp, vp, vp, vp, AP, vp, vp, ctypes.c_int64, vp]
    lib.acas2d_trace_step.argtypes = [PP, SP, vp, ctypes.POINTER(Trace), vp]
    lib.acas2d_inject_state.argtypes = [PP, SP, vp, vp, vp, vp, vp]
    lib.acas2d_extract_state.argtypes = [PP, SP, vp, vp, vp, vp, vp]
    lib.acas2d_rollout_random.argtypes = [PP, SP, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64, vp, vp]
    lib.acas2d_random_actions.argtypes = [SP, ctypes.c_uint64, ctypes.c_uint64, vp, vp]
    lib.acas2d_observe.argtypes = [PP, SP, vp, vp]
    lib.acas2d_render.argtypes = [PP, SP, ctypes.c_int64, vp, vp]
    lib.acas2d_launch_count.argtypes = []
    lib.acas2d_launch_count.restype = ctypes.c_int64
    lib.acas2d_set_tuning.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.acas2d_set_n1_kernel.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.acas2d_set_tiled_tuning.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.acas2d_policy_step.argtypes = [PP, SP, vp, ctypes.c_float, vp, vp, vp, vp, vp, vp, AP, ctypes.c_int32,
                                       ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, vp]
    lib.acas2d_policy_step_dyn.argtypes = [PP, SP, vp, ctypes.c_float, vp, vp, vp, vp, vp, vp, AP, ctypes.c_int32,
                                           ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, vp, vp, vp]
    CP = ctypes.POINTER(PpoConfig)
    lib.acas2d_ppo_values.argtypes = [vp, vp, ctypes.c_int64, vp, vp]
    lib.acas2d_ppo_gae.argtypes = [CP, vp, vp, vp, ctypes.c_int32, ctypes.c_int64, vp, vp, vp]
    lib.acas2d_ppo_grad.argtypes = [CP, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int64, vp, vp, vp, vp, vp]
    lib.acas2d_ppo_adam.argtypes = [CP, vp, vp, ctypes.c_float, vp, vp, vp, vp, vp]
    lib.acas2d_ppo_step.argtypes = [CP, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int64, vp, vp, vp, vp, vp, vp,
                                    ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(vp), vp]
    lib.acas2d_ppo_prepare.argtypes = []
    return lib


def load() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:                                        # build on demand (nvcc, sm_100a); still no CPU fallback
                build()
            except Exception:                           # noqa: BLE001
                pass
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The ACAS-2D batched step has no CPU fallback.")
        lib = declare(ctypes.CDLL(LIB_PATH))
        if lib.acas2d_abi_version() != ABI_VERSION:
            raise RuntimeError("libacas2d_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code == 0:
        return
    if code < 0:
        raise ValueError(f"{what}: {ERRORS.get(code, 'error')} (code {code})")
    raise RuntimeError(f"{what}: CUDA error {code}")


def params_from_settings(settings=None, n_traffic: Optional[int] = None, auto_reset: bool = False, **overrides) -> Params:
    """Build ``acas2d_params`` from a settings module / mapping (reference game.py:80-128)."""
    if settings is None:
        from gym_ACAS2D import settings as settings_mod
        settings = settings_mod
    get = (lambda k: overrides[k] if k in overrides else
           (settings[k] if isinstance(settings, dict) else getattr(settings, k)))
    lo, hi = int(get("MIN_TRAFFIC")), int(get("MAX_TRAFFIC"))
    if n_traffic is None:
        if lo != hi:
            raise NotImplementedError(
                "MIN_TRAFFIC != MAX_TRAFFIC: the reference pads the observation by 2 per missing intruder "
                "(game.py:213) and breaks its own Box shape; only a fixed traffic count is supported")
        n_traffic = hi
    n_traffic = int(n_traffic)
    if n_traffic < 1:
        raise ValueError("the reference requires at least one intruder (game.py:146-147,254-255)")
    if n_traffic > MAX_TRAFFIC:
        raise ValueError(f"n_traffic > {MAX_TRAFFIC} is not supported")
    p = Params()
    p.width, p.height, p.fps = float(get("WIDTH")), float(get("HEIGHT")), float(get("FPS"))
    p.max_steps = float(get("MAX_STEPS"))
    p.aircraft_size = float(get("AIRCRAFT_SIZE"))
    p.collision_radius = float(get("COLLISION_RADIUS"))
    p.goal_radius = float(get("GOAL_RADIUS"))
    p.safe_distance = float(get("SAFE_DISTANCE"))
    p.airspeed = float(get("AIRSPEED"))
    p.airspeed_factor_min = float(get("AIRSPEED_FACTOR_MIN"))
    p.airspeed_factor_max = float(get("AIRSPEED_FACTOR_MAX"))
    p.acc_lat_limit = float(get("ACC_LAT_LIMIT"))
    p.player_heading_lim = float(get("PLAYER_INITIAL_HEADING_LIM"))
    p.traffic_heading_lim = float(get("TRAFFIC_INITIAL_HEADING_LIM"))
    p.reward_goal, p.reward_collision = float(get("REWARD_GOAL")), float(get("REWARD_COLLISION"))
    p.goal_x = p.width - p.goal_radius                          # game.py:80
    p.goal_y = p.height / 2                                     # game.py:81
    p.player_x0 = p.collision_radius                            # game.py:85
    p.player_y0 = p.height / 2                                  # game.py:86
    p.player_psi_base = math.degrees(math.atan2(p.goal_y - p.player_y0, p.goal_x - p.player_x0) % (2 * math.pi))
    reach = (p.airspeed / p.fps) * p.max_steps
    p.d_goal_max = math.hypot(p.player_x0 - p.goal_x, p.player_y0 - p.goal_y) + reach     # game.py:120
    p.d_dev_max = reach                                         # game.py:122
    diag = math.sqrt(p.width ** 2 + p.height ** 2)
    p.d_separation_max = diag + 2 * reach                       # game.py:124
    p.d_cpa_max = diag                                          # game.py:126
    p.v_closing_max = 2 * (p.airspeed_factor_max * p.airspeed)  # game.py:128
    p.n_traffic = n_traffic
    p.auto_reset = 1 if auto_reset else 0
    return p
