"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle for the ACAS-2D environment step:

* ``Oracle``       -- ctypes wrapper over ``libacas2d_oracle.so`` (the float64 C
                      restatement in ``acas2d_oracle.c``), batched, used as the
                      parity checker for the CUDA path.
* ``PyPortGame``   -- a scalar pure-Python/numpy restatement with the reference's
                      own call pattern (numpy 2-vectors for every distance, each
                      quantity recomputed by observe / evaluate / is_done, per-step
                      record appends).  It exists so ``bench.py``'s CPU baseline is
                      timed on code shaped like the reference (a Python program),
                      and to replay the golden CSV through Python's ``random``.
* ``reference_spawn`` -- the reference's spawn draw order on a ``random.Random``
                      (game.py:41,88-114), needed to replay the golden CSV.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Parity status: PINNED (see
``acas2d_oracle.h``).  Citations are relative to the reference checkout
(``gym_ACAS2D/...``).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libacas2d_oracle.so")

# settings.py:9,15-17,31-48 (g = scipy.constants.g = 9.80665, settings.py:1)
DEFAULTS: Dict[str, float] = dict(
    MAX_STEPS=1000, WIDTH=1600, HEIGHT=1000, FPS=100,
    MIN_TRAFFIC=1, MAX_TRAFFIC=1,
    AIRCRAFT_SIZE=24, COLLISION_RADIUS=48, GOAL_RADIUS=144, SAFE_DISTANCE=192,
    AIRSPEED=200, AIRSPEED_FACTOR_MIN=1, AIRSPEED_FACTOR_MAX=1,
    ACC_LAT_LIMIT=20 * 9.80665,
    PLAYER_INITIAL_HEADING_LIM=3, TRAFFIC_INITIAL_HEADING_LIM=15,
    REWARD_GOAL=1000, REWARD_COLLISION=-1000,
)

FLAG_COLLISION, FLAG_GOAL, FLAG_TIMEOUT, FLAG_DONE = 1, 2, 4, 8


class _Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "width", "height", "fps", "max_steps", "aircraft_size", "collision_radius",
        "goal_radius", "safe_distance", "airspeed", "airspeed_factor_min",
        "airspeed_factor_max", "acc_lat_limit", "player_heading_lim", "traffic_heading_lim",
        "reward_goal", "reward_collision", "goal_x", "goal_y", "player_x0", "player_y0",
        "d_goal_max", "d_dev_max", "d_separation_max", "d_cpa_max", "v_closing_max")]


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc, seconds) and return the .so path."""
    src = os.path.join(_HERE, "acas2d_oracle.c")
    hdr = os.path.join(_HERE, "acas2d_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-s", "-C", _HERE, "-B", "libacas2d_oracle.so"], check=True)
    return _LIB_PATH


def make_constants(**overrides) -> Dict[str, float]:
    c = dict(DEFAULTS)
    for k, v in overrides.items():
        if k not in c:
            raise KeyError(k)
        c[k] = v
    return c


def derived(c: Dict[str, float]) -> Dict[str, float]:
    """Goal, player start and observation normalisers (game.py:80-89,120-128)."""
    goal_x = c["WIDTH"] - c["GOAL_RADIUS"]
    goal_y = c["HEIGHT"] / 2
    px0 = c["COLLISION_RADIUS"]
    py0 = c["HEIGHT"] / 2
    reach = (c["AIRSPEED"] / c["FPS"]) * c["MAX_STEPS"]
    diag = float(np.sqrt(c["WIDTH"] ** 2 + c["HEIGHT"] ** 2))
    return dict(
        goal_x=float(goal_x), goal_y=float(goal_y), player_x0=float(px0), player_y0=float(py0),
        d_goal_max=float(np.linalg.norm(np.array((px0, py0)) - np.array((goal_x, goal_y)), 2)) + reach,
        d_dev_max=reach,
        d_separation_max=diag + 2 * reach,
        d_cpa_max=diag,
        v_closing_max=2 * (c["AIRSPEED_FACTOR_MAX"] * c["AIRSPEED"]),
    )


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctype))


class Oracle:
    """Batched float64 oracle.  State arrays are plain numpy, owned by the caller."""

    def __init__(self, n_traffic: int = 1, **overrides):
        self.N = int(n_traffic)
        if self.N < 1:
            raise ValueError("the reference requires at least one intruder (game.py:146-147)")
        self.L = 5 + 3 * self.N
        self.c = make_constants(**overrides)
        self.d = derived(self.c)
        self.lib = ctypes.CDLL(build())
        p = _Params()
        p.width, p.height, p.fps = self.c["WIDTH"], self.c["HEIGHT"], self.c["FPS"]
        p.max_steps = self.c["MAX_STEPS"]
        p.aircraft_size = self.c["AIRCRAFT_SIZE"]
        p.collision_radius = self.c["COLLISION_RADIUS"]
        p.goal_radius = self.c["GOAL_RADIUS"]
        p.safe_distance = self.c["SAFE_DISTANCE"]
        p.airspeed = self.c["AIRSPEED"]
        p.airspeed_factor_min = self.c["AIRSPEED_FACTOR_MIN"]
        p.airspeed_factor_max = self.c["AIRSPEED_FACTOR_MAX"]
        p.acc_lat_limit = self.c["ACC_LAT_LIMIT"]
        p.player_heading_lim = self.c["PLAYER_INITIAL_HEADING_LIM"]
        p.traffic_heading_lim = self.c["TRAFFIC_INITIAL_HEADING_LIM"]
        p.reward_goal, p.reward_collision = self.c["REWARD_GOAL"], self.c["REWARD_COLLISION"]
        for k, v in self.d.items():
            setattr(p, k, v)
        self.params = p
        self._declare()

    # -- ctypes signatures ---------------------------------------------------------------
    def _declare(self):
        L = self.lib
        D, I32, U8, U32 = (ctypes.POINTER(t) for t in (ctypes.c_double, ctypes.c_int32, ctypes.c_uint8, ctypes.c_uint32))
        PP = ctypes.POINTER(_Params)
        i64, u64, ci = ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
        L.acas2d_oracle_observe.argtypes = [PP, i64, ci, D, D, I32, D]
        L.acas2d_oracle_step.argtypes = [PP, i64, ci, D, D, I32, D, U8, D, D, D, D, D, U8, U8]
        L.acas2d_oracle_rollout.argtypes = [PP, i64, ci, i64, D, D, I32, D, U8, D, D, D, D, D, U8, U8, D, D]
        L.acas2d_oracle_spawn_philox.argtypes = [PP, i64, ci, u64, u64, U32, D, D]
        L.acas2d_oracle_vec_step.argtypes = [PP, i64, ci, u64, u64, U32, D, D, I32, D, D, D, D, D, D, U8, U8, D, D, I32]
        L.acas2d_oracle_philox4x32_10.argtypes = [U32, U32, U32]
        for f in ("pymod", "distance", "relative_angle", "delta_heading", "heading_reward"):
            getattr(L, "acas2d_oracle_" + f).restype = ctypes.c_double
        L.acas2d_oracle_pymod.argtypes = [ctypes.c_double] * 2
        L.acas2d_oracle_distance.argtypes = [ctypes.c_double] * 4
        L.acas2d_oracle_relative_angle.argtypes = [ctypes.c_double] * 4
        L.acas2d_oracle_delta_heading.argtypes = [ctypes.c_double] * 2
        L.acas2d_oracle_heading_reward.argtypes = [ctypes.c_double] * 2
        for f in ("closest_approach_reward", "plan_deviation_reward", "goal_distance_reward", "step_reward_5"):
            getattr(L, "acas2d_oracle_" + f).restype = ctypes.c_double
        L.acas2d_oracle_closest_approach_reward.argtypes = [PP, ctypes.c_double, ctypes.c_double]
        L.acas2d_oracle_plan_deviation_reward.argtypes = [PP, ctypes.c_double]
        L.acas2d_oracle_goal_distance_reward.argtypes = [PP, ctypes.c_double]
        L.acas2d_oracle_step_reward_5.argtypes = [PP] + [ctypes.c_double] * 6

    # -- scalar reward terms (for the notebook known-answer tests) -------------------------
    def heading_reward(self, psi, phi):
        return self.lib.acas2d_oracle_heading_reward(psi, phi)

    def closest_approach_reward(self, v_closing, d_cpa):
        return self.lib.acas2d_oracle_closest_approach_reward(ctypes.byref(self.params), v_closing, d_cpa)

    def plan_deviation_reward(self, d_dev):
        return self.lib.acas2d_oracle_plan_deviation_reward(ctypes.byref(self.params), d_dev)

    def goal_distance_reward(self, d_goal):
        return self.lib.acas2d_oracle_goal_distance_reward(ctypes.byref(self.params), d_goal)

    def step_reward_5(self, v_closing, psi, phi, d_cpa, d_goal, d_dev):
        return self.lib.acas2d_oracle_step_reward_5(ctypes.byref(self.params), v_closing, psi, phi, d_cpa, d_goal, d_dev)

    def pymod(self, x, m):
        return self.lib.acas2d_oracle_pymod(x, m)

    # -- state containers --------------------------------------------------------------------
    def new_state(self, B: int) -> Dict[str, np.ndarray]:
        return dict(
            player=np.zeros((B, 5)), traffic=np.zeros((B, self.N, 4)),
            steps=np.zeros(B, np.int32), total_reward=np.zeros(B),
            running=np.ones(B, np.uint8), d_path=np.zeros(B), min_sep=np.full(B, np.inf),
            episode_idx=np.zeros(B, np.uint32),
        )

    @staticmethod
    def _check(st):
        for k, v in st.items():
            assert v.flags["C_CONTIGUOUS"], k

    def spawn_philox(self, st, seed: int, env_id_offset: int = 0):
        self._check(st)
        B = st["player"].shape[0]
        self.lib.acas2d_oracle_spawn_philox(
            ctypes.byref(self.params), B, self.N, seed, env_id_offset,
            _ptr(st["episode_idx"], ctypes.c_uint32), _ptr(st["player"], ctypes.c_double),
            _ptr(st["traffic"], ctypes.c_double))
        st["steps"][:] = 0
        st["total_reward"][:] = 0
        st["running"][:] = 1
        st["d_path"][:] = 0
        d = st["traffic"][:, :, :2] - st["player"][:, None, :2]
        st["min_sep"][:] = np.sqrt((d * d).sum(-1)).min(-1)

    def observe(self, st) -> np.ndarray:
        """game.observe(): increments ``steps`` (Q5) and returns obs [B, L]."""
        self._check(st)
        B = st["player"].shape[0]
        obs = np.empty((B, self.L))
        self.lib.acas2d_oracle_observe(ctypes.byref(self.params), B, self.N,
                                       _ptr(st["player"], ctypes.c_double), _ptr(st["traffic"], ctypes.c_double),
                                       _ptr(st["steps"], ctypes.c_int32), _ptr(obs, ctypes.c_double))
        return obs

    def step(self, st, actions: np.ndarray):
        self._check(st)
        B = st["player"].shape[0]
        a = np.ascontiguousarray(actions, np.float64).reshape(B)
        obs = np.empty((B, self.L)); rew = np.empty(B)
        flags = np.zeros(B, np.uint8); outcome = np.zeros(B, np.uint8)
        D = ctypes.c_double
        self.lib.acas2d_oracle_step(ctypes.byref(self.params), B, self.N,
                                    _ptr(st["player"], D), _ptr(st["traffic"], D), _ptr(st["steps"], ctypes.c_int32),
                                    _ptr(st["total_reward"], D), _ptr(st["running"], ctypes.c_uint8),
                                    _ptr(st["d_path"], D), _ptr(st["min_sep"], D), _ptr(a, D),
                                    _ptr(obs, D), _ptr(rew, D), _ptr(flags, ctypes.c_uint8), _ptr(outcome, ctypes.c_uint8))
        return obs, rew, flags, outcome

    def rollout(self, st, actions: np.ndarray, record_traffic: bool = False):
        """actions [T, B] -> dict of per-step records (finished envs repeat their terminal row)."""
        self._check(st)
        T, B = actions.shape
        a = np.ascontiguousarray(actions, np.float64)
        out = dict(obs=np.empty((T, B, self.L)), reward=np.empty((T, B)),
                   flags=np.zeros((T, B), np.uint8), outcome=np.zeros((T, B), np.uint8),
                   player=np.empty((T, B, 3)))
        out["traffic"] = np.empty((T, B, self.N, 2)) if record_traffic else None
        D = ctypes.c_double
        self.lib.acas2d_oracle_rollout(ctypes.byref(self.params), B, self.N, T,
                                       _ptr(st["player"], D), _ptr(st["traffic"], D), _ptr(st["steps"], ctypes.c_int32),
                                       _ptr(st["total_reward"], D), _ptr(st["running"], ctypes.c_uint8),
                                       _ptr(st["d_path"], D), _ptr(st["min_sep"], D), _ptr(a, D),
                                       _ptr(out["obs"], D), _ptr(out["reward"], D),
                                       _ptr(out["flags"], ctypes.c_uint8), _ptr(out["outcome"], ctypes.c_uint8),
                                       _ptr(out["player"], D), _ptr(out["traffic"], D))
        return out

    def vec_step(self, st, actions: np.ndarray, seed: int, env_id_offset: int = 0):
        """Auto-resetting step (SB3 DummyVecEnv semantics, Philox respawn)."""
        self._check(st)
        B = st["player"].shape[0]
        a = np.ascontiguousarray(actions, np.float64).reshape(B)
        obs = np.empty((B, self.L)); rew = np.empty(B)
        flags = np.zeros(B, np.uint8); outcome = np.zeros(B, np.uint8)
        term_obs = np.full((B, self.L), np.nan); ep_ret = np.full(B, np.nan); ep_len = np.zeros(B, np.int32)
        D = ctypes.c_double
        self.lib.acas2d_oracle_vec_step(ctypes.byref(self.params), B, self.N, seed, env_id_offset,
                                        _ptr(st["episode_idx"], ctypes.c_uint32),
                                        _ptr(st["player"], D), _ptr(st["traffic"], D), _ptr(st["steps"], ctypes.c_int32),
                                        _ptr(st["total_reward"], D), _ptr(st["d_path"], D), _ptr(st["min_sep"], D),
                                        _ptr(a, D), _ptr(obs, D), _ptr(rew, D), _ptr(flags, ctypes.c_uint8),
                                        _ptr(outcome, ctypes.c_uint8), _ptr(term_obs, D), _ptr(ep_ret, D),
                                        _ptr(ep_len, ctypes.c_int32))
        return obs, rew, flags, outcome, term_obs, ep_ret, ep_len

    def philox(self, ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        self.lib.acas2d_oracle_philox4x32_10(c, k, o)
        return tuple(int(x) for x in o)


# ----------------------------------------------------------------------------------------------
# Reference-order spawn on Python's MT19937 stream (needed to replay the golden CSV).
# ----------------------------------------------------------------------------------------------
def reference_spawn(rng, c: Dict[str, float], n_traffic_minmax=None):
    """Consume ``rng`` exactly like ``ACAS2DGame.__init__`` (game.py:41,88-114; Q18).

    Returns (player_row[5], traffic[N][4])."""
    lo, hi = n_traffic_minmax or (int(c["MIN_TRAFFIC"]), int(c["MAX_TRAFFIC"]))
    n = rng.randint(lo, hi)                                                    # game.py:41
    d = derived(c)
    rng.uniform(0, 360)                                                        # game.py:88 (discarded, Q18)
    base = math.degrees(math.atan2(d["goal_y"] - d["player_y0"], d["goal_x"] - d["player_x0"]) % (2 * math.pi))
    hl = c["PLAYER_INITIAL_HEADING_LIM"]
    psi = (base + rng.uniform(-hl, hl)) % 360                                  # game.py:91-92
    player = np.array([d["player_x0"], d["player_y0"], c["AIRSPEED"], psi, 0.0])
    traffic = np.zeros((n, 4))
    tl = c["TRAFFIC_INITIAL_HEADING_LIM"]
    for i in range(n):
        if i == 0:
            sd = rng.randint(0, 1)                                             # game.py:98
            x = c["WIDTH"] - c["COLLISION_RADIUS"]
            y = c["COLLISION_RADIUS"] + (sd * (c["HEIGHT"] - (2 * c["COLLISION_RADIUS"])))
            v = rng.uniform(c["AIRSPEED_FACTOR_MIN"], c["AIRSPEED_FACTOR_MAX"]) * c["AIRSPEED"]
            h = (145 + (sd * 70) + rng.uniform(-tl, tl)) % 360                 # game.py:105-106
        else:
            x = rng.uniform(0, c["WIDTH"] - c["AIRCRAFT_SIZE"])                # game.py:109
            y = rng.uniform(0, 3 * c["HEIGHT"] / 5)
            v = rng.uniform(c["AIRSPEED_FACTOR_MIN"], c["AIRSPEED_FACTOR_MAX"]) * c["AIRSPEED"]
            h = rng.uniform(0, 360)
        traffic[i] = (x, y, v, h)
    return player, traffic


# ----------------------------------------------------------------------------------------------
# Scalar Python port with the reference's call pattern (CPU-baseline stand-in).
# ----------------------------------------------------------------------------------------------
class _Craft:
    __slots__ = ("x", "y", "v_air", "psi", "a_lat")

    def __init__(self, x, y, v_air, psi):
        self.x, self.y, self.v_air, self.psi, self.a_lat = x, y, v_air, psi, 0


def _norm2(xa, ya, xb, yb):                                                    # kinematics.py:7-13
    return np.linalg.norm(np.array((xa, ya)) - np.array((xb, yb)), 2)


def _bearing(xa, ya, xb, yb):                                                  # kinematics.py:16-22
    return math.degrees(math.atan2(yb - ya, xb - xa) % (2 * math.pi))


class PyPortGame:
    """One episode of the reference game, float64 Python scalars (game.py:27-314)."""

    def __init__(self, rng, consts: Optional[Dict[str, float]] = None, keep_records: bool = True):
        c = self.c = consts or DEFAULTS
        self.dt = 1 / c["FPS"]
        d = derived(c)
        self.goal = (d["goal_x"], d["goal_y"])
        pl, tr = reference_spawn(rng, c)
        self.player = _Craft(c["COLLISION_RADIUS"], c["HEIGHT"] / 2, c["AIRSPEED"], float(pl[3]))
        self.traffic = [_Craft(float(r[0]), float(r[1]), float(r[2]), float(r[3])) for r in tr]
        self.steps, self.total_reward, self.outcome, self.running = 0, 0, None, True
        self.d_path = 0
        self.d_goal_max = self._d_goal() + (c["AIRSPEED"] / c["FPS"]) * c["MAX_STEPS"]
        self.d_dev_max, self.d_sep_max = d["d_dev_max"], d["d_separation_max"]
        self.d_cpa_max, self.v_closing_max = d["d_cpa_max"], d["v_closing_max"]
        self.keep = keep_records
        self.path = [(self.player.x, self.player.y)]
        self.traffic_paths = [[(t.x, t.y)] for t in self.traffic]
        self.rec = {k: [] for k in ("psi", "sep", "a_lat", "d_goal", "dh", "vc", "dcpa", "ddev",
                                    "r_goal", "r_head", "r_cpa", "r_dev", "r")}
        self._log_terms(discount=None)
        self.rec["psi"].append(self.player.psi); self.rec["sep"].append(self._min_sep()); self.rec["a_lat"].append(0)

    # geometry (game.py:162-192)
    def _min_sep(self):
        return np.min([_norm2(self.player.x, self.player.y, t.x, t.y) for t in self.traffic])

    def _d_goal(self):
        return _norm2(self.player.x, self.player.y, *self.goal)

    def _h_goal(self):
        return _bearing(self.player.x, self.player.y, *self.goal)

    def _dev(self):
        return self._d_goal() * np.sin((self._h_goal() / 360.0) * 2 * math.pi)

    def _hit(self):
        for t in self.traffic:
            if _norm2(self.player.x, self.player.y, t.x, t.y) < 2 * self.c["COLLISION_RADIUS"]:
                return True
        return False

    def _at_goal(self):
        return self._d_goal() < self.c["GOAL_RADIUS"]

    # kinematics (aircraft.py:16-26, kinematics.py:25-79)
    def _advance(self, a):
        dt = self.dt
        rate = a.a_lat / (a.v_air * dt)
        a.psi = (a.psi + (rate * dt)) % 360
        r = (a.psi / 360.0) * 2 * math.pi
        a.x = a.x + (a.v_air * math.cos(r) * dt)
        a.y = a.y + (a.v_air * math.sin(r) * dt)

    def _dcpa(self, a, b):
        d = _norm2(a.x, a.y, b.x, b.y)
        ar = (_bearing(a.x, a.y, b.x, b.y) / 360.0) * 2 * math.pi
        ra, rb = (a.psi / 360.0) * 2 * math.pi, (b.psi / 360.0) * 2 * math.pi
        vx = a.v_air * np.cos(ra) - b.v_air * np.cos(rb)
        vy = a.v_air * np.sin(ra) - b.v_air * np.sin(rb)
        with np.errstate(all="ignore"):
            return d * np.sin(ar - np.arctan(vy / vx))

    def _vclose(self, a, b):
        dt = self.dt
        ha = ((a.psi + ((a.a_lat / a.v_air) * dt)) % 360 / 360.0) * 2 * math.pi
        hb = ((b.psi + ((b.a_lat / b.v_air) * dt)) % 360 / 360.0) * 2 * math.pi
        pa = np.array([a.x + (a.v_air * math.cos(ha) * dt), a.y + (a.v_air * math.sin(ha) * dt)])
        ua = np.array([a.v_air * math.cos(ha) * dt, a.v_air * math.sin(ha) * dt])
        pb = np.array([b.x + (b.v_air * math.cos(hb) * dt), b.y + (b.v_air * math.sin(hb) * dt)])
        ub = np.array([b.v_air * math.cos(hb) * dt, a.v_air * math.sin(hb) * dt])   # Q3
        return (np.dot((ua - ub), (pa - pb)) / _norm2(pa[0], pa[1], pb[0], pb[1])) / dt

    # rewards (rewards.py:5-60)
    def _r_head(self, psi, phi):
        dh = min(abs(psi - phi), 360 - abs(psi - phi))
        return (1 - dh / 180) ** 4

    def _r_cpa(self, vc, dcpa):
        return 1 if vc > 0 else min(1, (dcpa / self.c["SAFE_DISTANCE"]) ** 4)

    def _r_dev(self, ddev):
        ddev = abs(ddev)
        lim = ((self.c["WIDTH"] - self.c["GOAL_RADIUS"]) - (2 * self.c["AIRCRAFT_SIZE"])) / 2
        return 0 if ddev > lim else (1 - ddev / lim) ** 0.5

    def _r_goal(self, dg):
        c = self.c
        top = ((c["WIDTH"] - c["GOAL_RADIUS"]) - (2 * c["AIRCRAFT_SIZE"])) + (c["AIRSPEED"] / c["FPS"]) * c["MAX_STEPS"]
        return min(1, (1 - dg / top) ** 4)

    def _shaped(self, vc, psi, phi, dcpa, dg, ddev):
        if vc <= 0:
            return self._r_head(psi, phi) * self._r_cpa(vc, dcpa) * self._r_dev(ddev)
        return self._r_head(psi, phi) * self._r_goal(dg)

    def _log_terms(self, discount):
        psi, phi = self.player.psi, self._h_goal()
        vc, dcpa = self._vclose(self.player, self.traffic[0]), self._dcpa(self.player, self.traffic[0])
        dg, ddev = self._d_goal(), self._dev()
        r = self._shaped(vc, psi, phi, dcpa, dg, ddev)
        if discount is not None:
            r = r * discount
        if self.keep:
            k = self.rec
            k["d_goal"].append(dg); k["dh"].append(min(abs(psi - phi), 360 - abs(psi - phi)))
            k["vc"].append(vc); k["dcpa"].append(dcpa); k["ddev"].append(ddev)
            k["r_goal"].append(self._r_goal(dg)); k["r_head"].append(self._r_head(psi, phi))
            k["r_cpa"].append(self._r_cpa(vc, dcpa)); k["r_dev"].append(self._r_dev(ddev)); k["r"].append(r)
        return r

    # game.observe / action / evaluate / is_done
    def observe(self):
        self.steps += 1
        p = self.player
        row = [self.steps / self.c["MAX_STEPS"], p.psi / 360, self._dev() / self.d_dev_max,
               self._d_goal() / self.d_goal_max, self._h_goal() / 360]
        for t in self.traffic:
            row.append(_norm2(p.x, p.y, t.x, t.y) / self.d_sep_max)
            row.append(self._dcpa(p, t) / self.d_cpa_max)
            row.append(self._vclose(p, t) / self.v_closing_max)
        return np.array(row).astype(np.float64)

    def action(self, act):
        p = self.player
        p.a_lat = act[0] * self.c["ACC_LAT_LIMIT"]
        xo, yo = p.x, p.y
        self._advance(p)
        if self.keep:
            self.path.append((p.x, p.y))
            for i, t in enumerate(self.traffic):
                self.traffic_paths[i].append((t.x, t.y))
            self.rec["psi"].append(p.psi); self.rec["sep"].append(self._min_sep()); self.rec["a_lat"].append(p.a_lat)
        self.d_path += _norm2(xo, yo, p.x, p.y)
        for t in self.traffic:
            if self.running:
                self._advance(t)

    def evaluate(self):
        r = self._log_terms(discount=1 - (self.steps / self.c["MAX_STEPS"]))
        if self._hit():
            r += self.c["REWARD_COLLISION"]
        if self._at_goal():
            r += self.c["REWARD_GOAL"]
        self.total_reward += r
        return r

    def is_done(self):
        if self.steps > self.c["MAX_STEPS"]:
            self.outcome = 3
        elif self._hit():
            self.outcome = 2
        elif self._at_goal():
            self.outcome = 1
        else:
            return False
        self.running = False
        return True

    def step(self, act):
        self.action(act)
        o = self.observe()
        r = self.evaluate()
        return o, r, self.is_done(), {}


def pyport_random_rollout(n_steps: int, seed: int = 13, consts=None) -> int:
    """Random-action rollout with resets on the Python port; returns env steps done."""
    import random
    rng = random.Random(seed)
    arng = np.random.default_rng(seed)
    g = PyPortGame(rng, consts); g.observe()
    acts = arng.uniform(-1, 1, size=(n_steps, 1))
    for k in range(n_steps):
        _, _, done, _ = g.step(acts[k])
        if done:
            g = PyPortGame(rng, consts); g.observe()
    return n_steps
