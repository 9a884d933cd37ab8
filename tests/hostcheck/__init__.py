"""TEST INFRASTRUCTURE: numpy-backed twin of ``BatchedACAS2D`` that drives the g++ build of the
product's per-env source (``hostcheck.cpp``).  Lets the CPU test-suite check the step logic
that the CUDA kernels inline against the oracle.  Never imported by the product."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, os.path.join(_ROOT, "gym-acas2d_b200"))

from gym_ACAS2D.envs import _native  # noqa: E402
from gym_ACAS2D.envs._native import Params, State, StepAux  # noqa: E402

_LIB = os.path.join(_HERE, "libacas2d_hostcheck.so")


def build() -> str:
    srcs = [os.path.join(_HERE, "hostcheck.cpp")] + [os.path.join(_native.CSRC_DIR, s) for s in ("acas2d_env.cuh", "acas2d_math.cuh", "acas2d_policy.cuh")]
    if not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", _LIB,
                        os.path.join(_HERE, "hostcheck.cpp"), "-lm"], check=True)
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data


class HostBatch:
    def __init__(self, num_envs, n_traffic=1, seed=13, env_id_offset=0, auto_reset=True, track_min_sep=False,
                 variant=None, **overrides):
        self.lib = ctypes.CDLL(build())
        vp, PP, SP, AP = ctypes.c_void_p, ctypes.POINTER(Params), ctypes.POINTER(State), ctypes.POINTER(StepAux)
        self.lib.hostcheck_reset.argtypes = [PP, SP, vp, vp]
        self.lib.hostcheck_step.argtypes = [PP, SP, vp, vp, vp, vp, AP, ctypes.c_int]
        self.lib.hostcheck_rollout_random.argtypes = [PP, SP, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64, vp]
        self.lib.hostcheck_random_actions.argtypes = [SP, ctypes.c_uint64, ctypes.c_uint64, vp]
        self.lib.hostcheck_inject.argtypes = [PP, SP, vp, vp, vp, vp]
        self.lib.hostcheck_observe.argtypes = [PP, SP, vp]
        self.lib.hostcheck_extract.argtypes = [PP, SP, vp, vp, vp, vp]
        self.lib.hostcheck_wrap360.argtypes = [ctypes.c_double]
        self.lib.hostcheck_wrap360.restype = ctypes.c_double
        self.params = _native.params_from_settings(None, n_traffic, auto_reset, **overrides)
        B, N = int(num_envs), int(n_traffic)
        self.num_envs, self.n_traffic, self.obs_dim = B, N, 5 + 3 * N
        self.variant = (0 if N == 1 else 1) if variant is None else variant
        f8, f4 = np.float64, np.float32
        self.ppos = np.zeros((B, 2), f8); self.paux = np.zeros((B, 2), f8)
        self.thot = np.zeros((B, N, 4), f4); self.tres = np.zeros((B, N, 4), f8)
        self.episode_idx = np.zeros(B, np.uint32)
        self.min_sep = np.zeros(B, f4) if track_min_sep else None
        self.stats = np.zeros((_native.STAT_SLOTS, _native.STAT_FIELDS), np.int64)
        self.obs = np.zeros((B, self.obs_dim), f4); self.reward = np.zeros(B, f4)
        self.done = np.zeros(B, np.uint8); self.flags = np.zeros(B, np.uint8); self.outcome = np.zeros(B, np.uint8)
        self.term_obs = np.full((B, self.obs_dim), np.nan, f4); self.ep_return = np.zeros(B, f4)
        self.ep_length = np.zeros(B, np.int32)
        self._state = State(num_envs=B, ppos=_p(self.ppos), paux=_p(self.paux), thot=_p(self.thot),
                            tres=_p(self.tres), episode_idx=_p(self.episode_idx), min_sep=_p(self.min_sep), stats=_p(self.stats),
                            seed=seed, env_id_offset=env_id_offset)
        self._aux = StepAux(flags=_p(self.flags), outcome=_p(self.outcome), term_obs=_p(self.term_obs),
                            ep_return=_p(self.ep_return), ep_length=_p(self.ep_length))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.lib.hostcheck_reset(ctypes.byref(self.params), ctypes.byref(self._state), _p(m), _p(self.obs))
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.num_envs)
        rc = self.lib.hostcheck_step(ctypes.byref(self.params), ctypes.byref(self._state), _p(a), _p(self.obs),
                                     _p(self.reward), _p(self.done), ctypes.byref(self._aux), self.variant)
        assert rc == 0, rc
        return self.obs, self.reward, self.done.view(np.bool_)

    def rollout_random(self, num_steps, action_seed=0, step0=0, reward_sum=None):
        self.lib.hostcheck_rollout_random(ctypes.byref(self.params), ctypes.byref(self._state), num_steps,
                                          action_seed, step0, _p(reward_sum))

    def random_actions(self, step_index, action_seed=0):
        out = np.zeros(self.num_envs, np.float32)
        self.lib.hostcheck_random_actions(ctypes.byref(self._state), action_seed, step_index, _p(out))
        return out

    def inject_state(self, player, traffic, steps=None, total_reward=None):
        B, N = self.num_envs, self.n_traffic
        pl = np.ascontiguousarray(player, np.float64).reshape(B, 3)
        tr = np.ascontiguousarray(traffic, np.float64).reshape(B, N, 4)
        st = np.ones(B, np.int32) if steps is None else np.ascontiguousarray(steps, np.int32)
        tot = np.zeros(B) if total_reward is None else np.ascontiguousarray(total_reward, np.float64)
        self.lib.hostcheck_inject(ctypes.byref(self.params), ctypes.byref(self._state), _p(pl), _p(tr), _p(st), _p(tot))

    def observe(self):
        self.lib.hostcheck_observe(ctypes.byref(self.params), ctypes.byref(self._state), _p(self.obs))
        return self.obs

    def extract_state(self):
        B, N = self.num_envs, self.n_traffic
        pl = np.zeros((B, 3)); tr = np.zeros((B, N, 4)); st = np.zeros(B, np.int32); tot = np.zeros(B)
        self.lib.hostcheck_extract(ctypes.byref(self.params), ctypes.byref(self._state), _p(pl), _p(tr), _p(st), _p(tot))
        out = dict(player=pl, traffic=tr, steps=st, total_reward=tot, episode_idx=self.episode_idx.copy())
        if self.min_sep is not None:
            out["min_sep"] = self.min_sep.copy()
        return out

    def episode_counters(self):
        return self.stats.sum(0)[: len(_native.STAT_NAMES)]
