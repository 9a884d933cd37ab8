"""``ACAS2DEnv`` -- the reference's gym surface (gym_ACAS2D/envs/environment.py:8-54) as a
B = 1 view over the batched CUDA core.

Same constructor (no arguments needed), same spaces (environment.py:18-27), same old-gym
4-tuple ``step`` and ``reset() -> obs``, observations returned as fresh ``float64`` arrays,
``render`` / ``close`` are no-ops (rendering is off the hot path).  Differences, all
deliberate (DESIGN.md "Boundary"):
  * no ``pygame.time.Clock.tick(FPS)``: the reference is wall-clock capped at 100 steps/s (Q14);
  * spawns come from Philox (seed, env id, episode) instead of Python's global ``random``;
  * no "Outcome: ..." print on episode end (Q16) unless ``verbose=True``.
"""
from __future__ import annotations

import numpy as np

from .batched import BatchedACAS2D
from .game import ACAS2DGame
from .spaces import GYM_FLAVOUR, action_box, observation_box

if GYM_FLAVOUR == "gym":                 # pragma: no cover
    import gym as _gym
    _EnvBase = _gym.Env
elif GYM_FLAVOUR == "gymnasium":         # pragma: no cover
    import gymnasium as _gym
    _EnvBase = _gym.Env
else:
    _EnvBase = object


class ACAS2DEnv(_EnvBase):
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, n_traffic=None, device="cuda", seed=None, env_id=0, verbose=False, settings=None):
        self._core = BatchedACAS2D(1, n_traffic=n_traffic, device=device, seed=seed, env_id_offset=env_id,
                                   auto_reset=False, settings=settings, host_zero_copy=True)
        self._verbose = verbose
        self._episode = 0
        n = self._core.n_traffic
        self.observation_space = observation_box(n, np.float64)      # environment.py:18-21
        self.action_space = action_box(np.float64)                   # environment.py:27
        # the reference builds the first game in __init__ (environment.py:13)
        self._core.reset()
        self.game = ACAS2DGame(self._core, 0)

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(-1)[:1]
        obs, reward, done = self._core.step_host(a)
        g = self.game
        g._a_lat = float(a[0]) * self._core.params.acc_lat_limit     # game.py:225
        g._step_calls += 1
        d = bool(done[0])
        if d:
            g.running = False
            g.outcome = int(self._core.host_buffers()["outcome"][0])     # came with the step's one packed copy
            if self._verbose:
                from gym_ACAS2D.settings import OUTCOME_NAMES
                print("Outcome: {:<10} - Time steps: {:<10} - Total Reward: {}".format(
                    OUTCOME_NAMES[g.outcome].upper(), g.steps, g.total_reward))
        return obs[0].astype(np.float64), float(reward[0]), d, {}

    def reset(self):
        obs = self._core.reset()
        self._episode += 1
        self.game = ACAS2DGame(self._core, 0)
        return obs[0].cpu().numpy().astype(np.float64)

    def render(self, mode="human"):
        """``mode="rgb_array"`` returns a uint8 [HEIGHT, WIDTH, 3] frame; ``"human"`` (the reference's pygame
        window, game.py:316-431) is a no-op: nothing is drawn on the hot path."""
        if mode == "rgb_array":
            return self._core.render(0).cpu().numpy()
        return None

    def close(self):
        return None
