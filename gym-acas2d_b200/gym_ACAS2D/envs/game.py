"""``ACAS2DGame`` -- a host-side VIEW of one game living in the batched device state.

The reference's ``ACAS2DGame`` (gym_ACAS2D/envs/game.py:8-314) owns the simulation.  Here the
simulation lives in HBM (``BatchedACAS2D``); this class only gives scripts the attribute
surface they reach through ``environment.game`` (baseline_main.py:36-58,
testing_main.py:84-105): ``outcome, total_reward, steps, d_path, episode, quit, running,
num_traffic, goal_x, goal_y, player, traffic`` and the normalisers.  Writing
``game.player.x = ...`` / ``game.traffic[i].psi = ...`` / ``game.steps = ...`` injects state
into the device batch, which is how the reference is poked in parity tests.
"""
from __future__ import annotations

import numpy as np


class _AircraftView:
    _FIELDS = ("x", "y", "v_air", "psi")

    def __init__(self, game: "ACAS2DGame", index: int):
        object.__setattr__(self, "_game", game)
        object.__setattr__(self, "_index", index)     # -1 = player

    def __getattr__(self, name):
        if name in self._FIELDS:
            return self._game._get_aircraft(self._index, name)
        if name == "a_lat":
            return self._game._a_lat if self._index < 0 else 0.0
        if name == "psi_dot":
            return 0.0
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name not in self._FIELDS:
            raise AttributeError(f"cannot set {name}")
        self._game._set_aircraft(self._index, name, float(value))

    def out_of_bounds(self, width, height):          # reference aircraft.py:28-29
        return self.x < 0 or self.x > width or self.y < 0 or self.y > height


class ACAS2DGame:
    def __init__(self, core, env_index: int = 0, episode=None):
        self._core, self._i = core, int(env_index)
        self.episode = episode
        self.quit = False
        self.manual = False
        self.outcome = None
        self.running = True
        self._a_lat = 0.0
        self._step_calls = 0
        p = core.params
        self.num_traffic = core.n_traffic
        self.goal_x, self.goal_y = p.goal_x, p.goal_y
        self.d_goal_max, self.d_dev_max = p.d_goal_max, p.d_dev_max
        self.d_separation_max, self.d_cpa_max, self.v_closing_max = p.d_separation_max, p.d_cpa_max, p.v_closing_max
        self.player = _AircraftView(self, -1)
        self.traffic = [_AircraftView(self, j) for j in range(core.n_traffic)]

    # ---- state access through extract / inject
    def _row(self):
        s = self._core.extract_state()
        i = self._i
        return s["player"][i], s["traffic"][i], int(s["steps"][i]), float(s["total_reward"][i]), s

    def _get_aircraft(self, j, name):
        pl, tr, _, _, _ = self._row()
        if j < 0:
            return {"x": pl[0], "y": pl[1], "psi": pl[2], "v_air": float(self._core.params.airspeed)}[name]
        return float(tr[j][{"x": 0, "y": 1, "v_air": 2, "psi": 3}[name]])

    def _set_aircraft(self, j, name, value):
        _, _, _, _, s = self._row()
        i = self._i
        if j < 0:
            if name == "v_air":
                if value != self._core.params.airspeed:
                    raise ValueError("the player flies at settings.AIRSPEED (game.py:87)")
                return
            s["player"][i][{"x": 0, "y": 1, "psi": 2}[name]] = value
        else:
            s["traffic"][i][j][{"x": 0, "y": 1, "v_air": 2, "psi": 3}[name]] = value
        self._core.inject_state(s["player"], s["traffic"], s["steps"], s["total_reward"])

    @property
    def steps(self) -> int:
        return self._row()[2]

    @steps.setter
    def steps(self, value: int):
        _, _, _, _, s = self._row()
        s["steps"][self._i] = int(value)
        self._core.inject_state(s["player"], s["traffic"], s["steps"], s["total_reward"])

    @property
    def total_reward(self) -> float:
        return self._row()[3]

    @total_reward.setter
    def total_reward(self, value: float):
        _, _, _, _, s = self._row()
        s["total_reward"][self._i] = float(value)
        self._core.inject_state(s["player"], s["traffic"], s["steps"], s["total_reward"])

    @property
    def d_path(self) -> float:
        """Path length: AIRSPEED/FPS px per step() call at constant speed (game.py:241)."""
        p = self._core.params
        return self._step_calls * (p.airspeed / p.fps)

    # ---- geometry predicates the scripts / notebooks call (game.py:162-192)
    def minimum_separation(self):
        pl, tr, _, _, _ = self._row()
        return float(np.min(np.hypot(tr[:, 0] - pl[0], tr[:, 1] - pl[1])))

    def distance_to_goal(self):
        pl = self._row()[0]
        return float(np.hypot(pl[0] - self.goal_x, pl[1] - self.goal_y))

    def check_timeout(self):
        return self.steps > self._core.params.max_steps

    def detect_collisions(self):
        return self.minimum_separation() < 2 * self._core.params.collision_radius

    def check_goal(self):
        return self.distance_to_goal() < self._core.params.goal_radius
