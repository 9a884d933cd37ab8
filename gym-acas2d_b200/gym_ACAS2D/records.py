"""Episode-record export in the reference's CSV layout.

The reference keeps per-step Python lists on ``ACAS2DGame`` (``envs/game.py:45-75``, appended in
``action`` ``:231-239`` and ``evaluate`` ``:266-276``) and its scripts dump them with pandas:
``baseline_main.py:62-74`` (Episode, Outcome, Total Reward, Time Steps, Path, Traffic Paths) and
``testing_main.py:113-138`` (+ Path Length, psi, d_sep, a_lat, d_goal, delta_heading, v_closing, d_cpa,
d_dev, r_d_goal, r_h_goal, r_d_cpa, r_d_dev, r_step).  ``record_episodes`` produces the same table for
E episodes stepped side by side on the GPU (one env per episode): positions / headings are the float64
device state, the diagnostics are the step's own float32 observation entries de-normalised, and the
reward decomposition applies the reference's term formulas (``envs/rewards.py:5-50``) to them.  An
analysis tool next to the hot path -- it copies the state to the host every step, so use small batches.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from gym_ACAS2D.envs.batched import BatchedACAS2D
from gym_ACAS2D.settings import OUTCOME_NAMES

BASELINE_COLUMNS = ["Episode", "Outcome", "Total Reward", "Time Steps", "Path", "Traffic Paths"]
TESTING_COLUMNS = BASELINE_COLUMNS[:4] + ["Path Length", "Path", "Traffic Paths", "psi", "d_sep", "a_lat", "d_goal",
                                           "delta_heading", "v_closing", "d_cpa", "d_dev", "r_d_goal", "r_h_goal",
                                           "r_d_cpa", "r_d_dev", "r_step"]
_KEYS = TESTING_COLUMNS[7:]


def _reward_terms(p, psi, phi, d_cpa, d_goal, d_dev, v_c):
    """rewards.py:5-60 on float64 arrays -> (delta_heading, r_d_goal, r_h_goal, r_d_cpa, r_d_dev, step_reward_5)."""
    dh = np.minimum(np.abs(psi - phi), 360 - np.abs(psi - phi))
    r_h = (1 - dh / 180) ** 4
    with np.errstate(invalid="ignore"):
        r_cpa = np.where(v_c > 0, 1.0, np.fmin(1.0, (d_cpa / p.safe_distance) ** 4))
    d_goal_init = (p.width - p.goal_radius) - 2 * p.aircraft_size
    r_dev = np.sqrt(np.maximum(0.0, 1 - np.abs(d_dev) / (d_goal_init / 2)))
    r_goal = np.minimum(1.0, (1 - d_goal / (d_goal_init + (p.airspeed / p.fps) * p.max_steps)) ** 4)
    return dh, r_goal, r_h, r_cpa, r_dev, np.where(v_c <= 0, r_h * r_cpa * r_dev, r_h * r_goal)


def record_episodes(env: BatchedACAS2D, policy: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                    start: Optional[Tuple[np.ndarray, np.ndarray]] = None, max_steps: Optional[int] = None) -> List[Dict]:
    """One episode per env of ``env`` (built with ``auto_reset=False``); returns one dict per episode
    keyed by the reference's column names.  ``policy(obs[B, L]) -> actions[B]`` works on device tensors
    (default: ``baseline_main.py``'s zero action).  ``start = (player[B,3] = x, y, psi,
    traffic[B,N,4] = x, y, v_air, psi)`` replaces the Philox spawn, e.g. to replay reference spawns."""
    if env.auto_reset:
        raise ValueError("record_episodes needs an env built with auto_reset=False")
    p, B, N, dev = env.params, env.num_envs, env.n_traffic, env.device
    env.reset()
    if start is not None:
        env.inject_state(start[0], start[1])
    obs = env.observe() if start is not None else env.obs
    st = env.extract_state()
    pl, tr = st["player"], st["traffic"]
    path = [[(float(pl[b, 0]), float(pl[b, 1]))] for b in range(B)]                 # plain floats: the CSV must literal_eval
    tpaths = [[[(float(tr[b, n, 0]), float(tr[b, n, 1]))] for n in range(N)] for b in range(B)]
    rec = {k: [[] for _ in range(B)] for k in _KEYS}

    def push(alive, o, psi, d_sep, a_lat, discount):
        o = o.astype(np.float64)
        d_goal, phi, d_dev = o[:, 3] * p.d_goal_max, o[:, 4] * 360, o[:, 2] * p.d_dev_max
        d_cpa, v_c = o[:, 6] * p.d_cpa_max, o[:, 7] * p.v_closing_max              # intruder 0 only (Q7)
        dh, r_goal, r_h, r_cpa, r_dev, r5 = _reward_terms(p, psi, phi, d_cpa, d_goal, d_dev, v_c)
        vals = dict(psi=psi, d_sep=d_sep, a_lat=a_lat, d_goal=d_goal, delta_heading=dh, v_closing=v_c, d_cpa=d_cpa,
                    d_dev=d_dev, r_d_goal=r_goal, r_h_goal=r_h, r_d_cpa=r_cpa, r_d_dev=r_dev, r_step=r5 * discount)
        for b in np.flatnonzero(alive):
            for k in _KEYS:
                rec[k][b].append(float(vals[k][b]))

    alive = np.ones(B, bool)
    sep = lambda player, traffic: np.hypot(traffic[:, :, 0] - player[:, None, 0], traffic[:, :, 1] - player[:, None, 1]).min(1)  # noqa: E731
    push(alive, obs.cpu().numpy(), pl[:, 2], sep(pl, tr), np.zeros(B), np.ones(B))   # game.py:132-160: initial rows, no discount
    outcome = np.zeros(B, int); total = np.zeros(B); steps = np.ones(B, int); d_path = np.zeros(B)
    zero = torch.zeros(B, device=dev)
    for _ in range(int(max_steps or p.max_steps)):
        act = zero if policy is None else policy(obs).to(dev, torch.float32).reshape(B)
        tr_before = tr
        obs, rew, done = env.step(act)
        ex = env.extract_state()
        pl, tr = ex["player"], ex["traffic"]
        for b in np.flatnonzero(alive):
            path[b].append((float(pl[b, 0]), float(pl[b, 1])))
            for n in range(N):
                tpaths[b][n].append((float(tr_before[b, n, 0]), float(tr_before[b, n, 1])))   # before the intruders move (Q10)
        push(alive, obs.cpu().numpy(), pl[:, 2], sep(pl, tr_before), act.cpu().numpy().astype(np.float64) * p.acc_lat_limit,
             1 - ex["steps"] / p.max_steps)
        d_path += np.where(alive, p.airspeed / p.fps, 0.0)                           # game.py:241
        dn = done.cpu().numpy() & alive
        if dn.any():
            outcome[dn] = env.outcome.cpu().numpy()[dn]
            total[dn] = env.ep_return.cpu().numpy()[dn]
            steps[dn] = env.ep_length.cpu().numpy()[dn]
        alive &= ~dn
        if not alive.any():
            break
    rows = []
    for b in range(B):
        row = {"Episode": b + 1, "Outcome": OUTCOME_NAMES.get(int(outcome[b]), "Unfinished"), "Total Reward": float(total[b]),
               "Time Steps": int(steps[b]), "Path Length": float(d_path[b]), "Path": path[b], "Traffic Paths": tpaths[b]}
        row.update({k: rec[k][b] for k in _KEYS})
        rows.append(row)
    return rows


def to_csv(rows: List[Dict], path: str, columns: Optional[List[str]] = None) -> None:
    """``log_df.to_csv(file, index=False)`` of the reference scripts; ``columns=BASELINE_COLUMNS`` gives
    ``baseline_main.py``'s six columns, the default ``testing_main.py``'s twenty."""
    import pandas as pd
    pd.DataFrame(rows)[columns or TESTING_COLUMNS].to_csv(path, index=False)
