"""Episode-record export in the reference's CSV layout.

The reference keeps per-step Python lists on ``ACAS2DGame`` (``envs/game.py:45-75``, appended in
``action`` ``:231-239`` and ``evaluate`` ``:266-276``) and its scripts dump them with pandas:
``baseline_main.py:62-74`` (Episode, Outcome, Total Reward, Time Steps, Path, Traffic Paths) and
``testing_main.py:113-138`` (+ Path Length, psi, d_sep, a_lat, d_goal, delta_heading, v_closing, d_cpa,
d_dev, r_d_goal, r_h_goal, r_d_cpa, r_d_dev, r_step).  ``record_episodes`` produces the same table for
E episodes stepped side by side on the GPU (one env per episode).  The per-step rows are written ON THE
DEVICE into ring buffers (``BatchedACAS2D.enable_trace`` -> ``acas2d_trace_step``: positions / headings are
the float64 device state, the diagnostics the step's own float32 quantities, the reward decomposition the
kernel's ``reward_terms``); the host reads them back ONCE, when every episode has ended.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from gym_ACAS2D.envs import _native
from gym_ACAS2D.envs.batched import BatchedACAS2D
from gym_ACAS2D.settings import OUTCOME_NAMES

BASELINE_COLUMNS = ["Episode", "Outcome", "Total Reward", "Time Steps", "Path", "Traffic Paths"]
TESTING_COLUMNS = BASELINE_COLUMNS[:4] + ["Path Length", "Path", "Traffic Paths", "psi", "d_sep", "a_lat", "d_goal",
                                           "delta_heading", "v_closing", "d_cpa", "d_dev", "r_d_goal", "r_h_goal",
                                           "r_d_cpa", "r_d_dev", "r_step"]
_KEYS = TESTING_COLUMNS[7:]


def record_episodes(env: BatchedACAS2D, policy: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                    start: Optional[Tuple[np.ndarray, np.ndarray]] = None, max_steps: Optional[int] = None) -> List[Dict]:
    """One episode per env of ``env`` (built with ``auto_reset=False``); returns one dict per episode
    keyed by the reference's column names.  ``policy(obs[B, L]) -> actions[B]`` works on device tensors
    (default: ``baseline_main.py``'s zero action).  ``start = (player[B,3] = x, y, psi,
    traffic[B,N,4] = x, y, v_air, psi)`` replaces the Philox spawn, e.g. to replay reference spawns."""
    if env.auto_reset:
        raise ValueError("record_episodes needs an env built with auto_reset=False")
    p, B, N, dev = env.params, env.num_envs, env.n_traffic, env.device
    horizon = int(max_steps or p.max_steps)
    env.reset()
    if start is not None:
        env.inject_state(start[0], start[1])
    obs = env.observe() if start is not None else env.obs
    nrec = min(N, _native.TRACE_MAX_TRAFFIC)
    env.enable_trace(B, 0, capacity=horizon + 2, n_traffic_rec=nrec)
    try:
        zero = torch.zeros(B, device=dev)
        finished = torch.zeros(B, dtype=torch.bool, device=dev)
        outcome = torch.zeros(B, dtype=torch.uint8, device=dev)
        total = torch.zeros(B, dtype=torch.float32, device=dev)
        steps = torch.ones(B, dtype=torch.int32, device=dev)
        count_at_end = torch.zeros(B, dtype=torch.int32, device=dev)
        for t in range(horizon):
            act = zero if policy is None else policy(obs).to(dev, torch.float32).reshape(B)
            obs, rew, done = env.step(act)
            newly = done & ~finished                      # a finished game is stepped on (harmlessly): its rows are cut below
            outcome = torch.where(newly, env.outcome, outcome)
            total = torch.where(newly, env.ep_return, total)
            steps = torch.where(newly, env.ep_length, steps)
            count_at_end = torch.where(newly, env._trace_cursor, count_at_end)
            finished |= done
            if t % 64 == 63 and bool(finished.all()):     # one host sync per 64 steps
                break
        rows_d, count = env.trace_rows()
        count = np.where(finished.cpu().numpy(), count_at_end.cpu().numpy(), count)
        outcome, total, steps = outcome.cpu().numpy(), total.cpu().numpy(), steps.cpu().numpy()
    finally:
        env.disable_trace()
    F = {k: i for i, k in enumerate(_native.TRACE_FIELDS)}
    D = _native.TRACE_DOUBLES
    rows = []
    for b in range(B):
        r = rows_d[b, : int(count[b])]                    # row 0 = the game's initial records, then one per step
        row = {"Episode": b + 1, "Outcome": OUTCOME_NAMES.get(int(outcome[b]), "Unfinished"), "Total Reward": float(total[b]),
               "Time Steps": int(steps[b]), "Path Length": float((len(r) - 1) * (p.airspeed / p.fps)),       # game.py:241
               "Path": [(float(x), float(y)) for x, y in r[:, :2]],                   # plain floats: the CSV must literal_eval
               "Traffic Paths": [[(float(x), float(y)) for x, y in r[:, D + 2 * n: D + 2 * n + 2]] for n in range(nrec)]}
        row.update({k: r[:, F[k]].tolist() for k in _KEYS})
        rows.append(row)
    return rows


def to_csv(rows: List[Dict], path: str, columns: Optional[List[str]] = None) -> None:
    """``log_df.to_csv(file, index=False)`` of the reference scripts; ``columns=BASELINE_COLUMNS`` gives
    ``baseline_main.py``'s six columns, the default ``testing_main.py``'s twenty."""
    import pandas as pd
    pd.DataFrame(rows)[columns or TESTING_COLUMNS].to_csv(path, index=False)
