"""Episode statistics of a sharded env batch.

The reference keeps per-episode records on the ``ACAS2DGame`` object and its scripts harvest
``outcome, total_reward, steps, ...`` at the end of every episode (game.py:28-38,
testing_main.py:84-105).  The batched step accumulates the same quantities on the device as
integer counters (``acas2d_b200.h`` ACAS2D_STAT_*); this module turns them into a summary and,
across GPUs, reduces them with the ONLY collective on this path: one all-reduce(SUM) of seven
int64 values per reporting interval (NCCL over NVLink on GPUs, gloo in the CPU tests).
Integer sums make the result independent of the reduction order and of the GPU count.
"""
from __future__ import annotations

from typing import Dict

import torch

from ._native import STAT_FX_SCALE, STAT_NAMES


def reduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """All-reduce(SUM) the int64 counter vector over the default process group, if any."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        counters = counters.clone()
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def summarise(counters: torch.Tensor, reduce: bool = True, track_min_sep: bool = False) -> Dict[str, float]:
    if reduce:
        counters = reduce_counters(counters)
    c = dict(zip(STAT_NAMES, (int(v) for v in counters.tolist())))
    n = max(c["episodes"], 1)
    out = dict(
        episodes=c["episodes"], goal=c["goal"], collision=c["collision"], timeout=c["timeout"],
        goal_rate=c["goal"] / n, collision_rate=c["collision"] / n, timeout_rate=c["timeout"] / n,
        mean_length=c["length_sum"] / n,
        mean_return=c["return_fx"] / STAT_FX_SCALE / n,
        # d_path: the player covers AIRSPEED/FPS px per step() call and an episode with
        # game.steps == s had s-1 calls (game.py:241, Q5)
        mean_step_calls=(c["length_sum"] - c["episodes"]) / n,
    )
    if track_min_sep:
        out["mean_min_separation"] = c["min_sep_fx"] / STAT_FX_SCALE / n
    return out
