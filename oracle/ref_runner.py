"""TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE -- the CPU arm of ``bench.py``.

Times the UNMODIFIED reference (``ACAS2DEnv.reset/step``, reference gym_ACAS2D/envs/environment.py:29-48)
on this machine's host cores: one env per process, random U(-1,1) actions, resets included, stdout
swallowed (the reference prints on every episode end, game.py:311-313).  ``gym`` / ``pygame`` are the
stand-ins of ``oracle/ref_shim.py`` -- so ``clock.tick(FPS)`` (environment.py:31, the 100 steps/s cap of
the shipped env) is a no-op and ``pygame.init`` / image loads cost nothing; everything else is the
reference's own code.

Before anything is timed the reference must pass the golden gate: ``random.seed(13)``, two discarded
games, zero-action episodes (the recipe that reproduces the reference's own
``models/logs/baseline_ACAS2D_PPO_11_100.csv``), compared with the committed condensed
fixture ``tests/golden/baseline_zero_action.npz`` -- outcome and time steps equal, last path point
bit-identical.

Runs in its own interpreter (``python -m oracle.ref_runner ...``) because the reference package has the
same name as the product package (``gym_ACAS2D``); prints one JSON object.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402


def golden_gate(n_episodes: int = 6) -> dict:
    """Replay the first ``n_episodes`` of the reference's golden CSV with the reference itself."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "baseline_zero_action.npz"))
    ref_shim.load(1)
    from gym_ACAS2D.envs.game import ACAS2DGame
    random.seed(13)
    bad = []
    with ref_shim.quiet():
        for _ in range(2):
            ACAS2DGame()
        for ep in range(n_episodes):
            game = ACAS2DGame()
            game.observe()
            done = False
            while not done:
                game.action(np.array([0]))
                game.observe()
                game.evaluate()
                done = game.is_done()
            last = np.array(game.path[-1], dtype=np.float64)
            ok = (int(game.outcome) == int(g["outcome"][ep]) and int(game.steps) == int(g["time_steps"][ep])
                  and np.array_equal(last, g["path_last"][ep]))
            if not ok:
                bad.append(ep)
    return {"episodes": n_episodes, "ok": not bad, "mismatched": bad}


def _worker(args):
    n_steps, seed, n_traffic = args
    import numpy as np
    ref_shim.load(n_traffic)
    from gym_ACAS2D.envs.environment import ACAS2DEnv
    random.seed(seed)
    arng = np.random.default_rng(seed)
    with ref_shim.quiet():
        env = ACAS2DEnv()
        env.reset()
        warm = max(10, n_steps // 20)
        acts = arng.uniform(-1, 1, size=(n_steps + warm, 1))
        episodes = 0
        for k in range(warm):
            if env.step(acts[k])[2]:
                env.reset()
        t0 = time.perf_counter()
        for k in range(warm, warm + n_steps):
            if env.step(acts[k])[2]:
                env.reset()
                episodes += 1
        dt = time.perf_counter() - t0
    return dt, episodes


def random_rollout(procs: int, steps_per_proc: int, n_traffic: int = 1) -> dict:
    jobs = [(steps_per_proc, 13 + 17 * i, n_traffic) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    return {"rate": procs * steps_per_proc / slowest, "procs": procs, "steps_per_proc": steps_per_proc,
            "n_traffic": n_traffic, "slowest_s": slowest, "wall_s": wall, "episodes": sum(r[1] for r in res)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000, help="env steps per process")
    ap.add_argument("--n-traffic", type=int, default=1)
    ap.add_argument("--gate", type=int, default=6, help="golden-CSV episodes replayed first (0 = skip)")
    args = ap.parse_args()
    if not ref_shim.available():
        print(json.dumps({"available": False, "why": "no reference tree and no oracle/_ref copy"}))
        return
    out = {"available": True, "root": ref_shim.REFERENCE_ROOT, "cpu_count": os.cpu_count()}
    if args.gate:
        out["gate"] = golden_gate(args.gate)
        if not out["gate"]["ok"]:
            print(json.dumps(out))
            return
    out.update(random_rollout(args.procs, args.steps, args.n_traffic))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
