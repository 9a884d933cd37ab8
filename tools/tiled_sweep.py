#!/usr/bin/env python
"""Development tool: time the tiled kernel (N_TRAFFIC > 1) over record format / lanes-per-env choices.
    python tools/tiled_sweep.py [--points 8:65536,64:65536,...] [--kin 0,1] [--per-lane 8,4]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-acas2d_b200")]
import torch
import bench
from gym_ACAS2D.envs import _native

ap = argparse.ArgumentParser()
ap.add_argument("--points", default="8:65536,64:65536,256:65536,8:1048576,64:262144")
ap.add_argument("--kin", default="0,1")
ap.add_argument("--per-lane", default="8")
ap.add_argument("--steps", type=int, default=128)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _native.load()
peak, _ = bench.measured_peak()


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(400000)
    e0.record()
    for k in range(iters):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for pt in args.points.split(","):
    n, b = (int(x) for x in pt.split(":"))
    for kin in (int(x) for x in args.kin.split(",")):
        for pl in (int(x) for x in args.per_lane.split(",")):
            lib.acas2d_set_tiled_tuning(kin, pl)
            r = bench.sweep_point(dev, n, b, args.steps, timed, 1, peak)
            print(json.dumps({"N": n, "B": b, "kin": kin, "per_lane": pl, "us_per_step": round(1e3 * r["ms_per_step"], 2),
                              "frac": round(r["roofline"]["frac"], 3), "episodes": r["episode_stats"].get("episodes"),
                              "mean_length": round(r["episode_stats"].get("mean_length", 0), 1)}), flush=True)
