#!/usr/bin/env python
"""ACAS-2D environment-step benchmark (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE pass of the hot path over one batch: ``step(actions[B])`` for all the envs a GPU
owns.  Workload (``config.workload``): BASELINE.json config 3's random-action rollout at default
N_TRAFFIC with auto-reset, at a per-GPU batch that saturates HBM (4 Mi envs / GPU -- SURVEY 8d:
131 072 envs / GPU would be launch-bound), sharded by global env id, weak scaling.

Prints ONE JSON line on rank 0.  ``value`` is env-steps/s with everything resident in HBM;
``e2e`` is the same metric through the host-buffer C-ABI call (H2D actions, D2H obs / reward /
done inside the timed region); ``roofline`` is algorithmic bytes / CUDA-event time against the
measured HBM peak; ``cpu_baseline`` times the oracle port on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "gym-acas2d_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "env-steps/sec (whole box, device-timed)"
UNIT = "env-steps/s"


def shard(total: int, world: int, rank: int):
    """Contiguous block of global env ids owned by ``rank``: (offset, count)."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def algorithmic_bytes(n_traffic: int) -> int:
    """A(N) = 69 + 32 N bytes per env-step (fp32 SoA, state read + written every step)."""
    return 69 + 32 * n_traffic


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(n_traffic: int, envs: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json), scaled to this run's envs per launch."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[f"n{n_traffic}"]
        return d["bytes_per_launch"] * envs / d["envs"]
    except Exception:  # noqa: BLE001
        return None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device_index: int, period_s: float = 0.004):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period_s
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU baselines
def _pyport_worker(args):
    n_steps, seed = args
    from oracle.acas2d_oracle import pyport_random_rollout
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        pyport_random_rollout(max(1, n_steps // 10), seed)          # warm-up
        t0 = time.perf_counter()
        pyport_random_rollout(n_steps, seed + 1)
        return time.perf_counter() - t0


def time_pyport(cores: int, steps_per_core: int):
    """Reference-shaped scalar Python port (oracle/acas2d_oracle.py PyPortGame), one env per
    process, random actions, resets included.  Returns env-steps/s summed over ``cores``."""
    import multiprocessing as mp
    if cores == 1:
        dt = _pyport_worker((steps_per_core, 13))
        return steps_per_core / dt
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.map(_pyport_worker, [(steps_per_core, 13 + 17 * i) for i in range(cores)])
    return cores * steps_per_core / max(times)


def time_c_oracle(envs: int, steps: int):
    """The C oracle (float64, one thread), auto-resetting random-action steps."""
    import numpy as np
    from oracle.acas2d_oracle import Oracle
    orc = Oracle(1)
    st = orc.new_state(envs)
    orc.spawn_philox(st, 13, 0)
    orc.observe(st)
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (steps, envs))
    orc.vec_step(st, acts[0], 13, 0)
    t0 = time.perf_counter()
    for k in range(steps):
        orc.vec_step(st, acts[k], 13, 0)
    return envs * steps / (time.perf_counter() - t0)


def run_ref_runner(procs: int, steps_per_proc: int, n_traffic: int = 1, gate: int = 6, timeout: float = 600.0):
    """The UNMODIFIED reference (oracle/_ref or /root/reference) in its own interpreter: the package name
    clashes with the product's, and the timed processes must not share this one's CUDA context.  Returns the
    runner's JSON dict, or None when no copy of the reference is present."""
    import subprocess
    cmd = [sys.executable, "-m", "oracle.ref_runner", "--procs", str(procs), "--steps", str(steps_per_proc),
           "--n-traffic", str(n_traffic), "--gate", str(gate)]
    try:
        out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, check=True).stdout
        d = json.loads(out.strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        return {"available": False, "why": repr(exc)[:200]}
    return d


REF_NOTE = ("the unmodified reference ACAS2DEnv.reset/step under gym/pygame stand-ins (clock.tick is a no-op: as shipped "
            "it is wall-clock capped at 100 steps/s, environment.py:31, and logged 69-89 steps/s, BASELINE.md), "
            "golden-CSV gate passed first")


def cpu_baseline_block(budget_s: float = 14.0):
    """rank 0, N=1: a bounded sample (~15-20 s of CPU work) of the same workload on ONE host core."""
    probe = run_ref_runner(1, 3000, gate=6)
    if probe.get("available") and probe.get("gate", {}).get("ok") and "rate" in probe:
        steps = int(min(max(probe["rate"] * budget_s, 5000), 200000))
        d = run_ref_runner(1, steps, gate=0)
        rate1, kind = d["rate"], "reference"
        sample = (f"{REF_NOTE}; 1 process x 1 env x {steps} random-action steps incl. {d['episodes']} resets, "
                  f"N_TRAFFIC=1, source {os.path.relpath(d['root'], ROOT) if d['root'].startswith(ROOT) else d['root']}")
    else:                                               # no copy of the reference on this machine: the oracle's port
        steps = 4000
        rate1 = time_pyport(1, steps)
        steps = int(min(max(rate1 * budget_s * 0.5, 2000), 200000))
        rate1, kind = time_pyport(1, steps), "port"
        sample = (f"oracle Python port (reference-shaped scalar float64 code; reference copy unavailable: "
                  f"{probe.get('why', probe.get('gate'))}), 1 env, {steps} random-action steps incl. resets, N_TRAFFIC=1")
    c_rate = time_c_oracle(8192, 64)
    return {"value": rate1, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
            "c_oracle_1core": {"value": c_rate, "sample": "oracle C restatement, 8192 envs x 64 auto-reset steps, 1 thread"}}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on every host core of this box
    (one env per process; oracle/_ref = the reference's files, unmodified; falls back to the oracle's Python
    port only when no copy of the reference is present)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    K_eff = min(K, 400)                              # bounded: the whole run ends within minutes
    per_step = max(256, -(-40000 // K_eff))          # env-steps per core per bench "step": >= 40 000 steps (~3 s) per core
    t0 = time.perf_counter()
    d = run_ref_runner(cores, per_step * K_eff, n_traffic=args.n_traffic, gate=6)
    if d.get("available") and d.get("gate", {}).get("ok") and "rate" in d:
        rate, kind = d["rate"], "reference"
        sample = (f"{REF_NOTE}; {cores} processes x 1 env x {per_step * K_eff} random-action steps incl. resets "
                  f"({d['episodes']} episodes), N_TRAFFIC={args.n_traffic}; steps capped at {K_eff} of the requested {K}")
    else:
        rate, kind = time_pyport(cores, per_step * K_eff), "port"
        sample = (f"{cores} processes x 1 env x {per_step * K_eff} random-action steps incl. resets (oracle Python port: "
                  f"no reference copy here, {d.get('why', d.get('gate'))}); steps capped at {K_eff} of the requested {K}")
    wall = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * cores * per_step / rate,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"random-action rollout, N_TRAFFIC={args.n_traffic}, resets included, one env per host core",
                       "envs": cores, "env_steps_per_step": cores * per_step},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(device_index: int):
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host buffer is allocated, so the
    end-to-end path's H2D / D2H copies do not cross sockets.  Returns what happened (the line reports it: on
    a single-node VM sysfs says -1 and there is nothing to bind to)."""
    info = {"node": None, "bound": False, "why": None, "host_nodes": None}
    try:
        import torch
        try:
            info["host_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        except OSError:
            info["host_nodes"] = None
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        path = f"/sys/bus/pci/devices/{bus}/numa_node"
        if not os.path.exists(path):
            info["why"] = f"{path} does not exist (virtualised PCI topology)"
            return info
        node = int(open(path).read())
        info["node"] = node
        if node < 0:
            info["why"] = "sysfs numa_node = -1: the platform exposes no NUMA affinity for this GPU (single-node VM)"
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
            info["why"] = f"bound to {len(cpus)} CPUs of node {node} before pinned allocations"
        else:
            info["why"] = "no allowed CPU on the GPU's node"
    except Exception as exc:  # noqa: BLE001
        info["why"] = repr(exc)[:160]
    return info


# ------------------------------------------------------------------------------------------ GPU arm
def age_batch(env, steps: int) -> int:
    """Bring a freshly reset batch to its steady state before anything is timed: `steps` auto-resetting
    random-action steps (not timed), so that the batch holds episodes of every age and some envs finish and
    respawn on EVERY step of the timed window, whatever --steps is.  N_TRAFFIC == 1: the fused K-step rollout
    kernel (in-kernel Philox actions); N > 1: plain step calls."""
    import torch
    if steps <= 0:
        return 0
    # all games start together at reset, and a quarter of the random-action games run into the 1000-step timeout: left
    # alone they keep ending in WAVES 1000 steps apart.  Stagger the step counters once (the games then time out
    # at different moments) and let the ageing steps mix the rest.
    # (Only where the ageing is long enough to play these first, displaced games out: N_TRAFFIC == 1.  With more
    # intruders episodes last a few steps and end by collision -- there are no timeout waves to break up.)
    if env.n_traffic == 1 and steps >= 1500:
        words = env.paux.view(torch.int32).view(env.num_envs, 4)[:, 2]               # paux.steps (flag bits above bit 27)
        gid = torch.arange(env.num_envs, device=env.device, dtype=torch.int64) + env.env_id_offset
        words.add_(((gid * 7919) % 997).to(torch.int32))
    if env.n_traffic == 1:
        env.rollout_random(steps, action_seed=77, step0=0)
        env.observe()                                   # the obs buffer of the aged state (a_lat taken as 0)
    else:
        a = torch.empty(env.num_envs, dtype=torch.float32, device=env.device)
        for k in range(steps):
            env.random_actions(k, action_seed=77, out=a)
            env.step(a, full_outputs=False)
    torch.cuda.synchronize(env.device)
    return steps


def sweep_point(dev, N: int, B: int, steps: int, timed, world: int, peak: float, offset: int = 0):
    """BASELINE config 4: one point of the traffic sweep (tiled kernel), CUDA-graph replays, steady state."""
    import torch
    from gym_ACAS2D.envs import BatchedACAS2D
    env = BatchedACAS2D(B, n_traffic=N, device=dev, seed=13, env_id_offset=offset, auto_reset=True)
    env.reset()
    KA = 4
    actions = torch.empty(KA, B, dtype=torch.float32, device=dev)
    for k in range(KA):
        env.random_actions(k, action_seed=2024, out=actions[k])
    age = age_batch(env, 48 if N >= 64 else 200)
    GK = 16
    graph = env.capture_steps(actions, full_outputs=False, num_steps=GK, warmup=True)
    graph.replay()
    reps = max(1, steps // GK)
    env.clear_stats()
    ms = timed(lambda k: graph.replay(), reps)
    stats = env.episode_stats(reduce=True)
    n_steps = reps * GK
    A = algorithmic_bytes(N)
    achieved = A * B * n_steps / (ms * 1e-3) / 1e9
    out = {"value": world * B * n_steps / (ms * 1e-3), "unit": UNIT, "envs_per_gpu": B, "n_traffic": N, "steps": n_steps,
           "aged_steps": age, "ms_per_step": ms / n_steps,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_env_step": A,
                        "state_and_outputs_mb": B * (16 * N + 32 + 4 * (5 + 3 * N) + 5) / 1e6},
           "episode_stats": {k: stats[k] for k in ("episodes", "mean_length", "goal_rate", "collision_rate", "timeout_rate") if k in stats}}
    del graph, env, actions
    torch.cuda.empty_cache()
    return out


def pcie_copy_leg(env, h_actions, timed, iters: int):
    """The ceiling of the end-to-end path: the same bytes between the same pinned host buffers and the same
    device buffers (H2D actions, D2H obs + reward + done), two streams, NO kernel; every rank at once."""
    import torch
    hb = env._host
    up, down = torch.cuda.Stream(env.device), torch.cuda.Stream(env.device)

    def copies(k):
        cur = torch.cuda.current_stream(env.device)
        up.wait_stream(cur); down.wait_stream(cur)
        with torch.cuda.stream(up):
            env._actions_dev.copy_(h_actions[k % h_actions.shape[0]], non_blocking=True)
        with torch.cuda.stream(down):
            hb["obs"].copy_(env.obs, non_blocking=True)
            hb["reward"].copy_(env.reward, non_blocking=True)
            hb["done"].copy_(env.done_u8, non_blocking=True)
        cur.wait_stream(up); cur.wait_stream(down)
        cur.synchronize()                               # the e2e call returns synchronised too

    for k in range(2):
        copies(k)
    return timed(copies, iters)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gym_ACAS2D.envs import BatchedACAS2D, _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the ACAS-2D step has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the ONE JSON line, nothing else: NCCL honours NCCL_DEBUG_FILE only above the VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    N = args.n_traffic
    L = 5 + 3 * N
    if args.total_envs:
        offset, B = shard(args.total_envs, world, rank)
        scaling = "strong"
    else:
        B = args.envs_per_gpu
        offset = rank * B
        scaling = "weak"
    K, W = args.steps, args.warmup
    lib = _native.load()
    if args.n1_occ:
        lib.acas2d_set_tuning(args.n1_occ, -1)
    lib.acas2d_set_n1_kernel(args.n1_tma, args.n1_stages)

    env = BatchedACAS2D(B, n_traffic=N, device=dev, seed=13, env_id_offset=offset, auto_reset=True)
    env.reset()
    age_steps = age_batch(env, args.age_steps)          # steady state: episodes of every age, resets on every step
    KA = 8
    actions = torch.empty(KA, B, dtype=torch.float32, device=dev)
    for k in range(KA):
        env.random_actions(k, action_seed=2024, out=actions[k])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters, prime=None):
        """CUDA-event time of ``iters`` calls of fn on the current stream, max over ranks (ms).  ``prime``: untimed
        work of the same kind queued right before the start event (after the barrier / synchronize), so that the
        timed region starts on a busy device instead of one that has just idled through a synchronisation."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # prime the launch queue: a ~0.2 ms spin kernel runs while the host enqueues the start event and the first
        # launches behind it, so the start event's timestamp is taken when the timed work is already queued -- the
        # region measures the device executing the K steps, not the host's latency to submit the first of them
        # (a CUDA-graph launch costs the host tens of microseconds: 3-4 % of a 20-step window)
        torch.cuda._sleep(400000)
        if prime is not None:
            prime()
        e0.record()
        for k in range(iters):
            fn(k)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput (the `value`): CUDA-graph replay of the step loop (the Python /
    #      ctypes launch path is timed separately as "eager")
    step_fn = lambda k: env.step(actions[k % KA], full_outputs=False)   # noqa: E731
    for k in range(W):
        step_fn(k)
    GK = min(K, 8 * KA)                                 # steps per graph replay
    replays, rest = divmod(K, GK)                       # exactly K steps: `replays` replays + `rest` eager steps
    if K < 8 * KA:
        # a short window: a graph launch costs ~40 us of device-side start-up, 3 % of 20 steps, while consecutive
        # eager launches of the step kernel chain through programmatic dependent launch without a gap
        replays, rest = 0, K
    graph = env.capture_steps(actions, full_outputs=False, num_steps=GK, warmup=False)
    graph.replay()

    def k_steps(_):
        for _r in range(replays):
            graph.replay()
        for k in range(rest):
            step_fn(k)

    clocks = ClockSampler(local)
    clocks.__enter__()                                  # sampled across the graph, eager and end-to-end regions
    env.clear_stats()
    def prime_steps():
        for k in range(4):
            step_fn(k)
        env.clear_stats()                               # the episode counters cover the timed window only

    ms = timed(k_steps, 1, prime=prime_steps)
    stats = env.episode_stats(reduce=True)              # the one collective of the path (NCCL): episodes that ended
    #                                                     inside the K timed steps
    if not stats["episodes"] > 0:
        raise SystemExit("bench: no episode finished inside the timed window -- the workload is not in steady state")
    launches = K                                         # one step kernel per env step (graph nodes + eager)
    value = world * B * K / (ms * 1e-3)
    peak, peak_src = measured_peak()
    A = algorithmic_bytes(N)
    achieved = A * B * K / (ms * 1e-3) / 1e9            # GB/s per GPU (max-over-ranks time)
    launches0 = lib.acas2d_launch_count()
    ms_eager = timed(step_fn, min(K, 400))
    eager_launches = lib.acas2d_launch_count() - launches0

    # ---- end to end through the host-buffer C-ABI call: actions start in pinned host memory, obs /
    #      reward / done end in pinned host memory, every step
    h_actions = actions[:KA].cpu().pin_memory()         # the steps' inputs live in pinned host memory
    Ke = max(3, min(args.e2e_steps, K))

    def e2e_step(k):
        env.step_host(h_actions[k % KA])                # H2D 4 B/env, kernel, D2H (4L+5) B/env, sync

    for k in range(3):
        e2e_step(k)
    ms_e2e = timed(e2e_step, Ke)
    e2e = world * B * Ke / (ms_e2e * 1e-3)
    e2e_bytes = (4 + 4 * L + 5) * B                     # per rank per step, both directions
    ms_copy = pcie_copy_leg(env, h_actions, timed, Ke)
    pcie_gbs = world * e2e_bytes * Ke / (ms_copy * 1e-3) / 1e9        # whole box, all ranks copying at once
    e2e_gbs = world * e2e_bytes * Ke / (ms_e2e * 1e-3) / 1e9
    clocks.__exit__(None, None, None)

    # ---- secondary workloads (reported, not the headline); each leg on its own so that one failure costs one key
    other = {}

    def leg(name, fn):
        try:
            fn()
        except Exception as exc:                       # noqa: BLE001
            other.setdefault("errors", {})[name] = repr(exc)[:300]
            torch.cuda.synchronize()

    def leg_fused_rollout():
        env.rollout_random(8, action_seed=1, step0=0)
        fused_k, reps = 64, 4
        ms_f = timed(lambda k: env.rollout_random(fused_k, action_seed=1, step0=100 + fused_k * k), reps)
        other["fused_rollout_64_steps_per_launch"] = {"value": world * B * fused_k * reps / (ms_f * 1e-3), "unit": UNIT,
                                                      "note": "in-kernel Philox actions, state in registers, no per-step outputs"}

    def leg_step_k():
        # open-loop K-step launches WITH every step's outputs (the caller holds the K actions): state traffic / K
        import ctypes
        ko, kr, kd = env.step_k(actions)
        kd8 = kd.view(torch.uint8)
        step_k_raw = lambda k: lib.acas2d_step_k(env._p(), env._s(), KA, actions.data_ptr(), ko.data_ptr(), kr.data_ptr(),   # noqa: E731
                                                 kd8.data_ptr(), ctypes.byref(env._aux_lean), torch.cuda.current_stream().cuda_stream)
        for k in range(2):
            step_k_raw(k)
        ms_k = timed(step_k_raw, 8)
        other["step_k_%d_steps_per_launch_with_outputs" % KA] = {
            "value": world * B * KA * 8 / (ms_k * 1e-3), "unit": UNIT, "us_per_env_step_batch": 1e3 * ms_k / (8 * KA),
            "note": "acas2d_step_k (C-ABI call, kernel only): obs / reward / done of every step written, state read and "
                    "written once per launch"}

    def leg_policy():
        # closed-loop rollout: the reference's trained actor (8-64-64-1 tanh MLP) fused with the env step
        from gym_ACAS2D.policy import MlpActor
        fixture = os.path.join(ROOT, "tests", "golden", "ppo_policy_1048576_11.npz")
        actor = MlpActor.from_file(fixture, dev) if os.path.exists(fixture) else MlpActor.random(0, dev)
        pol_steps = 16
        for k in range(3):
            env.policy_step(actor, deterministic=False, noise_seed=5, step_index=k, full_outputs=False)
        ms_p = timed(lambda k: env.policy_step(actor, deterministic=False, noise_seed=5, step_index=10 + k,
                                               full_outputs=False), pol_steps)
        other["policy_rollout_fused_mlp_env_step"] = {
            "value": world * B * pol_steps / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p / pol_steps,
            "note": "BASELINE config 5 inner loop: SB3 MlpPolicy actor (fp32, CUDA cores) + Gaussian noise + clip + env "
                    "step in one kernel per step; no host round trip"}
        for k in range(3):
            env.policy_step(actor, deterministic=False, noise_seed=5, step_index=k, full_outputs=False, tensor_cores=True)
        ms_t = timed(lambda k: env.policy_step(actor, deterministic=False, noise_seed=5, step_index=40 + k,
                                               full_outputs=False, tensor_cores=True), 4 * pol_steps)
        other["policy_rollout_tcgen05_mlp_env_step"] = {
            "value": world * B * 4 * pol_steps / (ms_t * 1e-3), "unit": UNIT, "ms_per_step": ms_t / (4 * pol_steps),
            "note": "same, hidden layers as tcgen05.mma kind::tf32 with TMEM accumulators (128 envs per CTA tile)"}

    def leg_vecenv():
        # the SB3-style numpy VecEnv surface (host arrays in / out, per-env info dicts) at 4096 envs, steady state
        from gym_ACAS2D.envs import ACAS2DVecEnv
        import numpy as np
        venv = ACAS2DVecEnv(4096, device=dev, seed=13)
        venv.reset()
        age_batch(venv.core, 1500)                      # episodes of every age: finished envs (info dicts) on most steps
        rng = np.random.default_rng(3)
        va = rng.uniform(-1, 1, (64, 4096, 1)).astype(np.float32)
        for k in range(20):
            venv.step(va[k % 64])
        n_it, fin = 1000, 0
        t0 = time.perf_counter()
        for k in range(n_it):
            fin += int(venv.step(va[k % 64])[2].sum())
        dt = time.perf_counter() - t0
        other["vecenv_numpy_surface_4096_envs"] = {
            "value": 4096 * n_it / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / n_it, "episodes_finished_per_step": fin / n_it,
            "note": "ACAS2DVecEnv.step(np actions) -> np obs / rewards / dones + list of info dicts (SB3 1.1.0 DummyVecEnv "
                    "semantics), random actions, steady state; wall clock, one process"}

    def leg_single_env():
        # the reference's own surface: gym ACAS2DEnv.step on ONE env (environment.py:29-42), wall clock
        from gym_ACAS2D.envs import ACAS2DEnv
        import numpy as np
        e1 = ACAS2DEnv(device=dev)
        e1.reset()
        rng = np.random.default_rng(5)
        acts = rng.uniform(-1, 1, (4096, 1))
        for k in range(200):
            if e1.step(acts[k])[2]:
                e1.reset()
        n_it = 4000
        t0 = time.perf_counter()
        for k in range(n_it):
            if e1.step(acts[k % 4096])[2]:
                e1.reset()
        dt = time.perf_counter() - t0
        other["gym_surface_single_env"] = {
            "value": n_it / dt, "unit": UNIT, "us_per_step": 1e6 * dt / n_it,
            "note": "ACAS2DEnv.step(action) -> (obs float64[8], reward, done, {}) on one env incl. resets, wall clock: "
                    "BASELINE config 1's counterpart through the drop-in gym surface (launch + sync bound)"}

    def leg_config2():
        small = BatchedACAS2D(4096, n_traffic=1, device=dev, seed=13, env_id_offset=0, auto_reset=True)
        small.reset()
        age_batch(small, 1500)
        sgraph = small.capture_steps(actions[:, :4096].contiguous(), num_steps=200)      # 200 steps per replay
        sgraph.replay()
        ms_s = timed(lambda k: sgraph.replay(), 10)
        other["config2_4096_envs_cuda_graph"] = {"value": world * 4096 * 200 * 10 / (ms_s * 1e-3), "unit": UNIT,
                                                 "note": "BASELINE config 2 batch: launch-bound, 200-step CUDA graph replay"}

    def leg_ppo_learner():
        # PPO learner (BASELINE config 5): one epoch of 32 minibatch gradient steps, this repo's
        # kernels vs the same arithmetic in torch autograd (both replayed from CUDA graphs)
        from gym_ACAS2D import ppo as _ppo
        n_l, mbs = 131072, 32
        data = [torch.rand(n_l, 8, device=dev) * 2 - 1, torch.randn(n_l, device=dev), -torch.rand(n_l, device=dev) - 0.5,
                torch.randn(n_l, device=dev), torch.randn(n_l, device=dev)]
        for name, cls in (("fused_kernels", _ppo.FusedLearner), ("torch_autograd_cuda_graph", _ppo.TorchLearner)):
            learner = cls(dev, None, None, cuda_graph=True)
            learner.bind(*data, mbs)
            learner.epoch()
            ms_l = timed(lambda k: learner.epoch(), 3)
            other["ppo_learner_" + name] = {
                "us_per_gradient_step": 1e3 * ms_l / (3 * mbs), "minibatch": n_l // mbs,
                "note": "advantage normalisation + forward/backward of actor and critic + grad-norm clip + Adam"}

    def leg_ppo_p2p():
        # N ranks: the data-parallel gradient exchange inside the update kernel over NVLink peer memory vs NCCL
        from gym_ACAS2D import ppo as _ppo
        n_l, mbs = 131072, 32
        g = torch.Generator(device=dev); g.manual_seed(100 + rank)
        data = [torch.rand(n_l, 8, device=dev, generator=g) * 2 - 1, torch.randn(n_l, device=dev, generator=g),
                -torch.rand(n_l, device=dev, generator=g) - 0.5, torch.randn(n_l, device=dev, generator=g),
                torch.randn(n_l, device=dev, generator=g)]
        res = {}
        for name in ("p2p", "nccl"):
            torch.manual_seed(1234)                      # the same initial parameters on every rank
            learner = _ppo.FusedLearner(dev, None, None, cuda_graph=True, exchange=name)
            learner.bind(*data, mbs)
            learner.epoch()
            ms_l = timed(lambda k: learner.epoch(), 3)
            res[name] = 1e3 * ms_l / (3 * mbs)
            if name == "p2p":
                blocks = [torch.empty_like(learner.params) for _ in range(world)]
                dist.all_gather(blocks, learner.params)
                res["params_identical_on_all_ranks"] = all(bool(torch.equal(blocks[0], x)) for x in blocks)
        other["ppo_p2p_step_us"] = {"value": res["p2p"], "nccl_allreduce_step_us": res["nccl"], "ranks": world,
                                    "params_identical_on_all_ranks": res.get("params_identical_on_all_ranks"),
                                    "minibatch_per_rank": n_l // mbs,
                                    "note": "one PPO gradient step = gradient kernel + update kernel; the update kernel sums "
                                            "every rank's gradient through NVLink peer memory (no NCCL call, whole epoch one "
                                            "CUDA graph) vs gradient -> NCCL all-reduce -> Adam"}

    def leg_config5():
        # BASELINE config 5: end-to-end PPO (rollout through the GPU VecEnv adapter + learner), SB3-default
        # hyper-parameters of the reference's model (training_main.py:44-52), every rank 1024 envs x 256 steps per iteration
        from gym_ACAS2D import ppo as _ppo
        t0 = time.perf_counter()
        hist = _ppo.train(num_envs=1024, n_steps=256, iterations=6, device=dev, seed=13, minibatches=64, n_epochs=10,
                          tensor_cores=True, cuda_graph=True, learner="fused", exchange="p2p", log=None)
        wall = time.perf_counter() - t0
        steady = hist[2:]                                   # the first iterations carry graph captures
        steps_it = 1024 * 256 * world
        roll = sum(h["rollout_s"] for h in steady) / len(steady)
        learn = sum(h["learn_s"] for h in steady) / len(steady)
        other["config5_ppo_vecenv"] = {
            "value": steps_it / (roll + learn), "unit": "env-steps/s (rollout + learning, whole job)", "ranks": world,
            "rollout_env_steps_per_s": steps_it / roll, "rollout_s_per_iteration": roll, "learn_s_per_iteration": learn,
            "gradient_steps_per_iteration": 640, "us_per_gradient_step": 1e6 * learn / 640, "wall_s_6_iterations": wall,
            "param_divergence_over_ranks": getattr(_ppo.train, "param_divergence", None),
            "note": "PPO from scratch through ACAS2DVecEnv.collect_rollout (fused tcgen05 actor + env step, one CUDA graph per "
                    "rollout) and the fused learner; N ranks: gradient exchange inside the update kernel over NVLink peer memory. "
                    "The reference's own run logged 69-89 env-steps/s (BASELINE.md)"}

    def leg_config3_strong():
        # BASELINE config 3 literally: 1 Mi envs split over the ranks (L2-resident shards: launch/latency-bound)
        total = 1 << 20
        off3, b3 = shard(total, world, rank)
        e3 = BatchedACAS2D(b3, n_traffic=1, device=dev, seed=13, env_id_offset=off3, auto_reset=True)
        e3.reset()
        age_batch(e3, 1500)
        a3 = torch.empty(KA, b3, dtype=torch.float32, device=dev)
        for k in range(KA):
            e3.random_actions(k, action_seed=2024, out=a3[k])
        g3 = e3.capture_steps(a3, full_outputs=False, num_steps=64, warmup=True)
        g3.replay()
        e3.clear_stats()
        ms3 = timed(lambda k: g3.replay(), 8)
        st3 = e3.episode_stats(reduce=True)
        ach = algorithmic_bytes(1) * b3 * 512 / (ms3 * 1e-3) / 1e9
        other["config3_1Mi_strong"] = {"value": total * 512 / (ms3 * 1e-3), "unit": UNIT, "total_envs": total, "envs_per_gpu": b3,
                                       "us_per_step": 1e3 * ms3 / 512, "scaling": "strong",
                                       "roofline_frac_per_gpu": ach / peak, "episodes": st3["episodes"],
                                       "note": "the shard's state + outputs sit in L2; the step is launch / latency-bound"}

    def leg_sweep():
        # BASELINE config 4: traffic sweep at its stated batch (65 536 envs) and at an HBM-saturating batch
        for n_t, b_t, st_t in ((8, 65536, 512), (64, 65536, 256), (256, 65536, 128),
                               (8, 1 << 20, 256), (64, 1 << 18, 128)):
            key = f"sweep_n{n_t}" if b_t == 65536 else f"sweep_n{n_t}_b{b_t}"
            other[key] = sweep_point(dev, n_t, b_t, st_t, timed, world, peak, offset=rank * b_t)

    if not args.skip_other:
        if N == 1:
            leg("fused_rollout", leg_fused_rollout)
            leg("step_k", leg_step_k)
            leg("policy", leg_policy)
            if world == 1:
                leg("config2", leg_config2)
                leg("vecenv", leg_vecenv)
                leg("single_env", leg_single_env)
                leg("ppo_learner", leg_ppo_learner)
        if world > 1:
            leg("config3_strong", leg_config3_strong)
            leg("ppo_p2p", leg_ppo_p2p)
        if N == 1:
            leg("config5", leg_config5)
        if world == 1 or args.sweep:
            leg("sweep", leg_sweep)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64 flag chain + f32 observation/reward", "data": "synthetic",
                "config": {"workload": "BASELINE config 3 random-action rollout at default N_TRAFFIC with auto-reset, "
                                       "HBM-saturating batch per GPU (SURVEY 8d), sharded by global env id; steady state: the "
                                       f"batch is aged by {age_steps} untimed random-action steps first, so episodes end and "
                                       "respawn on every timed step",
                           "envs_per_gpu": B, "total_envs": world * B, "n_traffic": N, "obs_dim": L,
                           "actions": f"pre-generated Philox U(-1,1) float32 [{KA},B] resident in HBM, cycled",
                           "timing": "CUDA events on the launching stream around exactly K steps, barrier + synchronize on both "
                                     "sides, max over ranks; the launch queue is primed (a 0.2 ms spin kernel + 4 untimed steps) before "
                                     "the start event: the window starts on a busy device with its first launch already queued",
                           "l2": "state + outputs per GPU = %.0f MB >> 126 MB L2 (no flush needed)" % (B * (A + 64) / 1e6),
                           "parallelism": f"env-sharded x{world}, one all-reduce of 7 int64 episode counters after the timed region"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(N, B), "peak_source": peak_src,
                             "algorithmic_bytes_per_env_step": A, "env_steps_per_launch": B,
                             "note": "per GPU; achieved = A(N) x envs per launch / CUDA-event time per launch"},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": (4 * L + 5) * B,
                        "steps": Ke, "ms_per_step": ms_e2e / Ke,
                        "path": "BatchedACAS2D.step_host -> acas2d_step_host (pinned host buffers)",
                        "achieved_gbs": e2e_gbs, "pcie_peak_gbs": pcie_gbs, "frac": e2e_gbs / pcie_gbs,
                        "pcie_peak_how": "same pinned host buffers, same device buffers, same bytes (H2D actions + D2H obs/reward/"
                                         "done) on two streams with NO kernel, all ranks at once, max over ranks; whole-box GB/s",
                        "numa": numa},
                "gpu_launches": int(launches), "clocks": clocks.summary(),
                "launch_mode": (f"CUDA graph, {GK} step kernels per replay x {replays} replays + {rest} eager steps" if replays else
                                f"{rest} eager launches (programmatic dependent launch between consecutive step kernels)"),
                "eager": {"ms_per_step": ms_eager / max(1, eager_launches), "launches": int(eager_launches),
                          "note": "same loop launched step by step from Python/ctypes"},
                "episode_stats": stats, "other": other}
        if world == 1 and not args.skip_cpu:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs-per-gpu", type=int, default=4 * 1024 * 1024)
    ap.add_argument("--total-envs", type=int, default=0, help="strong scaling: shard this many envs over the ranks")
    ap.add_argument("--n-traffic", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-other", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="run the BASELINE config 4 traffic sweep on every rank too (N>1)")
    ap.add_argument("--age-steps", type=int, default=2048, help="untimed steps that bring the batch to its steady state")
    ap.add_argument("--n1-occ", type=int, default=0, help="experiment: 3|4 resident blocks/SM for the N=1 kernel")
    ap.add_argument("--n1-tma", type=int, default=-1, help="experiment: 1 = TMA-ring persistent kernel, 0 = direct kernel")
    ap.add_argument("--n1-stages", type=int, default=0, help="experiment: TMA ring depth 2..5")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
