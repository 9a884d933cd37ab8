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


def cpu_baseline_block(budget_s: float = 12.0):
    steps = 4000
    rate1 = time_pyport(1, steps)
    steps = int(min(max(rate1 * budget_s * 0.5, 2000), 200000))
    rate1 = time_pyport(1, steps)
    c_rate = time_c_oracle(8192, 64)
    return {"value": rate1, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle Python port (reference-shaped scalar float64 code), 1 env, {steps} random-action "
                      f"steps incl. resets, N_TRAFFIC=1; the reference itself is wall-clock capped at "
                      f"100 steps/s (environment.py:31) and logged 69-89 steps/s (BASELINE.md)",
            "c_oracle_1core": {"value": c_rate, "sample": "oracle C restatement, 8192 envs x 64 auto-reset steps, 1 thread"}}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores.
    The reference is a Python program that cannot travel to the GPU box (no gym / pygame, and
    /root/reference is absent there), so the oracle's Python port stands in (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 256                                   # env-steps per core per bench "step"
    K, W = args.steps, args.warmup
    K_eff = min(K, 400)                              # bounded: the whole run ends within minutes
    t0 = time.perf_counter()
    rate = time_pyport(cores, per_step * K_eff)
    wall = time.perf_counter() - t0
    sample = (f"{cores} processes x 1 env x {per_step * K_eff} random-action steps incl. resets "
              f"(oracle Python port, N_TRAFFIC=1); steps capped at {K_eff} of the requested {K}")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * cores * per_step / rate,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "random-action rollout, default N_TRAFFIC=1, auto-reset, one env per host core",
                       "envs": cores, "env_steps_per_step": cores * per_step},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(device_index: int):
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host buffer is allocated, so the
    end-to-end path's H2D / D2H copies do not cross sockets (8 ranks on one node otherwise pile up)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # noqa: BLE001
        pass
    return None


# ------------------------------------------------------------------------------------------ GPU arm
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
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
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
    KA = 8
    actions = torch.empty(KA, B, dtype=torch.float32, device=dev)
    for k in range(KA):
        env.random_actions(k, action_seed=2024, out=actions[k])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters):
        """CUDA-event time of ``iters`` calls of fn on the current stream, max over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    graph = env.capture_steps(actions, full_outputs=False, num_steps=GK, warmup=False)
    graph.replay()

    def k_steps(_):
        for _r in range(replays):
            graph.replay()
        for k in range(rest):
            step_fn(k)

    clocks = ClockSampler(local)
    clocks.__enter__()                                  # sampled across the graph, eager and end-to-end regions
    ms = timed(k_steps, 1)
    launches = K                                         # one step kernel per env step (graph nodes + eager)
    value = world * B * K / (ms * 1e-3)
    peak, peak_src = measured_peak()
    A = algorithmic_bytes(N)
    achieved = A * B * K / (ms * 1e-3) / 1e9            # GB/s per GPU (max-over-ranks time)
    launches0 = lib.acas2d_launch_count()
    ms_eager = timed(step_fn, min(K, 400))
    eager_launches = lib.acas2d_launch_count() - launches0
    stats = env.episode_stats(reduce=True)              # the one collective of the path (NCCL)

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
    clocks.__exit__(None, None, None)

    # ---- secondary workloads (reported, not the headline)
    other = {}
    if N == 1 and not args.skip_other:
        try:                                           # the secondary lines must never cost the run its headline line
            env.rollout_random(8, action_seed=1, step0=0)
            fused_k, reps = 64, 4
            ms_f = timed(lambda k: env.rollout_random(fused_k, action_seed=1, step0=100 + fused_k * k), reps)
            other["fused_rollout_64_steps_per_launch"] = {"value": world * B * fused_k * reps / (ms_f * 1e-3), "unit": UNIT,
                                                          "note": "in-kernel Philox actions, state in registers, no per-step outputs"}
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
                        "written once per launch (80 / K + 41 bytes per env-step instead of 121)"}
            del ko, kr, kd, kd8
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
            # the SB3-style numpy VecEnv surface (host arrays in / out, per-env info dicts) at 4096 envs
            from gym_ACAS2D.envs import ACAS2DVecEnv
            import numpy as np
            venv = ACAS2DVecEnv(4096, device=dev, seed=13)
            venv.reset()
            ex = venv.core.extract_state()                 # stagger the episodes: finished envs (info dicts, terminal rows) every step
            ex["steps"][:] = 1 + (np.arange(4096) * 7) % 900
            venv.core.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
            va = np.zeros((4096, 1), np.float32)
            for _ in range(5):
                venv.step(va)
            t0 = time.perf_counter()
            for _ in range(200):
                venv.step(va)
            dt = time.perf_counter() - t0
            other["vecenv_numpy_surface_4096_envs"] = {"value": world * 4096 * 200 / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / 200,
                                                       "note": "ACAS2DVecEnv.step(np actions) -> np obs / rewards / dones + list of info dicts, ~7 episodes ending per step; wall clock"}
            small = BatchedACAS2D(4096, n_traffic=1, device=dev, seed=13, env_id_offset=0, auto_reset=True)
            small.reset()
            sgraph = small.capture_steps(actions[:, :4096].contiguous(), num_steps=200)      # 200 steps per replay
            sgraph.replay()
            ms_s = timed(lambda k: sgraph.replay(), 10)
            other["config2_4096_envs_cuda_graph"] = {"value": world * 4096 * 200 * 10 / (ms_s * 1e-3), "unit": UNIT,
                                                     "note": "BASELINE config 2 batch: launch-bound, 200-step CUDA graph replay"}

            if world == 1:
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
        except Exception as exc:                       # noqa: BLE001
            other["error"] = repr(exc)
            torch.cuda.synchronize()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64 flag chain + f32 observation/reward", "data": "synthetic",
                "config": {"workload": "BASELINE config 3 random-action rollout at default N_TRAFFIC with auto-reset, "
                                       "HBM-saturating batch per GPU (SURVEY 8d), sharded by global env id",
                           "envs_per_gpu": B, "total_envs": world * B, "n_traffic": N, "obs_dim": L,
                           "actions": f"pre-generated Philox U(-1,1) float32 [{KA},B] resident in HBM, cycled",
                           "l2": "state + outputs per GPU = %.0f MB >> 126 MB L2 (no flush needed)" % (B * (A + 64) / 1e6),
                           "parallelism": f"env-sharded x{world}, one all-reduce of 7 int64 episode counters after the timed region"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(N, B), "peak_source": peak_src,
                             "algorithmic_bytes_per_env_step": A, "env_steps_per_launch": B,
                             "note": "per GPU; achieved = A(N) x envs per launch / CUDA-event time per launch"},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": (4 * L + 5) * B,
                        "steps": Ke, "ms_per_step": ms_e2e / Ke,
                        "path": "BatchedACAS2D.step_host -> acas2d_step_host (pinned host buffers)",
                        "numa_node_rank0": numa},
                "gpu_launches": int(launches), "clocks": clocks.summary(),
                "launch_mode": f"CUDA graph, {GK} step kernels per replay x {replays} replays + {rest} eager steps",
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
