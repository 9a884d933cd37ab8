"""Host-side logic that needs no GPU: settings, spaces, registration, statistics and the
world_size-2 statistics reduction (gloo)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

import gym_ACAS2D
from gym_ACAS2D import settings
from gym_ACAS2D.envs import _native
from gym_ACAS2D.envs.spaces import action_box, observation_box
from gym_ACAS2D.envs.stats import summarise
from oracle import ref_shim
from oracle.acas2d_oracle import DEFAULTS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_settings_values():
    for k, v in DEFAULTS.items():
        assert getattr(settings, k) == v, k
    assert settings.TOTAL_STEPS == 1048576 and settings.EVAL_STEPS == 32768
    assert settings.OUTCOME_NAMES == {1: "Goal", 2: "Collision", 3: "Timeout"}


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_settings_equal_reference_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_settings", os.path.join(ref_shim.REFERENCE_ROOT, "gym_ACAS2D", "settings.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    names = [n for n in dir(ref) if n.isupper()]
    assert len(names) >= 38
    for n in names:
        assert getattr(settings, n) == getattr(ref, n), n


def test_spaces_follow_reference():
    """environment.py:18-27."""
    for n in (1, 8):
        ob = observation_box(n)
        assert ob.shape == (5 + 3 * n,) and ob.dtype == np.float64
        assert list(ob.low[:5]) == [0, 0, -1, 0, 0] and list(ob.low[5:8]) == [0, -1, -1]
        assert np.all(ob.high == 1)
    ab = action_box()
    assert ab.shape == (1,) and ab.low[0] == -1 and ab.high[0] == 1 and ab.dtype == np.float64
    assert ab.contains(ab.sample())


def test_registration_surface():
    assert gym_ACAS2D.ENV_ID == "ACAS2D-v0"
    assert gym_ACAS2D.ENTRY_POINT == "gym_ACAS2D.envs:ACAS2DEnv"        # reference __init__.py:3-6
    with pytest.raises(KeyError):
        gym_ACAS2D.make("nope-v0")
    from gym_ACAS2D.envs import ACAS2DEnv, ACAS2DGame                    # reference envs/__init__.py:1-2
    assert ACAS2DEnv.__name__ == "ACAS2DEnv" and ACAS2DGame.__name__ == "ACAS2DGame"


def test_params_reject_unsupported_traffic():
    with pytest.raises(ValueError):
        _native.params_from_settings(None, 0)
    with pytest.raises(NotImplementedError):
        _native.params_from_settings(None, None, MIN_TRAFFIC=1, MAX_TRAFFIC=3)   # Q11: reference breaks its own Box
    with pytest.raises(ValueError):
        _native.params_from_settings(None, _native.MAX_TRAFFIC + 1)
    p = _native.params_from_settings(None, None, MIN_TRAFFIC=4, MAX_TRAFFIC=4)
    assert p.n_traffic == 4


def test_summarise_counters():
    c = torch.tensor([10, 4, 5, 1, 5000, int(-700.5 * 1048576), int(1500.25 * 1048576)], dtype=torch.int64)
    s = summarise(c, reduce=False, track_min_sep=True)
    assert s["episodes"] == 10 and s["goal_rate"] == 0.4 and s["collision_rate"] == 0.5 and s["timeout_rate"] == 0.1
    assert s["mean_length"] == 500 and s["mean_step_calls"] == 499
    assert abs(s["mean_return"] + 70.05) < 1e-6 and abs(s["mean_min_separation"] - 150.025) < 1e-6
    assert summarise(torch.zeros(7, dtype=torch.int64), reduce=False)["episodes"] == 0


def test_stats_allreduce_world_size_2_gloo(tmp_path):
    """The only collective on the path: all-reduce(SUM) of the int64 episode counters."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, {os.path.join(ROOT, 'gym-acas2d_b200')!r})
        from gym_ACAS2D.envs.stats import summarise
        dist.init_process_group("gloo")
        r = dist.get_rank()
        c = torch.tensor([3 + r, 1 + r, 1, 1, 1000 * (r + 1), (r + 1) * 1048576, 0], dtype=torch.int64)
        s = summarise(c, reduce=True)
        assert s["episodes"] == 7 and s["goal"] == 3 and s["collision"] == 2, s
        assert s["mean_length"] == 3000 / 7 and abs(s["mean_return"] - 3 / 7) < 1e-12, s
        local = summarise(c, reduce=False)
        assert local["episodes"] == 3 + r
        dist.destroy_process_group()
        print("ok", r)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_sharded_rollout_world_size_2_gloo(tmp_path):
    """The multi-GPU path on CPU: two gloo ranks each own a contiguous block of global env ids (here
    stepped by the g++ host build of the device source), exchange nothing per step, and all-reduce the
    seven episode counters at the end.  The result must equal the single-process run of the whole batch:
    spawns, action streams and statistics depend on the global env id only."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'gym-acas2d_b200')!r})
        from tests.hostcheck import HostBatch
        from gym_ACAS2D.envs.stats import reduce_counters, summarise
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        TOTAL, T = 600, 260
        off, n = bench.shard(TOTAL, world, rank)
        mine = HostBatch(n, 1, seed=13, env_id_offset=off, auto_reset=True)
        mine.reset()
        ex = mine.extract_state(); ex["steps"][:] = 800 + (np.arange(off, off + n) % 190)      # plenty of episode ends
        mine.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
        for t in range(T):
            mine.step(mine.random_actions(t, action_seed=5))
        total = reduce_counters(torch.from_numpy(mine.episode_counters().copy()))
        if rank == 0:
            full = HostBatch(TOTAL, 1, seed=13, env_id_offset=0, auto_reset=True)
            full.reset()
            ex = full.extract_state(); ex["steps"][:] = 800 + (np.arange(TOTAL) % 190)
            full.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
            for t in range(T):
                full.step(full.random_actions(t, action_seed=5))
            assert np.array_equal(full.episode_counters(), total.numpy()), (full.episode_counters(), total)
            assert np.array_equal(full.ppos[:n], mine.ppos) and np.array_equal(full.obs[:n], mine.obs)
            s = summarise(torch.from_numpy(mine.episode_counters().copy()), reduce=True)
            assert s["episodes"] == int(total[0]) and s["episodes"] >= TOTAL // 2
        else:
            summarise(torch.from_numpy(mine.episode_counters().copy()), reduce=True)
        dist.destroy_process_group()
        print("ok", rank)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


def test_bench_sharding_arithmetic():
    """Contiguous global env ids per rank: union of shards == the whole batch."""
    sys.path.insert(0, ROOT)
    import bench
    for total, world in [(1 << 20, 8), (1000, 3), (7, 8)]:
        spans = [bench.shard(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][0] + spans[-1][1] == total
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 == o1
