"""Directed corner cases for the reference's bug-compat list:
one ``step()`` from an injected state, compared across
  * the unmodified reference (imported under the gym / pygame stand-ins; build container only),
  * the float64 oracle,
  * the product's per-env source compiled with g++ (``tests/hostcheck``),
  * the CUDA kernels through the C ABI (``-m gpu``).
The geometry of every case is made of small integers so that the strict thresholds (Q8) sit exactly on a
representable value in all four implementations."""
import numpy as np
import pytest

from oracle import ref_shim
from oracle.acas2d_oracle import FLAG_DONE, Oracle
from tests import parity
from tests.hostcheck import HostBatch

FLAG_COLLISION, FLAG_GOAL, FLAG_TIMEOUT, FLAG_OOB = 1, 2, 4, 16
GOAL = (1456.0, 500.0)                      # WIDTH - GOAL_RADIUS, HEIGHT / 2 (game.py:80-81)

# name -> dict(player=(x, y, psi), traffic=[(x, y, v, psi), ...], steps, action, expect=(flags & 15, outcome))
CASES = {
    # Q9: both bonuses on the same step (-1000 + 1000); outcome priority collision > goal
    "q9_collision_and_goal": dict(player=(1400.0, 500.0, 0.0), traffic=[(1452.0, 500.0, 200.0, 180.0)], steps=300,
                                  action=0.0, expect=(FLAG_COLLISION | FLAG_GOAL | 8, 2)),
    # Q5 / Q6 / Q9: the 1000th step() call times out (steps == 1001, discount -0.001); the collision bonus is
    # still added; outcome priority timeout > collision
    "q9_timeout_with_collision": dict(player=(700.0, 500.0, 0.0), traffic=[(752.0, 500.0, 200.0, 180.0)], steps=1000,
                                      action=0.0, expect=(FLAG_COLLISION | FLAG_TIMEOUT | 8, 3)),
    # Q8: strict thresholds -- separation exactly 96 is NOT a collision ...
    "q8_separation_exactly_96": dict(player=(500.0, 500.0, 0.0), traffic=[(600.0, 500.0, 200.0, 180.0)], steps=10,
                                     action=0.0, expect=(0, 0)),
    # ... one part in 1e11 closer is
    "q8_separation_just_below_96": dict(player=(500.0, 500.0, 0.0), traffic=[(600.0 - 1e-9, 500.0, 200.0, 180.0)], steps=10,
                                        action=0.0, expect=(FLAG_COLLISION | 8, 2)),
    # goal distance exactly 144 is NOT the goal
    "q8_goal_distance_exactly_144": dict(player=(1310.0, 500.0, 0.0), traffic=[(100.0, 100.0, 200.0, 90.0)], steps=10,
                                         action=0.0, expect=(0, 0)),
    "q8_goal_distance_just_below_144": dict(player=(1310.0 + 1e-9, 500.0, 0.0), traffic=[(100.0, 100.0, 200.0, 90.0)], steps=10,
                                            action=0.0, expect=(FLAG_GOAL | 8, 1)),
    # Q12: equal velocities -> arctan(0/0) -> d_cpa is NaN in the observation, min(1, nan) == 1 in the reward
    "q12_equal_velocities": dict(player=(400.0, 500.0, 10.0), traffic=[(700.0, 300.0, 200.0, 10.0)], steps=50,
                                 action=0.0, expect=(0, 0)),
    # Q12: arctan, not arctan2 -- relative velocity pointing in -x flips the sign convention of d_cpa
    "q12_negative_v12x": dict(player=(800.0, 500.0, 180.0), traffic=[(400.0, 380.0, 200.0, 20.0)], steps=50,
                              action=0.25, expect=(0, 0)),
    # Q13: leaving the map never ends an episode (informational flag only)
    "q13_out_of_map": dict(player=(1.0, 500.0, 180.0), traffic=[(900.0, 100.0, 200.0, 90.0)], steps=50,
                           action=0.0, expect=(0, 0)),
    # Q19 / Q2: actions are not clipped; the closing-speed look-ahead turns by a different amount than the aircraft
    "q19_unclipped_action": dict(player=(300.0, 400.0, 5.0), traffic=[(900.0, 700.0, 200.0, 200.0)], steps=50,
                                 action=3.7, expect=(0, 0)),
    "q2_full_deflection_converging": dict(player=(300.0, 500.0, 350.0), traffic=[(700.0, 450.0, 200.0, 175.0)], steps=50,
                                          action=-1.0, expect=(0, 0)),
    # Q7: only traffic[0] shapes the reward (two orders of the same two intruders), all intruders collide
    "q7_near_intruder_first": dict(player=(300.0, 500.0, 0.0), traffic=[(600.0, 520.0, 200.0, 185.0), (1500.0, 100.0, 200.0, 90.0)],
                                   steps=50, action=0.1, expect=(0, 0)),
    "q7_near_intruder_second": dict(player=(300.0, 500.0, 0.0), traffic=[(1500.0, 100.0, 200.0, 90.0), (600.0, 520.0, 200.0, 185.0)],
                                    steps=50, action=0.1, expect=(0, 0)),
    "q7_collision_with_second": dict(player=(300.0, 500.0, 0.0), traffic=[(1500.0, 100.0, 200.0, 90.0), (380.0, 500.0, 200.0, 180.0)],
                                     steps=50, action=0.0, expect=(FLAG_COLLISION | 8, 2)),
}


def run_oracle(case):
    n = len(case["traffic"])
    orc = Oracle(n)
    st = orc.new_state(1)
    x, y, psi = case["player"]
    st["player"][0] = (x, y, 200.0, psi, 0.0)
    st["traffic"][0] = case["traffic"]
    st["steps"][0] = case["steps"]
    with np.errstate(all="ignore"):
        obs, rew, flags, outcome = orc.step(st, np.array([case["action"]]))
    return obs[0], float(rew[0]), int(flags[0]), int(outcome[0])


def run_host(case, cls=HostBatch, **kw):
    n = len(case["traffic"])
    hb = cls(1, n, auto_reset=False, **kw)
    hb.inject_state(np.array([case["player"]]), np.array([case["traffic"]]), np.array([case["steps"]], np.int32), np.zeros(1))
    obs, rew, done = hb.step(np.array([case["action"]], np.float32))
    to = (lambda a: a.detach().cpu().numpy()) if not isinstance(obs, np.ndarray) else (lambda a: np.asarray(a))
    flags = int(to(hb.flags)[0])
    return to(obs)[0].astype(np.float64), float(to(rew)[0]), flags, int(to(hb.outcome)[0]) if flags & 8 else 0


def run_reference(case):
    """The unmodified reference: poke the game's attributes (what its own scripts do), one env.step."""
    n = len(case["traffic"])
    ref_shim.load(n)
    try:
        from gym_ACAS2D.envs.environment import ACAS2DEnv
        with ref_shim.quiet(), np.errstate(all="ignore"):
            env = ACAS2DEnv()
            env.reset()
            g = env.game
            g.player.x, g.player.y, g.player.psi = case["player"]
            g.player.a_lat = 0.0
            for t, (x, y, v, psi) in zip(g.traffic, case["traffic"]):
                t.x, t.y, t.v_air, t.psi = x, y, v, psi
            g.steps = case["steps"]
            g.total_reward = 0.0
            obs, rew, done, _ = env.step(np.array([case["action"]]))
        return np.asarray(obs, np.float64), float(rew), bool(done), int(g.outcome) if done else 0
    finally:
        ref_shim.unload()


def check_against_oracle(name, got, want):
    obs, rew, flags, outcome = got
    wobs, wrew, wflags, woutcome = want
    assert flags & 15 == wflags & 15 and outcome == woutcome, name
    assert np.array_equal(np.isnan(obs), np.isnan(wobs)), name
    with np.errstate(invalid="ignore"):
        err = np.abs(obs - wobs)
    assert np.nanmax(err[:6]) < parity.TOL_OBS_BASE and np.nanmax(err) < parity.TOL_OBS_CPA, (name, err)
    assert abs(rew - wrew) < parity.TOL_REWARD_ABS + parity.TOL_REWARD_REL * abs(wrew) + 1e-4 * (abs(wrew) >= 900), (name, rew, wrew)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_does_what_the_case_says(name):
    case = CASES[name]
    obs, rew, flags, outcome = run_oracle(case)
    assert (flags & 15, outcome) == case["expect"], (name, flags, outcome)
    if name == "q9_collision_and_goal":
        assert abs(rew) < 2.0                                  # shaped reward -1000 + 1000
    if name == "q9_timeout_with_collision":
        assert -1000.0 - 1e-2 < rew < -1000.0 + 1e-2 and abs(obs[0] - 1.001) < 1e-15     # Q5 / Q6
    if name == "q12_equal_velocities":
        assert np.isnan(obs[6]) and obs[7] == 0.0 and np.isfinite(rew) and rew > 0.0
    if name == "q13_out_of_map":
        assert not flags & FLAG_DONE
    if name == "q7_near_intruder_second":
        other = run_oracle(CASES["q7_near_intruder_first"])
        assert abs(other[1] - rew) > 1e-3 and other[2] == flags    # the reward depends on which intruder is traffic[0]
        assert np.allclose(other[0][5:8], obs[8:11]) and np.allclose(other[0][8:11], obs[5:8])


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_the_unmodified_reference(name):
    case = CASES[name]
    robs, rrew, rdone, routcome = run_reference(case)
    obs, rew, flags, outcome = run_oracle(case)
    assert rdone == bool(flags & FLAG_DONE) and routcome == outcome, name
    assert np.array_equal(np.isnan(robs), np.isnan(obs))
    with np.errstate(invalid="ignore"):
        assert np.nanmax(np.abs(robs - obs)) < 1e-13, (name, robs, obs)
    assert abs(rrew - rew) < 1e-12 * max(1.0, abs(rrew)), (name, rrew, rew)


@pytest.mark.parametrize("name", sorted(CASES))
def test_product_source_on_the_cpu(name):
    case = CASES[name]
    n = len(case["traffic"])
    for variant in ((0, 1) if n == 1 else (1,)):
        check_against_oracle(name, run_host(case, variant=variant), run_oracle(case))
    if name == "q13_out_of_map":
        assert run_host(case)[2] & FLAG_OOB


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_kernels(name):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D

    class Gpu(BatchedACAS2D):
        def __init__(self, B, n, auto_reset=False):
            super().__init__(B, n_traffic=n, device="cuda:0", auto_reset=auto_reset)

        def step(self, a):
            return super().step(torch.from_numpy(np.asarray(a, np.float32)).cuda())

    case = CASES[name]
    check_against_oracle(name, run_host(case, cls=Gpu), run_oracle(case))
    if name == "q13_out_of_map":
        assert run_host(case, cls=Gpu)[2] & FLAG_OOB


# ------------------------------------------------------------------------------------------------ random states
def random_states(B, n, seed):
    """States far outside what episodes visit: anywhere within half a map around the map, any heading, intruder
    speeds 100 .. 300, game.steps 1 .. 1000, unclipped actions in [-2, 2]."""
    rng = np.random.default_rng(seed)
    player = np.stack([rng.uniform(-800, 2400, B), rng.uniform(-500, 1500, B), rng.uniform(0, 360, B)], 1)
    traffic = np.stack([rng.uniform(-800, 2400, (B, n)), rng.uniform(-500, 1500, (B, n)),
                        rng.uniform(100, 300, (B, n)), rng.uniform(0, 360, (B, n))], 2)
    steps = rng.integers(1, 1001, B).astype(np.int32)
    actions = rng.uniform(-2, 2, B).astype(np.float32)
    return player, traffic, steps, actions


def compare_random_step(got_obs, got_rew, got_flags, got_outcome, orc, player, traffic, steps, actions):
    B, n = traffic.shape[:2]
    st = orc.new_state(B)
    st["player"][:, 0], st["player"][:, 1], st["player"][:, 2], st["player"][:, 3] = player[:, 0], player[:, 1], 200.0, player[:, 2]
    st["traffic"][:] = traffic
    st["steps"][:] = steps
    with np.errstate(all="ignore"):
        obs, rew, flags, outcome = orc.step(st, actions.astype(np.float64))
    assert np.array_equal(got_flags & 15, flags & 15)                        # collision / goal / timeout / done: bit-exact
    done = (flags & FLAG_DONE) > 0
    assert np.array_equal(got_outcome[done], outcome[done])
    g = np.asarray(got_obs, np.float64)
    d = np.abs(g[:, :5] - obs[:, :5])
    for c in (1, 4):
        d[:, c] = np.minimum(d[:, c], 1.0 - d[:, c])
    assert d.max() < parity.TOL_OBS_BASE
    assert np.abs(g[:, 5::3] - obs[:, 5::3]).max() < parity.TOL_OBS_SEP
    # float32 outputs: d_cpa is unbounded for near-parallel tracks, so its tolerance is relative beyond 1
    cpa, rcpa = g[:, 6::3], obs[:, 6::3]
    assert np.array_equal(np.isnan(cpa), np.isnan(rcpa))
    with np.errstate(invalid="ignore"):
        ok = np.abs(cpa - rcpa) <= parity.TOL_OBS_CPA * np.maximum(1.0, np.abs(rcpa))
        flip = np.abs(np.abs(cpa) - np.abs(rcpa)) <= parity.TOL_OBS_CPA * np.maximum(1.0, np.abs(rcpa))    # Q12 sign flip
    assert (ok | flip | np.isnan(rcpa)).all() and (~ok & flip).sum() <= 2
    assert np.nanmax(np.abs(g[:, 7::3] - obs[:, 7::3])) < parity.TOL_OBS_VC
    near_branch = (np.abs(obs[:, 7]) < 1e-6) | ~ok[:, 0]
    dr = np.abs(np.asarray(got_rew, np.float64) - rew)
    assert (dr[~near_branch] <= parity.TOL_REWARD_ABS + 1e-6 * np.abs(rew[~near_branch])).all(), dr[~near_branch].max()
    assert done.mean() > 0.01 and (flags & FLAG_COLLISION > 0).mean() > 0.001


@pytest.mark.parametrize("n,variant", [(1, 0), (1, 1), (5, 1)])
def test_random_states_single_step_product_source(n, variant):
    B = 20000
    player, traffic, steps, actions = random_states(B, n, seed=100 + n)
    hb = HostBatch(B, n, auto_reset=False, variant=variant)
    hb.inject_state(player, traffic, steps, np.zeros(B))
    obs, rew, done = hb.step(actions)
    compare_random_step(obs, rew, hb.flags, hb.outcome, Oracle(n), player, traffic, steps, actions)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 5, 64])
def test_random_states_single_step_cuda(n):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D
    B = 256 * 300 + 77 if n < 64 else 8192 + 3
    player, traffic, steps, actions = random_states(B, n, seed=100 + n)
    env = BatchedACAS2D(B, n_traffic=n, device="cuda:0", auto_reset=False)
    env.inject_state(player, traffic, steps, np.zeros(B))
    obs, rew, done = env.step(torch.from_numpy(actions).cuda())
    c = lambda t: t.detach().cpu().numpy()      # noqa: E731
    compare_random_step(c(obs), c(rew), c(env.flags), c(env.outcome), Oracle(n), player, traffic, steps, actions)


# ------------------------------------------------------------------------------------------------ other settings
# Every constant of settings.py the path reads, moved away from its default (a1: settings -> kernel parameters,
# including the derived normalisers of game.py:120-128 and the reward constants of rewards.py).
OTHER_SETTINGS = dict(WIDTH=1200, HEIGHT=800, FPS=50, MAX_STEPS=300, AIRCRAFT_SIZE=20, COLLISION_RADIUS=30,
                      GOAL_RADIUS=100, SAFE_DISTANCE=150, AIRSPEED=150, AIRSPEED_FACTOR_MIN=0.8,
                      AIRSPEED_FACTOR_MAX=1.25, ACC_LAT_LIMIT=120.0, PLAYER_INITIAL_HEADING_LIM=10,
                      TRAFFIC_INITIAL_HEADING_LIM=40, REWARD_GOAL=500, REWARD_COLLISION=-750)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_oracle_equals_the_unmodified_reference_under_other_settings():
    import random
    n = 2
    ref_shim.load(n, **OTHER_SETTINGS)
    try:
        from gym_ACAS2D.envs.environment import ACAS2DEnv
        orc = Oracle(n, **OTHER_SETTINGS)
        random.seed(5)
        arng = np.random.default_rng(5)
        outcomes = []
        with ref_shim.quiet(), np.errstate(all="ignore"):
            env = ACAS2DEnv()
            for ep in range(6):
                obs = env.reset()
                g = env.game
                st = orc.new_state(1)
                st["player"][0] = (g.player.x, g.player.y, g.player.v_air, g.player.psi, 0.0)
                for i in range(n):
                    t = g.traffic[i]
                    st["traffic"][0, i] = (t.x, t.y, t.v_air, t.psi)
                assert np.abs(orc.observe(st)[0] - obs).max() < 1e-14
                for k in range(400):
                    a = float(np.float32(arng.uniform(-1, 1))) * (0.2 if ep % 2 else 1.0)
                    ro, rr, rd, _ = env.step(np.array([a]))
                    co, cr, cf, coc = orc.step(st, np.array([a]))
                    assert bool(cf[0] & FLAG_DONE) == rd
                    assert np.nanmax(np.abs(ro - co[0])) < 1e-13 and abs(rr - cr[0]) < 1e-12 * max(1.0, abs(rr))
                    if rd:
                        assert g.outcome == coc[0] and g.steps == st["steps"][0]
                        outcomes.append(int(coc[0]))
                        break
        assert len(outcomes) == 6                  # every episode ended within the 400 steps: MAX_STEPS = 300 took effect
    finally:
        ref_shim.unload()


def _other_settings_rollout(make_env, to_np, B, n, steps):
    seed, off = 41, 77
    env = make_env(B, n, seed=seed, env_id_offset=off, auto_reset=True, **OTHER_SETTINGS)
    orc = Oracle(n, **OTHER_SETTINGS)
    st = orc.new_state(B)
    orc.spawn_philox(st, seed, off)
    assert np.abs(to_np(env.reset()) - orc.observe(st)).max() < parity.TOL_OBS_CPA
    rng = np.random.default_rng(8)
    rep = parity.ParityReport()
    episodes = 0
    for t in range(steps):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        obs, rew, done = env.step(a)
        o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, a.astype(np.float64), seed, off)
        d = f & FLAG_DONE > 0
        assert np.array_equal(to_np(done).astype(bool), d)
        parity.compare_step(rep, np.where(d[:, None], to_np(env.term_obs), to_np(obs)), to_np(rew), to_np(env.flags),
                            np.where(d[:, None], term, o), r, f)
        if d.any():
            assert np.array_equal(to_np(env.outcome)[d], oc[d]) and np.array_equal(to_np(env.ep_length)[d], ep_len[d])
            episodes += int(d.sum())
    parity.assert_flags_exact(rep)
    assert episodes > B                              # MAX_STEPS = 300: every env finished at least once


@pytest.mark.parametrize("n,variant", [(1, 0), (3, 1)])
def test_other_settings_product_source(n, variant):
    _other_settings_rollout(lambda B, n, **kw: HostBatch(B, n, variant=variant, **kw), np.asarray, 48, n, 700)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 8])
def test_other_settings_cuda(n):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D

    class Gpu(BatchedACAS2D):
        def __init__(self, B, n, **kw):
            super().__init__(B, n_traffic=n, device="cuda:0", **kw)

        def step(self, a):
            return super().step(torch.from_numpy(np.asarray(a, np.float32)).cuda())

    _other_settings_rollout(Gpu, lambda t: t.detach().cpu().numpy(), 256 * 2 + 9, n, 700)
