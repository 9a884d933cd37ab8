"""The oracle is pinned here: against the reference's golden CSV (condensed fixture), against
rollouts produced by the unmodified reference, against the notebook known-answer values, and --
when the reference tree is present (build container only) -- against the live reference."""
import os
import random

import numpy as np
import pytest

from oracle import ref_shim
from oracle.acas2d_oracle import (DEFAULTS, FLAG_DONE, Oracle, PyPortGame, reference_spawn)


@pytest.fixture(scope="module")
def orc():
    return Oracle(1)


def _replay_zero_action(orc, n_episodes):
    """SURVEY 8c recipe: random.seed(13), 2 discarded games, zero-action episodes."""
    rng = random.Random(13)
    for _ in range(2):
        reference_spawn(rng, DEFAULTS)
    for ep in range(n_episodes):
        pl, tr = reference_spawn(rng, DEFAULTS)
        st = orc.new_state(1)
        st["player"][0] = pl
        st["traffic"][0] = tr
        orc.observe(st)
        path, tpath = [(pl[0], pl[1])], [(tr[0, 0], tr[0, 1])]
        while True:
            tpath.append((st["traffic"][0, 0, 0], st["traffic"][0, 0, 1]))   # recorded before traffic moves (Q10)
            _, _, fl, oc = orc.step(st, np.zeros(1))
            path.append((st["player"][0, 0], st["player"][0, 1]))
            if fl[0] & FLAG_DONE:
                break
        yield ep, int(oc[0]), int(st["steps"][0]), float(st["total_reward"][0]), np.array(path), np.array(tpath)


def test_golden_csv_replay(orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "baseline_zero_action.npz"))
    stride = int(g["stride"])
    assert (g["outcome"] == 2).sum() == 58 and (g["outcome"] == 1).sum() == 42      # notebook cell 4
    assert abs(g["time_steps"].mean() - 494.65) < 1e-9
    assert abs(g["total_reward"].mean() - (-70.775)) < 1e-3
    for ep, oc, steps, total, path, tpath in _replay_zero_action(orc, 100):
        assert oc == g["outcome"][ep], ep
        assert steps == g["time_steps"][ep], ep
        assert len(path) == g["path_len"][ep], ep
        n = len(path[::stride])
        assert np.array_equal(path[::stride], g["path_samples"][ep][:n]), ep          # bit-identical
        assert np.array_equal(tpath[::stride], g["traffic_samples"][ep][:n]), ep
        assert np.array_equal(path[-1], g["path_last"][ep]) and np.array_equal(tpath[-1], g["traffic_last"][ep])
        assert np.array_equal(path.sum(0), g["path_sum"][ep]) and np.array_equal(tpath.sum(0), g["traffic_sum"][ep])
        assert abs(total - g["total_reward"][ep]) < 1e-11, ep                          # 1-ulp sin/cos differences


@pytest.mark.parametrize("n", [1, 8])
def test_oracle_matches_reference_rollouts(golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"ref_rollouts_n{n}.npz"))
    orc = Oracle(n)
    B, T = g["player0"].shape[0], g["actions"].shape[0]
    st = orc.new_state(B)
    st["player"][:] = g["player0"]; st["traffic"][:] = g["traffic0"]
    st["steps"][:] = g["steps0"]; st["total_reward"][:] = g["total0"]
    d = g["traffic0"][:, :, :2] - g["player0"][:, None, :2]
    st["min_sep"][:] = np.sqrt((d * d).sum(-1)).min(-1)
    out = orc.rollout(st, g["actions"].astype(np.float64), record_traffic=True)
    s = int(g["stride"])
    assert np.array_equal(out["flags"], g["flags"])
    assert np.array_equal(out["outcome"], g["outcome"])
    assert np.array_equal(out["player"][::s], g["player_strided"])                    # bit-identical positions
    assert np.array_equal(out["traffic"][::s], g["traffic_strided"])
    with np.errstate(invalid="ignore"):
        assert np.nanmax(np.abs(out["obs"][::s] - g["obs_strided"])) < 1e-14
        assert np.nanmax(np.abs(out["reward"] - g["reward"])) < 1e-14
    assert np.abs(st["total_reward"] - g["total_reward"]).max() < 1e-11
    assert np.array_equal(st["steps"], g["steps"])
    assert np.abs(st["min_sep"] - g["min_sep"]).max() < 1e-9
    assert np.abs(st["d_path"] - g["d_path"]).max() < 1e-9


def test_reward_known_answers():
    """notebooks/rewards.ipynb cells 11, 18, 23, 28.  The printed outputs are stale with respect
    to the notebook's own constants cell: cells 4-5 print d_goal_init = 856 and d_goal_max = 2296,
    and the GOAL_RADIUS answers only fit GOAL_RADIUS = 96.  Constants reproducing those prints:
    WIDTH - GOAL_RADIUS = 904, (AIRSPEED/FPS) * MAX_STEPS = 1440."""
    nb = Oracle(1, WIDTH=1000, GOAL_RADIUS=96, MAX_STEPS=720)
    kat = [(0, 1), (48, 0.9189622950516945), (96, 0.8429526657055691), (192, 0.7051725641148902), (2296.0, 0.0)]
    for d, want in kat:
        assert nb.goal_distance_reward(d) == pytest.approx(want, abs=1e-15)
    for psi, want in [(0, 1.0), (10, 0.7956199512269471), (350, 0.7956199512269471), (20, 0.624295076969974),
                      (340, 0.624295076969974), (30, 0.4822530864197532), (330, 0.4822530864197532),
                      (40, 0.3659503124523701), (320, 0.3659503124523701)]:
        assert nb.heading_reward(psi, 0) == pytest.approx(want, abs=1e-15)
    for d, want in [(0, 0.0), (48, 0.00390625), (-48, 0.00390625), (96, 0.0625), (-96, 0.0625),
                    (144, 0.31640625), (-144, 0.31640625), (192, 1), (-192, 1)]:
        assert nb.closest_approach_reward(-1, d) == pytest.approx(want, abs=1e-15)
    assert nb.closest_approach_reward(+1, 0.0) == 1
    for d, want in [(0, 1.0), (48, 0.9422581744350746), (-48, 0.9422581744350746), (96, 0.8807388571985678),
                    (192, 0.7425643872142526), (856 / 4, 0.7071067811865476), (856 / 3, 0.5773502691896258),
                    (856 / 2, 0.0), (-856 / 2, 0.0), (500, 0.0)]:
        assert nb.plan_deviation_reward(d) == pytest.approx(want, abs=1e-15)


def test_python_modulo_semantics(orc):
    for x in [0.0, 1.5, 359.999, 360.0, 360.5, 719.0, 725.25, -0.25, -1e-18, -360.0, -361.0, -725.5, 1e6 + 0.1]:
        assert orc.pymod(x, 360.0) == x % 360.0, x
    assert orc.pymod(-1e-18, 360.0) == 360.0           # the famous "result == divisor" case


def test_pyport_matches_c_oracle(orc):
    """The Python port (CPU-baseline stand-in) and the C oracle are the same function."""
    rng = random.Random(7)
    arng = np.random.default_rng(7)
    for _ in range(3):
        g = PyPortGame(rng)
        st = orc.new_state(1)
        st["player"][0] = (g.player.x, g.player.y, g.player.v_air, g.player.psi, 0.0)
        st["traffic"][0, 0] = (g.traffic[0].x, g.traffic[0].y, g.traffic[0].v_air, g.traffic[0].psi)
        o0 = g.observe(); o1 = orc.observe(st)
        assert np.abs(o0 - o1[0]).max() < 1e-14
        for k in range(1200):
            a = float(np.float32(arng.uniform(-1, 1)))
            po, pr, pd, _ = g.step(np.array([a]))
            co, cr, cf, coc = orc.step(st, np.array([a]))
            assert bool(cf[0] & FLAG_DONE) == pd
            assert np.nanmax(np.abs(po - co[0])) < 1e-13 and abs(pr - cr[0]) < 1e-12
            if pd:
                assert g.outcome == coc[0] and g.steps == st["steps"][0]
                assert abs(g.total_reward - st["total_reward"][0]) < 1e-10
                assert abs(g.d_path - st["d_path"][0]) < 1e-9
                break


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("n", [1, 3])
def test_oracle_matches_live_reference(n):
    """Random-action episodes through the unmodified reference and the oracle, same spawn stream."""
    ref_shim.load(n)
    try:
        from gym_ACAS2D.envs.environment import ACAS2DEnv
        import gym_ACAS2D.settings as rs
        consts = {k: getattr(rs, k) for k in DEFAULTS}
        orc = Oracle(n)
        random.seed(99)
        arng = np.random.default_rng(99)
        with ref_shim.quiet():
            env = ACAS2DEnv()
            for ep in range(4):
                obs = env.reset()
                g = env.game
                st = orc.new_state(1)
                st["player"][0] = (g.player.x, g.player.y, g.player.v_air, g.player.psi, 0.0)
                for i in range(n):
                    t = g.traffic[i]
                    st["traffic"][0, i] = (t.x, t.y, t.v_air, t.psi)
                assert np.abs(orc.observe(st)[0] - obs).max() < 1e-14
                for k in range(1100):
                    a = float(np.float32(arng.uniform(-1, 1)))
                    ro, rr, rd, _ = env.step(np.array([a]))
                    co, cr, cf, coc = orc.step(st, np.array([a]))
                    assert bool(cf[0] & FLAG_DONE) == rd
                    assert np.nanmax(np.abs(ro - co[0])) < 1e-13 and abs(rr - cr[0]) < 1e-12
                    assert (st["player"][0, 0], st["player"][0, 1], st["player"][0, 3]) == (g.player.x, g.player.y, g.player.psi)
                    if rd:
                        assert g.outcome == coc[0] and g.steps == st["steps"][0]
                        break
        assert consts == {k: DEFAULTS[k] if k not in ("MIN_TRAFFIC", "MAX_TRAFFIC") else n for k in DEFAULTS}
    finally:
        ref_shim.unload()


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference_with_unequal_speeds():
    """Q3 (kinematics.py:74): the closing-speed look-ahead multiplies the INTRUDER's sine by the PLAYER's
    airspeed.  Invisible at the default speed factors (1..1); here the unmodified reference runs with
    AIRSPEED_FACTOR_MIN / MAX = 0.6 / 1.4 so that intruders fly at other speeds than the player."""
    over = dict(AIRSPEED_FACTOR_MIN=0.6, AIRSPEED_FACTOR_MAX=1.4)
    ref_shim.load(2, **over)
    try:
        from gym_ACAS2D.envs.environment import ACAS2DEnv
        orc = Oracle(2, **over)
        random.seed(7)
        arng = np.random.default_rng(7)
        speeds = []
        with ref_shim.quiet():
            env = ACAS2DEnv()
            for ep in range(4):
                obs = env.reset()
                g = env.game
                st = orc.new_state(1)
                st["player"][0] = (g.player.x, g.player.y, g.player.v_air, g.player.psi, 0.0)
                for i in range(2):
                    t = g.traffic[i]
                    st["traffic"][0, i] = (t.x, t.y, t.v_air, t.psi)
                    speeds.append(t.v_air)
                assert np.abs(orc.observe(st)[0] - obs).max() < 1e-14
                for k in range(1100):
                    a = float(np.float32(arng.uniform(-1, 1)))
                    ro, rr, rd, _ = env.step(np.array([a]))
                    co, cr, cf, coc = orc.step(st, np.array([a]))
                    assert bool(cf[0] & FLAG_DONE) == rd
                    assert np.nanmax(np.abs(ro - co[0])) < 1e-13 and abs(rr - cr[0]) < 1e-12
                    if rd:
                        assert g.outcome == coc[0] and g.steps == st["steps"][0]
                        break
        assert max(abs(v - 200.0) for v in speeds) > 20.0           # the intruders really flew at other speeds
    finally:
        ref_shim.unload()
