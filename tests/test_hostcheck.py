"""CPU check of the PRODUCT's per-environment source: ``csrc/acas2d_env.cuh`` and
``csrc/acas2d_math.cuh`` (the __host__ __device__ bodies every CUDA kernel inlines) are
compiled with g++ by ``tests/hostcheck`` and compared with the float64 oracle and the
reference fixtures.  This is not a product path -- the library has no CPU fallback -- it is
how the step logic is verified on the GPU-less build machine.  The same comparisons run
against the real kernels in ``test_gpu_parity.py``."""
import os

import numpy as np
import pytest

from oracle.acas2d_oracle import FLAG_DONE, Oracle
from tests import parity
from tests.hostcheck import HostBatch


@pytest.mark.parametrize("n,variant", [(1, 0), (1, 1), (8, 1)])
def test_fixture_rollouts(golden_dir, n, variant):
    """Injected states x action sequences stepped by the reference itself (fixture)."""
    g = np.load(os.path.join(golden_dir, f"ref_rollouts_n{n}.npz"))
    B, T = g["player0"].shape[0], g["actions"].shape[0]
    hb = HostBatch(B, n, auto_reset=False, variant=variant)
    parity.inject_from_fixture(hb, g)
    rep = parity.ParityReport()
    alive = np.ones(B, bool)
    orc = Oracle(n)                      # fills the rows the strided fixture does not carry
    st = orc.new_state(B)
    st["player"][:] = g["player0"]; st["traffic"][:] = g["traffic0"]
    st["steps"][:] = g["steps0"]; st["total_reward"][:] = g["total0"]
    ref = orc.rollout(st, g["actions"].astype(np.float64))
    for t in range(T):
        obs, rew, _ = hb.step(g["actions"][t])
        fobs, frew, fflags = parity.fixture_rows(g, t)
        assert np.array_equal(ref["flags"][t], fflags)
        parity.compare_step(rep, obs, rew, hb.flags, fobs if fobs is not None else ref["obs"][t], frew, fflags, alive)
        if t % int(g["stride"]) == 0:
            ex = hb.extract_state()
            assert np.abs(ex["player"][alive] - g["player_strided"][t // int(g["stride"])][alive]).max(initial=0) < parity.TOL_POS
            assert np.abs(ex["traffic"][alive][:, :, :2] - g["traffic_strided"][t // int(g["stride"])][alive]).max(initial=0) < parity.TOL_POS
        newly = alive & (fflags & FLAG_DONE > 0)
        if newly.any():
            assert np.array_equal(hb.outcome[newly], g["outcome"][t][newly])
            assert np.array_equal(hb.ep_length[newly], g["steps"][newly])
            assert np.abs(hb.ep_return[newly] - g["total_reward"][newly]).max() < parity.TOL_RETURN
        alive &= ~(fflags & FLAG_DONE > 0)
    parity.assert_flags_exact(rep)
    assert rep.steps > 3000


@pytest.mark.parametrize("n,variant,steps", [(1, 0, 1300), (1, 1, 400), (4, 1, 500)])
def test_auto_reset_matches_oracle_vec_step(n, variant, steps):
    """SB3-style auto-reset with Philox respawns: whole rollouts incl. resets, episode bookkeeping."""
    B, seed, off = 96, 13, 1000
    hb = HostBatch(B, n, seed=seed, env_id_offset=off, auto_reset=True, track_min_sep=True, variant=variant)
    orc = Oracle(n)
    st = orc.new_state(B)
    orc.spawn_philox(st, seed, off)
    ref_obs0 = orc.observe(st)
    obs0 = hb.reset()
    assert np.abs(obs0 - ref_obs0).max() < parity.TOL_OBS_CPA
    ex = hb.extract_state()
    assert np.abs(ex["player"] - st["player"][:, [0, 1, 3]]).max() < 1e-12
    assert np.abs(ex["traffic"] - st["traffic"]).max() < 1e-12
    rng = np.random.default_rng(5)
    rep = parity.ParityReport()
    episodes = 0
    ret_sum = 0.0
    for t in range(steps):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        obs, rew, done = hb.step(a)
        o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, a.astype(np.float64), seed, off)
        d = f & FLAG_DONE > 0
        assert np.array_equal(done, d)
        # terminal rows are compared through term_obs, the rest through obs
        cmp_obs = np.where(d[:, None], hb.term_obs, obs)
        ref_cmp = np.where(d[:, None], term, o)
        parity.compare_step(rep, cmp_obs, rew, hb.flags, ref_cmp, r, f)
        if d.any():
            assert np.abs(obs[d] - o[d]).max() < parity.TOL_OBS_CPA            # reset observation
            assert np.array_equal(hb.outcome[d], oc[d])
            assert np.array_equal(hb.ep_length[d], ep_len[d])
            assert np.abs(hb.ep_return[d] - ep_ret[d]).max() < parity.TOL_RETURN
            episodes += int(d.sum()); ret_sum += float(ep_ret[d].sum())
    parity.assert_flags_exact(rep)
    ex = hb.extract_state()
    assert np.array_equal(ex["episode_idx"], st["episode_idx"] + 1)       # count of started episodes
    assert np.array_equal(ex["steps"], st["steps"])
    assert np.abs(ex["player"] - st["player"][:, [0, 1, 3]]).max() < parity.TOL_POS
    assert np.abs(ex["min_sep"] - st["min_sep"]).max() < 1e-3
    c = hb.episode_counters()
    assert c[0] == episodes and episodes > 10
    assert c[1] + c[2] + c[3] == episodes
    assert abs(c[5] / 1048576.0 - ret_sum) < 1e-2 * max(1, episodes) ** 0.5 + episodes * 1e-6


def test_fused_rollout_equals_stepwise():
    """acas2d_rollout_random (K steps in registers, in-kernel Philox actions) == K single steps
    fed by acas2d_random_actions."""
    B, K = 64, 700
    a = HostBatch(B, 1, seed=3, auto_reset=True)
    b = HostBatch(B, 1, seed=3, auto_reset=True)
    a.reset(); b.reset()
    rs = np.zeros(B, np.float32)
    a.rollout_random(K, action_seed=77, step0=5, reward_sum=rs)
    acc = np.zeros(B, np.float64)
    for k in range(K):
        _, r, _ = b.step(b.random_actions(5 + k, action_seed=77))
        acc += r
    ea, eb = a.extract_state(), b.extract_state()
    for key in ("player", "traffic", "steps", "episode_idx"):
        assert np.array_equal(ea[key], eb[key]), key
    assert np.array_equal(a.episode_counters(), b.episode_counters())
    assert np.abs(rs - acc).max() < 0.5          # float32 running sum vs float64 sum of float32 terms
    acts = b.random_actions(0, action_seed=77)
    assert acts.min() >= -1 and acts.max() <= 1 and abs(acts.mean()) < 0.3


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    hb = HostBatch(1, 1)
    import ctypes
    def px(ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        hb.lib.hostcheck_philox(c, k, o)
        return tuple(int(v) for v in o)
    orc = Oracle(1)
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        assert px(ctr, key) == want
        assert orc.philox(ctr, key) == want


def test_wrap360_is_python_modulo():
    hb = HostBatch(1, 1)
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(-2000, 2000, 2000)) + [0.0, 360.0, 720.0, -360.0, -720.0, 359.99999999999994,
                                                  -1e-18, -1e-300, 1e9 + 0.5, -1e9 - 0.5, 719.9999999999999]
    for x in xs:
        assert hb.lib.hostcheck_wrap360(x) == x % 360.0, x


def test_sincos_deg_against_50_digit_values():
    """sin / cos of a heading in degrees (csrc/acas2d_math.cuh sincos_deg, the one routine the float64 flag
    chain's accuracy rests on): |error| <= 1.5e-16 against 50-digit Decimal values, over the headings episodes
    visit, multiples of 45/128 degrees and their midpoints, and far outside [0, 360]."""
    import ctypes
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    pi = Decimal("3.14159265358979323846264338327950288419716939937510582097494459230781640628620899")

    def true_sincos(deg):
        x = Decimal(deg) * pi / 180                      # Decimal(float) is exact
        x = x - (x / (2 * pi)).to_integral_value() * 2 * pi
        s = c = Decimal(0)
        ts, tc, n, x2 = x, Decimal(1), 0, x * x
        while abs(ts) > Decimal(10) ** -50 or abs(tc) > Decimal(10) ** -50:
            c += tc; s += ts; n += 1
            tc = -tc * x2 / ((2 * n - 1) * (2 * n)); ts = -ts * x2 / ((2 * n) * (2 * n + 1))
        return s, c

    hb = HostBatch(1, 1)
    f = hb.lib.hostcheck_sincos_deg
    f.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    rng = np.random.default_rng(1)
    xs = list(rng.uniform(0, 360, 1500)) + list(rng.uniform(-2000, 2000, 300)) + [k * 0.3515625 for k in range(0, 1025, 37)] + \
        [(k + 0.5) * 0.3515625 for k in range(0, 1024, 41)] + [0.0, 90.0, 180.0, 270.0, 360.0, 359.99999999999994, 1e-300, -1e-18,
                                                               float(np.float32(136.41722591475224)), 45.0, 44.99999999999999]
    worst = 0.0
    for x in xs:
        s, c = ctypes.c_double(), ctypes.c_double()
        f(x, ctypes.byref(s), ctypes.byref(c))
        ts, tc = true_sincos(x)
        worst = max(worst, abs(float(Decimal(s.value) - ts)), abs(float(Decimal(c.value) - tc)))
    assert worst <= 1.5e-16, worst
    for x, (ws, wc) in [(0.0, (0.0, 1.0)), (90.0, (1.0, 0.0)), (180.0, (0.0, -1.0)), (270.0, (-1.0, 0.0)), (360.0, (0.0, 1.0))]:
        s, c = ctypes.c_double(), ctypes.c_double()
        f(x, ctypes.byref(s), ctypes.byref(c))
        assert (s.value, c.value) == (ws, wc), x         # the axes are exact


def test_inject_extract_roundtrip():
    rng = np.random.default_rng(1)
    B, N = 33, 5
    hb = HostBatch(B, N, auto_reset=False)
    pl = np.c_[rng.uniform(0, 1600, B), rng.uniform(0, 1000, B), rng.uniform(0, 360, B)]
    tr = np.stack([rng.uniform(0, 1600, (B, N)), rng.uniform(0, 1000, (B, N)),
                   np.full((B, N), 200.0), rng.uniform(0, 360, (B, N))], -1)
    steps = rng.integers(1, 1000, B).astype(np.int32)
    tot = rng.uniform(-5, 300, B)
    hb.inject_state(pl, tr, steps, tot)
    ex = hb.extract_state()
    assert np.array_equal(ex["player"], pl) and np.array_equal(ex["steps"], steps)
    assert np.abs(ex["traffic"] - tr).max() < 1e-9
    assert np.abs(ex["total_reward"] - tot).max() < 2e-5


def test_spawn_distribution_matches_reference_rules():
    """game.py:85-116: player at (48, 500) heading within +-3 deg of the goal bearing; intruder 0 at
    (1552, 48|952) heading 145+70*sd +- 15; others uniform in [0,1576]x[0,600], heading U(0,360)."""
    B, N = 4096, 3
    hb = HostBatch(B, N, seed=2)
    hb.reset()
    ex = hb.extract_state()
    pl, tr = ex["player"], ex["traffic"]
    assert np.all(pl[:, 0] == 48) and np.all(pl[:, 1] == 500)
    h = (pl[:, 2] + 180) % 360 - 180
    assert h.min() >= -3 and h.max() <= 3 and abs(h.mean()) < 0.15 and h.std() > 1.5
    assert np.all(tr[:, 0, 0] == 1552) and set(np.unique(tr[:, 0, 1])) == {48.0, 952.0}
    down = tr[:, 0, 1] == 952
    assert 0.45 < down.mean() < 0.55
    assert np.all(np.abs(tr[~down, 0, 3] - 145) <= 15) and np.all(np.abs(tr[down, 0, 3] - 215) <= 15)
    assert np.all(tr[:, :, 2] == 200)
    rest = tr[:, 1:]
    assert rest[..., 0].min() >= 0 and rest[..., 0].max() <= 1576 and rest[..., 1].max() <= 600
    assert abs(rest[..., 0].mean() - 788) < 20 and abs(rest[..., 1].mean() - 300) < 10
    assert abs(rest[..., 3].mean() - 180) < 6


def test_zero_action_outcome_rates_match_reference_baseline():
    """Distribution-level check of the Philox spawns.  The reference's zero-action baseline CSV
    (100 episodes) has 42 % goal / mean game.steps 494.65, but that is a small sample: 4000
    episodes spawned with the reference's own draw order on Python's MT19937 and stepped by the
    oracle give goal 54.6 %, mean steps 525.6 (binomial sigma 0.8 %).  Philox spawns draw from the
    same distributions, so a large batch must land on those rates."""
    B = 4000
    hb = HostBatch(B, 1, seed=13, auto_reset=True)
    hb.reset()
    zero = np.zeros(B, np.float32)
    for _ in range(760):                     # every first episode is over by then
        hb.step(zero)
    c = hb.episode_counters()
    n = c[0]
    assert n >= B
    assert c[3] == 0                                     # nobody times out flying straight
    assert 0.50 < c[1] / n < 0.59                        # goal rate (reference spawn rules: 0.546)
    assert 505 < c[4] / n < 545                          # mean game.steps (reference spawn rules: 525.6)


def test_unclipped_actions_general_wrap_path():
    """Q19 on the host build: large |action| exercises wrap360's fmod path and the full-range look-ahead."""
    rng = np.random.default_rng(3)
    B = 64
    hb = HostBatch(B, 1, auto_reset=False)
    pl = np.c_[rng.uniform(100, 1300, B), rng.uniform(100, 900, B), rng.uniform(0, 360, B)]
    tr = np.stack([rng.uniform(200, 1580, (B, 1)), rng.uniform(20, 980, (B, 1)), np.full((B, 1), 200.0), rng.uniform(0, 360, (B, 1))], -1)
    hb.inject_state(pl, tr)
    orc = Oracle(1); st = orc.new_state(B)
    st["player"][:, [0, 1, 3]] = pl; st["player"][:, 2] = 200.0; st["traffic"][:] = tr; st["steps"][:] = 1
    rep = parity.ParityReport(); alive = np.ones(B, bool)
    for t in range(40):
        a = (rng.uniform(-1, 1, B) * rng.choice([1.0, 50.0, 800.0, 5000.0], B)).astype(np.float32)
        obs, rew, _ = hb.step(a)
        o, r, f, oc = orc.step(st, a.astype(np.float64))
        parity.compare_step(rep, obs, rew, hb.flags, o, r, f, alive)
        alive &= ~(f & FLAG_DONE > 0)
    assert rep.flag_mismatch == 0


def test_observe_is_the_reset_observation_of_an_injected_state():
    """acas2d_observe: game.observe() without the counter increment == the oracle's observe() of the same state."""
    rng = np.random.default_rng(5)
    B, N = 40, 3
    hb = HostBatch(B, N, auto_reset=False)
    pl = np.c_[rng.uniform(0, 1600, B), rng.uniform(0, 1000, B), rng.uniform(0, 360, B)]
    tr = np.stack([rng.uniform(0, 1600, (B, N)), rng.uniform(0, 1000, (B, N)), np.full((B, N), 200.0), rng.uniform(0, 360, (B, N))], -1)
    steps = rng.integers(1, 900, B).astype(np.int32)
    hb.inject_state(pl, tr, steps)
    orc = Oracle(N); st = orc.new_state(B)
    st["player"][:, [0, 1, 3]] = pl; st["player"][:, 2] = 200.0; st["traffic"][:] = tr; st["steps"][:] = steps - 1
    ref = orc.observe(st)                               # increments to `steps`
    got = hb.observe().astype(np.float64)
    assert np.abs(got[:, :5] - ref[:, :5]).max() < parity.TOL_OBS_BASE
    assert np.nanmax(np.abs(got[:, 5:] - ref[:, 5:])) < parity.TOL_OBS_CPA
    assert np.array_equal(hb.extract_state()["steps"], steps)       # state untouched


@pytest.mark.parametrize("n,variant", [(1, 0), (3, 1)])
def test_unequal_speeds_q3(n, variant):
    """Q3 in the product's per-env source: intruders at 0.6 .. 1.4 x the player's airspeed (the oracle is pinned
    to the unmodified reference for this case by test_oracle_matches_live_reference_with_unequal_speeds).
    Spawned and injected states, auto-reset, flags bit-exact, observations / rewards within the tolerances."""
    over = dict(AIRSPEED_FACTOR_MIN=0.6, AIRSPEED_FACTOR_MAX=1.4)
    B, seed, off = 64, 29, 500
    hb = HostBatch(B, n, seed=seed, env_id_offset=off, auto_reset=True, variant=variant, **over)
    orc = Oracle(n, **over)
    st = orc.new_state(B)
    orc.spawn_philox(st, seed, off)
    assert np.abs(hb.reset() - orc.observe(st)).max() < parity.TOL_OBS_CPA
    assert np.abs(st["traffic"][:, :, 2] - 200.0).max() > 50.0
    # half of the envs continue from injected float64 states with arbitrary intruder speeds (residual path)
    rng = np.random.default_rng(3)
    ex = hb.extract_state()
    ex["traffic"][::2, :, 2] = rng.uniform(90.0, 310.0, ex["traffic"][::2, :, 2].shape)
    hb.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
    st["traffic"][::2, :, 2] = ex["traffic"][::2, :, 2]
    rep = parity.ParityReport()
    for t in range(500):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        obs, rew, done = hb.step(a)
        o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, a.astype(np.float64), seed, off)
        d = f & FLAG_DONE > 0
        assert np.array_equal(done, d)
        parity.compare_step(rep, np.where(d[:, None], hb.term_obs, obs), rew, hb.flags, np.where(d[:, None], term, o), r, f)
    parity.assert_flags_exact(rep)
    assert rep.steps == 500 * B
