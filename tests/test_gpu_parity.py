"""Parity tests proper: the CUDA path, called through the C ABI (``BatchedACAS2D`` -> ctypes ->
``libacas2d_b200.so``), against the float64 oracle and the reference fixtures.  Tolerances are
stated in ``tests/parity.py``; flags, outcomes and episode lengths must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle.acas2d_oracle import FLAG_DONE, Oracle
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import _native
    _native.load()
    return torch.device("cuda", 0)


def make(B, n=1, **kw):
    from gym_ACAS2D.envs import BatchedACAS2D
    return BatchedACAS2D(B, n_traffic=n, device="cuda:0", **kw)


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("n", [1, 8])
def test_reference_fixture_rollouts(cuda, golden_dir, n):
    """Injected states x action sequences recorded from the unmodified reference."""
    g = np.load(os.path.join(golden_dir, f"ref_rollouts_n{n}.npz"))
    B, T = g["player0"].shape[0], g["actions"].shape[0]
    env = make(B, n, auto_reset=False)
    parity.inject_from_fixture(env, g)
    orc = Oracle(n)
    st = orc.new_state(B)
    st["player"][:] = g["player0"]; st["traffic"][:] = g["traffic0"]
    st["steps"][:] = g["steps0"]; st["total_reward"][:] = g["total0"]
    ref = orc.rollout(st, g["actions"].astype(np.float64))
    rep = parity.ParityReport()
    alive = np.ones(B, bool)
    s = int(g["stride"])
    acts = torch.from_numpy(g["actions"]).cuda()
    for t in range(T):
        obs, rew, _ = env.step(acts[t])
        fobs, frew, fflags = parity.fixture_rows(g, t)
        parity.compare_step(rep, npy(obs), npy(rew), npy(env.flags), fobs if fobs is not None else ref["obs"][t],
                            frew, fflags, alive)
        if t % (4 * s) == 0:
            ex = env.extract_state()
            assert np.abs(ex["player"][alive] - g["player_strided"][t // s][alive]).max(initial=0) < parity.TOL_POS
            assert np.abs(ex["traffic"][alive][:, :, :2] - g["traffic_strided"][t // s][alive]).max(initial=0) < parity.TOL_POS
        newly = alive & (fflags & FLAG_DONE > 0)
        if newly.any():
            assert np.array_equal(npy(env.outcome)[newly], g["outcome"][t][newly])
            assert np.array_equal(npy(env.ep_length)[newly], g["steps"][newly])
            assert np.abs(npy(env.ep_return)[newly] - g["total_reward"][newly]).max() < parity.TOL_RETURN
        alive &= ~(fflags & FLAG_DONE > 0)
    parity.assert_flags_exact(rep)
    assert rep.steps > 3000


def _random_states(rng, B, N):
    """Random mid-air states + a share of threshold-grazing ones (collision / goal / timeout)."""
    pl = np.c_[rng.uniform(30, 1300, B), rng.uniform(100, 900, B), rng.uniform(-40, 40, B) % 360]
    tr = np.stack([rng.uniform(200, 1580, (B, N)), rng.uniform(20, 980, (B, N)),
                   np.full((B, N), 200.0), rng.uniform(0, 360, (B, N))], -1)
    steps = np.ones(B, np.int32)
    k = B // 4
    ang = rng.uniform(0, 2 * np.pi, k); d0 = rng.uniform(96.2, 130, k)          # closing head-on, just outside 96 px
    tr[:k, 0, 0] = pl[:k, 0] + d0 * np.cos(ang); tr[:k, 0, 1] = pl[:k, 1] + d0 * np.sin(ang)
    pl[:k, 2] = np.degrees(ang) % 360; tr[:k, 0, 3] = (np.degrees(ang) + 180) % 360
    ang = rng.uniform(0, 2 * np.pi, k); d0 = rng.uniform(144.2, 200, k)         # flying into the goal disc
    pl[k:2 * k, 0] = 1456 - d0 * np.cos(ang); pl[k:2 * k, 1] = 500 - d0 * np.sin(ang); pl[k:2 * k, 2] = np.degrees(ang) % 360
    steps[2 * k:2 * k + k // 2] = rng.integers(900, 1001, k // 2)               # about to time out
    return pl, tr, steps


def test_config2_4096_envs_1001_steps_vs_oracle(cuda):
    """BASELINE config 2: 4096 batched envs, default N_TRAFFIC, random actions, injected states,
    1001 steps with auto-reset: every flag of ~4.1 M env-steps bit-exact, floats within tolerance."""
    B, N, T, seed = 4096, 1, 1001, 13
    rng = np.random.default_rng(42)
    pl, tr, steps = _random_states(rng, B, N)
    env = make(B, N, seed=seed, auto_reset=True, track_min_sep=True)
    env.reset()
    env.inject_state(pl, tr, steps)
    orc = Oracle(N)
    st = orc.new_state(B)
    orc.spawn_philox(st, seed, 0)                      # episode 0 consumed, like env.reset()
    st["player"][:, [0, 1, 3]] = pl; st["player"][:, 2] = 200.0; st["traffic"][:] = tr; st["steps"][:] = steps
    d = tr[:, :, :2] - pl[:, None, :2]
    st["min_sep"][:] = np.sqrt((d * d).sum(-1)).min(-1)
    rep = parity.ParityReport()
    acts = rng.uniform(-1, 1, (T, B)).astype(np.float32)
    acts_d = torch.from_numpy(acts).cuda()
    episodes = 0
    for t in range(T):
        obs, rew, done = env.step(acts_d[t])
        o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, acts[t].astype(np.float64), seed, 0)
        dn = f & FLAG_DONE > 0
        assert np.array_equal(npy(done), dn)
        obs_n = npy(obs)
        cmp_obs = np.where(dn[:, None], npy(env.term_obs), obs_n)
        parity.compare_step(rep, cmp_obs, npy(rew), npy(env.flags), np.where(dn[:, None], term, o), r, f)
        if dn.any():
            assert np.abs(obs_n[dn] - o[dn]).max() < parity.TOL_OBS_CPA
            assert np.array_equal(npy(env.outcome)[dn], oc[dn])
            assert np.array_equal(npy(env.ep_length)[dn], ep_len[dn])
            assert np.abs(npy(env.ep_return)[dn] - ep_ret[dn]).max() < parity.TOL_RETURN
            episodes += int(dn.sum())
    parity.assert_flags_exact(rep)
    assert rep.steps == B * T
    ex = env.extract_state()
    assert np.array_equal(ex["steps"], st["steps"])
    assert np.array_equal(ex["episode_idx"], st["episode_idx"] + 1)
    assert np.abs(ex["player"] - st["player"][:, [0, 1, 3]]).max() < parity.TOL_POS
    assert np.abs(ex["traffic"][:, :, :2] - st["traffic"][:, :, :2]).max() < parity.TOL_POS
    assert np.abs(ex["min_sep"] - st["min_sep"]).max() < 1e-3
    c = npy(env.episode_counters())
    assert c[0] == episodes and c[1] + c[2] + c[3] == episodes and episodes > B
    print(rep)


@pytest.mark.parametrize("n,B,T", [(2, 512, 300), (8, 512, 300), (64, 128, 120), (256, 64, 40)])
def test_traffic_sweep_vs_oracle(cuda, n, B, T):
    """BASELINE config 4 traffic counts (plus N=2): auto-reset rollouts vs the oracle."""
    seed = 5
    env = make(B, n, seed=seed, env_id_offset=77, auto_reset=True, track_min_sep=True)
    orc = Oracle(n)
    st = orc.new_state(B)
    orc.spawn_philox(st, seed, 77)
    ref0 = orc.observe(st)
    assert np.nanmax(np.abs(npy(env.reset()) - ref0)) < parity.TOL_OBS_CPA
    rng = np.random.default_rng(n)
    rep = parity.ParityReport()
    for t in range(T):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        obs, rew, done = env.step(torch.from_numpy(a).cuda())
        o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, a.astype(np.float64), seed, 77)
        dn = f & FLAG_DONE > 0
        assert np.array_equal(npy(done), dn)
        obs_n = npy(obs)
        parity.compare_step(rep, np.where(dn[:, None], npy(env.term_obs), obs_n), npy(rew), npy(env.flags),
                            np.where(dn[:, None], term, o), r, f)
        if dn.any():
            assert np.nanmax(np.abs(obs_n[dn] - o[dn])) < parity.TOL_OBS_CPA
            assert np.array_equal(npy(env.ep_length)[dn], ep_len[dn])
    assert rep.flag_mismatch == 0, rep
    ex = env.extract_state()
    assert np.array_equal(ex["episode_idx"], st["episode_idx"] + 1)
    assert np.abs(ex["traffic"][:, :, :2] - st["traffic"][:, :, :2]).max() < parity.TOL_POS


@pytest.mark.parametrize("n,T", [(8, 60), (16, 40), (32, 40), (64, 40), (128, 24), (256, 24)])
def test_config4_parity_at_stated_batch(cuda, n, T):
    """BASELINE config 4 AT ITS STATED BATCH: 65 536 envs, N_TRAFFIC = 8 / 64 / 256, auto-reset, random actions.
    Eight 64-env windows spread over the batch (first / last warps and CTAs, odd offsets) are stepped by the
    oracle with the same global env ids: flags bit-exact, floats within the stated tolerances.  The lean
    launch the benchmark times (no flags / terminal rows) must agree with the full one bit for bit over
    the WHOLE batch (reference per-intruder loop: game.py:185-189,205-210)."""
    B, seed, W = 65536, 11, 64
    full = make(B, n, seed=seed, auto_reset=True)
    lean = make(B, n, seed=seed, auto_reset=True)
    o_full = full.reset().clone()
    assert torch.equal(o_full.view(torch.int32), lean.reset().view(torch.int32))
    offsets = [0, 64, 4096 - 32, 20000 + 7, 32768 - 16, 40001, 65536 - 128 - 3, 65536 - 64]
    orc = Oracle(n)
    states = []
    for off in offsets:
        st = orc.new_state(W)
        orc.spawn_philox(st, seed, off)
        ref0 = orc.observe(st)
        assert np.nanmax(np.abs(npy(o_full[off:off + W]) - ref0)) < parity.TOL_OBS_CPA
        states.append(st)
    rep = parity.ParityReport()
    episodes = 0
    for t in range(T):
        a = full.random_actions(t, action_seed=9)
        of, rf, df = full.step(a)
        ol, rl, dl = lean.step(a, full_outputs=False)
        assert torch.equal(of.view(torch.int32), ol.view(torch.int32))
        assert torch.equal(rf.view(torch.int32), rl.view(torch.int32)) and torch.equal(df, dl)
        a_n, of_n, rf_n, ff_n, tf_n = npy(a), npy(of), npy(rf), npy(full.flags), npy(full.term_obs)
        for off, st in zip(offsets, states):
            sl = slice(off, off + W)
            o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, a_n[sl].astype(np.float64), seed, off)
            dn = f & FLAG_DONE > 0
            assert np.array_equal(npy(df[sl]), dn)
            parity.compare_step(rep, np.where(dn[:, None], tf_n[sl], of_n[sl]), rf_n[sl], ff_n[sl],
                                np.where(dn[:, None], term, o), r, f)
            if dn.any():
                assert np.nanmax(np.abs(of_n[sl][dn] - o[dn])) < parity.TOL_OBS_CPA
                assert np.array_equal(npy(full.ep_length[sl])[dn], ep_len[dn])
                assert np.array_equal(npy(full.outcome[sl])[dn], oc[dn])
                episodes += int(dn.sum())
    assert rep.flag_mismatch == 0 and rep.steps == T * W * len(offsets), rep
    assert torch.equal(full.ppos, lean.ppos) and torch.equal(full.paux, lean.paux) and torch.equal(full.thot, lean.thot)
    assert torch.equal(full.episode_counters(), lean.episode_counters())
    ex = full.extract_state()
    for off, st in zip(offsets, states):
        assert np.array_equal(ex["episode_idx"][off:off + W], st["episode_idx"] + 1)
        assert np.abs(ex["traffic"][off:off + W, :, :2] - st["traffic"][:, :, :2]).max() < parity.TOL_POS
        assert np.abs(ex["player"][off:off + W] - st["player"][:, [0, 1, 3]]).max() < parity.TOL_POS
    assert episodes > (0 if n == 8 else 100)
    print(rep)


@pytest.mark.parametrize("n", [2, 3, 8, 12, 16, 24, 64, 96, 256])
def test_tiled_kernel_equals_loop_kernel(cuda, n):
    """The shared-memory tiled kernel (G lanes per env, cp.async staging, shuffle reductions) and the
    one-thread-per-env kernel are the same function: bit-identical state and outputs."""
    from gym_ACAS2D.envs import _native
    lib = _native.load()
    B = 1000 if n <= 24 else 150                  # not a multiple of the envs-per-warp: ragged last warp
    a = make(B, n, seed=21, auto_reset=True, track_min_sep=True)
    b = make(B, n, seed=21, auto_reset=True, track_min_sep=True)
    a.reset(); b.reset()
    try:
        for t in range(60):
            act = a.random_actions(t, 5)
            lib.acas2d_set_tuning(0, 0)
            oa, ra, da = a.step(act)
            lib.acas2d_set_tuning(0, 1)
            ob, rb, db = b.step(act)
            assert torch.equal(oa.view(torch.int32), ob.view(torch.int32))           # NaNs included
            assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(a.flags, b.flags)
            assert torch.equal(a.term_obs[da].view(torch.int32), b.term_obs[db].view(torch.int32))
    finally:
        lib.acas2d_set_tuning(0, 0)
    assert torch.equal(a.ppos, b.ppos) and torch.equal(a.paux, b.paux) and torch.equal(a.thot, b.thot)
    assert torch.equal(a.min_sep, b.min_sep) and torch.equal(a.episode_idx, b.episode_idx)
    assert torch.equal(a.episode_counters(), b.episode_counters())


@pytest.mark.parametrize("stages,occ", [(2, 2), (2, 4), (3, 1), (4, 3), (5, 3)])
def test_tma_ring_kernel_equals_direct_kernel(cuda, stages, occ):
    """N_TRAFFIC == 1: the persistent kernel fed by the TMA bulk-copy ring and the direct
    one-thread-per-env kernel are the same function, bit for bit -- ragged batch (B % 256 != 0),
    injected float64 states (residual path) and auto-resets included."""
    from gym_ACAS2D.envs import _native
    lib = _native.load()
    B = 148 * 4 * 256 * 2 + 256 * 3 + 77          # several tiles per CTA, then a ragged tail
    a = make(B, 1, seed=4, auto_reset=True); b = make(B, 1, seed=4, auto_reset=True)
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    for e in (a, b):
        ex = e.extract_state()
        ex["steps"][:] = 1 + (np.arange(B) % 1000)
        ex["traffic"][::3, 0, 0] += 0.123456789        # not float32-representable -> residual path
        e.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
    try:
        for t in range(30):
            act = a.random_actions(t, 9)
            lib.acas2d_set_tuning(occ, -1); lib.acas2d_set_n1_kernel(1, stages)
            oa, ra, da = a.step(act)
            lib.acas2d_set_n1_kernel(0, 0)
            ob, rb, db = b.step(act)
            assert torch.equal(oa.view(torch.int32), ob.view(torch.int32))
            assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(a.flags, b.flags)
            assert torch.equal(a.term_obs[da], b.term_obs[db]) and torch.equal(a.ep_length[da], b.ep_length[db])
    finally:
        lib.acas2d_set_tuning(2, -1); lib.acas2d_set_n1_kernel(1, 2)
    assert torch.equal(a.ppos, b.ppos) and torch.equal(a.paux, b.paux) and torch.equal(a.thot, b.thot)
    assert torch.equal(a.episode_idx, b.episode_idx) and torch.equal(a.episode_counters(), b.episode_counters())
    assert a.episode_counters()[0].item() > 1000


def test_deferred_respawn_queue_overflow(cuda):
    """The persistent N == 1 kernel queues finished envs per CTA (256 slots) and respawns them after its last tile;
    more than that in one launch -- here EVERY env of a 400 000-env batch times out on the same step, ~1350 per
    CTA -- must take the in-line path and still equal the direct kernel bit for bit; then a second mass ending of
    the freshly respawned (compact-record) games by collision-free timeouts again."""
    from gym_ACAS2D.envs import _native
    lib = _native.load()
    B = 400_000 + 123
    a = make(B, 1, seed=6, auto_reset=True); b = make(B, 1, seed=6, auto_reset=True)
    a.reset(); b.reset()
    try:
        for rnd in range(2):
            for e in (a, b):
                ex = e.extract_state()
                ex["steps"][:] = 1000                      # the next step is step() call number 1000: timeout (Q5)
                if rnd == 1:
                    ex["steps"][::2] = 999                 # ... half of them one step later
                e.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
            for t in range(3):
                act = a.random_actions(10 * rnd + t, 2)
                lib.acas2d_set_n1_kernel(1, 2)
                oa, ra, da = a.step(act)
                lib.acas2d_set_n1_kernel(0, 0)
                ob, rb, db = b.step(act)
                assert torch.equal(oa.view(torch.int32), ob.view(torch.int32)) and torch.equal(ra, rb) and torch.equal(da, db)
                assert torch.equal(a.term_obs[da], b.term_obs[db]) and torch.equal(a.outcome[da], b.outcome[db])
                if t == 0:
                    assert int(da.sum()) >= (B if rnd == 0 else B // 2)       # all (half) time out; a few also collide / land
    finally:
        lib.acas2d_set_n1_kernel(1, 2)
    assert torch.equal(a.ppos, b.ppos) and torch.equal(a.paux, b.paux) and torch.equal(a.thot, b.thot)
    assert torch.equal(a.tpsi0, b.tpsi0) and torch.equal(a.episode_idx, b.episode_idx)
    assert torch.equal(a.episode_counters(), b.episode_counters()) and int(a.episode_counters()[0]) >= 2 * B


@pytest.mark.parametrize("stages,occ", [(2, 2), (5, 3)])
def test_tma_ring_stress_at_full_size(cuda, stages, occ):
    """Regression for a cross-proxy WAR race: at 4 Mi envs every CTA refills each ring stage ~10 times
    per launch; without `fence.proxy.async` before the barrier whole warps read the next tile's
    records (seen as 16-64 corrupted rows in ~1/3 of the steps).  60 steps must be bit-identical to
    the direct kernel."""
    from gym_ACAS2D.envs import _native
    lib = _native.load()
    B = 4 << 20
    a = make(B, 1, seed=4, auto_reset=True); b = make(B, 1, seed=4, auto_reset=True)
    a.reset(); b.reset()
    try:
        for t in range(60):
            act = a.random_actions(t, 9)
            lib.acas2d_set_tuning(occ, -1); lib.acas2d_set_n1_kernel(1, stages)
            a.step(act, full_outputs=False)
            lib.acas2d_set_n1_kernel(0, 0)
            b.step(act, full_outputs=False)
            assert torch.equal(a.ppos, b.ppos) and torch.equal(a.paux, b.paux), t
            assert torch.equal(a.obs.view(torch.int32), b.obs.view(torch.int32)) and torch.equal(a.reward, b.reward), t
    finally:
        lib.acas2d_set_tuning(2, -1); lib.acas2d_set_n1_kernel(1, 2)


def test_sharding_invariance_and_determinism(cuda):
    """1 Mi envs (BASELINE config 3 size): two half-batches addressed by global env id reproduce the
    full batch bit for bit (state, outputs, integer episode counters) -- the multi-GPU property."""
    B, T = 1 << 20, 24
    full = make(B, 1, seed=9, auto_reset=True)
    lo = make(B // 2, 1, seed=9, env_id_offset=0, auto_reset=True)
    hi = make(B // 2, 1, seed=9, env_id_offset=B // 2, auto_reset=True)
    o_full = full.reset().clone()
    assert torch.equal(o_full[: B // 2], lo.reset()) and torch.equal(o_full[B // 2:], hi.reset())
    # age the envs so that the short horizon crosses plenty of episode ends
    for e in (full, lo, hi):
        ex = e.extract_state()
        ex["steps"][:] = 1 + (np.arange(e.env_id_offset, e.env_id_offset + e.num_envs) % 997)
        e.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
    for t in range(T):
        a = full.random_actions(t, action_seed=4)
        o, r, d = full.step(a)
        o1, r1, d1 = lo.step(a[: B // 2])
        o2, r2, d2 = hi.step(a[B // 2:])
        assert torch.equal(o[: B // 2], o1) and torch.equal(o[B // 2:], o2)
        assert torch.equal(r[: B // 2], r1) and torch.equal(d[B // 2:], d2)
    assert torch.equal(full.episode_counters(), lo.episode_counters() + hi.episode_counters())
    assert full.episode_counters()[0].item() > 10000
    assert torch.equal(full.ppos[B // 2:], hi.ppos) and torch.equal(full.paux[: B // 2], lo.paux)
    # action stream: global-id addressed too
    assert torch.equal(full.random_actions(3, 8)[B // 2:], hi.random_actions(3, 8))


def test_properties_at_full_size(cuda):
    """Size-independent properties at 4 Mi envs (the bench batch): observation bounds of the declared
    Box (apart from Q5's 1.001), reward range, done == any flag, timeout exactly at game.steps 1001,
    straight flight covers AIRSPEED/FPS px per step."""
    B = 4 << 20
    env = make(B, 1, seed=1, auto_reset=True)
    obs = env.reset()
    assert obs.shape == (B, 8) and float(obs[:, 0].max()) == pytest.approx(0.001)
    p0 = env.ppos.clone()
    zero = torch.zeros(B, device="cuda")
    for _ in range(5):
        obs, rew, done = env.step(zero)
    moved = (env.ppos - p0).norm(dim=1)
    assert float((moved - 10.0).abs().max()) < 1e-9
    lo = torch.tensor([0, 0, -1, 0, 0, 0, -1, -1], device="cuda", dtype=torch.float32)
    for t in range(40):
        obs, rew, done = env.step(env.random_actions(t, 3))
        f = env.flags
        assert torch.equal(done, (f & 7) != 0)
        assert bool((obs >= lo).all()) and bool((obs <= 1.0011).all())
        shaped = rew - 1000.0 * ((f & 2) != 0) + 1000.0 * ((f & 1) != 0)
        assert float(shaped.min()) >= -2e-3 and float(shaped.max()) <= 1.0 + 1e-4
    from gym_ACAS2D.envs import _native
    ex_steps = env.paux.view(torch.int32)[:, 2] & _native.STEPS_MASK          # the word's top bits are record flags
    assert int(ex_steps.min()) >= 1 and int(ex_steps.max()) <= 1001


def test_edge_sizes(cuda):
    """Empty batch, single env, ragged sizes around the 256-env tile, and the maximum traffic count."""
    e0 = make(0, 1, auto_reset=True)
    assert e0.reset().shape == (0, 8)
    o, r, d = e0.step(torch.zeros(0, device="cuda"))
    assert o.shape == (0, 8) and r.shape == (0,) and d.shape == (0,)
    assert e0.episode_stats()["episodes"] == 0
    for B in (1, 31, 255, 256, 257, 513):
        a = make(B, 1, seed=2, auto_reset=True); orc = Oracle(1)
        st = orc.new_state(B); orc.spawn_philox(st, 2, 0); ref0 = orc.observe(st)
        assert np.abs(npy(a.reset()) - ref0).max() < parity.TOL_OBS_CPA
        for t in range(5):
            act = np.full(B, 0.25 * t - 0.5, np.float32)
            o, r, d = a.step(torch.from_numpy(act).cuda())
            ro, rr, rf, *_ = orc.vec_step(st, act.astype(np.float64), 2, 0)
            assert np.abs(npy(o) - ro).max() < parity.TOL_OBS_CPA and np.array_equal(npy(d), rf & FLAG_DONE > 0)
    from gym_ACAS2D.envs import _native
    big = make(3, _native.MAX_TRAFFIC, seed=1, auto_reset=True)           # 1024 intruders, obs row of 3077 floats
    orc = Oracle(_native.MAX_TRAFFIC); st = orc.new_state(3); orc.spawn_philox(st, 1, 0); ref0 = orc.observe(st)
    assert np.nanmax(np.abs(npy(big.reset()) - ref0)) < parity.TOL_OBS_CPA
    o, r, d = big.step(torch.zeros(3, device="cuda"))
    ro, rr, rf, *_ = orc.vec_step(st, np.zeros(3), 1, 0)
    assert np.nanmax(np.abs(npy(o) - ro)) < parity.TOL_OBS_CPA and np.array_equal(npy(d), rf & FLAG_DONE > 0)
    with pytest.raises(ValueError):
        make(4, _native.MAX_TRAFFIC + 1)
    with pytest.raises(ValueError):
        make(4, 0)                                                        # Q20: the reference needs >= 1 intruder
    with pytest.raises(ValueError):
        make(8, 1).step(torch.zeros(7, device="cuda"))                    # wrong number of actions


def test_unclipped_actions_follow_the_reference(cuda):
    """Q19: the env does not clip actions.  |a| up to 2000 turns the heading by up to 5 revolutions per
    step (general fmod path of the heading wrap, full-range look-ahead rotation)."""
    B, T = 512, 60
    rng = np.random.default_rng(11)
    pl, tr, steps = _random_states(rng, B, 1)
    env = make(B, 1, auto_reset=False); env.reset(); env.inject_state(pl, tr, steps)
    orc = Oracle(1); st = orc.new_state(B)
    st["player"][:, [0, 1, 3]] = pl; st["player"][:, 2] = 200.0; st["traffic"][:] = tr; st["steps"][:] = steps
    scale = np.array([1.0, 3.0, 40.0, 400.0, 2000.0])[rng.integers(0, 5, B)]
    rep = parity.ParityReport()
    alive = np.ones(B, bool)
    for t in range(T):
        a = (rng.uniform(-1, 1, B) * scale).astype(np.float32)
        obs, rew, done = env.step(torch.from_numpy(a).cuda())
        o, r, f, oc = orc.step(st, a.astype(np.float64))
        parity.compare_step(rep, npy(obs), npy(rew), npy(env.flags), o, r, f, alive)
        alive &= ~(f & FLAG_DONE > 0)
    assert rep.flag_mismatch == 0 and rep.steps > B * 10, rep
    ex = env.extract_state()
    psi_err = np.abs(ex["player"][:, 2] - st["player"][:, 3]); psi_err = np.minimum(psi_err, 360 - psi_err)
    assert psi_err[alive].max(initial=0) < 1e-9 and np.abs(ex["player"][alive, :2] - st["player"][alive][:, :2]).max(initial=0) < 1e-9


def test_timeout_is_exactly_1000_step_calls(cuda):
    """Q5/Q6: an episode is at most 1000 step() calls; final game.steps == 1001, obs[0] == 1.001, and the
    time discount is -0.001 on that step."""
    env = make(4, 1, auto_reset=False)
    env.reset()
    pl = np.tile([200.0, 500.0, 90.0], (4, 1)); pl[:, 1] = [100, 300, 500, 700]
    tr = np.tile([1500.0, 50.0, 200.0, 0.0], (4, 1, 1))
    env.inject_state(pl, tr, np.full(4, 998, np.int32))
    a = torch.zeros(4, device="cuda")
    o, r, d = env.step(a); assert not d.any() and float(o[0, 0]) == pytest.approx(0.999)
    o, r, d = env.step(a); assert not d.any() and float(o[0, 0]) == pytest.approx(1.0)
    o, r, d = env.step(a); assert d.all() and float(o[0, 0]) == pytest.approx(1.001)
    assert (env.flags & 15 == 12).all() and (env.outcome == 3).all() and (env.ep_length == 1001).all()
    assert float(r.max()) <= 0.0 and float(r.min()) > -0.0011


def test_host_buffer_step_equals_device_step(cuda):
    a = make(2048, 1, seed=2, auto_reset=True); b = make(2048, 1, seed=2, auto_reset=True)
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    for t in range(50):
        act = rng.uniform(-1, 1, 2048).astype(np.float32)
        o1, r1, d1 = a.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = b.step_host(act)
        assert np.array_equal(npy(o1), o2) and np.array_equal(npy(r1), r2) and np.array_equal(npy(d1), d2)


@pytest.mark.parametrize("n,B", [(1, 600_000 + 77), (16, 600_000), (256, 530_000)])
def test_chunked_host_step_equals_device_step(cuda, n, B):
    """Large batches take acas2d_step_host's chunk pipeline (256 Ki-env chunks rotating over three streams, each chunk a
    sub-batch view of every state array -- records, kinematic cache, compact headings, pre-pass scratch, spawn
    separations -- with its own global env id offset): bit-identical to the one-launch device step, state included."""
    a = make(B, n, seed=12, auto_reset=True); b = make(B, n, seed=12, auto_reset=True)
    assert B > a.PACKED_HOST_LIMIT and B >= 2 * 262144
    a.reset(); b.reset()
    if n == 1:                                           # age the single-intruder games: episodes end on every step
        a.rollout_random(1200, action_seed=1); b.rollout_random(1200, action_seed=1)
        a.clear_stats(); b.clear_stats()
    for t in range(6 if n == 256 else 12):
        act = a.random_actions(t, action_seed=4)
        o1, r1, d1 = a.step(act)
        o2, r2, d2 = b.step_host(act.cpu().numpy())
        assert np.array_equal(npy(o1).view(np.int32), o2.view(np.int32)), t
        assert np.array_equal(npy(r1), r2) and np.array_equal(npy(d1), d2), t
    assert torch.equal(a.ppos, b.ppos) and torch.equal(a.paux, b.paux) and torch.equal(a.thot, b.thot)
    assert torch.equal(a.episode_idx, b.episode_idx) and torch.equal(a.episode_counters(), b.episode_counters())
    if n > 1:
        assert torch.equal(a.spawn_sep, b.spawn_sep)
    assert int(a.episode_counters()[0]) > 1000


def test_fused_rollout_equals_stepwise_and_graph(cuda):
    B, K = 8192, 300
    a = make(B, 1, seed=3, auto_reset=True); b = make(B, 1, seed=3, auto_reset=True); c = make(B, 1, seed=3, auto_reset=True)
    for e in (a, b, c):
        e.reset()
    rs = torch.zeros(B, device="cuda")
    a.rollout_random(K, action_seed=77, step0=0, reward_sum=rs)
    acts = torch.stack([b.random_actions(k, 77) for k in range(K)])
    acc = torch.zeros(B, device="cuda", dtype=torch.float64)
    for k in range(K):
        acc += b.step(acts[k])[1]
    assert torch.equal(a.ppos, b.ppos) and torch.equal(a.paux, b.paux) and torch.equal(a.thot, b.thot)
    assert torch.equal(a.episode_counters(), b.episode_counters())
    assert float((rs.double() - acc).abs().max()) < 0.5
    # CUDA-graph replay of the step loop == eager (capture runs one warm-up step with actions[0])
    c.step(acts[0])
    graph = c.capture_steps(acts[1:])
    # capture_steps' warm-up advanced c by acts[1]; rebuild c to the post-acts[0] state and replay
    c2 = make(B, 1, seed=3, auto_reset=True); c2.reset(); c2.step(acts[0])
    c.load_state_dict(c2.state_dict())
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(c.ppos, b.ppos) and torch.equal(c.paux, b.paux)


def test_state_dict_roundtrip(cuda):
    a = make(1000, 3, seed=8, auto_reset=True)
    a.reset()
    for t in range(30):
        a.step(a.random_actions(t))
    sd = a.state_dict()
    b = make(1000, 3, seed=8, auto_reset=True)
    b.load_state_dict(sd)
    for t in range(30, 60):
        act = a.random_actions(t)
        oa = a.step(act)[0].clone()
        assert torch.equal(oa, b.step(act)[0])


def test_state_dict_resumes_a_closed_loop_rollout(cuda, golden_dir):
    """Checkpoint in the middle of a policy-driven rollout: the actor's next input is the ``obs`` buffer, so the
    checkpoint must carry it (and the Philox stream identity: seed, env id offset).  The resumed batch -- built
    with another seed and offset on purpose -- continues bit for bit."""
    from gym_ACAS2D.policy import MlpActor
    actor = MlpActor.from_file(os.path.join(golden_dir, "ppo_policy_1048576_11.npz"), "cuda:0")
    a = make(2048, 1, seed=8, env_id_offset=4096, auto_reset=True)
    a.reset()
    ex = a.extract_state()
    ex["steps"][:] = 1 + (np.arange(2048) * 13) % 990                  # episodes end (and respawn) along the way
    a.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
    a.observe()
    for t in range(40):
        a.policy_step(actor, deterministic=False, noise_seed=3, step_index=t)
    sd = a.state_dict()
    b = make(2048, 1, seed=99, env_id_offset=0, auto_reset=True)
    b.load_state_dict(sd)
    for t in range(40, 80):
        oa, ra, da = a.policy_step(actor, deterministic=False, noise_seed=3, step_index=t)
        ob, rb, db = b.policy_step(actor, deterministic=False, noise_seed=3, step_index=t)
        assert torch.equal(oa.view(torch.int32), ob.view(torch.int32)) and torch.equal(ra, rb) and torch.equal(da, db)
    assert torch.equal(a.episode_idx, b.episode_idx) and torch.equal(a.paux, b.paux)
    assert int(a.episode_counters()[0]) > 50


def test_gym_surface_single_env(cuda):
    """The reference's gym API (environment.py:8-54): spaces, reset -> float64[8], 4-tuple step,
    env.game attributes; and an episode replayed through the oracle from the extracted state."""
    import gym_ACAS2D
    env = gym_ACAS2D.make("ACAS2D-v0")
    assert env.observation_space.shape == (8,) and env.action_space.shape == (1,)
    obs = env.reset()
    assert obs.dtype == np.float64 and obs.shape == (8,) and env.observation_space.contains(obs)
    assert obs[0] == pytest.approx(0.001) and env.game.steps == 1 and env.game.outcome is None
    assert env.game.player.x == 48 and env.game.player.y == 500 and env.game.traffic[0].x == 1552
    orc = Oracle(1)
    st = orc.new_state(1)
    ex = env._core.extract_state()
    st["player"][0] = (ex["player"][0, 0], ex["player"][0, 1], 200.0, ex["player"][0, 2], 0.0)
    st["traffic"][0] = ex["traffic"][0]; st["steps"][0] = 1
    rng = np.random.default_rng(3)
    for k in range(1100):
        a = np.array([rng.uniform(-1, 1)], dtype=np.float32)
        o, r, d, info = env.step(a)
        ro, rr, rf, roc = orc.step(st, a.astype(np.float64))
        assert isinstance(r, float) and isinstance(d, bool) and info == {} and o.dtype == np.float64
        assert d == bool(rf[0] & FLAG_DONE)
        assert np.abs(o - ro[0]).max() < 2e-6 and abs(r - rr[0]) < 2e-4
        if d:
            assert env.game.outcome == roc[0] and env.game.steps == st["steps"][0]
            assert abs(env.game.total_reward - st["total_reward"][0]) < parity.TOL_RETURN
            assert env.game.d_path == pytest.approx(st["d_path"][0], abs=1e-6)
            break
    assert d
    # state injection through the game view (how the reference is poked)
    env.reset()
    env.game.player.x = 300.0; env.game.player.psi = 10.0; env.game.traffic[0].y = 640.0; env.game.steps = 17
    assert env.game.player.x == 300.0 and env.game.player.psi == 10.0 and env.game.steps == 17
    assert env.game.traffic[0].y == pytest.approx(640.0, abs=1e-9)
    env.render(); env.close()


def test_vec_env_adapter_semantics(cuda):
    """SB3 DummyVecEnv semantics: terminal_observation + episode info on done, reset obs returned."""
    from gym_ACAS2D.envs import ACAS2DVecEnv
    B = 64
    venv = ACAS2DVecEnv(B, seed=13)
    assert venv.num_envs == B and venv.observation_space.shape == (8,) and venv.action_space.shape == (1,)
    obs = venv.reset()
    assert obs.shape == (B, 8) and obs.dtype == np.float32
    orc = Oracle(1)
    st = orc.new_state(B); orc.spawn_philox(st, 13, 0); orc.observe(st)
    seen = 0
    for t in range(800):
        act = np.zeros((B, 1), np.float32)
        obs, rew, dones, infos = venv.step(act)
        o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, act[:, 0].astype(np.float64), 13, 0)
        assert rew.dtype == np.float32 and dones.dtype == np.bool_ and len(infos) == B
        assert np.array_equal(dones, f & FLAG_DONE > 0)
        for i in np.flatnonzero(dones):
            assert np.abs(infos[i]["terminal_observation"] - term[i]).max() < 2e-6
            assert infos[i]["episode"]["l"] == ep_len[i] - 1 and abs(infos[i]["episode"]["r"] - ep_ret[i]) < parity.TOL_RETURN
            assert abs(obs[i, 0] - 0.001) < 1e-7                                   # first obs of the next episode
            seen += 1
        assert all(infos[i] == {} for i in np.flatnonzero(~dones))
    assert seen >= B
    assert venv.env_is_wrapped(object) == [False] * B and venv.get_attr("num_envs")[0] == B
    s = venv.core.episode_stats()
    assert s["episodes"] == seen and s["timeout"] == 0


def test_golden_csv_replayed_on_the_gpu(cuda, golden_dir):
    """The reference's own golden artefact, through the CUDA path: the 100 zero-action episodes of
    baseline_ACAS2D_PPO_11_100.csv, spawned with the reference's draw order on MT19937 (seed 13, two
    discarded games) and injected, then recorded in the reference's CSV schema by records.record_episodes.
    Outcome and Time Steps bit-exact for 100/100, paths within 1e-9 px, total reward within 5e-3."""
    import random
    from gym_ACAS2D import records
    from oracle.acas2d_oracle import DEFAULTS, reference_spawn
    g = np.load(os.path.join(golden_dir, "baseline_zero_action.npz"))
    rng = random.Random(13)
    for _ in range(2):
        reference_spawn(rng, DEFAULTS)
    spawns = [reference_spawn(rng, DEFAULTS) for _ in range(100)]
    player = np.array([[s[0][0], s[0][1], s[0][3]] for s in spawns]); traffic = np.array([s[1] for s in spawns])
    env = make(100, 1, auto_reset=False)
    rows = records.record_episodes(env, start=(player, traffic))
    names = {1: "Goal", 2: "Collision", 3: "Timeout"}
    stride = int(g["stride"])
    for ep, row in enumerate(rows):
        assert row["Outcome"] == names[int(g["outcome"][ep])], ep
        assert row["Time Steps"] == g["time_steps"][ep], ep
        path = np.array(row["Path"]); tpath = np.array(row["Traffic Paths"][0])
        assert len(path) == g["path_len"][ep] and len(tpath) == len(path), ep
        n = len(path[::stride])
        assert np.abs(path[::stride] - g["path_samples"][ep][:n]).max() < 1e-9, ep
        assert np.abs(tpath[::stride] - g["traffic_samples"][ep][:n]).max() < 1e-9, ep
        assert np.array_equal(tpath[0], tpath[1])                                   # Q10
        assert abs(row["Total Reward"] - g["total_reward"][ep]) < parity.TOL_RETURN, ep
        assert row["Path Length"] == pytest.approx(2.0 * (row["Time Steps"] - 1))
        assert len(row["psi"]) == len(path) and len(row["r_step"]) == len(path)
    assert sum(r["Outcome"] == "Collision" for r in rows) == 58 and sum(r["Outcome"] == "Goal" for r in rows) == 42
    import tempfile, pandas as pd
    with tempfile.TemporaryDirectory() as d:
        records.to_csv(rows[:3], os.path.join(d, "t.csv"))
        df = pd.read_csv(os.path.join(d, "t.csv"))
        assert list(df.columns) == records.TESTING_COLUMNS and len(df) == 3
        import ast
        assert ast.literal_eval(df["Path"][0])[0] == (48.0, 500.0)                  # parses like the reference's CSV
        assert len(ast.literal_eval(df["Traffic Paths"][0])[0]) == len(ast.literal_eval(df["Path"][0]))
        records.to_csv(rows[:3], os.path.join(d, "b.csv"), records.BASELINE_COLUMNS)
        assert list(pd.read_csv(os.path.join(d, "b.csv")).columns) == records.BASELINE_COLUMNS


def test_on_device_episode_records_match_the_reference_lists(cuda):
    """SURVEY 8f-3 as specified: per-step records written by a kernel into ring buffers in HBM and read back once
    (acas2d_trace_step), against the reference's per-step lists (game.py:45-75, 231-239, 266-276) as the oracle's
    Python port keeps them -- random actions, N_TRAFFIC = 1 and 3, every recorded quantity."""
    import random
    from gym_ACAS2D import records
    from oracle.acas2d_oracle import DEFAULTS, PyPortGame
    for n in (1, 3):
        consts = dict(DEFAULTS, MIN_TRAFFIC=n, MAX_TRAFFIC=n)
        rng = random.Random(100 + n)
        games = [PyPortGame(rng, consts) for _ in range(6)]
        player = np.array([[g.player.x, g.player.y, g.player.psi] for g in games])
        traffic = np.array([[[t.x, t.y, t.v_air, t.psi] for t in g.traffic] for g in games])
        arng = np.random.default_rng(n)
        acts = arng.uniform(-1, 1, (1001, 6)).astype(np.float32)
        for b, g in enumerate(games):
            g.observe()
            for k in range(1001):
                if g.step(np.array([float(acts[k, b])]))[2]:
                    break
        env = make(6, n, auto_reset=False)
        step_no = {"k": 0}

        def policy(obs):
            a = torch.from_numpy(acts[step_no["k"]]).cuda()
            step_no["k"] += 1
            return a

        rows = records.record_episodes(env, policy=policy, start=(player, traffic))
        for b, (row, g) in enumerate(zip(rows, games)):
            assert row["Time Steps"] == g.steps and row["Outcome"] == {1: "Goal", 2: "Collision", 3: "Timeout"}[g.outcome]
            assert np.abs(np.array(row["Path"]) - np.array(g.path)).max() < 1e-9
            for i in range(n):
                assert np.abs(np.array(row["Traffic Paths"][i]) - np.array(g.traffic_paths[i])).max() < 1e-9
            want = dict(psi=g.rec["psi"], d_sep=g.rec["sep"], a_lat=g.rec["a_lat"], d_goal=g.rec["d_goal"],
                        delta_heading=g.rec["dh"], v_closing=g.rec["vc"], d_cpa=g.rec["dcpa"], d_dev=g.rec["ddev"],
                        r_d_goal=g.rec["r_goal"], r_h_goal=g.rec["r_head"], r_d_cpa=g.rec["r_cpa"], r_d_dev=g.rec["r_dev"],
                        r_step=g.rec["r"])
            tol = dict(psi=1e-9, d_sep=2e-3, a_lat=2e-5, d_goal=2e-3, delta_heading=1e-4, v_closing=1e-3, d_cpa=5e-3,
                       d_dev=1e-4, r_d_goal=2e-6, r_h_goal=2e-6, r_d_cpa=2e-5, r_d_dev=2e-5, r_step=2e-5)
            for key, ref in want.items():
                got, ref = np.array(row[key], np.float64), np.array(ref, np.float64)
                assert got.shape == ref.shape, (key, got.shape, ref.shape)
                near_branch = np.abs(np.array(g.rec["vc"])) < 1e-3          # r_d_cpa / r_step switch on v_closing's sign
                d = np.abs(got - ref)
                if key in ("r_d_cpa", "r_step"):
                    d = np.where(near_branch, 0.0, d)
                if key == "delta_heading":
                    d = np.minimum(d, 360 - d)
                assert np.nanmax(d) <= tol[key], (n, b, key, float(np.nanmax(d)))
            assert abs(row["Total Reward"] - g.total_reward) < parity.TOL_RETURN


def test_episode_records_across_auto_resets(cuda):
    """The trace kernel under SB3 semantics (auto-reset inside the step): a finished game's last row carries the done
    flag, the next rows are the NEW game's initial row (steps == 1, a_lat == 0, reward 0: game.py:132-160) and its
    first step (steps == 2); rewards and flags in the rows are the step's own outputs."""
    from gym_ACAS2D.envs import _native
    B, T = 256, 80
    env = make(B, 1, seed=31, auto_reset=True)
    env.reset()
    ex = env.extract_state()
    ex["steps"][:] = 1000 - (np.arange(B) % 60)                 # every game times out somewhere inside the window
    env.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
    env.enable_trace(B, 0, capacity=4 * T)
    F = {k: i for i, k in enumerate(_native.TRACE_FIELDS)}
    rewards, flags = [], []
    for t in range(T):
        o, r, d = env.step(env.random_actions(t, action_seed=8))
        rewards.append(npy(r).copy()); flags.append(npy(env.flags).copy())
    rows, count = env.trace_rows()
    rewards, flags = np.array(rewards), np.array(flags)
    ended = 0
    for b in range(B):
        rb = rows[b, : count[b]]
        step_rows = rb[~((rb[:, F["steps"]] == 1))]             # initial rows have steps == 1; step rows follow the step order
        assert len(step_rows) == T
        assert np.array_equal(step_rows[:, F["flags"]].astype(np.uint8) & 15, flags[:, b] & 15)
        assert np.abs(step_rows[:, F["reward"]] - rewards[:, b]).max() == 0.0
        for k in np.flatnonzero(rb[:-1, F["flags"]].astype(np.int64) & 8):          # a done row ...
            nxt = rb[k + 1]
            assert nxt[F["steps"]] == 1 and nxt[F["a_lat"]] == 0 and nxt[F["reward"]] == 0            # ... then a new game's initial row
            assert (nxt[F["x"]], nxt[F["y"]]) == (48.0, 500.0)
            if k + 2 < len(rb):
                assert rb[k + 2][F["steps"]] == 2
            ended += 1
    assert ended >= B - 4                                        # (a game ending on the window's last step has no successor row)


def test_headless_render_frame(cuda):
    """debug frame of one env -- sky, goal disc + yellow GOAL_RADIUS ring, player disc + red
    COLLISION_RADIUS ring, intruder disc + ring, at the positions of the device state."""
    env = make(3, 2, auto_reset=False)
    env.reset()
    pl = np.array([[300.0, 400.0, 0.0]] * 3); tr = np.tile([[900.0, 700.0, 200.0, 180.0], [1200.0, 200.0, 200.0, 90.0]], (3, 1, 1))
    env.inject_state(pl, tr)
    img = npy(env.render(1))
    assert img.shape == (1000, 1600, 3) and img.dtype == np.uint8
    assert tuple(img[5, 5]) == (60, 150, 220)                       # SKY_RGB
    assert tuple(img[400, 300]) == (0, 0, 0)                        # player
    assert tuple(img[400, 300 + 48]) == (255, 0, 0)                 # COLLISION_RADIUS ring
    assert tuple(img[700, 900]) == (90, 90, 90) and tuple(img[200, 1200]) == (90, 90, 90)
    assert tuple(img[500, 1456]) == (0, 255, 0)                     # goal
    assert tuple(img[500, 1456 - 144]) == (255, 255, 0)             # GOAL_RADIUS ring
    import gym_ACAS2D
    e = gym_ACAS2D.make("ACAS2D-v0")
    frame = e.render(mode="rgb_array")
    assert frame.shape == (1000, 1600, 3) and tuple(frame[500, 48]) == (0, 0, 0) and e.render() is None


def test_tiled_kernel_sharding_invariance_at_scale(cuda):
    """N_TRAFFIC = 8 at 262 144 envs (tiled kernel, warp-cooperative respawns): two half-batches addressed
    by global env id reproduce the full batch bit for bit, and a slice agrees with the oracle."""
    B, N, T = 1 << 18, 8, 40
    full = make(B, N, seed=3, auto_reset=True, track_min_sep=True)
    lo = make(B // 2, N, seed=3, env_id_offset=0, auto_reset=True, track_min_sep=True)
    hi = make(B // 2, N, seed=3, env_id_offset=B // 2, auto_reset=True, track_min_sep=True)
    o = full.reset().clone()
    assert torch.equal(o[: B // 2], lo.reset()) and torch.equal(o[B // 2:], hi.reset())
    orc = Oracle(N); K = 512
    st = orc.new_state(K); orc.spawn_philox(st, 3, 0); orc.observe(st)
    for t in range(T):
        a = full.random_actions(t, action_seed=6)
        of, rf, df = full.step(a)
        o1, r1, d1 = lo.step(a[: B // 2]); o2, r2, d2 = hi.step(a[B // 2:])
        assert torch.equal(of[: B // 2].view(torch.int32), o1.view(torch.int32))
        assert torch.equal(of[B // 2:].view(torch.int32), o2.view(torch.int32))
        assert torch.equal(rf[B // 2:], r2) and torch.equal(df[: B // 2], d1)
        ro, rr, rfl, *_ = orc.vec_step(st, npy(a[:K]).astype(np.float64), 3, 0)
        assert np.array_equal(npy(df[:K]), rfl & FLAG_DONE > 0)
        assert np.nanmax(np.abs(npy(of[:K]) - ro)) < parity.TOL_OBS_CPA
    assert torch.equal(full.episode_counters(), lo.episode_counters() + hi.episode_counters())
    assert torch.equal(full.min_sep[B // 2:], hi.min_sep) and torch.equal(full.thot[: B // 2], lo.thot)
    assert full.episode_counters()[0].item() > 10000


@pytest.mark.parametrize("n", [1, 8])
def test_unequal_speeds_q3_on_the_gpu(cuda, n):
    """Q3 (kinematics.py:74) through the CUDA kernels: intruders at 0.6 .. 1.4 x the player's airspeed, spawned
    and injected (float64 residual path) states, auto-reset; flags bit-exact, floats within the tolerances.
    The oracle is pinned to the unmodified reference for this setting in tests/test_oracle_golden.py."""
    over = dict(AIRSPEED_FACTOR_MIN=0.6, AIRSPEED_FACTOR_MAX=1.4)
    B, seed, off = 256 * 3 + 5, 29, 500
    env = make(B, n, seed=seed, env_id_offset=off, auto_reset=True, **over)
    orc = Oracle(n, **over)
    st = orc.new_state(B)
    orc.spawn_philox(st, seed, off)
    assert np.abs(npy(env.reset()) - orc.observe(st)).max() < parity.TOL_OBS_CPA
    rng = np.random.default_rng(3)
    ex = env.extract_state()
    assert np.abs(ex["traffic"][:, :, 2] - 200.0).max() > 50.0
    ex["traffic"][::2, :, 2] = rng.uniform(90.0, 310.0, ex["traffic"][::2, :, 2].shape)
    env.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
    st["traffic"][::2, :, 2] = ex["traffic"][::2, :, 2]
    rep = parity.ParityReport()
    for t in range(400):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        obs, rew, done = env.step(torch.from_numpy(a).cuda())
        o, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, a.astype(np.float64), seed, off)
        d = f & FLAG_DONE > 0
        assert np.array_equal(npy(done), d)
        parity.compare_step(rep, np.where(d[:, None], npy(env.term_obs), npy(obs)), npy(rew), npy(env.flags),
                            np.where(d[:, None], term, o), r, f)
    parity.assert_flags_exact(rep)
    assert rep.steps == 400 * B


def test_step_k_equals_k_single_steps(cuda):
    """acas2d_step_k (K open-loop steps per launch, state in registers in between) writes exactly what K calls
    of acas2d_step write -- every step's obs / reward / done, the state, the episode counters -- with
    auto-reset inside the window."""
    B, K = 256 * 5 + 19, 12
    a, b = make(B, seed=4, auto_reset=True), make(B, seed=4, auto_reset=True)
    for env in (a, b):
        env.reset()
        ex = env.extract_state(); ex["steps"][:] = 985 + (np.arange(B) % 14)          # plenty of timeouts inside the window
        env.inject_state(ex["player"], ex["traffic"], ex["steps"], ex["total_reward"])
    for rep in range(3):
        acts = torch.stack([a.random_actions(rep * K + k, action_seed=9) for k in range(K)])
        obs_k, rew_k, done_k = a.step_k(acts)
        for k in range(K):
            o, r, d = b.step(acts[k], full_outputs=False)
            assert torch.equal(obs_k[k], o) and torch.equal(rew_k[k], r) and torch.equal(done_k[k], d), (rep, k)
        assert torch.equal(a.ppos, b.ppos) and torch.equal(a.paux, b.paux) and torch.equal(a.thot, b.thot)
        assert torch.equal(a.episode_idx, b.episode_idx) and torch.equal(a.episode_counters(), b.episode_counters())
        assert torch.equal(a.obs, b.obs)
    assert int(a.episode_counters()[0]) >= B
    assert a.step_k(acts[:0])[0].shape[0] == 0                                          # K = 0: nothing to do
