"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes, next to this file:

* ``baseline_zero_action.npz`` -- the reference's own golden artefact
  ``gym_ACAS2D/models/logs/baseline_ACAS2D_PPO_11_100.csv`` (100 zero-action
  episodes, written by ``baseline_main.py:53-74``) condensed to outcomes, time
  steps, total rewards, strided path samples and full-path sums.
* ``ref_rollouts_n{1,8}.npz`` -- injected states (random + adversarial) and random
  action sequences stepped through the reference itself (imported under gym/pygame
  stand-ins, ``oracle/ref_shim.py``); every reward/flag, strided obs and positions.
* ``ref_reset_obs.npz`` -- SURVEY App. A probe: first game after ``random.seed(13)``.

The script also cross-checks the C oracle against what it just generated and prints
the worst differences, so a drifting oracle is caught at fixture time.
"""
from __future__ import annotations

import ast
import csv
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.acas2d_oracle import Oracle, FLAG_COLLISION, FLAG_GOAL, FLAG_TIMEOUT, FLAG_DONE  # noqa: E402

CSV_REL = "gym_ACAS2D/models/logs/baseline_ACAS2D_PPO_11_100.csv"
OUTCOME_CODE = {"Goal": 1, "Collision": 2, "Timeout": 3}
STRIDE = 8


def condense_csv():
    path = os.path.join(ref_shim.REFERENCE_ROOT, CSV_REL)
    csv.field_size_limit(1 << 30)
    rows = list(csv.DictReader(open(path)))
    assert len(rows) == 100
    out = dict(outcome=[], total_reward=[], time_steps=[], path_len=[], path_sum=[], traffic_sum=[],
               path_last=[], traffic_first=[], traffic_last=[])
    samples_p, samples_t = [], []
    for r in rows:
        p = np.array(ast.literal_eval(r["Path"]), dtype=np.float64)
        t = np.array(ast.literal_eval(r["Traffic Paths"]), dtype=np.float64)  # [N=1][len][2]
        out["outcome"].append(OUTCOME_CODE[r["Outcome"]])
        out["total_reward"].append(float(r["Total Reward"]))
        out["time_steps"].append(int(r["Time Steps"]))
        out["path_len"].append(len(p))
        out["path_sum"].append(p.sum(0))
        out["traffic_sum"].append(t[0].sum(0))
        out["path_last"].append(p[-1])
        out["traffic_first"].append(t[0][0])
        out["traffic_last"].append(t[0][-1])
        sp = np.full((1001 // STRIDE + 1, 2), np.nan); st = sp.copy()
        sp[: len(p[::STRIDE])] = p[::STRIDE]; st[: len(t[0][::STRIDE])] = t[0][::STRIDE]
        samples_p.append(sp); samples_t.append(st)
    out = {k: np.array(v) for k, v in out.items()}
    out["path_samples"] = np.array(samples_p)
    out["traffic_samples"] = np.array(samples_t)
    out["stride"] = np.array(STRIDE)
    np.savez_compressed(os.path.join(HERE, "baseline_zero_action.npz"), **out)
    return rows, out


def injected_states(rng: np.random.Generator, B: int, N: int):
    """Random + adversarial initial states.  Positions/headings are arbitrary float64."""
    player = np.zeros((B, 5)); traffic = np.zeros((B, N, 4)); steps = np.ones(B, np.int32)
    total = np.zeros(B)
    for b in range(B):
        kind = b % 8
        px, py = rng.uniform(30, 900), rng.uniform(150, 850)
        psi = rng.uniform(0, 360) if kind == 1 else (rng.uniform(-25, 25) % 360)
        player[b] = (px, py, 200.0, psi, 0.0)
        for i in range(N):
            traffic[b, i] = (rng.uniform(400, 1580), rng.uniform(20, 980), 200.0, rng.uniform(0, 360))
        if kind == 0:      # reference-like spawn
            player[b] = (48.0, 500.0, 200.0, rng.uniform(-3, 3) % 360, 0.0)
            sd = int(rng.integers(0, 2))
            traffic[b, 0] = (1552.0, 48.0 + sd * 904.0, 200.0, (145 + 70 * sd + rng.uniform(-15, 15)) % 360)
        elif kind == 2:    # head-on, starts a few px outside the 96 px collision threshold
            ang = rng.uniform(0, 2 * np.pi); d0 = rng.uniform(96.5, 140.0)
            traffic[b, 0, 0] = px + d0 * np.cos(ang); traffic[b, 0, 1] = py + d0 * np.sin(ang)
            player[b, 3] = np.degrees(ang) % 360
            traffic[b, 0, 3] = (np.degrees(ang) + 180.0) % 360
        elif kind == 3:    # just outside the 144 px goal disc, flying at it
            ang = rng.uniform(0, 2 * np.pi); d0 = rng.uniform(144.5, 190.0)
            player[b, 0] = 1456.0 - d0 * np.cos(ang); player[b, 1] = 500.0 - d0 * np.sin(ang)
            player[b, 3] = np.degrees(ang) % 360
            for i in range(N):
                traffic[b, i, :2] = (rng.uniform(100, 600), rng.uniform(20, 300))
        elif kind == 4:    # about to time out (Q5/Q6), far from everything
            steps[b] = int(rng.integers(985, 1001)); total[b] = rng.uniform(50, 300)
            player[b, :2] = (rng.uniform(100, 400), rng.uniform(400, 600))
            player[b, 3] = rng.uniform(100, 260)
            for i in range(N):
                traffic[b, i, :2] = (rng.uniform(1200, 1580), rng.uniform(20, 980))
                traffic[b, i, 3] = rng.uniform(-60, 60) % 360
        elif kind == 5:    # heading wrap at 0/360 and y crossing the goal line (phi wrap)
            player[b, 3] = rng.choice([359.7, 0.2, 359.99, 0.0])
            player[b, 1] = 500.0 + rng.uniform(-0.5, 0.5)
        elif kind == 6:    # relative velocity x-component near zero (Q12 sign flip region)
            traffic[b, 0, 3] = (player[b, 3] + rng.choice([-1, 1]) * rng.uniform(0.0, 2.0)) % 360
            traffic[b, 0, 3] = (360.0 - traffic[b, 0, 3]) % 360 if rng.random() < 0.5 else traffic[b, 0, 3]
        # kind 7: plain random
    return player, traffic, steps, total


def run_reference(pkg, N, player, traffic, steps, total, actions):
    """Step the reference env from the injected states; returns per-step records."""
    from gym_ACAS2D.envs.environment import ACAS2DEnv
    from gym_ACAS2D.envs.game import ACAS2DGame
    T, B = actions.shape
    L = 5 + 3 * N
    obs = np.zeros((T, B, L)); rew = np.zeros((T, B)); flags = np.zeros((T, B), np.uint8)
    outcome = np.zeros((T, B), np.uint8); pl = np.zeros((T, B, 3)); tr = np.zeros((T, B, N, 2))
    total_out = np.zeros(B); steps_out = np.zeros(B, np.int32); minsep = np.zeros(B); dpath = np.zeros(B)
    with ref_shim.quiet():
        env = ACAS2DEnv()
        for b in range(B):
            env.game = g = ACAS2DGame()
            assert g.num_traffic == N
            g.player.x, g.player.y, g.player.v_air, g.player.psi, g.player.a_lat = (float(v) for v in player[b])
            for i in range(N):
                t = g.traffic[i]
                t.x, t.y, t.v_air, t.psi = (float(v) for v in traffic[b, i])
            g.steps = int(steps[b]); g.total_reward = float(total[b])
            g.d_sep_record = [g.minimum_separation()]
            done = False
            for k in range(T):
                if not done:
                    o, r, done, _ = env.step(np.array([float(actions[k, b])]))
                    f = (FLAG_COLLISION * g.detect_collisions()) | (FLAG_GOAL * g.check_goal()) | \
                        (FLAG_TIMEOUT * g.check_timeout()) | (FLAG_DONE * bool(done))
                    obs[k, b], rew[k, b], flags[k, b] = o, r, f
                    outcome[k, b] = g.outcome or 0
                else:
                    obs[k, b], rew[k, b], flags[k, b], outcome[k, b] = obs[k - 1, b], rew[k - 1, b], flags[k - 1, b], outcome[k - 1, b]
                pl[k, b] = (g.player.x, g.player.y, g.player.psi)
                for i in range(N):
                    tr[k, b, i] = (g.traffic[i].x, g.traffic[i].y)
            total_out[b] = g.total_reward; steps_out[b] = g.steps
            minsep[b] = np.min(g.d_sep_record); dpath[b] = g.d_path
    return dict(obs=obs, reward=rew, flags=flags, outcome=outcome, player=pl, traffic=tr,
                total_reward=total_out, steps=steps_out, min_sep=minsep, d_path=dpath)


def make_rollouts(N: int, B: int, T: int, seed: int):
    pkg = ref_shim.load(N)
    rng = np.random.default_rng(seed)
    player, traffic, steps, total = injected_states(rng, B, N)
    actions = rng.uniform(-1, 1, size=(T, B)).astype(np.float32)
    actions[:, ::5] = 0.0                                   # some straight flyers
    ref = run_reference(pkg, N, player, traffic, steps, total, actions.astype(np.float64))

    # cross-check the C oracle on the identical inputs
    orc = Oracle(N)
    st = orc.new_state(B)
    st["player"][:] = player; st["traffic"][:] = traffic; st["steps"][:] = steps; st["total_reward"][:] = total
    d = traffic[:, :, :2] - player[:, None, :2]
    st["min_sep"][:] = np.sqrt((d * d).sum(-1)).min(-1)
    got = orc.rollout(st, actions.astype(np.float64), record_traffic=True)
    assert np.array_equal(got["flags"], ref["flags"]), "oracle flags differ from the reference"
    assert np.array_equal(got["outcome"], ref["outcome"])
    with np.errstate(invalid="ignore"):
        print(f"N={N}: oracle-vs-reference max|d obs|={np.nanmax(np.abs(got['obs'] - ref['obs'])):.3e} "
              f"reward={np.nanmax(np.abs(got['reward'] - ref['reward'])):.3e} "
              f"player={np.abs(got['player'] - ref['player']).max():.3e} "
              f"traffic={np.abs(got['traffic'] - ref['traffic']).max():.3e} "
              f"total={np.abs(st['total_reward'] - ref['total_reward']).max():.3e} "
              f"minsep={np.abs(st['min_sep'] - ref['min_sep']).max():.3e} "
              f"dpath={np.abs(st['d_path'] - ref['d_path']).max():.3e}; "
              f"done envs={int((ref['flags'][-1] & FLAG_DONE > 0).sum())}/{B}")

    np.savez_compressed(
        os.path.join(HERE, f"ref_rollouts_n{N}.npz"),
        n_traffic=np.array(N), stride=np.array(STRIDE),
        player0=player, traffic0=traffic, steps0=steps, total0=total, actions=actions,
        reward=ref["reward"], flags=ref["flags"], outcome=ref["outcome"],
        obs_strided=ref["obs"][::STRIDE], player_strided=ref["player"][::STRIDE],
        traffic_strided=ref["traffic"][::STRIDE],
        obs_last=ref["obs"][-1], player_last=ref["player"][-1], traffic_last=ref["traffic"][-1],
        total_reward=ref["total_reward"], steps=ref["steps"], min_sep=ref["min_sep"], d_path=ref["d_path"])


def make_reset_probe():
    pkg = ref_shim.load(1)
    from gym_ACAS2D.envs.environment import ACAS2DEnv
    random.seed(13)
    with ref_shim.quiet():
        env = ACAS2DEnv()
        g = env.game
        state = np.array([g.player.x, g.player.y, g.player.v_air, g.player.psi,
                          g.traffic[0].x, g.traffic[0].y, g.traffic[0].v_air, g.traffic[0].psi])
        # ACAS2DEnv() already built the first game; observe() is what reset() returns for it
        obs0 = g.observe()
        obs1, r1, d1, _ = env.step(np.array([0.5]))
    np.savez(os.path.join(HERE, "ref_reset_obs.npz"), state=state, obs0=obs0, obs1=obs1,
             reward1=np.array(r1), done1=np.array(d1))
    print("reset probe:", obs0, obs1, r1, d1)


def replay_csv_with_oracle(rows):
    """SURVEY 8c recipe: seed 13, discard 2 games, 100 zero-action episodes."""
    from oracle.acas2d_oracle import reference_spawn, DEFAULTS
    orc = Oracle(1)
    rng = random.Random(13)
    for _ in range(2):
        reference_spawn(rng, DEFAULTS)
    worst_r = 0.0
    for ep, row in enumerate(rows):
        pl, tr = reference_spawn(rng, DEFAULTS)
        st = orc.new_state(1)
        st["player"][0] = pl; st["traffic"][0] = tr
        orc.observe(st)
        path = [(pl[0], pl[1])]; tpath = [(tr[0, 0], tr[0, 1])]
        while True:
            tpath.append((st["traffic"][0, 0, 0], st["traffic"][0, 0, 1]))
            _, _, fl, oc = orc.step(st, np.zeros(1))
            path.append((st["player"][0, 0], st["player"][0, 1]))
            if fl[0] & FLAG_DONE:
                break
        ref_path = np.array(ast.literal_eval(row["Path"]))
        ref_tp = np.array(ast.literal_eval(row["Traffic Paths"]))[0]
        assert OUTCOME_CODE[row["Outcome"]] == oc[0], ep
        assert int(row["Time Steps"]) == st["steps"][0], ep
        assert np.array_equal(np.array(path), ref_path), ep
        assert np.array_equal(np.array(tpath), ref_tp), ep
        worst_r = max(worst_r, abs(float(row["Total Reward"]) - st["total_reward"][0]))
    print(f"golden CSV replay through the C oracle: 100/100 outcome, steps, paths bit-identical; "
          f"max |total reward diff| = {worst_r:.3e}")


def make_policy_fixture():
    """The reference's trained agent (models/best_model_1048576_11/best_model.zip, an artefact the
    reference commits): its policy.pth tensors as a plain npz, plus the reference's own summary of that
    agent (notebooks/simulation_ACAS2D_PPO_1048576_11_100.ipynb cell 4; SURVEY 8c trained-policy replay)."""
    import io
    import zipfile
    import torch
    z = zipfile.ZipFile(os.path.join(ref_shim.REFERENCE_ROOT, "gym_ACAS2D/models/best_model_1048576_11/best_model.zip"))
    sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    arrays = {k: v.numpy() for k, v in sd.items()}
    np.savez_compressed(os.path.join(HERE, "ppo_policy_1048576_11.npz"), **arrays)
    print("policy fixture:", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    if not ref_shim.available():
        sys.exit("reference tree not found; fixtures can only be generated in the build container")
    rows, _ = condense_csv()
    replay_csv_with_oracle(rows)
    make_reset_probe()
    make_policy_fixture()
    make_rollouts(N=1, B=48, T=1001, seed=2024)
    make_rollouts(N=8, B=24, T=400, seed=2025)
    ref_shim.unload()
