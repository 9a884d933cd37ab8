"""The reference's trained agent (next row): policy-file compatibility and the fused
policy + env-step kernel.  Fixture: ``tests/golden/ppo_policy_1048576_11.npz`` = the tensors of the
reference's ``models/best_model_1048576_11/best_model.zip/policy.pth``.  External pins: the reference's
own summary of that agent -- 100/100 Goal, mean game.steps 704.35, mean return 1210.07
(``notebooks/simulation_ACAS2D_PPO_1048576_11_100.ipynb`` cell 4; SURVEY 8c replay: 706.7 / 1208.2)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from gym_ACAS2D.policy import MlpActor
from oracle.acas2d_oracle import FLAG_DONE, Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "ppo_policy_1048576_11.npz")


def host_policy_mean(actor, obs):
    from tests.hostcheck import HostBatch
    lib = HostBatch(1, 1).lib
    w = actor.packed.cpu().numpy()
    o = np.ascontiguousarray(obs, np.float32)
    out = np.zeros(len(o), np.float32)
    lib.hostcheck_policy_mean.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    lib.hostcheck_policy_mean(w.ctypes.data, o.ctypes.data, len(o), out.ctypes.data)
    return out


def test_policy_file_loads_and_matches_torch_reference():
    actor = MlpActor.from_file(FIXTURE)
    assert actor.packed.numel() == 4804 and abs(actor.log_std - (-1.6)) < 1.5      # a trained log_std, not the 0 init
    rng = np.random.default_rng(0)
    obs = rng.uniform(-1, 1, (512, 8)).astype(np.float32)
    ref = actor.reference_mean(torch.from_numpy(obs)).numpy()
    got = host_policy_mean(actor, obs)                     # the product's policy_mean(), g++ build
    assert np.abs(got - ref).max() < 2e-6


def test_trained_policy_on_the_oracle_reproduces_reference_statistics():
    """Deterministic trained actor + float64 oracle env, Philox spawns: the agent must do what the
    reference says it does (all episodes reach the goal in ~705 steps for ~1210 return)."""
    actor = MlpActor.from_file(FIXTURE)
    orc = Oracle(1)
    B = 256
    st = orc.new_state(B)
    orc.spawn_philox(st, 13, 0)
    obs = orc.observe(st)
    finished = np.zeros(B, bool); outcome = np.zeros(B, int); length = np.zeros(B, int); ret = np.zeros(B)
    for _ in range(1001):
        a = np.clip(host_policy_mean(actor, obs.astype(np.float32)).astype(np.float64), -1, 1)
        obs, r, f, oc, term, ep_ret, ep_len = orc.vec_step(st, a, 13, 0)
        d = (f & FLAG_DONE > 0) & ~finished
        outcome[d], length[d], ret[d] = oc[d], ep_len[d], ep_ret[d]
        finished |= d
        if finished.all():
            break
    assert finished.all()
    assert (outcome == 1).mean() >= 0.97                    # reference: 100 % Goal
    assert abs(length.mean() - 705) < 15                    # reference: 704.35 (notebook), 706.7 (replay)
    assert abs(ret.mean() - 1209) < 25                      # reference: 1210.07 (notebook), 1208.2 (replay)


@pytest.mark.gpu
def test_fused_policy_step_matches_torch_policy_plus_env_step():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D
    actor = MlpActor.from_file(FIXTURE, "cuda:0")
    B = 256 * 40 + 33
    a = BatchedACAS2D(B, seed=5, auto_reset=True); b = BatchedACAS2D(B, seed=5, auto_reset=True)
    oa = a.reset(); ob = b.reset()
    acts = torch.zeros(B, device="cuda")
    for t in range(300):
        ref_mean = actor.reference_mean(oa)                 # torch float32 reference on the same observations
        oa, ra, da = a.policy_step(actor, deterministic=True, actions_out=acts)
        assert float((acts - ref_mean).abs().max()) < 5e-6
        ob, rb, db = b.step(acts.clamp(-1, 1))              # the kernel's own action through the plain step
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    assert torch.equal(a.ppos, b.ppos) and torch.equal(a.episode_counters(), b.episode_counters())
    # stochastic head: N(mean, exp(log_std)^2) samples and their log-probabilities
    logp = torch.zeros(B, device="cuda")
    mean = actor.reference_mean(a.obs)
    a.policy_step(actor, deterministic=False, noise_seed=7, step_index=3, actions_out=acts, logp_out=logp)
    eps = (acts - mean) / np.exp(actor.log_std)
    assert abs(float(eps.mean())) < 0.05 and abs(float(eps.std()) - 1.0) < 0.05
    want = -0.5 * eps ** 2 - actor.log_std - 0.5 * np.log(2 * np.pi)
    assert float((logp - want).abs().max()) < 1e-3


@pytest.mark.gpu
def test_tensor_core_policy_step_matches_fp32_reference():
    """tcgen05 TF32 path (TMEM accumulators, tanh.approx): action mean within 3e-3 of the torch float32
    forward pass (stated tolerance: TF32 keeps 10 mantissa bits), and the env-step half of the fused
    kernel is bit-identical to the plain step fed with the kernel's own actions.  Ragged batch."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D
    actor = MlpActor.from_file(FIXTURE, "cuda:0")
    B = 128 * 148 * 3 * 2 + 128 * 5 + 41
    a = BatchedACAS2D(B, seed=6, auto_reset=True); b = BatchedACAS2D(B, seed=6, auto_reset=True)
    oa = a.reset(); b.reset()
    acts = torch.zeros(B, device="cuda")
    worst = 0.0
    for t in range(200):
        ref_mean = actor.reference_mean(oa)
        oa, ra, da = a.policy_step(actor, deterministic=True, actions_out=acts, tensor_cores=True)
        worst = max(worst, float((acts - ref_mean).abs().max()))
        ob, rb, db = b.step(acts.clamp(-1, 1))
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), t
    assert worst < 3e-3, worst
    assert torch.equal(a.ppos, b.ppos) and torch.equal(a.episode_counters(), b.episode_counters())
    print(f"tcgen05 policy: max |action mean - fp32 reference| = {worst:.2e}")


@pytest.mark.gpu
@pytest.mark.parametrize("tensor_cores", [False, True])
def test_trained_policy_statistics_on_gpu(tensor_cores):
    """Same external pin as the oracle test, through the fused kernel: 8192 first episodes."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D
    actor = MlpActor.from_file(FIXTURE, "cuda:0")
    B = 8192
    env = BatchedACAS2D(B, seed=13, auto_reset=True)
    env.reset()
    finished = torch.zeros(B, dtype=torch.bool, device="cuda")
    outcome = torch.zeros(B, dtype=torch.uint8, device="cuda"); length = torch.zeros(B, dtype=torch.int32, device="cuda")
    ret = torch.zeros(B, device="cuda")
    for _ in range(1001):
        _, _, d = env.policy_step(actor, deterministic=True, tensor_cores=tensor_cores)
        new = d & ~finished
        outcome[new] = env.outcome[new]; length[new] = env.ep_length[new]; ret[new] = env.ep_return[new]
        finished |= new
    assert bool(finished.all())
    goal_rate = float((outcome == 1).float().mean())
    assert goal_rate >= 0.97, goal_rate
    assert abs(float(length.float().mean()) - 705) < 10
    assert abs(float(ret.mean()) - 1209) < 20
    print(f"trained policy on GPU env (tensor_cores={tensor_cores}): goal {goal_rate:.4f}, mean steps {float(length.float().mean()):.1f}, mean return {float(ret.mean()):.1f}")


@pytest.mark.gpu
def test_collect_rollout_buffers_are_consistent():
    """PPO-style collection: [T, B] buffers written in place by the fused kernel equal a step-by-step
    rollout with the same noise stream, and obs[t+1] is what step t produced."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D
    actor = MlpActor.from_file(FIXTURE, "cuda:0")
    B, T = 3000, 48
    a = BatchedACAS2D(B, seed=8, auto_reset=True); b = BatchedACAS2D(B, seed=8, auto_reset=True)
    a.reset(); b.reset()
    buf = a.collect_rollout(actor, T, noise_seed=21, step0=100, tensor_cores=False)
    acts = torch.zeros(B, device="cuda"); logp = torch.zeros(B, device="cuda")
    for t in range(T):
        assert torch.equal(buf["obs"][t], b.obs)
        o, r, d = b.policy_step(actor, deterministic=False, noise_seed=21, step_index=100 + t, actions_out=acts, logp_out=logp)
        assert torch.equal(buf["actions"][t], acts) and torch.equal(buf["logp"][t], logp)
        assert torch.equal(buf["rewards"][t], r) and torch.equal(buf["dones"][t].bool(), d)
    assert torch.equal(buf["obs"][T], b.obs) and torch.equal(a.obs, b.obs) and torch.equal(a.ppos, b.ppos)
    std = float(np.exp(actor.log_std))
    assert 0.5 * std < float((buf["actions"] - buf["actions"].mean()).std()) < 10 * std + 1.0


@pytest.mark.gpu
def test_ppo_training_loop_mechanics():
    """Two tiny PPO iterations (rollout collection with the fused kernels + the torch autograd learner, i.e.
    the numerics reference of the learner kernels, end to end): finite numbers, episodes finish, parameters
    move.  The same loop on the learner kernels: tests/test_ppo_learner.py.  Full learning curves (100 % goal
    after 70 iterations of 1024 envs x 1024 steps): profiles/r01_ppo_training_curve*.json."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D import ppo
    hist = ppo.train(num_envs=512, n_steps=600, iterations=2, minibatches=8, n_epochs=2, tensor_cores=True,
                     learner="torch", log=None)
    assert len(hist) == 2 and hist[1]["episodes"] > 0
    assert all(np.isfinite(v) for r in hist for v in r.values())
    assert hist[1]["env_steps"] == 2 * 512 * 600
    assert abs(hist[1]["log_std"]) > 0.0            # the learner updated the policy
