"""PPO learner (BASELINE config 5): this repo's kernels (``acas2d_ppo_values / gae / grad / adam``)
against the same arithmetic in plain torch float32 (``gym_ACAS2D.ppo.reference_loss`` + autograd +
``torch.optim.Adam``), and the data-parallel gradient exchange on two gloo ranks.

The reference trains with stable-baselines3 1.1.0 (``gym_ACAS2D/training_main.py:44-52``), which is neither
vendored by the reference nor installed here: parity of the learner is against SB3's published loss restated
in torch, not against SB3 itself ("parity unpinned" for this row; the end-to-end pin is the learning curve --
the agent must reach the reference's 100 % Goal / ~1.2 k return)."""
import math
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from gym_ACAS2D import ppo
from gym_ACAS2D.envs import _native
from gym_ACAS2D.envs._native import PpoConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synthetic_rollout(n, seed=0, device="cpu"):
    """Rollout-shaped inputs with every branch of the loss populated: ratios on both sides of the clip range,
    advantages of both signs, observation rows in the env's range."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.rand(n, 8, generator=g) * 2 - 1
    actions = torch.randn(n, generator=g) * 0.7
    old_logp = -0.5 * torch.randn(n, generator=g) ** 2 - 0.9 + 0.3 * torch.randn(n, generator=g)
    adv = torch.randn(n, generator=g) * 3 + 0.5
    ret = torch.randn(n, generator=g) * 5
    return [x.to(device).contiguous() for x in (obs, actions, old_logp, adv, ret)]


def init_state_dict(seed=3, log_std=-0.4):
    torch.manual_seed(seed)
    sd = ppo.ActorCritic().sb3_state_dict()
    sd = {k: v.clone() for k, v in sd.items()}
    sd["action_net.weight"] = sd["action_net.weight"] * 30          # SB3 starts the action head at gain 0.01
    sd["mlp_extractor.policy_net.0.bias"] = torch.randn(64) * 0.1   # non-zero biases
    sd["mlp_extractor.value_net.2.bias"] = torch.randn(64) * 0.1
    sd["log_std"] = torch.tensor([log_std])
    return sd


def test_param_block_round_trip_and_reference_shapes():
    sd = init_state_dict()
    block = ppo.pack_params(sd)
    assert block.numel() == _native.PPO_PARAM_FLOATS == 9612
    back = ppo.unpack_params(block)
    for k, v in sd.items():
        assert torch.equal(back[k].reshape(v.shape), v.float()), k
    # the actor half is exactly the block acas2d_policy_step reads
    from gym_ACAS2D.policy import MlpActor
    assert torch.equal(MlpActor(sd).packed[:4801], block[:4801])
    obs, actions, old_logp, adv, ret = synthetic_rollout(257)
    loss, stats = ppo.reference_loss(back, obs, actions, old_logp, adv, ret, PpoConfig.sb3_defaults())
    assert loss.ndim == 0 and math.isfinite(float(loss)) and 0 < stats["clip_fraction"] < 1


def test_reference_gae_matches_hand_computation():
    r = torch.tensor([[1.0], [2.0], [3.0]]); d = torch.tensor([[0], [1], [0]], dtype=torch.uint8)
    v = torch.tensor([[0.5], [0.25], [0.125], [4.0]])
    adv, ret = ppo.reference_gae(r, d, v, 0.99, 0.95)
    a2 = 3 + 0.99 * 4.0 - 0.125
    a1 = 2 - 0.25                                     # episode ended at t = 1: no bootstrap, no carry-over
    a0 = 1 + 0.99 * 0.25 - 0.5 + 0.99 * 0.95 * a1
    assert torch.allclose(adv.reshape(-1), torch.tensor([a0, a1, a2]), atol=1e-6)
    assert torch.allclose(ret, adv + v[:3])


def test_data_parallel_gradient_exchange_world_size_2_gloo(tmp_path):
    """BASELINE config 5 on N GPUs, host logic on CPU: every rank takes the gradient of ITS minibatch, the ranks
    SUM-all-reduce the flat 9612-float gradient and scale by 1/world -- that must equal the single-process
    gradient over the concatenated minibatch, and after clip + Adam the parameter blocks must be identical on
    both ranks (torch reference learner; the CUDA learner shares ``allreduce_gradient_``)."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'gym-acas2d_b200')!r})
        from gym_ACAS2D import ppo
        from gym_ACAS2D.envs._native import PpoConfig
        from tests.test_ppo_learner import synthetic_rollout, init_state_dict
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        cfg = PpoConfig.sb3_defaults(normalize_advantage=0)
        n = 512
        data = synthetic_rollout(n, seed=11)
        lo, hi = rank * n // world, (rank + 1) * n // world
        mine = [x[lo:hi].contiguous() for x in data]
        L = ppo.TorchLearner("cpu", cfg, init_state_dict(), cuda_graph=False)
        for step in range(3):
            g = L.gradient(*mine, None)
            scale = ppo.allreduce_gradient_(g)
            assert scale == 1.0 / world
            L.apply(scale)
        blocks = [torch.zeros_like(L.params.detach()) for _ in range(world)]
        dist.all_gather(blocks, L.params.detach())
        assert torch.equal(blocks[0], blocks[1]), "ranks diverged"
        if rank == 0:
            F = ppo.TorchLearner("cpu", cfg, init_state_dict(), cuda_graph=False)
            for step in range(3):
                F.gradient(*data, None)
                F.apply(1.0)
            err = float((F.params.detach() - L.params.detach()).abs().max())
            assert err < 2e-6, err
        dist.destroy_process_group()
        print("ok", rank)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


def test_learner_checkpoint_resume_cpu():
    """Checkpoint / resume of the learner (parameter block, Adam moments, step count): a resumed learner continues
    exactly like the one that never stopped (torch learner on the CPU; the CUDA learner shares the keys)."""
    cfg = PpoConfig.sb3_defaults()
    data = synthetic_rollout(2048, seed=4)
    a = ppo.TorchLearner("cpu", cfg, init_state_dict(), cuda_graph=False)
    for _ in range(5):
        a.gradient(*data, None); a.apply(1.0)
    ckpt = {k: v.clone() for k, v in a.state_dict().items()}
    assert int(ckpt["adam_step"]) == 5 and ckpt["params"].numel() == _native.PPO_PARAM_FLOATS
    b = ppo.TorchLearner("cpu", cfg, init_state_dict(seed=99), cuda_graph=False)
    b.load_state_dict(ckpt)
    for L in (a, b):
        for _ in range(4):
            L.gradient(*data, None); L.apply(1.0)
    assert torch.equal(a.params.detach(), b.params.detach())
    assert int(b.state_dict()["adam_step"]) == 9


# ------------------------------------------------------------------------------------------------ GPU
def _group_errors(got: torch.Tensor, ref: torch.Tensor):
    """max |got - ref| per tensor of the block, relative to that tensor's largest reference entry."""
    g, r = ppo.unpack_params(got.cpu()), ppo.unpack_params(ref.cpu())
    return {k: float((g[k] - r[k]).abs().max() / (r[k].abs().max() + 1e-12)) for k in r}


@pytest.mark.gpu
@pytest.mark.parametrize("mb,normalize", [(4096, 1), (1000, 1), (64 * 148 * 2 + 17, 0), (37, 1)])
def test_fused_gradient_matches_torch_autograd(mb, normalize):
    """One minibatch gradient, kernels vs autograd of the same float32 loss: every tensor of both networks and
    log_std within 2e-4 of its largest entry, logged statistics within 1e-4; ragged / multi-tile-per-CTA /
    sub-tile minibatches included."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = "cuda:0"
    cfg = PpoConfig.sb3_defaults(normalize_advantage=normalize, ent_coef=0.01)
    n = mb + 500
    data = synthetic_rollout(n, seed=mb, device=dev)
    idx = torch.randperm(n, device=dev)[:mb].contiguous()
    sd = init_state_dict()
    F = ppo.FusedLearner(dev, cfg, sd, cuda_graph=False)
    T = ppo.TorchLearner(dev, cfg, sd, cuda_graph=False)
    got = F.gradient(*data, idx.data_ptr(), mb).clone()
    ref = T.gradient(*data, idx).detach().clone()
    errs = _group_errors(got, ref)
    assert max(errs.values()) < 2e-4, errs
    assert float(got[_native.PPO_LOG_STD + 1:].abs().max()) == 0.0 and float(got[4801:4804].abs().max()) == 0.0
    logged, want = F.logged(), T.logged()
    for k in ("policy_loss", "value_loss", "approx_kl", "clip_fraction"):
        assert abs(logged[k] - want[k]) < 1e-4 * max(1.0, abs(want[k])), (k, logged[k], want[k])
    # deterministic: a second evaluation is bit-identical
    again = F.gradient(*data, idx.data_ptr(), mb)
    assert torch.equal(again, got)
    assert int(F.adam_step) == 2
    # rows 0..mb-1 when no index list is given
    g0 = F.gradient(*data, None, mb).clone()
    r0 = T.gradient(*[x[:mb] for x in data], None).detach()
    assert max(_group_errors(g0, r0).values()) < 2e-4
    # the fused two-kernel step after split-path calls on the same learner: its barriers number the launches
    # themselves, not the Adam step count, so mixing the two paths is legal; it applies the same gradient
    before = F.params.clone()
    F.step(*data, None, mb, grad_out=True)
    torch.cuda.synchronize()
    assert max(_group_errors(F.grad, g0).values()) < 1e-5          # same rows, the four-lane summation order differs
    assert int(F.adam_step) == 4 and not torch.equal(F.params, before)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["split", "fused"])
def test_fused_adam_steps_track_torch_adam(path):
    """40 clipped Adam steps on a fixed minibatch sequence: parameter blocks stay within 1e-4 of the torch
    learner's (Adam's first steps move every parameter by ~lr regardless of gradient scale, so this is tight).
    "split" = gradient / reduction kernels then the Adam kernel (the NCCL-exchange path), "fused" = the
    two-kernel step with reduction + clip + Adam in one launch."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = "cuda:0"
    cfg = PpoConfig.sb3_defaults()
    n, mb = 8192, 1024
    data = synthetic_rollout(n, seed=5, device=dev)
    sd = init_state_dict()
    F = ppo.FusedLearner(dev, cfg, sd, cuda_graph=False)
    T = ppo.TorchLearner(dev, cfg, sd, cuda_graph=False)
    g = torch.Generator(device=dev).manual_seed(1)
    for step in range(40):
        idx = torch.randperm(n, device=dev, generator=g)[:mb].contiguous()
        if path == "split":
            F.gradient(*data, idx.data_ptr(), mb)
            F.apply(1.0)
        else:
            F.step(*data, idx.data_ptr(), mb, grad_out=True)
        ref_grad = T.gradient(*data, idx).detach().clone()
        T.apply(1.0)
        if step == 0:
            assert F.logged()["grad_norm"] > cfg.max_grad_norm            # the clip is active in this test
            assert abs(F.logged()["grad_norm"] - float(ref_grad.norm())) < 1e-4 * float(ref_grad.norm())
        if path == "fused":
            assert max(_group_errors(F.grad, ref_grad).values()) < 5e-4   # the gradient the fused kernel applied
    assert int(F.adam_step) == 40
    err = float((F.params - T.params.detach()).abs().max())
    assert err < 1e-4, err
    assert float((F.params - ppo.pack_params(sd).to(dev)).abs().max()) > 5e-3     # and they did move


@pytest.mark.gpu
def test_fused_values_and_gae_match_torch():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = "cuda:0"
    cfg = PpoConfig.sb3_defaults()
    sd = init_state_dict()
    F = ppo.FusedLearner(dev, cfg, sd, cuda_graph=False)
    T, B = 37, 1000 + 13
    g = torch.Generator().manual_seed(2)
    obs = (torch.rand(T + 1, B, 8, generator=g) * 2 - 1).to(dev)
    values = F.values(obs).view(T + 1, B)
    ref_v = ppo.reference_forward({k: v.to(dev) for k, v in sd.items()}, "vf", obs.reshape(-1, 8)).view(T + 1, B)
    assert float((values - ref_v).abs().max()) < 2e-6
    rewards = torch.randn(T, B, generator=g).to(dev)
    dones = (torch.rand(T, B, generator=g) < 0.05).to(torch.uint8).to(dev)
    adv, ret = F.gae(rewards, dones, values)
    ref_adv, ref_ret = ppo.reference_gae(rewards, dones, values, cfg.gamma, cfg.gae_lambda)
    assert float((adv - ref_adv).abs().max()) < 1e-5 and float((ret - ref_ret).abs().max()) < 1e-5


@pytest.mark.gpu
def test_epoch_graph_equals_eager_steps():
    """The captured epoch (minibatches x 4 kernels, one replay) leaves exactly the parameters the same steps
    launched one by one leave."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = "cuda:0"
    n, minibatches = 4096, 8
    data = synthetic_rollout(n, seed=9, device=dev)
    sd = init_state_dict()
    out = []
    for graph in (True, False):
        L = ppo.FusedLearner(dev, PpoConfig.sb3_defaults(), sd, cuda_graph=graph)
        L.bind(*data, minibatches)
        torch.manual_seed(77)
        for _ in range(3):
            L.epoch()
        assert int(L.adam_step) == 3 * minibatches
        out.append(L.params.clone())
    assert torch.equal(out[0], out[1])


@pytest.mark.gpu
def test_graphed_rollout_equals_eager_rollout_while_the_policy_changes():
    """``collect_rollout(graph=True)`` -- T captured policy-step launches replayed as one CUDA graph -- fills the
    same buffers bit for bit as T eager launches, over two consecutive rollouts between which the learner's
    parameter block (weights AND log_std, read from device memory by the captured launches) is modified."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from gym_ACAS2D.envs import BatchedACAS2D
    dev = torch.device("cuda", 0)
    T, B = 40, 256 * 3 + 17
    out = []
    for graph in (False, True):
        block = ppo.pack_params(init_state_dict()).to(dev)
        actor = ppo.BlockActor(block)
        env = BatchedACAS2D(B, device=dev, seed=21, auto_reset=True)
        env.reset()
        buf = None
        snaps = []
        for it in range(2):
            buf = env.collect_rollout(actor, T, noise_seed=3, step0=it * T, tensor_cores=False, buffers=buf, graph=graph)
            snaps.append({k: v.clone() for k, v in buf.items()})
            block[:4801] *= 1.01                                     # the learner moves the policy in place ...
            block[_native.PPO_LOG_STD] -= 0.25                       # ... and its log_std
        out.append((snaps, env.ppos.clone(), env.episode_counters().clone()))
    for a, b in zip(out[0][0], out[1][0]):
        for k in a:
            assert torch.equal(a[k], b[k]), k
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    assert not torch.equal(out[0][0][0]["actions"], out[0][0][1]["actions"])
    std0 = float((out[0][0][0]["logp"]).mean()); std1 = float((out[0][0][1]["logp"]).mean())
    assert std1 > std0 + 0.2                                         # smaller log_std -> larger log-densities: it was read live


@pytest.mark.gpu
def test_learner_checkpoint_resume_across_learners():
    """A FusedLearner checkpoint resumes a FusedLearner bit for bit, and a TorchLearner to within rounding."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = "cuda:0"
    cfg = PpoConfig.sb3_defaults()
    data = synthetic_rollout(4096, seed=6, device=dev)
    a = ppo.FusedLearner(dev, cfg, init_state_dict(), cuda_graph=False)
    for _ in range(5):
        a.step(*data, None, 4096)
    ckpt = a.state_dict()
    assert int(ckpt["adam_step"]) == 5
    b = ppo.FusedLearner(dev, cfg, init_state_dict(seed=99), cuda_graph=False)
    b.load_state_dict(ckpt)
    t = ppo.TorchLearner(dev, cfg, init_state_dict(seed=98), cuda_graph=False)
    t.load_state_dict(ckpt)
    for _ in range(4):
        a.step(*data, None, 4096); b.step(*data, None, 4096)
        t.gradient(*data, None); t.apply(1.0)
    assert torch.equal(a.params, b.params) and int(b.adam_step) == 9
    assert float((t.params.detach() - a.params).abs().max()) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["load_then_bind", "bind_then_load"])
def test_torch_learner_resume_with_cuda_graph(order):
    """TorchLearner with cuda_graph=True (the default on CUDA): a checkpoint loaded BEFORE bind() must survive
    bind()'s warm-up, and one loaded AFTER bind() must reach the storages the captured graph updates.  Either
    way the resumed learner continues like the one that never stopped."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    dev = "cuda:0"
    cfg = PpoConfig.sb3_defaults()
    n, mbs = 4096, 4
    data = synthetic_rollout(n, seed=8, device=dev)

    def run_epochs(L, k):
        torch.manual_seed(123)                              # same minibatch permutations for both learners
        for _ in range(k):
            L.epoch()

    a = ppo.TorchLearner(dev, cfg, init_state_dict(), cuda_graph=True)
    a.bind(*data, mbs)
    run_epochs(a, 2)
    ckpt = {k: v.clone() for k, v in a.state_dict().items()}
    assert int(ckpt["adam_step"]) == 2 * mbs
    b = ppo.TorchLearner(dev, cfg, init_state_dict(seed=77), cuda_graph=True)
    if order == "load_then_bind":
        b.load_state_dict(ckpt); b.bind(*data, mbs)
    else:
        b.bind(*data, mbs); b.load_state_dict(ckpt)
    assert int(b.state_dict()["adam_step"]) == 2 * mbs
    assert torch.equal(b.state_dict()["adam_m"], ckpt["adam_m"]) and torch.equal(b.params.detach(), ckpt["params"])
    run_epochs(a, 2); run_epochs(b, 2)
    assert int(b.state_dict()["adam_step"]) == 4 * mbs
    assert float((a.params.detach() - b.params.detach()).abs().max()) < 1e-6


@pytest.mark.gpu
def test_ppo_training_loop_fused_learner():
    """Short end-to-end run on the fused learner: finite statistics, the policy moves, episodes finish."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    hist = ppo.train(num_envs=512, n_steps=600, iterations=2, minibatches=8, n_epochs=2, tensor_cores=True,
                     learner="fused", log=None)
    assert len(hist) == 2 and hist[1]["episodes"] > 0
    assert all(np.isfinite(v) for r in hist for v in r.values())
    assert abs(hist[1]["log_std"]) > 0.0 and hist[1]["grad_norm"] > 0.0


@pytest.mark.gpu
def test_peer_memory_gradient_exchange_two_gpus(tmp_path):
    """BASELINE config 5 on N GPUs: the gradient exchange inside the update kernel (NVLink peer memory,
    release/acquire flags, rank-ordered sum) against the NCCL all-reduce path, two ranks with different
    minibatches: parameters bit-identical across ranks, both paths within 1e-6 of each other, and the captured
    epoch graph replays across GPUs.  Needs two GPUs (skipped on a one-GPU box)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'gym-acas2d_b200')!r})
        from gym_ACAS2D import ppo
        from gym_ACAS2D.envs._native import PpoConfig
        from tests.test_ppo_learner import synthetic_rollout, init_state_dict
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        rank, world = dist.get_rank(), dist.get_world_size()
        cfg = PpoConfig.sb3_defaults()
        n, minibatches = 4096, 8
        data = synthetic_rollout(n, seed=100 + rank, device=dev)          # every rank its own rollout
        out = {{}}
        for exchange, graph in (("nccl", False), ("p2p", False), ("p2p", True)):
            L = ppo.FusedLearner(dev, cfg, init_state_dict(), cuda_graph=graph, exchange=exchange)
            L.bind(*data, minibatches)
            torch.manual_seed(5)
            for _ in range(3):
                L.epoch()
            torch.cuda.synchronize()
            assert int(L.adam_step) == 3 * minibatches
            p0 = L.params.clone()
            dist.broadcast(p0, 0)
            assert torch.equal(p0, L.params), f"ranks diverged with {{exchange}}"
            out[(exchange, graph)] = L.params.clone()
        assert torch.equal(out[("p2p", False)], out[("p2p", True)])
        err = float((out[("p2p", True)] - out[("nccl", False)]).abs().max())
        assert err < 1e-6, err
        moved = float((out[("p2p", True)] - ppo.pack_params(init_state_dict()).to(dev)).abs().max())
        assert moved > 1e-3
        dist.barrier()
        dist.destroy_process_group()
        print("ok", rank, err)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29549", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert out.stdout.count("ok") == 2
