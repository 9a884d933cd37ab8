"""PPO on the batched CUDA environment (BASELINE config 5; the reference's ``training_main.py``).

The reference trains ``stable_baselines3.PPO('MlpPolicy', env, seed=13)`` for 1 048 576 steps of ONE
environment (``gym_ACAS2D/training_main.py:44-52``; 245 minutes, 71 env-steps/s).  Here the
rollout side -- actor forward, exploration noise, clipping, environment step, auto-reset -- is one fused
kernel per step over the whole batch (``BatchedACAS2D.collect_rollout``), and only the learner (GAE,
clipped surrogate, Adam) is ordinary PyTorch on the same device.  Hyper-parameters are the SB3 1.1.0
defaults stored in the reference's ``best_model.zip/data``: gamma 0.99, GAE lambda 0.95, clip 0.2,
10 epochs, lr 3e-4, vf_coef 0.5, ent_coef 0, max_grad_norm 0.5, separate 8-64-64 tanh actor / critic.
The minibatch is scaled with the batch (SB3 used 64 samples for a 2048-sample rollout, i.e. 32
minibatches per epoch; the same 32 are used here).
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from gym_ACAS2D.envs import BatchedACAS2D
from gym_ACAS2D.policy import HIDDEN, OBS_DIM, MlpActor


class ActorCritic(nn.Module):
    """SB3 ``MlpPolicy`` layout (same parameter names as ``policy.pth``)."""

    def __init__(self):
        super().__init__()
        self.policy_net = nn.Sequential(nn.Linear(OBS_DIM, HIDDEN), nn.Tanh(), nn.Linear(HIDDEN, HIDDEN), nn.Tanh())
        self.value_net_body = nn.Sequential(nn.Linear(OBS_DIM, HIDDEN), nn.Tanh(), nn.Linear(HIDDEN, HIDDEN), nn.Tanh())
        self.action_net = nn.Linear(HIDDEN, 1)
        self.value_net = nn.Linear(HIDDEN, 1)
        self.log_std = nn.Parameter(torch.zeros(1))
        for seq, gain in ((self.policy_net, math.sqrt(2)), (self.value_net_body, math.sqrt(2))):   # SB3 ortho_init
            for m in seq:
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain)
                    nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.action_net.weight, 0.01); nn.init.zeros_(self.action_net.bias)
        nn.init.orthogonal_(self.value_net.weight, 1.0); nn.init.zeros_(self.value_net.bias)

    def mean(self, obs):
        return self.action_net(self.policy_net(obs)).squeeze(-1)

    def value(self, obs):
        return self.value_net(self.value_net_body(obs)).squeeze(-1)

    def sb3_state_dict(self) -> Dict[str, torch.Tensor]:
        """Tensors under the names SB3 / ``MlpActor`` use."""
        p, v = self.policy_net, self.value_net_body
        return {"log_std": self.log_std.detach(),
                "mlp_extractor.policy_net.0.weight": p[0].weight.detach(), "mlp_extractor.policy_net.0.bias": p[0].bias.detach(),
                "mlp_extractor.policy_net.2.weight": p[2].weight.detach(), "mlp_extractor.policy_net.2.bias": p[2].bias.detach(),
                "mlp_extractor.value_net.0.weight": v[0].weight.detach(), "mlp_extractor.value_net.0.bias": v[0].bias.detach(),
                "mlp_extractor.value_net.2.weight": v[2].weight.detach(), "mlp_extractor.value_net.2.bias": v[2].bias.detach(),
                "action_net.weight": self.action_net.weight.detach(), "action_net.bias": self.action_net.bias.detach(),
                "value_net.weight": self.value_net.weight.detach(), "value_net.bias": self.value_net.bias.detach()}


class _GraphedUpdate:
    """One PPO gradient step (gather minibatch, losses, backward, grad clip, Adam) as a CUDA graph."""

    def __init__(self, net, opt, flat_obs, flat_act, flat_logp, n, mb, clip_range, vf_coef, ent_coef, max_grad_norm,
                 use_graph=True):
        dev = flat_obs.device
        self.net, self.opt = net, opt
        self.obs, self.act, self.logp = flat_obs, flat_act, flat_logp          # views of the reused rollout buffers
        self.adv = torch.zeros(n, device=dev); self.ret = torch.zeros(n, device=dev)
        self.idx = torch.zeros(mb, dtype=torch.long, device=dev)
        self.cfg = (clip_range, vf_coef, ent_coef, max_grad_norm)
        self.graph = None
        if use_graph:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                state = ([p.detach().clone() for p in net.parameters()], opt.state_dict())
                for _ in range(3):                                              # warm-up (allocations, Adam state)
                    self._body()
                with torch.no_grad():                                           # the warm-up must not train
                    for p, q in zip(net.parameters(), state[0]):
                        p.copy_(q)
                for st in opt.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self._body()

    def _body(self):
        clip_range, vf_coef, ent_coef, max_grad_norm = self.cfg
        net, idx = self.net, self.idx
        o, a, lp_old, ad, rt = self.obs[idx], self.act[idx], self.logp[idx], self.adv[idx], self.ret[idx]
        ad = (ad - ad.mean()) / (ad.std() + 1e-8)
        mean = net.mean(o)
        logp = -0.5 * ((a - mean) / net.log_std.exp()) ** 2 - net.log_std - 0.5 * math.log(2 * math.pi)
        ratio = (logp - lp_old).exp()
        pg = -torch.min(ad * ratio, ad * ratio.clamp(1 - clip_range, 1 + clip_range)).mean()
        vl = torch.nn.functional.mse_loss(net.value(o), rt)
        ent = (net.log_std + 0.5 + 0.5 * math.log(2 * math.pi)).sum()
        loss = pg + vf_coef * vl - ent_coef * ent
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        nn.utils.clip_grad_norm_(net.parameters(), max_grad_norm)
        self.opt.step()

    def load(self, adv, ret):
        self.adv.copy_(adv); self.ret.copy_(ret)

    def step(self, idx):
        self.idx.copy_(idx)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()


def train(num_envs: int = 4096, n_steps: int = 128, iterations: int = 20, device="cuda", seed: int = 13,
          gamma: float = 0.99, gae_lambda: float = 0.95, clip_range: float = 0.2, n_epochs: int = 10,
          minibatches: int = 32, lr: float = 3e-4, vf_coef: float = 0.5, ent_coef: float = 0.0,
          max_grad_norm: float = 0.5, tensor_cores: bool = True, cuda_graph: bool = True,
          log=print) -> List[Dict[str, float]]:
    """Train from scratch; returns one record per iteration (episode statistics of that iteration's
    rollout, timings).  All tensors stay on ``device``."""
    torch.manual_seed(seed)
    dev = torch.device(device)
    env = BatchedACAS2D(num_envs, device=dev, seed=seed, auto_reset=True)
    env.reset()
    net = ActorCritic().to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=lr, eps=1e-5, capturable=True)
    T, B = n_steps, num_envs
    buffers = None
    upd = None
    history: List[Dict[str, float]] = []
    for it in range(iterations):
        # ---- rollout: one fused policy + env kernel per step, written into the [T, B] buffers
        actor = MlpActor(net.sb3_state_dict(), dev)
        env.clear_stats()
        t0 = time.perf_counter()
        buffers = env.collect_rollout(actor, T, noise_seed=seed, step0=it * T, tensor_cores=tensor_cores, buffers=buffers)
        torch.cuda.synchronize(dev)
        t_roll = time.perf_counter() - t0
        stats = env.episode_stats(reduce=True)

        # ---- GAE (SB3 RolloutBuffer.compute_returns_and_advantage)
        with torch.no_grad():
            obs, acts, old_logp = buffers["obs"], buffers["actions"], buffers["logp"]
            rewards, dones = buffers["rewards"], buffers["dones"].float()
            values = net.value(obs.reshape(-1, OBS_DIM)).reshape(T + 1, B)
            adv = torch.zeros(T, B, device=dev)
            last = torch.zeros(B, device=dev)
            for t in reversed(range(T)):
                nonterminal = 1.0 - dones[t]
                delta = rewards[t] + gamma * values[t + 1] * nonterminal - values[t]
                last = delta + gamma * gae_lambda * nonterminal * last
                adv[t] = last
            returns = adv + values[:T]
            flat_obs = obs[:T].reshape(-1, OBS_DIM); flat_act = acts.reshape(-1)
            flat_logp = old_logp.reshape(-1); flat_adv = adv.reshape(-1); flat_ret = returns.reshape(-1)

        # ---- clipped-surrogate updates: one CUDA-graph replay per gradient step (the eager step is
        #      ~60 tiny kernels, i.e. host-launch-bound); the graph reads the rollout through static
        #      storages (the [T, B] buffers are reused) and a static index tensor
        t0 = time.perf_counter()
        n = T * B
        mb = n // minibatches
        if upd is None:
            upd = _GraphedUpdate(net, opt, flat_obs, flat_act, flat_logp, n, mb, clip_range, vf_coef, ent_coef,
                                 max_grad_norm, use_graph=cuda_graph)
        upd.load(flat_adv, flat_ret)
        for _ in range(n_epochs):
            perm = torch.randperm(n, device=dev)
            for k in range(minibatches):
                upd.step(perm[k * mb:(k + 1) * mb])
        torch.cuda.synchronize(dev)
        t_learn = time.perf_counter() - t0
        rec = dict(iteration=it, env_steps=(it + 1) * n, episodes=stats["episodes"], mean_return=stats["mean_return"],
                   goal_rate=stats["goal_rate"], collision_rate=stats["collision_rate"], timeout_rate=stats["timeout_rate"],
                   mean_length=stats["mean_length"], log_std=float(net.log_std.detach()), rollout_s=t_roll, learn_s=t_learn,
                   rollout_env_steps_per_s=n / t_roll)
        history.append(rec)
        if log:
            log("it {iteration:3d}  steps {env_steps:>10d}  episodes {episodes:6d}  return {mean_return:8.1f}  goal {goal_rate:.3f}  "
                "coll {collision_rate:.3f}  tout {timeout_rate:.3f}  len {mean_length:6.1f}  log_std {log_std:+.2f}  "
                "rollout {rollout_s:.3f}s ({rollout_env_steps_per_s:.3g} steps/s)  learn {learn_s:.2f}s".format(**rec))
    train.last_policy = net
    return history


def evaluate(actor: MlpActor, episodes: int = 4096, device="cuda", seed: int = 99, tensor_cores: bool = False) -> Dict[str, float]:
    """Deterministic evaluation, ``testing_main.py:62-108`` style (``model.predict(obs, deterministic=True)``):
    the first episode of ``episodes`` envs, statistics as in the reference's notebooks."""
    dev = torch.device(device)
    env = BatchedACAS2D(episodes, device=dev, seed=seed, auto_reset=True)
    env.reset()
    finished = torch.zeros(episodes, dtype=torch.bool, device=dev)
    outcome = torch.zeros(episodes, dtype=torch.uint8, device=dev)
    length = torch.zeros(episodes, dtype=torch.int32, device=dev)
    ret = torch.zeros(episodes, device=dev)
    for _ in range(int(env.params.max_steps) + 1):
        _, _, d = env.policy_step(actor, deterministic=True, tensor_cores=tensor_cores)
        new = d & ~finished
        outcome[new] = env.outcome[new]; length[new] = env.ep_length[new]; ret[new] = env.ep_return[new]
        finished |= new
    return dict(episodes=episodes, goal_rate=float((outcome == 1).float().mean()),
                collision_rate=float((outcome == 2).float().mean()), timeout_rate=float((outcome == 3).float().mean()),
                mean_steps=float(length.float().mean()), mean_return=float(ret.mean()), std_return=float(ret.std()))


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--minibatches", type=int, default=32)
    ap.add_argument("--fp32", action="store_true", help="CUDA-core float32 actor instead of tcgen05 TF32")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    hist = train(a.envs, a.n_steps, a.iterations, minibatches=a.minibatches, tensor_cores=not a.fp32)
    import os
    result = {"training": hist}
    trained = MlpActor(train.last_policy.sb3_state_dict(), "cuda")
    result["eval_trained_here"] = evaluate(trained)
    print("deterministic eval, policy trained here:      ", result["eval_trained_here"])
    fixture = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                           "tests", "golden", "ppo_policy_1048576_11.npz")
    if os.path.exists(fixture):
        result["eval_reference_agent"] = evaluate(MlpActor.from_file(fixture, "cuda"))
        print("deterministic eval, the reference's saved agent:", result["eval_reference_agent"])
    if a.out:
        json.dump(result, open(a.out, "w"), indent=1)
