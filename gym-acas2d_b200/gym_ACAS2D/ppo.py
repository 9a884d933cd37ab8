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

Two learners share the loss (``reference_loss``): ``FusedLearner`` -- this repo's own kernels behind
``acas2d_ppo_values / gae / grad / adam`` (``csrc/acas2d_ppo.cuh``), a whole epoch of gradient steps replayed
as one CUDA graph -- and ``TorchLearner``, the plain torch float32 autograd version the kernels are checked
against.  Under ``torchrun`` every rank owns a shard of the env batch (global env ids, so spawns do not
depend on the GPU count), draws its minibatches from its own rollout and the ranks average one 38 KB
gradient per gradient step -- inside the update kernel itself, through NVLink peer memory
(``exchange="p2p"``: symmetric-memory blocks, release/acquire flags, rank-ordered sum, no NCCL call, so the
epoch graph also replays on N GPUs), or with an NCCL all-reduce between the gradient and the Adam kernels
(``exchange="nccl"``).  Either way the parameters stay bit-identical on all ranks.
"""
from __future__ import annotations

import ctypes
import math
import time
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from gym_ACAS2D.envs import BatchedACAS2D, _native
from gym_ACAS2D.envs._native import PpoConfig
from gym_ACAS2D.policy import HIDDEN, OBS_DIM, MlpActor

_NET_KEYS = {"pi": ("mlp_extractor.policy_net.0", "mlp_extractor.policy_net.2", "action_net"),
             "vf": ("mlp_extractor.value_net.0", "mlp_extractor.value_net.2", "value_net")}


def pack_params(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """SB3-named tensors -> the flat float32 block of ``acas2d_ppo_*`` (include/acas2d_b200.h):
    actor W1|b1|W2|b2|W3|b3|pad, critic likewise, log_std, pad."""
    out = torch.zeros(_native.PPO_PARAM_FLOATS, dtype=torch.float32)
    for n, names in enumerate((_NET_KEYS["pi"], _NET_KEYS["vf"])):
        parts = []
        for name in names:
            parts += [sd[name + ".weight"].detach().float().cpu().reshape(-1), sd[name + ".bias"].detach().float().cpu().reshape(-1)]
        flat = torch.cat(parts)
        out[n * _native.POLICY_FLOATS: n * _native.POLICY_FLOATS + flat.numel()] = flat
    out[_native.PPO_LOG_STD] = sd["log_std"].detach().float().cpu().reshape(-1)[0]
    return out


def unpack_params(block: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Inverse of ``pack_params`` (views of ``block``), under the names SB3 / ``MlpActor`` use."""
    sd: Dict[str, torch.Tensor] = {}
    shapes = ((HIDDEN, OBS_DIM), (HIDDEN, HIDDEN), (1, HIDDEN))
    for n, names in enumerate((_NET_KEYS["pi"], _NET_KEYS["vf"])):
        o = n * _native.POLICY_FLOATS
        for name, shp in zip(names, shapes):
            k = shp[0] * shp[1]
            sd[name + ".weight"] = block[o:o + k].view(shp); o += k
            sd[name + ".bias"] = block[o:o + shp[0]]; o += shp[0]
    sd["log_std"] = block[_native.PPO_LOG_STD:_native.PPO_LOG_STD + 1]
    return sd


def reference_forward(sd: Dict[str, torch.Tensor], net: str, obs: torch.Tensor) -> torch.Tensor:
    a, b, c = _NET_KEYS[net]
    h = torch.tanh(obs @ sd[a + ".weight"].T + sd[a + ".bias"])
    h = torch.tanh(h @ sd[b + ".weight"].T + sd[b + ".bias"])
    return (h @ sd[c + ".weight"].T + sd[c + ".bias"]).squeeze(-1)


def reference_loss(sd: Dict[str, torch.Tensor], obs, actions, old_logp, adv, ret, cfg: PpoConfig):
    """SB3 1.1.0 ``PPO.train()`` loss of one minibatch in plain torch float32 -- the numerics reference of
    ``acas2d_ppo_grad``.  Returns (loss, dict of the logged statistics)."""
    if cfg.normalize_advantage:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    log_std = sd["log_std"].reshape(())
    mean = reference_forward(sd, "pi", obs)
    logp = -0.5 * ((actions - mean) / log_std.exp()) ** 2 - log_std - 0.5 * math.log(2 * math.pi)
    log_ratio = logp - old_logp
    ratio = log_ratio.exp()
    pg = -torch.min(adv * ratio, adv * ratio.clamp(1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
    vl = torch.nn.functional.mse_loss(reference_forward(sd, "vf", obs), ret)
    ent = log_std + 0.5 + 0.5 * math.log(2 * math.pi)
    loss = pg + cfg.vf_coef * vl - cfg.ent_coef * ent
    with torch.no_grad():
        stats = dict(policy_loss=float(pg), value_loss=float(vl), approx_kl=float(((ratio - 1) - log_ratio).mean()),
                     clip_fraction=float(((ratio - 1).abs() > cfg.clip_range).float().mean()))
    return loss, stats


def reference_gae(rewards, dones, values, gamma: float, lam: float):
    """SB3 ``RolloutBuffer.compute_returns_and_advantage`` (torch, any device): rewards / dones [T, B],
    values [T+1, B] -> advantages, returns [T, B]."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    d = dones.float()
    for t in reversed(range(T)):
        nonterminal = 1.0 - d[t]
        delta = rewards[t] + gamma * values[t + 1] * nonterminal - values[t]
        last = delta + gamma * lam * nonterminal * last
        adv[t] = last
    return adv, adv + values[:T]


def allreduce_gradient_(grad: torch.Tensor) -> float:
    """Data-parallel exchange of one gradient step: SUM all-reduce of the flat gradient over the ranks
    (NCCL on GPUs, gloo in the CPU tests).  Returns the scale (1 / world size) that turns the sum of the
    ranks' minibatch-mean gradients into the mean over the global minibatch."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1.0
    dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()


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


class BlockActor:
    """The actor half of a learner's live parameter block, in the shape ``BatchedACAS2D.policy_step`` expects
    from an ``MlpActor`` (the block starts with the ``acas2d_policy_step`` weight layout, so nothing is copied).
    ``log_std_ptr`` lets the rollout kernels read the live log_std from the block itself."""

    def __init__(self, block: torch.Tensor):
        self.packed, self.device = block, block.device
        self.log_std_ptr = block.data_ptr() + 4 * _native.PPO_LOG_STD

    @property
    def log_std(self) -> float:
        return float(self.packed[_native.PPO_LOG_STD])           # a host read; the graphed rollout does not need it

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("the learner's parameter block lives on " + str(self.device))
        return self


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class FusedLearner:
    """PPO learner on this repo's kernels (``csrc/acas2d_ppo.cuh`` behind ``acas2d_ppo_*``): critic forward,
    GAE, and per gradient step two kernels -- minibatch gradient of both networks, then reduction + gradient
    exchange + grad-norm clip + Adam.  float32, deterministic.  A whole epoch (``minibatches`` gradient
    steps) is one CUDA-graph replay, on one GPU and -- with the peer-memory exchange -- on N."""

    def __init__(self, device, cfg: Optional[PpoConfig] = None, init: Optional[Dict[str, torch.Tensor]] = None,
                 cuda_graph: bool = True, exchange: str = "p2p"):
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("FusedLearner runs on CUDA devices only; there is no CPU fallback (use TorchLearner)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.device = dev
        self.lib = _native.load()
        self.cfg = cfg or PpoConfig.sb3_defaults()
        self.params = pack_params(init if init is not None else ActorCritic().sb3_state_dict()).to(dev)
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)   # noqa: E731
        self.adam_m, self.adam_v, self.grad = z(_native.PPO_PARAM_FLOATS), z(_native.PPO_PARAM_FLOATS), z(_native.PPO_PARAM_FLOATS)
        self.workspace = z(_native.PPO_WORKSPACE_FLOATS)
        self.loss_stats = z(_native.PPO_LOSS_STATS)
        self.sync = z(4, dtype=torch.int32)              # [0] = Adam step count, [1] = barrier arrivals
        self.adam_step = self.sync[0:1]
        self.cuda_graph = cuda_graph
        self.rank, self.world = _world()
        self.exchange = exchange if self.world > 1 else "none"
        self._peer_ptrs = None
        if self.exchange == "p2p":
            self._open_peer_blocks()
        with torch.cuda.device(dev):
            _native.check(self.lib.acas2d_ppo_prepare(), "acas2d_ppo_prepare")
        self.launches = 0
        self._bound = None
        self._graph = None

    def _open_peer_blocks(self) -> None:
        """One exchange block per rank in symmetric memory (CUDA VMM, peer-mapped over NVLink): ``buffer_ptrs``
        are every rank's block as seen from this process."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if self.world > _native.PPO_MAX_RANKS:
            raise ValueError(f"the peer-memory exchange supports up to {_native.PPO_MAX_RANKS} ranks")
        group = dist.group.WORLD
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:  # noqa: BLE001  (newer torch enables every group implicitly)
            pass
        with torch.cuda.device(self.device):
            self._xbuf = symm.empty(_native.PPO_EXCHANGE_FLOATS, dtype=torch.float32, device=self.device)
            self._xbuf.zero_()
            torch.cuda.synchronize(self.device)
            self._xhdl = symm.rendezvous(self._xbuf, group)
        ptrs = [int(x) for x in self._xhdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self._xbuf.data_ptr()
        self._peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        dist.barrier()                                   # every block is zeroed before anyone signals

    # ---- plumbing
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _c(self):
        return ctypes.byref(self.cfg)

    def actor(self) -> BlockActor:
        return BlockActor(self.params)

    def sb3_state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.clone() for k, v in unpack_params(self.params).items()}

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Checkpoint of the learner: parameter block, Adam moments and step count (same keys as ``TorchLearner``)."""
        return dict(params=self.params.detach().clone(), adam_m=self.adam_m.clone(), adam_v=self.adam_v.clone(),
                    adam_step=self.adam_step.clone().to(torch.int64))

    def load_state_dict(self, d: Dict[str, torch.Tensor]) -> None:
        """Resume in place (captured graphs stay valid: the storages do not move)."""
        self.params.copy_(d["params"]); self.adam_m.copy_(d["adam_m"]); self.adam_v.copy_(d["adam_v"])
        self.adam_step.copy_(d["adam_step"].to(torch.int32))

    # ---- value head and GAE
    def values(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        obs = obs.reshape(-1, OBS_DIM)
        assert obs.is_contiguous() and obs.dtype == torch.float32 and obs.device == self.device
        n = obs.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_ppo_values(self.params.data_ptr(), obs.data_ptr(), n, out.data_ptr(),
                                                     self._stream()), "acas2d_ppo_values")
        self.launches += 1
        return out

    def gae(self, rewards, dones, values, adv=None, ret=None):
        T, B = rewards.shape
        assert values.shape == (T + 1, B) and dones.dtype == torch.uint8
        assert rewards.is_contiguous() and dones.is_contiguous() and values.is_contiguous()
        adv = torch.empty_like(rewards) if adv is None else adv
        ret = torch.empty_like(rewards) if ret is None else ret
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_ppo_gae(self._c(), rewards.data_ptr(), dones.data_ptr(), values.data_ptr(),
                                                  T, B, adv.data_ptr(), ret.data_ptr(), self._stream()), "acas2d_ppo_gae")
        self.launches += 1
        return adv, ret

    # ---- gradient steps
    def gradient(self, obs, actions, old_logp, adv, ret, idx_ptr: Optional[int], mb: int) -> torch.Tensor:
        """``grad`` <- this rank's gradient of the PPO loss over rows ``idx`` (device pointer to int64[mb]; None =
        first mb rows).  Two kernels (gradient, reduction); increments the Adam step count."""
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_ppo_grad(
                self._c(), self.params.data_ptr(), obs.data_ptr(), actions.data_ptr(), old_logp.data_ptr(),
                adv.data_ptr(), ret.data_ptr(), idx_ptr, int(mb), self.workspace.data_ptr(), self.grad.data_ptr(),
                self.loss_stats.data_ptr(), self.sync.data_ptr(), self._stream()), "acas2d_ppo_grad")
        self.launches += 2
        return self.grad

    def apply(self, grad_scale: float = 1.0) -> None:
        """Clip + Adam on ``grad * grad_scale`` (one kernel)."""
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_ppo_adam(self._c(), self.params.data_ptr(), self.grad.data_ptr(), float(grad_scale),
                                                   self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.sync.data_ptr(),
                                                   self.loss_stats.data_ptr(), self._stream()), "acas2d_ppo_adam")
        self.launches += 1

    def step(self, obs, actions, old_logp, adv, ret, idx_ptr: Optional[int], mb: int, grad_out: bool = False) -> None:
        """One whole gradient step: two kernels with the fused update (``exchange`` "none" / "p2p"), or gradient ->
        NCCL all-reduce -> Adam (``exchange`` "nccl")."""
        if self.exchange == "nccl":
            self.gradient(obs, actions, old_logp, adv, ret, idx_ptr, mb)
            self.apply(allreduce_gradient_(self.grad))
            return
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_ppo_step(
                self._c(), self.params.data_ptr(), obs.data_ptr(), actions.data_ptr(), old_logp.data_ptr(),
                adv.data_ptr(), ret.data_ptr(), idx_ptr, int(mb), self.workspace.data_ptr(), self.adam_m.data_ptr(),
                self.adam_v.data_ptr(), self.sync.data_ptr(), self.loss_stats.data_ptr(),
                self.grad.data_ptr() if grad_out else None, self.rank, self.world, self._peer_ptrs, self._stream()),
                "acas2d_ppo_step")
        self.launches += 2

    def _step(self, k: int) -> None:
        obs, act, logp, adv, ret, perm, mb = self._bound
        self.step(obs, act, logp, adv, ret, perm.data_ptr() + 8 * k * mb, mb)

    def bind(self, obs, actions, old_logp, adv, ret, minibatches: int) -> None:
        """Fix the (reused) flattened rollout storages the epochs read and capture one epoch as a CUDA graph
        (not with the NCCL exchange: its all-reduces stay eager)."""
        n = actions.numel()
        for x in (obs, actions, old_logp, adv, ret):
            assert x.is_contiguous() and x.dtype == torch.float32 and x.device == self.device
        perm = torch.arange(n, dtype=torch.int64, device=self.device)
        self._bound = (obs, actions, old_logp, adv, ret, perm, n // minibatches)
        self.minibatches = minibatches
        self._graph = None
        if self.cuda_graph and self.exchange != "nccl":
            torch.cuda.synchronize(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    for k in range(minibatches):
                        self._step(k)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph = graph

    def epoch(self) -> None:
        """One pass over the bound rollout in ``minibatches`` shuffled minibatches."""
        perm = self._bound[5]
        torch.randperm(perm.numel(), device=self.device, out=perm)
        if self._graph is not None:
            self._graph.replay()
            self.launches += 2 * self.minibatches
        else:
            for k in range(self.minibatches):
                self._step(k)

    def check(self) -> None:
        """Raises if the update kernel recorded a timed-out grid barrier / peer flag (``sync[1]``, see
        ``acas2d_ppo_step``): the kernel's spins are bounded, so a dead or late peer rank ends in this error
        instead of a hung GPU."""
        code = int(self.sync[1].item())
        if code:
            what = {1: "a grid barrier of the update kernel timed out",
                    2: "a peer rank's gradient did not arrive (dead or late rank)"}.get(code, f"error {code}")
            raise RuntimeError(f"FusedLearner: {what}; the last update is invalid")

    def logged(self) -> Dict[str, float]:
        self.check()
        s = self.loss_stats.tolist()
        return dict(policy_loss=s[0], value_loss=s[1], approx_kl=s[2], clip_fraction=s[3], grad_norm=s[4])


class TorchLearner:
    """The same learner in plain torch float32 (autograd, ``clip_grad_norm_``, ``torch.optim.Adam``): the numerics
    reference of ``FusedLearner`` and the round-1 implementation (one CUDA graph per gradient step)."""

    def __init__(self, device, cfg: Optional[PpoConfig] = None, init: Optional[Dict[str, torch.Tensor]] = None,
                 cuda_graph: bool = True):
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.cfg = cfg or PpoConfig.sb3_defaults()
        block = pack_params(init if init is not None else ActorCritic().sb3_state_dict()).to(dev)
        self.params = block.requires_grad_(True)
        self.opt = torch.optim.Adam([self.params], lr=self.cfg.lr, betas=(self.cfg.beta1, self.cfg.beta2),
                                    eps=self.cfg.adam_eps, capturable=dev.type == "cuda")
        self.cuda_graph = cuda_graph and dev.type == "cuda"
        self.stats: Dict[str, float] = {}
        self._bound = None
        self._graph = None

    def actor(self):
        return MlpActor({k: v.detach() for k, v in unpack_params(self.params).items()}, self.device)

    def sb3_state_dict(self):
        return {k: v.detach().clone() for k, v in unpack_params(self.params).items()}

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Checkpoint with the keys of ``FusedLearner.state_dict`` (the two learners can resume each other)."""
        st = self.opt.state.get(self.params, {})
        zeros = torch.zeros_like(self.params.detach())
        step = st.get("step", torch.zeros(()))
        return dict(params=self.params.detach().clone(), adam_m=st.get("exp_avg", zeros).clone(),
                    adam_v=st.get("exp_avg_sq", zeros).clone(),
                    adam_step=torch.as_tensor(step).detach().reshape(1).to(torch.int64).cpu())

    def load_state_dict(self, d: Dict[str, torch.Tensor]) -> None:
        """Resume.  Existing optimiser-state tensors are overwritten IN PLACE: a CUDA graph captured by ``bind``
        keeps updating those storages, so replacing them would silently drop the restored moments."""
        with torch.no_grad():
            self.params.copy_(d["params"])
            capturable = self.device.type == "cuda"
            step = d["adam_step"].reshape(()).to(torch.float32)
            st = self.opt.state.get(self.params)
            if st and all(k in st for k in ("step", "exp_avg", "exp_avg_sq")):
                st["exp_avg"].copy_(d["adam_m"]); st["exp_avg_sq"].copy_(d["adam_v"])
                st["step"].copy_(step) if torch.is_tensor(st["step"]) else st.__setitem__("step", step.cpu())
            else:
                self.opt.state[self.params] = dict(step=step.to(self.device) if capturable else step.cpu(),
                                                   exp_avg=d["adam_m"].to(self.device).clone(),
                                                   exp_avg_sq=d["adam_v"].to(self.device).clone())

    def values(self, obs, out=None):
        with torch.no_grad():
            v = reference_forward(unpack_params(self.params), "vf", obs.reshape(-1, OBS_DIM))
        return v if out is None else out.copy_(v)

    def gae(self, rewards, dones, values, adv=None, ret=None):
        a, r = reference_gae(rewards, dones, values, self.cfg.gamma, self.cfg.gae_lambda)
        return (a, r) if adv is None else (adv.copy_(a), ret.copy_(r))

    def gradient(self, obs, actions, old_logp, adv, ret, idx: Optional[torch.Tensor]) -> torch.Tensor:
        sel = (lambda x: x) if idx is None else (lambda x: x[idx])
        loss, self.stats = reference_loss(unpack_params(self.params), sel(obs), sel(actions), sel(old_logp), sel(adv),
                                          sel(ret), self.cfg)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        return self.params.grad

    def apply(self, grad_scale: float = 1.0) -> None:
        if grad_scale != 1.0:
            self.params.grad.mul_(grad_scale)
        if self.cfg.max_grad_norm > 0:
            nn.utils.clip_grad_norm_([self.params], self.cfg.max_grad_norm)
        self.opt.step()

    def _body(self):
        obs, act, logp, adv, ret, _, _ = self._bound
        o, a, lp, ad, rt = obs[self._idx], act[self._idx], logp[self._idx], adv[self._idx], ret[self._idx]
        loss, _ = _loss_no_sync(unpack_params(self.params), o, a, lp, ad, rt, self.cfg)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if _world()[1] > 1:
            self.params.grad.mul_(allreduce_gradient_(self.params.grad))
        if self.cfg.max_grad_norm > 0:
            nn.utils.clip_grad_norm_([self.params], self.cfg.max_grad_norm)
        self.opt.step()

    def bind(self, obs, actions, old_logp, adv, ret, minibatches: int) -> None:
        n = actions.numel()
        perm = torch.arange(n, dtype=torch.int64, device=self.device)
        mb = n // minibatches
        self._bound = (obs, actions, old_logp, adv, ret, perm, mb)
        self.minibatches = minibatches
        self._idx = torch.zeros(mb, dtype=torch.int64, device=self.device)
        self._graph = None
        if self.cuda_graph and _world()[1] == 1:
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                # the warm-up (allocations, Adam state creation) must neither train nor lose a resumed state:
                # parameters and the optimiser moments / step count are put back afterwards, into the same storages
                prior = self.opt.state.get(self.params)
                saved_params = self.params.detach().clone()
                saved_state = {k: v.detach().clone() for k, v in prior.items() if torch.is_tensor(v)} if prior else None
                for _ in range(3):
                    self._body()
                with torch.no_grad():
                    self.params.copy_(saved_params)
                    for k, v in self.opt.state[self.params].items():
                        if torch.is_tensor(v):
                            v.copy_(saved_state[k]) if saved_state is not None and k in saved_state else v.zero_()
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self._graph):
                self._body()

    def epoch(self) -> None:
        perm, mb = self._bound[5], self._bound[6]
        torch.randperm(perm.numel(), device=self.device, out=perm)
        for k in range(self.minibatches):
            self._idx.copy_(perm[k * mb:(k + 1) * mb])
            if self._graph is not None:
                self._graph.replay()
            else:
                self._body()

    def logged(self) -> Dict[str, float]:
        return dict(self.stats)


def _loss_no_sync(sd, obs, actions, old_logp, adv, ret, cfg):
    """``reference_loss`` without the host reads of its logged statistics (CUDA-graph capturable)."""
    if cfg.normalize_advantage:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    log_std = sd["log_std"].reshape(())
    mean = reference_forward(sd, "pi", obs)
    logp = -0.5 * ((actions - mean) / log_std.exp()) ** 2 - log_std - 0.5 * math.log(2 * math.pi)
    ratio = (logp - old_logp).exp()
    pg = -torch.min(adv * ratio, adv * ratio.clamp(1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
    vl = torch.nn.functional.mse_loss(reference_forward(sd, "vf", obs), ret)
    ent = log_std + 0.5 + 0.5 * math.log(2 * math.pi)
    return pg + cfg.vf_coef * vl - cfg.ent_coef * ent, None


def train(num_envs: int = 4096, n_steps: int = 128, iterations: int = 20, device="cuda", seed: int = 13,
          gamma: float = 0.99, gae_lambda: float = 0.95, clip_range: float = 0.2, n_epochs: int = 10,
          minibatches: int = 32, lr: float = 3e-4, vf_coef: float = 0.5, ent_coef: float = 0.0,
          max_grad_norm: float = 0.5, tensor_cores: bool = True, cuda_graph: bool = True, learner: str = "fused",
          exchange: str = "p2p", log=print) -> List[Dict[str, float]]:
    """Train from scratch; returns one record per iteration (episode statistics of that iteration's rollout
    over ALL ranks, timings).  ``num_envs`` is per rank; all tensors stay on ``device``.  ``learner``:
    "fused" = this repo's kernels, "torch" = the autograd reference."""
    torch.manual_seed(seed)                                   # same initial parameters on every rank
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    rank, world = _world()
    cfg = PpoConfig.sb3_defaults(gamma=gamma, gae_lambda=gae_lambda, clip_range=clip_range, vf_coef=vf_coef,
                                 ent_coef=ent_coef, max_grad_norm=max_grad_norm, lr=lr)
    # the GPU VecEnv adapter (BASELINE config 5: "PPO rollout via GPU VecEnv adapter"); its zero-copy tensor surface
    from gym_ACAS2D.envs.vec_env import ACAS2DVecEnv
    venv = ACAS2DVecEnv(num_envs, device=dev, seed=seed, env_id_offset=rank * num_envs)
    venv.reset_tensor()
    env = venv.core
    init = ActorCritic().sb3_state_dict()
    if learner == "fused":
        L = FusedLearner(dev, cfg, init, cuda_graph=cuda_graph, exchange=exchange)
    else:
        L = TorchLearner(dev, cfg, init, cuda_graph=cuda_graph)
    T, B = n_steps, num_envs
    n = T * B
    buffers = None
    values = torch.empty(T + 1, B, device=dev)
    adv, ret = torch.empty(T, B, device=dev), torch.empty(T, B, device=dev)
    history: List[Dict[str, float]] = []
    for it in range(iterations):
        # ---- rollout: one fused policy + env kernel per step, written into the [T, B] buffers
        actor = L.actor()
        env.clear_stats()
        t0 = time.perf_counter()
        buffers = venv.collect_rollout(actor, T, noise_seed=seed, step0=it * T, tensor_cores=tensor_cores, buffers=buffers,
                                       graph=cuda_graph and learner == "fused")
        torch.cuda.synchronize(dev)
        t_roll = time.perf_counter() - t0
        stats = env.episode_stats(reduce=True)

        # ---- critic forward over the T+1 observation rows, GAE (SB3 RolloutBuffer.compute_returns_and_advantage)
        t0 = time.perf_counter()
        L.values(buffers["obs"], out=values.view(-1))
        L.gae(buffers["rewards"], buffers["dones"], values, adv, ret)
        if L._bound is None:                                   # the [T, B] storages are reused by every iteration
            L.bind(buffers["obs"][:T].reshape(-1, OBS_DIM), buffers["actions"].reshape(-1), buffers["logp"].reshape(-1),
                   adv.reshape(-1), ret.reshape(-1), minibatches)
        # ---- clipped-surrogate epochs
        for _ in range(n_epochs):
            L.epoch()
        torch.cuda.synchronize(dev)
        t_learn = time.perf_counter() - t0
        rec = dict(iteration=it, env_steps=(it + 1) * n * world, episodes=stats["episodes"], mean_return=stats["mean_return"],
                   goal_rate=stats["goal_rate"], collision_rate=stats["collision_rate"], timeout_rate=stats["timeout_rate"],
                   mean_length=stats["mean_length"], log_std=float(L.params.detach()[_native.PPO_LOG_STD]), rollout_s=t_roll,
                   learn_s=t_learn, rollout_env_steps_per_s=n * world / t_roll, **L.logged())
        history.append(rec)
        if log and rank == 0:
            log("it {iteration:3d}  steps {env_steps:>10d}  episodes {episodes:6d}  return {mean_return:8.1f}  goal {goal_rate:.3f}  "
                "coll {collision_rate:.3f}  tout {timeout_rate:.3f}  len {mean_length:6.1f}  log_std {log_std:+.2f}  "
                "rollout {rollout_s:.3f}s ({rollout_env_steps_per_s:.3g} steps/s)  learn {learn_s:.3f}s".format(**rec))
    train.last_policy = L
    train.param_divergence = 0.0
    if world > 1:                                              # the ranks must hold bit-identical parameters
        import torch.distributed as dist
        p0 = L.params.detach().clone()
        dist.broadcast(p0, 0)
        diff = (p0 - L.params.detach()).abs().max()
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        train.param_divergence = float(diff)
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
    # single GPU:  python -m gym_ACAS2D.ppo --envs 1024 --n-steps 1024 --iterations 70 --minibatches 256
    # N GPUs:      python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m gym_ACAS2D.ppo ...
    import argparse
    import json
    import os
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="envs per rank")
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--minibatches", type=int, default=32)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--learner", default="fused", choices=("fused", "torch"))
    ap.add_argument("--fp32", action="store_true", help="CUDA-core float32 actor instead of tcgen05 TF32")
    ap.add_argument("--exchange", default="p2p", choices=("p2p", "nccl"),
                    help="N > 1: gradient exchange inside the update kernel over NVLink peer memory, or an NCCL all-reduce")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    dev = f"cuda:{torch.cuda.current_device()}"
    t_start = time.perf_counter()
    hist = train(a.envs, a.n_steps, a.iterations, device=dev, minibatches=a.minibatches, n_epochs=a.epochs,
                 tensor_cores=not a.fp32, learner=a.learner, exchange=a.exchange)
    wall = time.perf_counter() - t_start
    if _world()[0] == 0:
        result = {"training": hist, "wall_s": wall, "world_size": world, "learner": a.learner,
                  "param_divergence_over_ranks": train.param_divergence, "exchange": a.exchange if world > 1 else "none",
                  "rollout_s": sum(h["rollout_s"] for h in hist), "learn_s": sum(h["learn_s"] for h in hist)}
        print(f"world {world}  parameter divergence over ranks {train.param_divergence}")
        print(f"wall {wall:.1f} s  (rollouts {result['rollout_s']:.2f} s, learner {result['learn_s']:.2f} s)")
        trained = MlpActor(train.last_policy.sb3_state_dict(), dev)
        result["eval_trained_here"] = evaluate(trained, device=dev)
        print("deterministic eval, policy trained here:      ", result["eval_trained_here"])
        fixture = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                               "tests", "golden", "ppo_policy_1048576_11.npz")
        if os.path.exists(fixture):
            result["eval_reference_agent"] = evaluate(MlpActor.from_file(fixture, dev), device=dev)
            print("deterministic eval, the reference's saved agent:", result["eval_reference_agent"])
        if a.out:
            json.dump(result, open(a.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
