"""``BatchedACAS2D`` -- B independent ACAS-2D games stepped by one CUDA kernel per call.

This is the vectorised surface the north star adds next to the reference's single-env
API: ``step(actions[B]) -> (obs[B, 5+3N], reward[B], done[B])`` on device tensors, with the
semantics of ``ACAS2DEnv.step`` (reference gym_ACAS2D/envs/environment.py:29-42) applied to
every env, and of ``ACAS2DEnv.reset`` (environment.py:44-48) on reset / auto-reset.

PyTorch is used for what it is good at here -- device memory, streams, CUDA graphs and
``torch.distributed`` -- and nothing else: all arithmetic happens in
``csrc/acas2d_kernels.cu`` behind the C ABI of ``include/acas2d_b200.h``.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _native
from ._native import Params, State, StepAux


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("BatchedACAS2D needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"BatchedACAS2D runs on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class BatchedACAS2D:
    """SoA state of ``num_envs`` games in HBM + the kernels that advance it.

    Parameters
    ----------
    num_envs       B, envs owned by this process / GPU.
    n_traffic      intruders per env (default: ``settings.MAX_TRAFFIC``).
    seed           Philox key for spawns (default ``settings.RANDOM_SEED``).
    env_id_offset  global id of local env 0; spawns depend on the global id only, so a batch
                   sharded over R ranks reproduces the single-GPU batch exactly.
    auto_reset     True: SB3 VecEnv semantics (finished envs respawn inside ``step`` and the
                   returned obs row is the reset observation).  False: reference ``ACAS2DEnv``
                   semantics (a finished game stays in place until ``reset``).
    track_min_sep  keep the per-episode minimum separation (reference ``d_sep_record``,
                   game.py:237); costs 8 B/env-step of extra HBM traffic.
    """

    def __init__(self, num_envs: int, n_traffic: Optional[int] = None, device="cuda", seed: Optional[int] = None,
                 env_id_offset: int = 0, auto_reset: bool = True, track_min_sep: bool = False,
                 settings=None, host_zero_copy: bool = False, **setting_overrides):
        self.lib = _native.load()
        self.device = _require_cuda(device)
        self.params: Params = _native.params_from_settings(settings, n_traffic, auto_reset, **setting_overrides)
        if seed is None:
            from gym_ACAS2D.settings import RANDOM_SEED
            seed = RANDOM_SEED
        self.num_envs = B = int(num_envs)
        self.n_traffic = N = int(self.params.n_traffic)
        self.obs_dim = L = 5 + 3 * N
        self.seed, self.env_id_offset = int(seed), int(env_id_offset)
        self.auto_reset = bool(auto_reset)
        dev, f64, f32 = self.device, torch.float64, torch.float32
        with torch.cuda.device(dev):
            # ---- state (layout: include/acas2d_b200.h)
            self.ppos = torch.zeros(B, 2, dtype=f64, device=dev)
            self.paux = torch.zeros(B, 2, dtype=f64, device=dev)          # 16 B records {psi, steps, ep_return}
            self.thot = torch.zeros(B, N, 4, dtype=f32, device=dev)          # {x0, y0, psi, v}: all a step reads
            self.tres = torch.zeros(B, N, 4, dtype=f64, device=dev)          # cold float64 remainders (injected states)
            # kinematic cache (N > 1): 24 B / intruder {x0, y0 float32; dx, dy float64}, written at spawn / injection
            self.tkin = torch.zeros(B, N, 3, dtype=f64, device=dev) if N > 1 else None
            self.tpsi0 = torch.zeros(B, dtype=f32, device=dev) if N == 1 else None     # compact intruder-0 headings
            # per-step scratch of the player pre-pass (N > 1): 7 x 16 B per env, structure of arrays
            self.pstage = torch.zeros(_native.PSTAGE_BYTES // 16, B, 4, dtype=f32, device=dev) if N > 1 else None
            # minimum separation of each game at its spawn (N > 1): proves first-step collisions without reading intruders
            self.spawn_sep = torch.full((B,), float("inf"), dtype=f32, device=dev) if N > 1 else None
            self.episode_idx = torch.zeros(B, dtype=torch.int32, device=dev)
            self.min_sep = torch.zeros(B, dtype=f32, device=dev) if track_min_sep else None
            self.stats = torch.zeros(_native.STAT_SLOTS, _native.STAT_FIELDS, dtype=torch.int64, device=dev)
            # ---- per-step outputs (overwritten by every step): ONE packed allocation, so that the host-buffer
            #      step of a small batch moves everything with a single device-to-host copy
            self._layout, self._out_bytes = self._packed_layout(B, L)
            self._out = torch.zeros(self._out_bytes, dtype=torch.uint8, device=dev)
            for name, (off, nbytes, dtype, shape) in self._layout.items():
                setattr(self, name, self._out[off:off + nbytes].view(dtype).view(shape))
        self._state = State(
            num_envs=B, ppos=self.ppos.data_ptr(), paux=self.paux.data_ptr(), thot=self.thot.data_ptr(),
            tres=self.tres.data_ptr(), episode_idx=self.episode_idx.data_ptr(),
            min_sep=self.min_sep.data_ptr() if track_min_sep else None,
            stats=self.stats.data_ptr(), seed=self.seed, env_id_offset=self.env_id_offset,
            tkin=self.tkin.data_ptr() if self.tkin is not None else None,
            tpsi0=self.tpsi0.data_ptr() if self.tpsi0 is not None else None,
            pstage=self.pstage.data_ptr() if self.pstage is not None else None,
            spawn_sep=self.spawn_sep.data_ptr() if self.spawn_sep is not None else None)
        self._aux_full = StepAux(flags=self.flags.data_ptr(), outcome=self.outcome.data_ptr(),
                                 term_obs=self.term_obs.data_ptr(), ep_return=self.ep_return.data_ptr(),
                                 ep_length=self.ep_length.data_ptr())
        self._aux_lean = StepAux(flags=None, outcome=self.outcome.data_ptr(), term_obs=None,
                                 ep_return=self.ep_return.data_ptr(), ep_length=self.ep_length.data_ptr())
        self._host = None          # pinned staging buffers of step_host
        # tiny batches (the single-env gym surface): step_host lets the kernel read the action from and write its outputs
        # straight to the pinned (device-mapped) host block -- one launch + one sync, no copies.  The device-side
        # output buffers (obs, reward, ...) are then NOT updated by step_host.
        # (N == 1 and fewer envs than one 256-env tile only: those launches touch the host block with plain loads / stores;
        #  the bulk-copy engines are kept away from mapped host memory)
        self._zero_copy = bool(host_zero_copy) and N == 1 and B < 256
        self._aux_host = None
        self._trace = None         # on-device episode records (enable_trace)
        self._actions_dev = torch.zeros(B, dtype=f32, device=dev)
        self.launches = 0          # kernels launched through this object

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _packed_layout(B: int, L: int):
        """Sections of the packed per-step output buffer, each 256-byte aligned: the small per-env arrays first,
        then obs, then term_obs (so a prefix of the buffer holds everything but the terminal rows)."""
        sections = (("reward", torch.float32, (B,)), ("done_u8", torch.uint8, (B,)), ("flags", torch.uint8, (B,)),
                    ("outcome", torch.uint8, (B,)), ("ep_return", torch.float32, (B,)), ("ep_length", torch.int32, (B,)),
                    ("obs", torch.float32, (B, L)), ("term_obs", torch.float32, (B, L)))
        layout, off = {}, 0
        for name, dtype, shape in sections:
            n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
            layout[name] = (off, n, dtype, shape)
            off = (off + n + 255) & ~255
        return layout, max(off, 256)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _p(self):
        return ctypes.byref(self.params)

    def _s(self):
        return ctypes.byref(self._state)

    @property
    def done(self) -> torch.Tensor:
        return self.done_u8.view(torch.bool)

    def _as_actions(self, actions) -> torch.Tensor:
        a = actions
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a), device=self.device)
        if a.device != self.device or a.dtype != torch.float32:
            a = a.to(device=self.device, dtype=torch.float32)
        a = a.reshape(-1)
        if a.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {a.numel()}")
        return a.contiguous()

    # ------------------------------------------------------------------ reset / step
    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """New game for every env (or those with ``mask`` set); returns the obs buffer [B, L]."""
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_reset(self._p(), self._s(), mptr, self.obs.data_ptr(), self._stream()),
                          "acas2d_reset")
        self.launches += 1
        return self.obs

    # ------------------------------------------------------------------ on-device episode records
    def enable_trace(self, num_envs: Optional[int] = None, first_env: int = 0, capacity: Optional[int] = None,
                     n_traffic_rec: Optional[int] = None) -> None:
        """Record per-step episode data (the reference's per-step lists, game.py:45-75) for the window
        ``[first_env, first_env + num_envs)`` into ring buffers in HBM: ``capacity`` rows per env (default
        MAX_STEPS + 2, a whole episode).  ``step`` / ``step_host`` then launch the trace kernel before the step
        kernel (``acas2d_trace_step``); read with ``trace_rows()``.  ``policy_step`` / ``rollout_random`` /
        ``step_k`` choose or hold their actions inside the kernel and are not traced."""
        E = self.num_envs - first_env if num_envs is None else int(num_envs)
        cap = int(self.params.max_steps) + 2 if capacity is None else int(capacity)
        nrec = min(self.n_traffic, _native.TRACE_MAX_TRAFFIC) if n_traffic_rec is None else int(n_traffic_rec)
        width = _native.TRACE_DOUBLES + 2 * nrec
        self._trace_rows = torch.zeros(E, cap, width, dtype=torch.float64, device=self.device)
        self._trace_cursor = torch.zeros(E, dtype=torch.int32, device=self.device)
        self._trace = _native.Trace(first_env=int(first_env), num_envs=E, capacity=cap, n_traffic_rec=nrec,
                                    cursor=self._trace_cursor.data_ptr(), rows=self._trace_rows.data_ptr())

    def disable_trace(self) -> None:
        self._trace = None

    def _trace_step(self, a_ptr: int) -> None:
        _native.check(self.lib.acas2d_trace_step(self._p(), self._s(), a_ptr, ctypes.byref(self._trace), self._stream()),
                      "acas2d_trace_step")
        self.launches += 1

    def trace_rows(self):
        """(rows float64 [E, capacity, 18 + 2 n_rec], count int32 [E]) as numpy; row t of env w is
        ``rows[w, t % capacity]``, fields ``_native.TRACE_FIELDS`` then the recorded intruders' x, y."""
        return self._trace_rows.cpu().numpy(), self._trace_cursor.cpu().numpy()

    def step(self, actions, full_outputs: bool = True):
        """One environment step for all envs.  Returns views of the internal (obs, reward, done)
        buffers; they are overwritten by the next call.  With ``full_outputs`` the per-env
        ``flags`` and ``term_obs`` buffers are also written (``outcome`` / ``ep_return`` /
        ``ep_length`` always are, where done)."""
        a = self._as_actions(actions)
        aux = self._aux_full if full_outputs else self._aux_lean
        with torch.cuda.device(self.device):
            if self._trace is not None:
                self._trace_step(a.data_ptr())
            _native.check(self.lib.acas2d_step(self._p(), self._s(), a.data_ptr(), self.obs.data_ptr(),
                                               self.reward.data_ptr(), self.done_u8.data_ptr(),
                                               ctypes.byref(aux), self._stream()), "acas2d_step")
        self.launches += 1
        return self.obs, self.reward, self.done

    def step_k(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, reward: Optional[torch.Tensor] = None,
               done: Optional[torch.Tensor] = None):
        """K consecutive steps in one launch for actions known in advance (open loop): ``actions`` float32 [K, B]
        -> (obs [K, B, L], reward [K, B], done bool [K, B]), bit-identical to K ``step`` calls; the state is
        read and written once per launch instead of once per step (N_TRAFFIC == 1)."""
        a = actions.to(device=self.device, dtype=torch.float32).contiguous()
        assert a.dim() == 2 and a.shape[1] == self.num_envs
        K, B, L, dev = a.shape[0], self.num_envs, self.obs_dim, self.device
        obs = torch.empty(K, B, L, dtype=torch.float32, device=dev) if obs is None else obs
        reward = torch.empty(K, B, dtype=torch.float32, device=dev) if reward is None else reward
        done = torch.empty(K, B, dtype=torch.uint8, device=dev) if done is None else done
        if done.dtype == torch.bool:                       # the buffer this method returned earlier, passed back in
            done = done.view(torch.uint8)
        with torch.cuda.device(dev):
            _native.check(self.lib.acas2d_step_k(self._p(), self._s(), K, a.data_ptr(), obs.data_ptr(), reward.data_ptr(),
                                                 done.data_ptr(), ctypes.byref(self._aux_lean), self._stream()),
                          "acas2d_step_k")
        self.launches += 1
        if K:
            self.obs.copy_(obs[K - 1]); self.reward.copy_(reward[K - 1]); self.done_u8.copy_(done[K - 1])
        return obs, reward, done.view(torch.bool)

    PACKED_HOST_LIMIT = 262144       # envs: below this the host-buffer step is latency-bound -> one packed D2H copy

    def host_buffers(self) -> Dict[str, np.ndarray]:
        """Pinned host staging buffers of ``step_host`` as numpy views: ``actions`` float32[B] (write
        your actions here and call ``step_host()`` to skip the pageable->pinned copy), ``obs``,
        ``reward``, ``done`` and -- valid where ``done`` -- ``term_obs``, ``ep_return``, ``ep_length``,
        ``outcome``, plus ``flags``.  All but ``actions`` are sections of one pinned block that mirrors the
        packed device block."""
        if self._host is None:
            B = self.num_envs
            self._host_actions = torch.zeros(B, dtype=torch.float32).pin_memory()
            self._host_out = torch.zeros(self._out_bytes, dtype=torch.uint8).pin_memory()
            self._host = {name: self._host_out[off:off + n].view(dtype).view(shape)
                          for name, (off, n, dtype, shape) in self._layout.items()}
            self._host["actions"] = self._host_actions
            self._host["done"] = self._host["done_u8"]
            self._host_np = {k: v.numpy() for k, v in self._host.items()}
        return self._host_np

    def step_host(self, actions: Optional[np.ndarray] = None):
        """Host-buffer step (the end-to-end path): ``actions`` float32[B] in host memory (``None`` =
        already written into ``host_buffers()["actions"]``) -> (obs float32[B, L], reward float32[B],
        done bool[B]) numpy views of pinned buffers.  The H2D copy, the kernel, the D2H copies and one
        stream sync happen inside the call: small batches move the whole packed output block (terminal rows
        and finished-episode records included) with ONE copy; large batches are chunk-pipelined over streams
        and copy obs / reward / done only."""
        hb = self.host_buffers()
        h = self._host
        a_ptr = h["actions"].data_ptr()
        if isinstance(actions, torch.Tensor) and actions.device.type == "cpu" and actions.is_pinned() \
                and actions.dtype == torch.float32 and actions.is_contiguous() and actions.numel() == self.num_envs:
            a_ptr = actions.data_ptr()                      # caller-owned pinned memory: no staging copy
        elif actions is not None:
            np.copyto(hb["actions"], np.asarray(actions, dtype=np.float32).reshape(-1))
        if self._zero_copy and self._trace is None:
            if self._aux_host is None:
                self._aux_host = StepAux(flags=h["flags"].data_ptr(), outcome=h["outcome"].data_ptr(),
                                         term_obs=h["term_obs"].data_ptr(), ep_return=h["ep_return"].data_ptr(),
                                         ep_length=h["ep_length"].data_ptr())
                self._mapped_args = (self._p(), self._s(), None, h["obs"].data_ptr(), h["reward"].data_ptr(),
                                     h["done_u8"].data_ptr(), ctypes.byref(self._aux_host))
                self._mapped_out = (hb["obs"], hb["reward"], hb["done"].view(np.bool_))
                # the stream current at the first call serves every later one: looking it up costs more than the launch
                self._mapped_stream = self._stream()
            m = self._mapped_args
            if torch.cuda.current_device() != self.device.index:
                with torch.cuda.device(self.device):
                    rc = self.lib.acas2d_step_mapped(m[0], m[1], a_ptr, m[3], m[4], m[5], m[6], self._mapped_stream)
            else:
                rc = self.lib.acas2d_step_mapped(m[0], m[1], a_ptr, m[3], m[4], m[5], m[6], self._mapped_stream)
            if rc:
                _native.check(rc, "acas2d_step_mapped")
            self.launches += 1
            return self._mapped_out
        with torch.cuda.device(self.device):
            if self._trace is not None:                     # the trace kernel reads the actions on the device
                src = actions if a_ptr != h["actions"].data_ptr() else h["actions"]
                self._actions_dev.copy_(src.reshape(-1), non_blocking=True)
                self._trace_step(self._actions_dev.data_ptr())
            if self.num_envs <= self.PACKED_HOST_LIMIT:
                _native.check(self.lib.acas2d_step_host_packed(
                    self._p(), self._s(), a_ptr, self._actions_dev.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                    self.done_u8.data_ptr(), ctypes.byref(self._aux_full), self._out.data_ptr(), self._host_out.data_ptr(),
                    self._out_bytes, self._stream()), "acas2d_step_host_packed")
            else:
                _native.check(self.lib.acas2d_step_host(
                    self._p(), self._s(), a_ptr, h["obs"].data_ptr(), h["reward"].data_ptr(),
                    h["done"].data_ptr(), self._actions_dev.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                    self.done_u8.data_ptr(), ctypes.byref(self._aux_full), self._stream()), "acas2d_step_host")
        self.launches += 1
        return hb["obs"], hb["reward"], hb["done"].view(np.bool_)

    @property
    def host_records_valid(self) -> bool:
        """True when ``step_host`` also brings terminal rows / episode records to the host buffers."""
        return self.num_envs <= self.PACKED_HOST_LIMIT

    # ------------------------------------------------------------------ synthetic rollouts
    def random_actions(self, step_index: int, action_seed: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """U(-1,1) actions from the same Philox stream ``rollout_random`` draws in-kernel."""
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_random_actions(self._s(), action_seed, step_index, out.data_ptr(),
                                                         self._stream()), "acas2d_random_actions")
        self.launches += 1
        return out

    def rollout_random(self, num_steps: int, action_seed: int = 0, step0: int = 0,
                       reward_sum: Optional[torch.Tensor] = None) -> None:
        """``num_steps`` fused auto-resetting steps with in-kernel random actions (N_TRAFFIC == 1)."""
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_rollout_random(
                self._p(), self._s(), int(num_steps), action_seed, step0,
                reward_sum.data_ptr() if reward_sum is not None else None, self._stream()), "acas2d_rollout_random")
        self.launches += 1

    def policy_step(self, actor, deterministic: bool = True, noise_seed: int = 0, step_index: int = 0,
                    actions_out: Optional[torch.Tensor] = None, logp_out: Optional[torch.Tensor] = None,
                    obs_in: Optional[torch.Tensor] = None, full_outputs: bool = True, tensor_cores: bool = False,
                    obs_out: Optional[torch.Tensor] = None, reward_out: Optional[torch.Tensor] = None,
                    done_out: Optional[torch.Tensor] = None, log_std_ptr: Optional[int] = None,
                    step_base_ptr: Optional[int] = None):
        """Closed-loop step: ``actor`` (``gym_ACAS2D.policy.MlpActor`` on this device) is evaluated on the
        current observation rows (default: this object's ``obs`` buffer, i.e. the previous step's output),
        its action -- ``model.predict(obs, deterministic=...)`` semantics, clipped to the Box -- is applied,
        all in one kernel.  Returns the (obs, reward, done) buffers; the unclipped action sample and its
        log-probability go to ``actions_out`` / ``logp_out`` when given.  ``tensor_cores`` runs the two
        hidden layers as tcgen05 TF32 MMAs (TMEM accumulators) instead of float32 on the CUDA cores.
        ``obs_out`` / ``reward_out`` / ``done_out`` (uint8) redirect the step's outputs, e.g. into row t+1 /
        row t of rollout buffers, so collecting a rollout copies nothing.  ``log_std_ptr`` / ``step_base_ptr``
        (device pointers to a float32 / an int64) make the launch read log_std and the base of its noise
        counter at run time -- what a captured rollout graph needs to stay valid while the policy is trained."""
        if actor.packed.device != self.device:
            actor.to(self.device)
        src = self.obs if obs_in is None else obs_in
        o_dst = self.obs if obs_out is None else obs_out
        r_dst = self.reward if reward_out is None else reward_out
        d_dst = self.done_u8 if done_out is None else done_out
        aux = self._aux_full if full_outputs else self._aux_lean
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_policy_step_dyn(
                self._p(), self._s(), actor.packed.data_ptr(), 0.0 if log_std_ptr else float(actor.log_std), src.data_ptr(),
                actions_out.data_ptr() if actions_out is not None else None,
                logp_out.data_ptr() if logp_out is not None else None,
                o_dst.data_ptr(), r_dst.data_ptr(), d_dst.data_ptr(), ctypes.byref(aux),
                0 if deterministic else 1, int(noise_seed), int(step_index), 1 if tensor_cores else 0,
                log_std_ptr, step_base_ptr, self._stream()), "acas2d_policy_step")
        self.launches += 1
        return o_dst, r_dst, d_dst.view(torch.bool)

    def collect_rollout(self, actor, n_steps: int, noise_seed: int = 0, step0: int = 0, tensor_cores: bool = True,
                        buffers: Optional[Dict[str, torch.Tensor]] = None, graph: bool = False) -> Dict[str, torch.Tensor]:
        """PPO-style rollout collection (what SB3's ``collect_rollouts`` does for the reference's
        ``PPO(...).learn``, training_main.py:44-52): ``n_steps`` stochastic closed-loop steps from the
        current observations, written straight into [T, B] buffers -- ``obs`` [T+1, B, L] (row t is what
        the actor saw at step t; row T bootstraps the value), ``actions``, ``logp``, ``rewards`` [T, B]
        float32, ``dones`` [T, B] uint8.  One fused kernel per step, no copies, no host round trip.
        ``graph``: the T launches are captured once into a CUDA graph (per actor block / buffers) and
        replayed -- the rollout of a small batch is launch-bound otherwise; the launches read the actor's
        weights, its log_std (``actor.log_std_ptr``) and the noise-counter base from device memory, so
        the graph stays valid while the learner updates the policy in place."""
        T, B, L, dev = int(n_steps), self.num_envs, self.obs_dim, self.device
        if buffers is None:
            buffers = dict(obs=torch.empty(T + 1, B, L, dtype=torch.float32, device=dev),
                           actions=torch.empty(T, B, dtype=torch.float32, device=dev),
                           logp=torch.empty(T, B, dtype=torch.float32, device=dev),
                           rewards=torch.empty(T, B, dtype=torch.float32, device=dev),
                           dones=torch.empty(T, B, dtype=torch.uint8, device=dev))

        def launch_all(step_index0: int, log_std_ptr, step_base_ptr):
            buffers["obs"][0].copy_(self.obs)
            for t in range(T):
                self.policy_step(actor, deterministic=False, noise_seed=noise_seed, step_index=step_index0 + t,
                                 actions_out=buffers["actions"][t], logp_out=buffers["logp"][t],
                                 obs_in=buffers["obs"][t], obs_out=buffers["obs"][t + 1],
                                 reward_out=buffers["rewards"][t], done_out=buffers["dones"][t],
                                 full_outputs=False, tensor_cores=tensor_cores,
                                 log_std_ptr=log_std_ptr, step_base_ptr=step_base_ptr)
            self.obs.copy_(buffers["obs"][T])

        if not graph:
            launch_all(step0, getattr(actor, "log_std_ptr", None), None)
            return buffers
        key = (T, actor.packed.data_ptr(), int(noise_seed), bool(tensor_cores), buffers["obs"].data_ptr(),
               buffers["actions"].data_ptr(), getattr(actor, "log_std_ptr", None))
        if getattr(self, "_rollout_key", None) != key:
            if actor.packed.device != dev:
                actor.to(dev)
            self._rollout_step_base = torch.zeros(1, dtype=torch.int64, device=dev)
            with torch.cuda.device(dev):
                _native.check(self.lib.acas2d_ppo_prepare(), "acas2d_ppo_prepare")    # attributes / module load before capture
            torch.cuda.synchronize(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            g = torch.cuda.CUDAGraph()
            launches_before = self.launches
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    launch_all(0, getattr(actor, "log_std_ptr", None), self._rollout_step_base.data_ptr())
            torch.cuda.current_stream(dev).wait_stream(side)
            self.launches = launches_before
            self._rollout_graph, self._rollout_key = g, key
            self._rollout_buffers = buffers                                           # keep the captured storages alive
        self._rollout_step_base.fill_(int(step0))
        self._rollout_graph.replay()
        self.launches += T
        return buffers

    def capture_steps(self, actions: torch.Tensor, full_outputs: bool = False, num_steps: Optional[int] = None,
                      warmup: bool = True) -> "torch.cuda.CUDAGraph":
        """Capture consecutive steps into one CUDA graph (replay with ``graph.replay()``): step k uses
        ``actions[k % len(actions)]`` (actions [K, B], resident in HBM), ``num_steps`` steps in total
        (default K).  Launch-bound batches and host-overhead-free rollouts need this.  With ``warmup``
        one eager step with ``actions[0]`` runs first (kernels must be loaded before capture)."""
        actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        assert actions.dim() == 2 and actions.shape[1] == self.num_envs
        self._graph_actions = actions
        n = int(num_steps or actions.shape[0])
        stream = torch.cuda.Stream(self.device)
        stream.wait_stream(torch.cuda.current_stream(self.device))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            if warmup:
                self.step(actions[0], full_outputs)
            torch.cuda.current_stream(self.device).synchronize()
            with torch.cuda.graph(graph, stream=stream):
                for k in range(n):
                    self.step(actions[k % actions.shape[0]], full_outputs)
        torch.cuda.current_stream(self.device).wait_stream(stream)
        return graph

    # ------------------------------------------------------------------ state injection / extraction
    def inject_state(self, player, traffic, steps=None, total_reward=None) -> None:
        """Overwrite the games: ``player`` [B,3] = x, y, psi; ``traffic`` [B,N,4] = x, y, v_air, psi
        (current position); ``steps`` [B] = game.steps (default 1, i.e. just reset);
        ``total_reward`` [B].  Mirrors poking ``game.player`` / ``game.traffic`` in the reference."""
        B, N, dev = self.num_envs, self.n_traffic, self.device
        pl = torch.as_tensor(np.asarray(player, dtype=np.float64), device=dev).reshape(B, 3).contiguous()
        tr = torch.as_tensor(np.asarray(traffic, dtype=np.float64), device=dev).reshape(B, N, 4).contiguous()
        st = torch.ones(B, dtype=torch.int32, device=dev) if steps is None else \
            torch.as_tensor(np.asarray(steps, dtype=np.int32), device=dev).reshape(B).contiguous()
        tot = torch.zeros(B, dtype=torch.float64, device=dev) if total_reward is None else \
            torch.as_tensor(np.asarray(total_reward, dtype=np.float64), device=dev).reshape(B).contiguous()
        with torch.cuda.device(dev):
            _native.check(self.lib.acas2d_inject_state(self._p(), self._s(), pl.data_ptr(), tr.data_ptr(),
                                                       st.data_ptr(), tot.data_ptr(), self._stream()),
                          "acas2d_inject_state")
            torch.cuda.current_stream(dev).synchronize()            # staging tensors die with this frame
        self.launches += 1

    def observe(self) -> torch.Tensor:
        """Observation rows of the games as they stand (no step, no ``steps`` increment; last lateral
        acceleration taken as 0): the reset observation of a freshly injected state.  Written into and
        returned as the ``obs`` buffer."""
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_observe(self._p(), self._s(), self.obs.data_ptr(), self._stream()),
                          "acas2d_observe")
        self.launches += 1
        return self.obs

    def render(self, env_index: int = 0) -> torch.Tensor:
        """Debug frame of one env: uint8 [HEIGHT, WIDTH, 3] device tensor (the scene of the reference's
        ``game.view()`` without sprites / HUD text).  Off the hot path."""
        H, W = int(self.params.height), int(self.params.width)
        img = torch.empty(H, W, 3, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(self.lib.acas2d_render(self._p(), self._s(), int(env_index), img.data_ptr(), self._stream()),
                          "acas2d_render")
        self.launches += 1
        return img

    def extract_state(self) -> Dict[str, np.ndarray]:
        """Current games as numpy float64: player [B,3], traffic [B,N,4], steps, total_reward, episode_idx."""
        B, N, dev = self.num_envs, self.n_traffic, self.device
        pl = torch.empty(B, 3, dtype=torch.float64, device=dev)
        tr = torch.empty(B, N, 4, dtype=torch.float64, device=dev)
        st = torch.empty(B, dtype=torch.int32, device=dev)
        tot = torch.empty(B, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _native.check(self.lib.acas2d_extract_state(self._p(), self._s(), pl.data_ptr(), tr.data_ptr(),
                                                        st.data_ptr(), tot.data_ptr(), self._stream()),
                          "acas2d_extract_state")
        self.launches += 1
        out = dict(player=pl.cpu().numpy(), traffic=tr.cpu().numpy(), steps=st.cpu().numpy(),
                   total_reward=tot.cpu().numpy(), episode_idx=self.episode_idx.cpu().numpy().astype(np.uint32))
        if self.min_sep is not None:
            out["min_sep"] = self.min_sep.cpu().numpy()
        return out

    # the games AND the per-step buffers: policy_step / collect_rollout read ``obs`` as the actor's input, so a
    # resumed closed-loop rollout needs it (it cannot be rebuilt exactly: the row depends on the last action)
    _STATE_TENSORS = ("ppos", "paux", "thot", "tres", "episode_idx", "stats", "obs", "reward", "done_u8")

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Checkpoint of the env batch (the reference never saves env state)."""
        d = {k: getattr(self, k).detach().clone() for k in self._STATE_TENSORS}
        for k in ("min_sep", "tkin", "tpsi0", "spawn_sep"):
            if getattr(self, k) is not None:
                d[k] = getattr(self, k).clone()
        d["meta"] = torch.tensor([self.num_envs, self.n_traffic, self.seed, self.env_id_offset], dtype=torch.int64)
        return d

    def load_state_dict(self, d: Dict[str, torch.Tensor]) -> None:
        meta = d["meta"].tolist()
        if meta[0] != self.num_envs or meta[1] != self.n_traffic:
            raise ValueError("state_dict was taken from a batch of a different shape")
        if meta[2] != self.seed or meta[3] != self.env_id_offset:
            # respawns are Philox(seed, global env id, episode): adopt the checkpoint's stream, or they diverge
            self.seed, self.env_id_offset = int(meta[2]), int(meta[3])
            self._state.seed, self._state.env_id_offset = self.seed, self.env_id_offset
        for k in self._STATE_TENSORS:
            getattr(self, k).copy_(d[k])
        for k in ("min_sep", "tkin", "tpsi0", "spawn_sep"):
            if getattr(self, k) is not None:
                if k not in d:
                    if k == "min_sep":
                        continue
                    raise ValueError(f"state_dict lacks '{k}' (taken by an older version): re-inject the state instead")
                getattr(self, k).copy_(d[k])

    # ------------------------------------------------------------------ episode statistics
    def episode_counters(self) -> torch.Tensor:
        """int64[7] device tensor: episodes, goal, collision, timeout, sum(length), sum(return)*2^20,
        sum(min_sep)*2^20 -- finished episodes of THIS rank since construction / ``clear_stats``."""
        return self.stats.sum(0)[: len(_native.STAT_NAMES)]

    def clear_stats(self) -> None:
        self.stats.zero_()

    def episode_stats(self, reduce: bool = True) -> Dict[str, float]:
        """Finished-episode statistics; summed over all ranks with one NCCL all-reduce of seven
        int64 counters when ``torch.distributed`` is initialised and ``reduce`` is set."""
        from .stats import summarise
        return summarise(self.episode_counters(), reduce=reduce, track_min_sep=self.min_sep is not None)
