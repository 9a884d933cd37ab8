"""``ACAS2DVecEnv`` -- stable-baselines3 ``VecEnv`` adapter over the batched CUDA core.

Duck-types SB3 1.1.0's ``VecEnv`` (the version the reference's trained model was saved with;
SB3 itself is not vendored by the reference nor installed here) with ``DummyVecEnv``
semantics: when an env finishes inside ``step``, ``infos[i]["terminal_observation"]`` holds
its last observation, ``infos[i]["episode"] = {"r", "l"}`` its Monitor-style return / length,
and the returned observation row is the first observation of the next episode.

Two surfaces:
  * numpy (SB3): ``reset() / step_async / step_wait / step`` -- host buffers, copies inside;
  * torch, zero-copy: ``step_tensor(actions[B]) -> (obs, reward, done)`` device tensors.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from .batched import BatchedACAS2D
from .spaces import action_box, observation_box

try:                                                   # pragma: no cover - SB3 is not in the image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:                                      # noqa: BLE001
    _VecEnvBase = object


class ACAS2DVecEnv(_VecEnvBase):
    def __init__(self, num_envs: int, n_traffic: Optional[int] = None, device="cuda", seed: Optional[int] = None,
                 env_id_offset: int = 0, track_min_sep: bool = False, settings=None):
        self.core = BatchedACAS2D(num_envs, n_traffic=n_traffic, device=device, seed=seed,
                                  env_id_offset=env_id_offset, auto_reset=True, track_min_sep=track_min_sep,
                                  settings=settings)
        obs_space = observation_box(self.core.n_traffic, np.float32)
        act_space = action_box(np.float32)
        if _VecEnvBase is not object:                  # pragma: no cover
            super().__init__(num_envs, obs_space, act_space)
        else:
            self.num_envs, self.observation_space, self.action_space = int(num_envs), obs_space, act_space
        self._pending: Optional[np.ndarray] = None
        self._fin = None                               # pinned staging of finished-episode data
        self.metadata = {"render.modes": []}

    # ------------------------------------------------------------------ SB3 numpy surface
    def reset(self) -> np.ndarray:
        return self.core.reset().cpu().numpy()

    def step_async(self, actions: np.ndarray) -> None:
        self._pending = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, -1)[:, 0]

    def step_wait(self):
        if self._pending is None:
            raise RuntimeError("step_wait() without step_async()")
        obs, reward, done = self.core.step_host(self._pending)
        self._pending = None
        infos: List[dict] = [{} for _ in range(self.num_envs)]
        idx = np.flatnonzero(done)
        if idx.size:
            # finished episodes: the small per-env arrays whole, the terminal rows gathered, all four copies
            # queued into pinned buffers behind one synchronisation
            core, n = self.core, int(idx.size)
            if self._fin is None:
                B, L = self.num_envs, core.obs_dim
                self._fin = dict(ret=torch.empty(B, dtype=torch.float32).pin_memory(),
                                 length=torch.empty(B, dtype=torch.int32).pin_memory(),
                                 outcome=torch.empty(B, dtype=torch.uint8).pin_memory(),
                                 term=torch.empty(B, L, dtype=torch.float32).pin_memory(),
                                 sel=torch.empty(B, dtype=torch.int64).pin_memory())
            f = self._fin
            f["sel"][:n] = torch.from_numpy(idx)
            sel = f["sel"][:n].to(core.device, non_blocking=True)
            f["term"][:n].copy_(core.term_obs.index_select(0, sel), non_blocking=True)
            f["ret"].copy_(core.ep_return, non_blocking=True)
            f["length"].copy_(core.ep_length, non_blocking=True)
            f["outcome"].copy_(core.outcome, non_blocking=True)
            torch.cuda.current_stream(core.device).synchronize()
            term = f["term"][:n].numpy().copy()
            ep_r, ep_l, oc = f["ret"].numpy(), f["length"].numpy(), f["outcome"].numpy()
            for k, i in enumerate(idx):
                infos[i] = {"terminal_observation": term[k],
                            "episode": {"r": float(ep_r[i]), "l": int(ep_l[i]) - 1},   # l = step() calls
                            "outcome": int(oc[i])}
        return obs.copy(), reward.copy(), done.copy(), infos

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    # ------------------------------------------------------------------ zero-copy torch surface
    def reset_tensor(self) -> torch.Tensor:
        return self.core.reset()

    def step_tensor(self, actions: torch.Tensor):
        """Device tensors in, device tensors out (views of the core's buffers; overwritten by the
        next step).  ``core.term_obs / ep_return / ep_length / outcome`` hold the finished-episode
        data for the rows where ``done`` is set."""
        return self.core.step(actions)

    # ------------------------------------------------------------------ VecEnv plumbing
    def close(self) -> None:
        return None

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        if seed is not None:
            self.core.seed = int(seed)
            self.core._state.seed = int(seed)
        return [seed] * self.num_envs

    def _indices(self, indices) -> Sequence[int]:
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        return [getattr(self, attr_name) if hasattr(self, attr_name) else getattr(self.core, attr_name)
                for _ in self._indices(indices)]

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        return [getattr(self.core, method_name)(*args, **kwargs) for _ in self._indices(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode: str = "human"):
        return None
