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
                 env_id_offset: int = 0, track_min_sep: bool = False, settings=None, copy: bool = True):
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
        self._copy = bool(copy)
        self._empty: List[dict] = [{} for _ in range(self.num_envs)]     # one (reused) empty info dict per env
        self._infos: List[dict] = list(self._empty)
        self._dirty: List[int] = []                                      # envs whose entry is not their empty dict
        self.metadata = {"render.modes": []}

    # ------------------------------------------------------------------ SB3 numpy surface
    def reset(self) -> np.ndarray:
        return self.core.reset().cpu().numpy()

    def step_async(self, actions: np.ndarray) -> None:
        a = np.asarray(actions, dtype=np.float32)
        self._pending = a.reshape(self.num_envs, -1)[:, 0] if a.ndim > 1 else a

    def step_wait(self):
        """SB3 1.1.0 ``DummyVecEnv.step_wait`` semantics: fresh obs / rewards / dones arrays (``copy=False`` at
        construction returns the pinned staging views instead) and one info dict per env -- empty unless the
        env finished, then ``terminal_observation``, ``episode`` = {r, l} (Monitor) and ``outcome``."""
        if self._pending is None:
            raise RuntimeError("step_wait() without step_async()")
        core = self.core
        obs, reward, done = core.step_host(self._pending)
        self._pending = None
        infos = self._infos
        for i in self._dirty:                          # last step's finished envs: back to their empty dict
            infos[i] = self._empty[i]
        idx = np.flatnonzero(done)
        self._dirty = idx.tolist()
        if self._dirty:
            if core.host_records_valid:                # the packed host block already holds rows and records
                hb = core.host_buffers()
                term, ep_r, ep_l, oc = hb["term_obs"][idx], hb["ep_return"][idx], hb["ep_length"][idx], hb["outcome"][idx]
            else:                                      # large batch: gather the finished envs' rows on the device
                sel = torch.from_numpy(idx).to(core.device, non_blocking=True)
                term = core.term_obs.index_select(0, sel).cpu().numpy()
                ep_r = core.ep_return.index_select(0, sel).cpu().numpy()
                ep_l = core.ep_length.index_select(0, sel).cpu().numpy()
                oc = core.outcome.index_select(0, sel).cpu().numpy()
            ep_r, ep_l, oc = ep_r.tolist(), ep_l.tolist(), oc.tolist()
            for k, i in enumerate(self._dirty):
                infos[i] = {"terminal_observation": term[k],
                            "episode": {"r": ep_r[k], "l": ep_l[k] - 1},     # l = step() calls
                            "outcome": oc[k]}
        if self._copy:
            return obs.copy(), reward.copy(), done.copy(), list(infos)
        return obs, reward, done, infos

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

    def collect_rollout(self, actor, n_steps: int, **kw):
        """On-device counterpart of SB3's ``OnPolicyAlgorithm.collect_rollouts`` over this VecEnv (BASELINE
        config 5): ``n_steps`` closed-loop steps of the fused actor + env-step kernel written straight into
        [T, B] rollout buffers (``BatchedACAS2D.collect_rollout``); nothing crosses the host."""
        return self.core.collect_rollout(actor, n_steps, **kw)

    def episode_stats(self, reduce: bool = True):
        return self.core.episode_stats(reduce=reduce)

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
