"""The reference agent's actor on the GPU, fused with the environment step.

The reference trains / evaluates a stable-baselines3 1.1.0 ``PPO('MlpPolicy')`` on ``ACAS2D-v0``
(``gym_ACAS2D/training_main.py:44-52``, ``testing_main.py:55-78``).  SB3 is not needed here: the saved
``best_model.zip`` is a zip whose ``policy.pth`` member is a plain ``torch`` state dict
(``mlp_extractor.policy_net.{0,2}``, ``action_net``, ``log_std``, and the separate value network).
``MlpActor`` packs the actor into the 19 KB weight block ``acas2d_policy_step`` reads, and
``BatchedACAS2D.policy_step`` runs actor + exploration noise + clip + env step as ONE kernel per step, so
closed-loop rollouts (evaluation, or PPO's ``collect_rollouts``) never leave the device.
"""
from __future__ import annotations

import io
import zipfile
from typing import Dict, Optional

import numpy as np
import torch

from gym_ACAS2D.envs._native import POLICY_FLOATS

OBS_DIM, HIDDEN = 8, 64
_KEYS = ("mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias",
         "mlp_extractor.policy_net.2.weight", "mlp_extractor.policy_net.2.bias",
         "action_net.weight", "action_net.bias")


def load_sb3_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """State dict of an SB3 model zip (``best_model.zip``), of a bare ``policy.pth``, or of an ``.npz``
    holding the same keys."""
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k]) for k in z.files}
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            if "policy.pth" in z.namelist():
                return torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    return torch.load(path, map_location="cpu", weights_only=True)


class MlpActor:
    """8 -> 64 -> 64 -> 1 tanh actor with a state-independent ``log_std`` (SB3 ``MlpPolicy`` defaults)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cpu"):
        sd = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in state_dict.items()}
        w1, b1, w2, b2, w3, b3 = (sd[k] for k in _KEYS)
        if tuple(w1.shape) != (HIDDEN, OBS_DIM) or tuple(w2.shape) != (HIDDEN, HIDDEN) or tuple(w3.shape) != (1, HIDDEN):
            raise ValueError("only the reference architecture (8 -> 64 -> 64 -> 1) is supported")
        self.tensors = dict(w1=w1, b1=b1, w2=w2, b2=b2, w3=w3, b3=b3)
        self.log_std = float(sd["log_std"].reshape(-1)[0]) if "log_std" in sd else 0.0
        packed = torch.cat([w1.reshape(-1), b1, w2.reshape(-1), b2, w3.reshape(-1), b3.reshape(-1),
                            torch.zeros(3, device=w1.device)])
        assert packed.numel() == POLICY_FLOATS
        self.device = torch.device(device)
        self.packed = packed.contiguous().to(self.device)

    @classmethod
    def from_file(cls, path: str, device="cpu") -> "MlpActor":
        return cls(load_sb3_state_dict(path), device)

    @classmethod
    def random(cls, seed: int = 0, device="cpu") -> "MlpActor":
        """Random weights of the right shapes (synthetic benchmarks: no checkpoint travels to the box)."""
        g = torch.Generator().manual_seed(seed)
        sd = {_KEYS[0]: torch.randn(HIDDEN, OBS_DIM, generator=g) * 0.5, _KEYS[1]: torch.zeros(HIDDEN),
              _KEYS[2]: torch.randn(HIDDEN, HIDDEN, generator=g) * 0.2, _KEYS[3]: torch.zeros(HIDDEN),
              _KEYS[4]: torch.randn(1, HIDDEN, generator=g) * 0.1, _KEYS[5]: torch.zeros(1),
              "log_std": torch.zeros(1)}
        return cls(sd, device)

    def to(self, device) -> "MlpActor":
        self.device = torch.device(device)
        self.packed = self.packed.to(self.device)
        return self

    def reference_mean(self, obs: torch.Tensor) -> torch.Tensor:
        """Plain torch float32 forward pass (the numerics reference for the fused kernel)."""
        t = {k: v.to(obs.device) for k, v in self.tensors.items()}
        h = torch.tanh(obs.float() @ t["w1"].T + t["b1"])
        h = torch.tanh(h @ t["w2"].T + t["b2"])
        return (h @ t["w3"].T + t["b3"]).reshape(-1)
