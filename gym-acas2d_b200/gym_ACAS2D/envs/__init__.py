"""Reference ``gym_ACAS2D/envs/__init__.py:1-2`` re-exports ``ACAS2DEnv`` and ``ACAS2DGame``;
the batched core and the VecEnv adapter are the additions."""
from gym_ACAS2D.envs.environment import ACAS2DEnv
from gym_ACAS2D.envs.game import ACAS2DGame
from gym_ACAS2D.envs.batched import BatchedACAS2D
from gym_ACAS2D.envs.vec_env import ACAS2DVecEnv

__all__ = ["ACAS2DEnv", "ACAS2DGame", "BatchedACAS2D", "ACAS2DVecEnv"]
