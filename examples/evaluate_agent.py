"""The reference's ``testing_main.py``: load a stable-baselines3 PPO model (``best_model.zip`` or a bare
``policy.pth`` / ``.npz``) and run deterministic episodes -- here thousands at once, actor fused with the
environment step on the GPU.  Default model: the reference's own saved agent (test fixture)."""
import argparse
import os

import _path
from gym_ACAS2D import ppo
from gym_ACAS2D.policy import MlpActor

default = os.path.join(_path.ROOT, "tests", "golden", "ppo_policy_1048576_11.npz")
ap = argparse.ArgumentParser()
ap.add_argument("--model", default=default)
ap.add_argument("--episodes", type=int, default=4096)
ap.add_argument("--tensor-cores", action="store_true", help="tcgen05 TF32 actor instead of float32")
args = ap.parse_args()

stats = ppo.evaluate(MlpActor.from_file(args.model, "cuda"), args.episodes, tensor_cores=args.tensor_cores)
print({k: round(v, 3) if isinstance(v, float) else v for k, v in stats.items()})
