"""The reference's ``training_main.py`` (PPO with the SB3 default hyper-parameters) with on-device rollouts.
Saves the trained network under SB3's parameter names so ``evaluate_agent.py --model`` can load it."""
import argparse

import numpy as np

import _path  # noqa: F401
from gym_ACAS2D import ppo
from gym_ACAS2D.policy import MlpActor

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1024)
ap.add_argument("--n-steps", type=int, default=1024)
ap.add_argument("--iterations", type=int, default=70)
ap.add_argument("--minibatches", type=int, default=256)
ap.add_argument("--out", default="ppo_acas2d.npz")
args = ap.parse_args()

ppo.train(args.envs, args.n_steps, args.iterations, minibatches=args.minibatches)
state = {k: v.cpu().numpy() for k, v in ppo.train.last_policy.sb3_state_dict().items()}
np.savez(args.out, **state)
print("saved", args.out, ppo.evaluate(MlpActor(ppo.train.last_policy.sb3_state_dict(), "cuda"), 2048))
