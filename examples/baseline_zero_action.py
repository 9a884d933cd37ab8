"""The reference's ``baseline_main.py`` (100 episodes with action = [0], CSV with Episode, Outcome, Total
Reward, Time Steps, Path, Traffic Paths) on the CUDA environment: all episodes run side by side."""
import argparse

import _path  # noqa: F401
from gym_ACAS2D import records
from gym_ACAS2D.envs import BatchedACAS2D
from gym_ACAS2D.settings import MODEL_VERSION, RANDOM_SEED, TEST_EPISODES

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=TEST_EPISODES)
ap.add_argument("--out", default=f"baseline_ACAS2D_PPO_{MODEL_VERSION}_{TEST_EPISODES}.csv")
args = ap.parse_args()

env = BatchedACAS2D(args.episodes, seed=RANDOM_SEED, auto_reset=False)
rows = records.record_episodes(env)                       # policy=None -> the zero action
records.to_csv(rows, args.out, records.BASELINE_COLUMNS)
goal = sum(r["Outcome"] == "Goal" for r in rows)
print(f"{len(rows)} episodes -> {args.out}: {goal} Goal, {len(rows) - goal} other; "
      f"mean steps {sum(r['Time Steps'] for r in rows) / len(rows):.1f}, "
      f"mean reward {sum(r['Total Reward'] for r in rows) / len(rows):.1f}")
