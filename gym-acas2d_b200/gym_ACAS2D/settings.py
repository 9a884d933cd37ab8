"""Constants of the ACAS-2D environment (drop-in for the reference ``gym_ACAS2D/settings.py``).

Every public name and value of the reference module is kept, so ``from gym_ACAS2D.settings
import *`` behaves the same.  The batched CUDA step does not read this module at run time:
``envs/_native.params_from_settings`` turns it into the POD ``acas2d_params`` struct once, at
environment construction (so traffic sweeps only need ``n_traffic=...``, not a re-import).
"""

# --- episode / training bookkeeping (reference settings.py:3-12)
MODEL_VERSION = 11
OUTCOME_NAMES = {1: "Goal", 2: "Collision", 3: "Timeout"}
TEST_EPISODES = 100
EVAL_EPISODES = 10
MAX_STEPS = 1000                      # an episode is at most 1000 step() calls (game.py:182-183)
N_STEPS = 2048                        # PPO rollout length used by training_main.py
TOTAL_STEPS = N_STEPS * 512           # 1,048,576
EVAL_STEPS = TOTAL_STEPS / 32

# --- airspace (reference settings.py:15-25); the window constants are kept for scripts that
#     import them although nothing in the batched path draws
WIDTH, HEIGHT = 1600, 1000            # px
FPS = 100                             # dt = 1/FPS
CAPTION = "ACAS-2D"
FONT_NAME, FONT_SIZE = "freesansbold.ttf", 14
BLACK_RGB, SKY_RGB = (0, 0, 0), (60, 150, 220)
GREEN_RGB, RED_RGB, YELLOW_RBG = (0, 255, 0), (255, 0, 0), (255, 255, 0)

RANDOM_SEED = 13                      # reference settings.py:28

# --- aircraft geometry (reference settings.py:31-36)
MIN_TRAFFIC = MAX_TRAFFIC = 1         # intruders per episode; the reference needs MIN == MAX >= 1
AIRCRAFT_SIZE = 24
COLLISION_RADIUS = 2 * AIRCRAFT_SIZE  # 48; collision iff centre distance < 2*COLLISION_RADIUS
GOAL_RADIUS = 6 * AIRCRAFT_SIZE       # 144
SAFE_DISTANCE = 4 * COLLISION_RADIUS  # 192

# --- kinematics (reference settings.py:39-44)
_STANDARD_GRAVITY = 9.80665           # scipy.constants.g
AIRSPEED = 200
AIRSPEED_FACTOR_MIN = AIRSPEED_FACTOR_MAX = 1
ACC_LAT_LIMIT = 20 * _STANDARD_GRAVITY
PLAYER_INITIAL_HEADING_LIM = 3        # deg
TRAFFIC_INITIAL_HEADING_LIM = 15      # deg

# --- terminal rewards (reference settings.py:47-48)
REWARD_GOAL, REWARD_COLLISION = 1000, -1000

# --- sprite files of the reference renderer (settings.py:51-54); unused by the batched path
LOGO = "png/004-compass.png"
PLAYER_IMG = "png/001-plane.png"
TRAFFIC_IMG = "png/002-travelling.png"
GOAL_IMG = "png/003-army.png"
