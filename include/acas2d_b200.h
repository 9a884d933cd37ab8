/*
 * acas2d_b200.h -- C ABI of the B200-native batched ACAS-2D environment step.
 *
 * This library replaces ONE path of Christos-14/gym-ACAS2D: `ACAS2DEnv.reset/step`
 * (reference gym_ACAS2D/envs/environment.py:29-48 and everything it calls in
 * envs/game.py, envs/aircraft.py, envs/kinematics.py, envs/rewards.py).  The
 * reference has no FFI layer -- its boundary is the Python gym.Env API -- so the
 * entry points below are what a ctypes binding added to the reference's
 * `environment.py` would call (see INTEGRATION.md for that stub).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - every pointer in `acas2d_state` and every array argument is a DEVICE pointer
 *     owned by the caller (PyTorch allocates them), except in `acas2d_step_host`.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     allocate, never synchronise (again except `acas2d_step_host`, which syncs the
 *     stream once so the host buffers are valid on return).
 *   - return value: 0 = ok, >0 = cudaError_t from the launch, <0 = ACAS2D_E_* below.
 *   - one CUDA context per process/GPU; calls on different state objects are
 *     independent and re-entrant.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Data layout in HBM (B = num_envs, N = n_traffic, L = 5 + 3N):
 *   ppos        double[B][2]    player x, y                    (aircraft.py:9-10)
 *   paux        16 B / env      { double psi; int32 steps; float ep_return }
 *                               psi in degrees (aircraft.py:12); steps == game.steps
 *                               (game.py:30,197) | ACAS2D_STEPS_RESIDUAL_BIT; ep_return ==
 *                               game.total_reward so far (game.py:32,287)
 *   thot        float[B][N][4]  intruder record {x0, y0, psi, v}: position at steps == 1, heading [deg],
 *                               speed.  The intruder flies a straight line (game.py:243-245, a_lat is
 *                               always 0): position after k moves = (x0, y0) + k * v*dt*(cos psi, sin psi),
 *                               evaluated in float64.  Spawned intruders are float32-representable by
 *                               construction, so this 16-byte record is exact and is all a step reads.
 *   tres        double[B][N][4] full - float32(full) of the same four values.  Cold: written by
 *                               acas2d_inject_state, read by step / extract only for envs whose
 *                               paux.steps carries ACAS2D_STEPS_RESIDUAL_BIT (injected float64 states).
 *   tkin        24 B / intruder { float x0, y0; double dx, dy }  OPTIONAL kinematic cache (N > 1): the intruder's
 *                               displacement per step v*dt*(cos psi, sin psi) in float64, written once at spawn /
 *                               injection next to thot.  With it the step reads 24 B per intruder and evaluates no
 *                               sin / cos per intruder per step (a heading never changes, game.py:243-245); without
 *                               it (NULL) the step derives the same values from thot every step.  Same results.
 *   tpsi0       float[B]        OPTIONAL compact form of intruder 0 for N == 1: the reference spawns intruder 0 at
 *                               x = WIDTH - COLLISION_RADIUS, y in {CR, HEIGHT - CR} with speed AIRSPEED * factor
 *                               (game.py:97-106), so a spawned (not injected) env is described by its heading and
 *                               one bit.  Envs whose paux.steps carries ACAS2D_STEPS_COMPACT_BIT (+ ACAS2D_STEPS_DOWN_BIT)
 *                               are stepped from these 4 bytes instead of the 16-byte thot record (which stays valid).
 *   episode_idx uint32[B]       episodes started by this env (Philox counter word 2)
 *   min_sep     float[B]        running minimum separation of the episode, or NULL
 *   stats       int64[ACAS2D_STAT_SLOTS][ACAS2D_STAT_FIELDS]  finished-episode counters
 */
#ifndef ACAS2D_B200_H
#define ACAS2D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACAS2D_ABI_VERSION 2

#define ACAS2D_E_NULL        (-1)  /* required pointer is NULL */
#define ACAS2D_E_BAD_TRAFFIC (-2)  /* n_traffic < 1 (reference precondition, game.py:146-147) or > ACAS2D_MAX_TRAFFIC */
#define ACAS2D_E_BAD_SIZE    (-3)  /* num_envs < 0 or exceeds int32 element indexing of an array */
#define ACAS2D_E_NO_DEVICE   (-4)  /* no CUDA device / not an sm_100 device */

#define ACAS2D_MAX_TRAFFIC 1024
#define ACAS2D_PSTAGE_BYTES 112

/* bit of paux.steps: this env's intruder records need their float64 residuals (tres) */
#define ACAS2D_STEPS_RESIDUAL_BIT 0x40000000
/* bits of paux.steps (N == 1, state->tpsi0 given): intruder 0 is in its spawn pattern, heading in tpsi0[env];
 * DOWN = it started at the bottom edge (starts_down, game.py:98-101) */
#define ACAS2D_STEPS_COMPACT_BIT  0x20000000
#define ACAS2D_STEPS_DOWN_BIT     0x10000000
#define ACAS2D_STEPS_MASK         0x0fffffff

/* step flags (uint8 per env) */
#define ACAS2D_FLAG_COLLISION 1u   /* game.py:185-189  d < 2*COLLISION_RADIUS for any intruder */
#define ACAS2D_FLAG_GOAL      2u   /* game.py:191-192 */
#define ACAS2D_FLAG_TIMEOUT   4u   /* game.py:182-183 */
#define ACAS2D_FLAG_DONE      8u   /* game.py:294-310 */
#define ACAS2D_FLAG_OOB       16u  /* aircraft.py:28-29, informational only: never ends an episode */

/* outcome codes, settings.py:6 */
#define ACAS2D_OUTCOME_GOAL 1
#define ACAS2D_OUTCOME_COLLISION 2
#define ACAS2D_OUTCOME_TIMEOUT 3

/* finished-episode counters: stats[slot][field]; sum over slots on the host. */
#define ACAS2D_STAT_SLOTS  128
#define ACAS2D_STAT_FIELDS 16      /* 128 B per slot */
#define ACAS2D_STAT_EPISODES   0
#define ACAS2D_STAT_GOAL       1
#define ACAS2D_STAT_COLLISION  2
#define ACAS2D_STAT_TIMEOUT    3
#define ACAS2D_STAT_LENGTH     4   /* sum of game.steps at episode end */
#define ACAS2D_STAT_RETURN_FX  5   /* sum of total_reward, fixed point: value * 2^20 (order-independent) */
#define ACAS2D_STAT_MINSEP_FX  6   /* sum of per-episode minimum separation * 2^20 (when tracked) */
#define ACAS2D_STAT_FX_SCALE   1048576.0

/* Constants of gym_ACAS2D/settings.py and the derived values of game.py:80-128. */
typedef struct acas2d_params {
    double width, height, fps;                 /* settings.py:15-17 */
    double max_steps;                          /* settings.py:9 */
    double aircraft_size;                      /* settings.py:33 */
    double collision_radius;                   /* settings.py:34 */
    double goal_radius;                        /* settings.py:35 */
    double safe_distance;                      /* settings.py:36 */
    double airspeed;                           /* settings.py:39 */
    double airspeed_factor_min;                /* settings.py:40 */
    double airspeed_factor_max;                /* settings.py:41 */
    double acc_lat_limit;                      /* settings.py:42 */
    double player_heading_lim;                 /* settings.py:43 */
    double traffic_heading_lim;                /* settings.py:44 */
    double reward_goal, reward_collision;      /* settings.py:47-48 */
    double goal_x, goal_y;                     /* game.py:80-81 */
    double player_x0, player_y0;               /* game.py:85-86 */
    double player_psi_base;                    /* game.py:91 relative_angle(player -> goal), degrees */
    double d_goal_max;                         /* game.py:120 */
    double d_dev_max;                          /* game.py:122 */
    double d_separation_max;                   /* game.py:124 */
    double d_cpa_max;                          /* game.py:126 */
    double v_closing_max;                      /* game.py:128 */
    int32_t n_traffic;                         /* MIN_TRAFFIC == MAX_TRAFFIC, settings.py:31-32 */
    int32_t auto_reset;                        /* 0: reference ACAS2DEnv semantics; 1: SB3 VecEnv auto-reset */
} acas2d_params;

typedef struct acas2d_state {
    int64_t   num_envs;
    void     *ppos;
    void     *paux;
    void     *thot;
    void     *tres;
    uint32_t *episode_idx;
    float    *min_sep;        /* optional */
    int64_t  *stats;          /* optional */
    uint64_t  seed;           /* Philox key */
    uint64_t  env_id_offset;  /* global id of env 0 (rank sharding: spawns do not depend on the GPU count) */
    void     *tkin;           /* optional (N > 1): 24 B / intruder kinematic cache, see above */
    float    *tpsi0;          /* optional (N == 1): compact intruder-0 headings, see above */
    void     *pstage;         /* optional scratch (N > 1), ACAS2D_PSTAGE_BYTES per env: with it the float64 player update of
                                 a step runs as its own one-thread-per-env launch and the tiled kernel's lanes read its
                                 result, instead of every lane of an env's group redoing it; contents are per-step */
    float    *spawn_sep;      /* optional (N > 1), float[B]: the game's minimum player-intruder separation AT ITS SPAWN (+inf after
                                 an injection).  Aircraft move at bounded speed, so spawn_sep + k * (largest relative displacement per
                                 step) < 2 * COLLISION_RADIUS proves a collision at step k without looking at a single intruder:
                                 with N >= 256 the reference's spawn rule ends nearly every game that way on its first step */
} acas2d_state;

/* Optional per-step outputs (any pointer may be NULL). */
typedef struct acas2d_step_aux {
    uint8_t *flags;      /* [B]     ACAS2D_FLAG_* of this step */
    uint8_t *outcome;    /* [B]     written only where done */
    float   *term_obs;   /* [B][L]  terminal observation, written only where done (auto_reset) */
    float   *ep_return;  /* [B]     finished episode's total reward, written only where done */
    int32_t *ep_length;  /* [B]     finished episode's game.steps, written only where done */
} acas2d_step_aux;

int acas2d_abi_version(void);

/* settings.py defaults (+ derived values) for `n_traffic` intruders. */
int acas2d_params_default(acas2d_params *params, int32_t n_traffic);

/* Replaces ACAS2DEnv.reset() (environment.py:44-48 -> ACAS2DGame.__init__, game.py:27-160
 * and game.observe, game.py:194-220) for every env whose mask byte is non-zero (all envs
 * if mask == NULL).  Spawns come from Philox4x32-10 keyed by state->seed with counter
 * (global env id, episode_idx, draw slot).  Writes the reset observation rows into obs
 * (float[B][L], rows of unselected envs untouched; obs may be NULL). */
int acas2d_reset(const acas2d_params *params, const acas2d_state *state,
                 const uint8_t *mask, float *obs, void *stream);

/* Replaces ACAS2DEnv.step() (environment.py:29-42: game.action -> observe -> evaluate ->
 * is_done) for all B envs at once.  actions float[B] in [-1,1] (not clipped, game.py:225);
 * obs float[B][L]; reward float[B]; done uint8[B].  With params->auto_reset the env is
 * respawned on done and obs holds the reset observation (SB3 DummyVecEnv semantics). */
int acas2d_step(const acas2d_params *params, const acas2d_state *state,
                const float *actions, float *obs, float *reward, uint8_t *done,
                const acas2d_step_aux *aux, void *stream);

/* Same as acas2d_step but with HOST buffers for actions / obs / reward / done: copies
 * actions host->device, steps, copies the three results device->host and synchronises
 * `stream`.  d_* are caller-owned device staging buffers of the same shapes; h_* should be
 * pinned for full PCIe rate. */
int acas2d_step_host(const acas2d_params *params, const acas2d_state *state,
                     const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done,
                     float *d_actions, float *d_obs, float *d_reward, uint8_t *d_done,
                     const acas2d_step_aux *aux, void *stream);

/* acas2d_step on buffers the DEVICE can address directly (pinned, mapped host memory under unified addressing), then
 * one stream synchronisation: the kernel reads the actions from and writes every output straight to host memory --
 * one launch, no copy.  For tiny batches (the single-env gym surface, ACAS2DEnv.step: environment.py:29-42). */
int acas2d_step_mapped(const acas2d_params *params, const acas2d_state *state, const float *actions, float *obs,
                       float *reward, uint8_t *done, const acas2d_step_aux *aux, void *stream);

/* Host-buffer step for small (latency-bound) batches: the caller keeps obs / reward / done and every aux array
 * inside ONE device block [d_packed, d_packed + packed_bytes); after the step the whole block goes to the pinned
 * host block h_packed with a single copy (terminal rows and finished-episode records included), then the stream
 * is synchronised.  obs / reward / done / aux are the device pointers the step writes (inside the block). */
int acas2d_step_host_packed(const acas2d_params *params, const acas2d_state *state, const float *h_actions,
                            float *d_actions, float *obs, float *reward, uint8_t *done, const acas2d_step_aux *aux,
                            const void *d_packed, void *h_packed, int64_t packed_bytes, void *stream);

/* On-device episode records (reference: the per-step lists of game.py:45-75, appended at
 * game.py:231-239 and 266-276, dumped by testing_main.py:113-138).  A window of envs gets one row of float64
 * values per step in a ring buffer in HBM:
 *   [0] x  [1] y  [2] psi  [3] a_lat  [4] d_sep (minimum separation, new player / OLD traffic, Q10)  [5] d_goal
 *   [6] delta_heading  [7] v_closing  [8] d_cpa  [9] d_dev  [10] r_d_goal  [11] r_h_goal  [12] r_d_cpa  [13] r_d_dev
 *   [14] r_step (shaped reward incl. time discount, no terminal bonus)  [15] game.steps after the step
 *   [16] reward of the step  [17] ACAS2D_FLAG_* of the step;  then x, y of the first n_traffic_rec intruders as the
 *   reference records them (before they move).  A game at steps == 1 first gets its initial row (a_lat = 0, no
 *   discount, reward 0: game.py:132-160).
 * acas2d_trace_step is called BEFORE acas2d_step with the same actions: it computes the rows that step is about to
 * produce without touching the state (the hot kernels carry no tracing code).  cursor int32[num_envs] counts the
 * rows written per env (row index = count % capacity); rows double[num_envs][capacity][18 + 2 n_traffic_rec]. */
#define ACAS2D_TRACE_DOUBLES 18
#define ACAS2D_TRACE_MAX_TRAFFIC 16
typedef struct acas2d_trace {
    int64_t  first_env;      /* traced window: envs [first_env, first_env + num_envs) of the batch */
    int64_t  num_envs;
    int32_t  capacity;       /* rows per env */
    int32_t  n_traffic_rec;  /* intruders whose positions are recorded, <= min(n_traffic, ACAS2D_TRACE_MAX_TRAFFIC) */
    int32_t *cursor;
    double  *rows;
} acas2d_trace;
int acas2d_trace_step(const acas2d_params *params, const acas2d_state *state, const float *actions,
                      const acas2d_trace *trace, void *stream);

/* State injection / extraction (the reference's tests poke game.player / game.traffic
 * attributes directly).  player double[B][3] = x, y, psi; traffic
 * double[B][N][4] = x, y, v_air, psi (CURRENT position); steps int32[B] = game.steps;
 * total_reward double[B].  Device pointers.  inject leaves episode_idx untouched and
 * restarts min_sep from the injected geometry. */
int acas2d_inject_state(const acas2d_params *params, const acas2d_state *state,
                        const double *player, const double *traffic,
                        const int32_t *steps, const double *total_reward, void *stream);
int acas2d_extract_state(const acas2d_params *params, const acas2d_state *state,
                         double *player, double *traffic,
                         int32_t *steps, double *total_reward, void *stream);

/* Observation rows float[B][L] of the games as they stand (game.observe, game.py:194-220, without its
 * steps increment and with the player's last lateral acceleration taken as 0 -- i.e. exactly the
 * reset observation for a state injected at steps == 1).  Does not modify the state. */
int acas2d_observe(const acas2d_params *params, const acas2d_state *state, float *obs, void *stream);

/* Off-path debug frame of env `env_index`: uint8 rgb[HEIGHT][WIDTH][3] (device pointer), the scene of
 * ACAS2DGame.view() (game.py:323-347) without sprites and HUD text.  Not part of the hot path. */
int acas2d_render(const acas2d_params *params, const acas2d_state *state, int64_t env_index, uint8_t *rgb, void *stream);

/* num_steps consecutive steps in ONE launch, for callers that already hold the actions of all of them (open
 * loop: random-action rollouts, replays of recorded action sequences; N_TRAFFIC == 1, no min_sep tracking).
 * actions float[K][B]; obs float[K][B][L], reward float[K][B], done uint8[K][B] receive every step's outputs
 * exactly as K calls of acas2d_step would write them (bit-identical, auto-reset included); the [B]-shaped aux
 * arrays are overwritten by each step (the last finished episode per env stays).  The state is read and
 * written once per launch instead of once per step. */
int acas2d_step_k(const acas2d_params *params, const acas2d_state *state, int32_t num_steps, const float *actions,
                  float *obs, float *reward, uint8_t *done, const acas2d_step_aux *aux, void *stream);

/* Synthetic benchmark path: K consecutive auto-resetting steps per launch with actions
 * drawn in-kernel, a ~ U(-1,1) from Philox(key = action_seed, counter = (global env id,
 * step0 + k)).  State stays in registers between the K steps; nothing but the state, the
 * episode statistics and (optionally) reward_sum float[B] (+= sum of the K rewards) is
 * written. */
int acas2d_rollout_random(const acas2d_params *params, const acas2d_state *state,
                          int32_t num_steps, uint64_t action_seed, uint64_t step0,
                          float *reward_sum, void *stream);

/* Fills actions float[B] with the same U(-1,1) stream acas2d_rollout_random uses for step
 * `step_index` (so the per-step path can be checked against the fused path). */
int acas2d_random_actions(const acas2d_state *state, uint64_t action_seed,
                          uint64_t step_index, float *actions, void *stream);

/* Closed-loop rollout step for the reference's trained agent (N_TRAFFIC == 1): evaluates the SB3
 * MlpPolicy actor saved by gym_ACAS2D/training_main.py:44-52 (policy.pth: mlp_extractor.policy_net
 * 8 -> 64 -> 64 tanh, action_net 64 -> 1) on obs_in float[B][8], optionally adds N(0, exp(log_std)^2)
 * exploration noise from Philox(key = noise_seed, counter = (global env id, step_index)), clips to the
 * action Box and performs acas2d_step with that action -- one kernel.  This is what
 * `model.predict(obs, deterministic=True)` + `env.step` (testing_main.py:74-78) or SB3's rollout
 * collection does per step.  weights: float[ACAS2D_POLICY_FLOATS] device block packed as
 * W1[64][8] | b1[64] | W2[64][64] | b2[64] | W3[64] | b3[1] | 3 pad (rows as SB3 stores them).
 * actions_out float[B] (the unclipped sample; may be NULL), logp_out float[B] (log-probability of
 * the sample; stochastic only; may be NULL).  obs_in and obs_out may be the same buffer.
 * tensor_cores: 0 = float32 on the CUDA cores (matches a torch float32 forward to ~1e-6);
 * 1 = both hidden layers as tcgen05.mma kind::tf32 with TMEM accumulators (action mean within ~2e-3). */
#define ACAS2D_POLICY_FLOATS 4804
int acas2d_policy_step(const acas2d_params *params, const acas2d_state *state, const float *weights,
                       float log_std, const float *obs_in, float *actions_out, float *logp_out, float *obs_out,
                       float *reward, uint8_t *done, const acas2d_step_aux *aux, int32_t stochastic,
                       uint64_t noise_seed, uint64_t step_index, int32_t tensor_cores, void *stream);

/* ------------------------------------------------------------------------------------------------
 * PPO learner for the reference's agent (BASELINE config 5; gym_ACAS2D/training_main.py:44-52:
 * stable_baselines3 PPO('MlpPolicy') with the SB3 1.1.0 defaults stored in best_model.zip/data).
 *
 * params float[ACAS2D_PPO_PARAM_FLOATS]: actor block (the ACAS2D_POLICY_FLOATS layout acas2d_policy_step
 * reads, so `params` itself can be passed there as `weights`) | critic block in the same layout
 * (mlp_extractor.value_net.{0,2}, value_net) | log_std | 3 pad.  All pointers are device pointers; every
 * call is asynchronous on `stream`, allocates nothing and is CUDA-graph capturable. */
#define ACAS2D_PPO_PARAM_FLOATS   (2 * ACAS2D_POLICY_FLOATS + 4)
#define ACAS2D_PPO_PARTIAL_FLOATS 4816     /* one partial-gradient row */
#define ACAS2D_PPO_MAX_CTAS       148      /* partial rows per network */
#define ACAS2D_PPO_WORKSPACE_HEAD  64       /* per-CTA norm partials of the fused update */
#define ACAS2D_PPO_WORKSPACE_FLOATS (ACAS2D_PPO_WORKSPACE_HEAD + 2 * ACAS2D_PPO_MAX_CTAS * ACAS2D_PPO_PARTIAL_FLOATS)
#define ACAS2D_PPO_MAX_RANKS      16       /* data-parallel ranks of the peer-memory gradient exchange */
#define ACAS2D_PPO_EXCHANGE_FLOATS (2 * ACAS2D_PPO_PARAM_FLOATS + ACAS2D_PPO_MAX_RANKS)
#define ACAS2D_PPO_LOSS_STATS     8        /* policy loss, value loss, approx KL, clip fraction, grad norm, 3 spare */

typedef struct acas2d_ppo_config {
    float gamma, gae_lambda;                   /* 0.99, 0.95 */
    float clip_range, vf_coef, ent_coef;       /* 0.2, 0.5, 0.0 */
    float max_grad_norm;                       /* 0.5 (<= 0: no clipping) */
    float lr, beta1, beta2, adam_eps;          /* 3e-4, 0.9, 0.999, 1e-5 */
    int32_t normalize_advantage;               /* 1: (adv - mean) / (std + 1e-8) over the minibatch */
    int32_t reserved;
} acas2d_ppo_config;

/* values[n] = critic(obs[n][8]). */
int acas2d_ppo_values(const float *params, const float *obs, int64_t n, float *values, void *stream);

/* SB3 RolloutBuffer.compute_returns_and_advantage on a [T][B] rollout: rewards float[T][B], dones
 * uint8[T][B] (step t ended the episode), values float[T+1][B] -> advantages, returns float[T][B]. */
int acas2d_ppo_gae(const acas2d_ppo_config *cfg, const float *rewards, const uint8_t *dones, const float *values,
                   int32_t n_steps, int64_t num_envs, float *advantages, float *returns, void *stream);

/* Gradient of  policy_loss + vf_coef * value_loss - ent_coef * entropy  (SB3 ppo.py train()) over the
 * minibatch rows `indices` int64[minibatch] (NULL: rows 0..minibatch-1) of the flattened rollout
 * (obs float[n][8], actions = unclipped samples, old_logp, advantages, returns: float[n]).
 * workspace float[ACAS2D_PPO_WORKSPACE_FLOATS]; grad float[ACAS2D_PPO_PARAM_FLOATS] (overwritten);
 * loss_stats float[ACAS2D_PPO_LOSS_STATS] or NULL; adam_step int32[1] or NULL, incremented by one.
 * Deterministic (no floating-point atomics).  Launches two kernels. */
int acas2d_ppo_grad(const acas2d_ppo_config *cfg, const float *params, const float *obs, const float *actions,
                    const float *old_logp, const float *advantages, const float *returns, const int64_t *indices,
                    int64_t minibatch, float *workspace, float *grad, float *loss_stats, int32_t *adam_step,
                    void *stream);

/* clip_grad_norm_(max_grad_norm) + Adam step on grad * grad_scale (pass 1/world_size after a SUM
 * all-reduce of `grad` over the data-parallel ranks, 1 otherwise).  adam_m / adam_v float[PARAM_FLOATS],
 * adam_step = the counter acas2d_ppo_grad incremented.  loss_stats[4] receives the pre-clip norm. */
int acas2d_ppo_adam(const acas2d_ppo_config *cfg, float *params, const float *grad, float grad_scale,
                    float *adam_m, float *adam_v, const int32_t *adam_step, float *loss_stats, void *stream);

/* One whole gradient step in two kernels: the gradient of acas2d_ppo_grad, then -- fused in one kernel --
 * the fixed-order reduction, the data-parallel gradient exchange, clip_grad_norm_ and the Adam step.
 * sync int32[4] (16-byte aligned), zero-initialised, owned by the learner: [0] = Adam step count (incremented
 * here), [1] = error word (0, or ACAS2D_PPO_ERR_* once a grid barrier / a peer's flag did not arrive within the
 * kernel's bounded spin -- the update of that step is then invalid; the kernel never hangs),
 * [2..3] = a 64-bit count of barrier arrivals that numbers the launches of the update kernel (its grid
 * barriers, the exchange-buffer parity and the peer flags derive from it, not from the step count, so this call
 * may be mixed with acas2d_ppo_grad / acas2d_ppo_adam on the same learner).
 * grad_out float[PARAM_FLOATS] or NULL: the (averaged) gradient that was applied, before clipping.
 *
 * world > 1 (one process per GPU, same call on every rank, same number of calls): peer_exchange is a HOST
 * array of `world` DEVICE pointers, entry r = rank r's exchange block of ACAS2D_PPO_EXCHANGE_FLOATS floats
 * (zero-initialised) mapped into this process (CUDA IPC / VMM symmetric memory; NVLink peer access).  Each
 * rank publishes its gradient in its own block (double-buffered by step parity), signals every peer with a
 * release store, waits for all peers' signals and sums the blocks in rank order -- every rank applies the
 * bit-identical mean gradient and no NCCL call is made.  world == 1: rank 0, peer_exchange NULL.
 * Stream ordering note: the gradient kernel is launched as a PROGRAMMATIC DEPENDENT (it overlaps its gathers of obs /
 * actions / old_logp / advantages / returns / indices with the tail of the kernel before it when that kernel signals
 * early: this library's update kernel and its environment step kernels do).  Do not pass, as those six arrays, buffers
 * that the IMMEDIATELY preceding launch on the stream writes (e.g. the obs buffer of an acas2d_step issued right
 * before); rollout buffers filled by acas2d_policy_step* and anything separated by another launch are fine.
 * The update kernel is a COOPERATIVE launch (its CTAs wait on each other): if the device cannot hold all of them
 * at once the call returns the launch error (cudaErrorCooperativeLaunchTooLarge) instead of dead-locking. */
#define ACAS2D_PPO_ERR_BARRIER 1   /* sync[1]: a grid barrier of the update kernel timed out */
#define ACAS2D_PPO_ERR_PEER    2   /* sync[1]: a peer rank's gradient flag did not arrive */
int acas2d_ppo_step(const acas2d_ppo_config *cfg, float *params, const float *obs, const float *actions,
                    const float *old_logp, const float *advantages, const float *returns, const int64_t *indices,
                    int64_t minibatch, float *workspace, float *adam_m, float *adam_v, int32_t *sync,
                    float *loss_stats, float *grad_out, int32_t rank, int32_t world, void *const *peer_exchange,
                    void *stream);

/* Sets the function attributes of the PPO and policy-step kernels and loads them on the current device
 * (optional; makes a first acas2d_ppo_* / acas2d_policy_step* call legal inside a CUDA-graph capture). */
int acas2d_ppo_prepare(void);

/* acas2d_policy_step with two launch arguments optionally read from device memory at run time, so that a
 * CUDA graph of T captured policy steps (a whole PPO rollout) can be replayed while the learner changes the
 * policy: log_std_dev (NULL: use log_std) points at the live log_std (e.g. params + 2 * ACAS2D_POLICY_FLOATS
 * of the PPO parameter block; `weights` = the same block is re-read by every launch anyway), step_base_dev
 * (NULL: 0) is added to step_index to form the noise counter of the launch. */
int acas2d_policy_step_dyn(const acas2d_params *params, const acas2d_state *state, const float *weights,
                           float log_std, const float *obs_in, float *actions_out, float *logp_out, float *obs_out,
                           float *reward, uint8_t *done, const acas2d_step_aux *aux, int32_t stochastic,
                           uint64_t noise_seed, uint64_t step_index, int32_t tensor_cores,
                           const float *log_std_dev, const uint64_t *step_base_dev, void *stream);

/* Kernels launched by this library since load (all entry points). */
int64_t acas2d_launch_count(void);

/* Experiment knobs (process-wide; also settable through the environment variables
 * ACAS2D_N1_OCC and ACAS2D_FORCE_LOOP before the first call).  n1_occupancy: 1..4 resident
 * blocks per SM for the N_TRAFFIC == 1 kernel (default 2; other values: unchanged).  force_loop: 1 routes
 * N_TRAFFIC > 1 to the one-thread-per-env kernel that the shared-memory tiled kernel is
 * checked against, 0 back to the tiled kernel, negative: unchanged. */
int acas2d_set_tuning(int32_t n1_occupancy, int32_t force_loop);

/* N_TRAFFIC == 1 kernel choice (process-wide; env ACAS2D_N1_TMA / ACAS2D_N1_STAGES): use_tma 1 = the
 * persistent kernel fed by a TMA bulk-copy ring of `stages` (2..5) input tiles (default), 0 = the
 * direct one-thread-per-env kernel, negative = unchanged.  Both are bit-identical in results. */
int acas2d_set_n1_kernel(int32_t use_tma, int32_t stages);

/* N_TRAFFIC > 1 kernel tuning (process-wide; env ACAS2D_TILED_KIN / ACAS2D_TILED_PER_LANE): kin_mode 1 = read the
 * kinematic cache (state->tkin) whenever it is there, 0 = always the 16-byte records, -1 = choose by N (default);
 * per_lane = intruders per lane the lanes-per-env choice aims at (default 8), 0 = unchanged.  Results do not change. */
int acas2d_set_tiled_tuning(int32_t kin_mode, int32_t per_lane);

#ifdef __cplusplus
}
#endif
#endif /* ACAS2D_B200_H */
