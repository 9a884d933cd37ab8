/*
 * acas2d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see acas2d_oracle.h).
 *
 * Scalar float64 restatement of the reference environment step.  Every function
 * names the reference lines it follows.  The arithmetic keeps the reference's
 * operation order (no algebraic shortcuts, no FMA contraction: build with
 * -ffp-contract=off) and recomputes shared quantities wherever the reference
 * does, so that it is both the parity checker and an honest "port" baseline.
 *
 * Parity status: PINNED against the reference's golden CSV and against outputs
 * of the unmodified reference (tests/test_oracle_golden.py).
 */
#include "acas2d_oracle.h"

#include <math.h>
#include <string.h>

#define TWO_PI_PY (2.0 * 3.141592653589793)   /* `2 * math.pi` */
#define RAD2DEG_PY (180.0 / 3.141592653589793) /* CPython math.degrees factor */

/* ------------------------------------------------------------------ helpers */

/* Python float `%` (floor-mod): the result takes the sign of the divisor and may
 * round up to exactly m for tiny negative x.  Used at aircraft.py:22,
 * kinematics.py:20,58,68 and game.py:92,105-106. */
double acas2d_oracle_pymod(double x, double m)
{
    double r = fmod(x, m);
    if (r != 0.0) {
        if ((m < 0.0) != (r < 0.0)) r += m;
    } else {
        r = copysign(0.0, m);
    }
    return r;
}

/* kinematics.py:7-13 -- np.linalg.norm of the 2-vector difference. */
double acas2d_oracle_distance(double x1, double y1, double x2, double y2)
{
    double dx = x1 - x2, dy = y1 - y2;
    return sqrt(dx * dx + dy * dy);
}

/* kinematics.py:16-22 -- bearing from (x1,y1) to (x2,y2) in degrees, [0,360]. */
double acas2d_oracle_relative_angle(double x1, double y1, double x2, double y2)
{
    double rads = acas2d_oracle_pymod(atan2(y2 - y1, x2 - x1), TWO_PI_PY);
    return rads * RAD2DEG_PY;
}

static double deg2rad_ref(double deg)
{
    /* `(psi / 360.0) * 2 * math.pi`, left to right (aircraft.py:23, kinematics.py:29,33,45,59,69) */
    return ((deg / 360.0) * 2.0) * 3.141592653589793;
}

/* kinematics.py:25-37 */
static void relative_speed(const acas2d_oracle_aircraft *a1, const acas2d_oracle_aircraft *a2,
                           double *v12x, double *v12y)
{
    double r1 = deg2rad_ref(a1->psi), r2 = deg2rad_ref(a2->psi);
    *v12x = a1->v_air * cos(r1) - a2->v_air * cos(r2);
    *v12y = a1->v_air * sin(r1) - a2->v_air * sin(r2);
}

/* kinematics.py:40-49 (quirk Q12: arctan, not arctan2; /0 gives inf or nan) */
double acas2d_oracle_distance_closest_approach(const acas2d_oracle_params *P,
        const acas2d_oracle_aircraft *a1, const acas2d_oracle_aircraft *a2)
{
    (void)P;
    double d = acas2d_oracle_distance(a1->x, a1->y, a2->x, a2->y);
    double a_rel = acas2d_oracle_relative_angle(a1->x, a1->y, a2->x, a2->y);
    double a_rel_rad = deg2rad_ref(a_rel);
    double v12x, v12y;
    relative_speed(a1, a2, &v12x, &v12y);
    double h_rel_rad = atan(v12y / v12x);
    return d * sin(a_rel_rad - h_rel_rad);
}

/* kinematics.py:52-79 (quirks Q2, Q3, Q4) */
double acas2d_oracle_closing_speed(const acas2d_oracle_params *P,
        const acas2d_oracle_aircraft *a1, const acas2d_oracle_aircraft *a2)
{
    double dt = 1.0 / P->fps;

    double psi_dot_1 = a1->a_lat / a1->v_air;                                  /* :57  (Q2) */
    double psi_1 = acas2d_oracle_pymod(a1->psi + psi_dot_1 * dt, 360.0);       /* :58 */
    double r1 = deg2rad_ref(psi_1);
    double u1x = a1->v_air * cos(r1) * dt, u1y = a1->v_air * sin(r1) * dt;      /* :64 */
    double x1 = a1->x + u1x, y1 = a1->y + u1y;                                  /* :60-61 */

    double psi_dot_2 = a2->a_lat / a2->v_air;                                  /* :67 */
    double psi_2 = acas2d_oracle_pymod(a2->psi + psi_dot_2 * dt, 360.0);       /* :68 */
    double r2 = deg2rad_ref(psi_2);
    double x2 = a2->x + a2->v_air * cos(r2) * dt;                               /* :70 */
    double y2 = a2->y + a2->v_air * sin(r2) * dt;                               /* :71 */
    double u2x = a2->v_air * cos(r2) * dt;
    double u2y = a1->v_air * sin(r2) * dt;                                      /* :74  (Q3: aircraft1's speed) */

    double dvx = u1x - u2x, dvy = u1y - u2y;
    double dpx = x1 - x2, dpy = y1 - y2;
    double dot = dvx * dpx + dvy * dpy;
    return (dot / acas2d_oracle_distance(x1, y1, x2, y2)) / dt;                 /* :77  (Q4) */
}

/* kinematics.py:82-83 */
double acas2d_oracle_delta_heading(double psi, double phi)
{
    double a = fabs(psi - phi), b = 360.0 - fabs(psi - phi);
    return b < a ? b : a; /* Python min(a, b) */
}

/* Python min(1, x): x only if x < 1 (NaN keeps 1). rewards.py:16,48 */
static double pymin1(double x) { return x < 1.0 ? x : 1.0; }

/* rewards.py:5-9 (the ValueError branch is unreachable for headings in [0,360]) */
double acas2d_oracle_heading_reward(double psi, double phi)
{
    return pow(1.0 - acas2d_oracle_delta_heading(psi, phi) / 180.0, 4.0);
}

/* rewards.py:12-16 */
double acas2d_oracle_closest_approach_reward(const acas2d_oracle_params *P, double v_closing, double d_cpa)
{
    if (v_closing > 0.0) return 1.0;
    return pymin1(pow(d_cpa / P->safe_distance, 4.0));
}

/* rewards.py:19-27 */
double acas2d_oracle_plan_deviation_reward(const acas2d_oracle_params *P, double d_dev)
{
    d_dev = fabs(d_dev);
    double d_goal_init = (P->width - P->goal_radius) - (2.0 * P->aircraft_size);
    double d_dev_max = d_goal_init / 2.0;
    if (d_dev > d_dev_max) return 0.0;
    return pow(1.0 - d_dev / d_dev_max, 0.5);
}

/* rewards.py:44-50 */
double acas2d_oracle_goal_distance_reward(const acas2d_oracle_params *P, double d_goal)
{
    double d_goal_init = (P->width - P->goal_radius) - (2.0 * P->aircraft_size);
    double d_goal_max = d_goal_init + (P->airspeed / P->fps) * P->max_steps;
    return pymin1(pow(1.0 - d_goal / d_goal_max, 4.0));
}

/* rewards.py:53-60 (NaN v_closing takes the else branch, as in Python) */
double acas2d_oracle_step_reward_5(const acas2d_oracle_params *P, double v_closing, double psi,
        double phi, double d_cpa, double d_goal, double d_dev)
{
    if (v_closing <= 0.0)
        return acas2d_oracle_heading_reward(psi, phi) *
               acas2d_oracle_closest_approach_reward(P, v_closing, d_cpa) *
               acas2d_oracle_plan_deviation_reward(P, d_dev);
    return acas2d_oracle_heading_reward(psi, phi) * acas2d_oracle_goal_distance_reward(P, d_goal);
}

/* aircraft.py:16-26 (quirk Q1: heading advances a_lat/v_air DEGREES per step) */
void acas2d_oracle_update_state(const acas2d_oracle_params *P, acas2d_oracle_aircraft *a)
{
    double dt = 1.0 / P->fps;
    double psi_dot = a->a_lat / (a->v_air * dt);
    a->psi = acas2d_oracle_pymod(a->psi + psi_dot * dt, 360.0);
    double r = deg2rad_ref(a->psi);
    a->x = a->x + (a->v_air * cos(r) * dt);
    a->y = a->y + (a->v_air * sin(r) * dt);
}

/* ------------------------------------------------------------ game predicates */

typedef struct {
    acas2d_oracle_aircraft player;
    const double *traffic; /* [N][4] x,y,v,psi */
    int n;
} game_view;

static acas2d_oracle_aircraft traffic_at(const game_view *g, int i)
{
    acas2d_oracle_aircraft t;
    t.x = g->traffic[4 * i + 0];
    t.y = g->traffic[4 * i + 1];
    t.v_air = g->traffic[4 * i + 2];
    t.psi = g->traffic[4 * i + 3];
    t.a_lat = 0.0;
    return t;
}

/* game.py:162-166 */
static double minimum_separation(const game_view *g)
{
    double m = INFINITY;
    for (int i = 0; i < g->n; ++i) {
        double d = acas2d_oracle_distance(g->player.x, g->player.y, g->traffic[4 * i], g->traffic[4 * i + 1]);
        if (d < m) m = d;
    }
    return m;
}

/* game.py:168-169 */
static double distance_to_goal(const acas2d_oracle_params *P, const game_view *g)
{
    return acas2d_oracle_distance(g->player.x, g->player.y, P->goal_x, P->goal_y);
}

/* game.py:171-173 */
static double heading_to_goal(const acas2d_oracle_params *P, const game_view *g)
{
    return acas2d_oracle_relative_angle(g->player.x, g->player.y, P->goal_x, P->goal_y);
}

/* game.py:175-180 */
static double plan_deviation(const acas2d_oracle_params *P, const game_view *g)
{
    double d_goal = distance_to_goal(P, g);
    double h_goal_rad = deg2rad_ref(heading_to_goal(P, g));
    return d_goal * sin(h_goal_rad);
}

/* game.py:185-189 (quirk Q8: strict <) */
static int detect_collisions(const acas2d_oracle_params *P, const game_view *g)
{
    for (int i = 0; i < g->n; ++i)
        if (acas2d_oracle_distance(g->player.x, g->player.y, g->traffic[4 * i], g->traffic[4 * i + 1]) <
            2.0 * P->collision_radius)
            return 1;
    return 0;
}

/* game.py:191-192 */
static int check_goal(const acas2d_oracle_params *P, const game_view *g)
{
    return distance_to_goal(P, g) < P->goal_radius;
}

/* game.py:194-220; the caller has already done `steps += 1` (game.py:197). */
static void observe_row(const acas2d_oracle_params *P, const game_view *g, int steps, double *obs)
{
    obs[0] = (double)steps / P->max_steps;
    obs[1] = g->player.psi / 360.0;
    obs[2] = plan_deviation(P, g) / P->d_dev_max;
    obs[3] = distance_to_goal(P, g) / P->d_goal_max;
    obs[4] = heading_to_goal(P, g) / 360.0;
    for (int i = 0; i < g->n; ++i) {
        acas2d_oracle_aircraft t = traffic_at(g, i);
        obs[5 + 3 * i + 0] = acas2d_oracle_distance(g->player.x, g->player.y, t.x, t.y) / P->d_separation_max;
        obs[5 + 3 * i + 1] = acas2d_oracle_distance_closest_approach(P, &g->player, &t) / P->d_cpa_max;
        obs[5 + 3 * i + 2] = acas2d_oracle_closing_speed(P, &g->player, &t) / P->v_closing_max;
    }
}

/* game.py:249-292 without the record appends (quirks Q6, Q7, Q9) */
static double evaluate(const acas2d_oracle_params *P, const game_view *g, int steps)
{
    acas2d_oracle_aircraft t0 = traffic_at(g, 0);
    double psi = g->player.psi;
    double phi = heading_to_goal(P, g);
    double v_closing = acas2d_oracle_closing_speed(P, &g->player, &t0);
    double d_cpa = acas2d_oracle_distance_closest_approach(P, &g->player, &t0);
    double d_goal = distance_to_goal(P, g);
    double d_dev = plan_deviation(P, g);
    double r_step = acas2d_oracle_step_reward_5(P, v_closing, psi, phi, d_cpa, d_goal, d_dev);
    double tdf = 1.0 - ((double)steps / P->max_steps);
    double reward = r_step * tdf;
    if (detect_collisions(P, g)) reward += P->reward_collision;
    if (check_goal(P, g)) reward += P->reward_goal;
    return reward;
}

/* ------------------------------------------------------------ batched drivers */

static void load_player(const double *row, acas2d_oracle_aircraft *p)
{
    p->x = row[0]; p->y = row[1]; p->v_air = row[2]; p->psi = row[3]; p->a_lat = row[4];
}
static void store_player(double *row, const acas2d_oracle_aircraft *p)
{
    row[0] = p->x; row[1] = p->y; row[2] = p->v_air; row[3] = p->psi; row[4] = p->a_lat;
}

void acas2d_oracle_observe(const acas2d_oracle_params *P, int64_t B, int N,
        const double *player, const double *traffic, int32_t *steps, double *obs)
{
    const int L = 5 + 3 * N;
    for (int64_t b = 0; b < B; ++b) {
        game_view g;
        load_player(player + 5 * b, &g.player);
        g.traffic = traffic + 4 * (int64_t)N * b;
        g.n = N;
        steps[b] += 1; /* game.py:197 (Q5) */
        observe_row(P, &g, steps[b], obs + (int64_t)L * b);
    }
}

/* One env: game.action (game.py:222-247) then observe, evaluate, is_done
 * (environment.py:33-39). */
static void step_one(const acas2d_oracle_params *P, int N,
        double *player, double *traffic, int32_t *steps, double *total_reward,
        uint8_t *running, double *d_path, double *min_sep,
        double action, double *obs, double *reward, uint8_t *flags, uint8_t *outcome)
{
    game_view g;
    load_player(player, &g.player);
    g.traffic = traffic;
    g.n = N;

    /* --- game.action */
    g.player.a_lat = action * P->acc_lat_limit;                 /* game.py:225 (Q19: no clip) */
    double x_old = g.player.x, y_old = g.player.y;
    acas2d_oracle_update_state(P, &g.player);                    /* game.py:229 */
    if (min_sep) {
        double m = minimum_separation(&g);                       /* game.py:237 (Q10: old traffic) */
        if (m < *min_sep) *min_sep = m;
    }
    if (d_path) *d_path += acas2d_oracle_distance(x_old, y_old, g.player.x, g.player.y); /* game.py:241 */
    if (!running || *running) {                                  /* game.py:243-245 (Q17) */
        for (int i = 0; i < N; ++i) {
            acas2d_oracle_aircraft t;
            t.x = traffic[4 * i]; t.y = traffic[4 * i + 1]; t.v_air = traffic[4 * i + 2];
            t.psi = traffic[4 * i + 3]; t.a_lat = 0.0;
            acas2d_oracle_update_state(P, &t);
            traffic[4 * i] = t.x; traffic[4 * i + 1] = t.y; traffic[4 * i + 3] = t.psi;
        }
    }
    store_player(player, &g.player);

    /* --- game.observe */
    *steps += 1;                                                 /* game.py:197 */
    observe_row(P, &g, *steps, obs);

    /* --- game.evaluate */
    double r = evaluate(P, &g, *steps);
    *total_reward += r;                                          /* game.py:287 */
    *reward = r;

    /* --- game.is_done (game.py:294-310; priority timeout > collision > goal, Q9) */
    int timeout = (double)*steps > P->max_steps;                 /* game.py:182-183 */
    int coll = detect_collisions(P, &g);
    int goal = check_goal(P, &g);
    uint8_t f = 0, oc = 0;
    if (coll) f |= ACAS2D_ORACLE_COLLISION;
    if (goal) f |= ACAS2D_ORACLE_GOAL;
    if (timeout) f |= ACAS2D_ORACLE_TIMEOUT;
    if (timeout) oc = 3; else if (coll) oc = 2; else if (goal) oc = 1;
    if (oc) { f |= ACAS2D_ORACLE_DONE; if (running) *running = 0; }
    *flags = f;
    *outcome = oc;
}

void acas2d_oracle_step(const acas2d_oracle_params *P, int64_t B, int N,
        double *player, double *traffic, int32_t *steps, double *total_reward,
        uint8_t *running, double *d_path, double *min_sep,
        const double *actions,
        double *obs, double *reward, uint8_t *flags, uint8_t *outcome)
{
    const int L = 5 + 3 * N;
    for (int64_t b = 0; b < B; ++b)
        step_one(P, N, player + 5 * b, traffic + 4 * (int64_t)N * b, steps + b, total_reward + b,
                 running ? running + b : 0, d_path ? d_path + b : 0, min_sep ? min_sep + b : 0,
                 actions[b], obs + (int64_t)L * b, reward + b, flags + b, outcome + b);
}

void acas2d_oracle_rollout(const acas2d_oracle_params *P, int64_t B, int N, int64_t T,
        double *player, double *traffic, int32_t *steps, double *total_reward,
        uint8_t *running, double *d_path, double *min_sep,
        const double *actions,
        double *obs_out, double *reward_out, uint8_t *flags_out, uint8_t *outcome_out,
        double *player_out, double *traffic_out)
{
    const int L = 5 + 3 * N;
    for (int64_t b = 0; b < B; ++b) {
        int finished = 0;
        for (int64_t t = 0; t < T; ++t) {
            int64_t tb = t * B + b;
            if (!finished) {
                step_one(P, N, player + 5 * b, traffic + 4 * (int64_t)N * b, steps + b, total_reward + b,
                         running ? running + b : 0, d_path ? d_path + b : 0, min_sep ? min_sep + b : 0,
                         actions[tb], obs_out + L * tb, reward_out + tb, flags_out + tb, outcome_out + tb);
                if (flags_out[tb] & ACAS2D_ORACLE_DONE) finished = 1;
            } else {
                int64_t pb = (t - 1) * B + b;
                memcpy(obs_out + L * tb, obs_out + L * pb, sizeof(double) * (size_t)L);
                reward_out[tb] = reward_out[pb];
                flags_out[tb] = flags_out[pb];
                outcome_out[tb] = outcome_out[pb];
            }
            if (player_out) {
                player_out[3 * tb + 0] = player[5 * b + 0];
                player_out[3 * tb + 1] = player[5 * b + 1];
                player_out[3 * tb + 2] = player[5 * b + 3];
            }
            if (traffic_out)
                for (int i = 0; i < N; ++i) {
                    traffic_out[(tb * N + i) * 2 + 0] = traffic[(b * N + i) * 4 + 0];
                    traffic_out[(tb * N + i) * 2 + 1] = traffic[(b * N + i) * 4 + 1];
                }
        }
    }
}

/* ---------------------------------------------------------------- Philox spawn */

/* Philox4x32-10 (Salmon et al., SC'11; Random123 constants). */
void acas2d_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

/* New-framework spawn spec (DESIGN.md "Spawn"): distributions of game.py:85-116,
 * draws from Philox4x32-10 with key = seed, counter = (env_id lo, env_id hi,
 * episode_idx, slot); slot 0 feeds the player and intruder 0, slot i intruder i. */
static void spawn_one(const acas2d_oracle_params *P, int N, uint64_t seed, uint64_t env_id,
                      uint32_t episode, double *player, double *traffic)
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t ctr[4] = { (uint32_t)env_id, (uint32_t)(env_id >> 32), episode, 0u };
    uint32_t r[4];
    acas2d_oracle_philox4x32_10(ctr, key, r);

    double hl = P->player_heading_lim, tl = P->traffic_heading_lim;
    double fmin = P->airspeed_factor_min, fmax = P->airspeed_factor_max;
    double base = acas2d_oracle_relative_angle(P->player_x0, P->player_y0, P->goal_x, P->goal_y);
    player[0] = P->player_x0;
    player[1] = P->player_y0;
    player[2] = P->airspeed;
    player[3] = acas2d_oracle_pymod(base + (-hl + (hl - (-hl)) * u01(r[0])), 360.0); /* game.py:91-92 */
    player[4] = 0.0;

    double sd = (double)(r[1] >> 31);                                               /* game.py:98 */
    traffic[0] = P->width - P->collision_radius;                                     /* game.py:100 */
    traffic[1] = P->collision_radius + (sd * (P->height - (2.0 * P->collision_radius))); /* :101 */
    traffic[2] = (fmin + (fmax - fmin) * u01(r[2])) * P->airspeed;                   /* :103 */
    traffic[3] = acas2d_oracle_pymod(145.0 + (sd * 70.0) + (-tl + (tl - (-tl)) * u01(r[3])), 360.0); /* :105-106 */

    for (int i = 1; i < N; ++i) {
        /* intruders n > 0 are stored as float32, so the new framework draws them in float32:
         * u = 24 random bits * 2^-24, distributions of game.py:109-114 */
        ctr[3] = (uint32_t)i;
        acas2d_oracle_philox4x32_10(ctr, key, r);
        const float two24 = 5.9604644775390625e-08f;
        float u0 = (float)(r[0] >> 8) * two24, u1 = (float)(r[1] >> 8) * two24;
        float u2 = (float)(r[2] >> 8) * two24, u3 = (float)(r[3] >> 8) * two24;
        traffic[4 * i + 0] = (double)((float)(P->width - P->aircraft_size) * u0);                  /* game.py:109 */
        traffic[4 * i + 1] = (double)((float)(3.0 * P->height / 5.0) * u1);                       /* :110 */
        traffic[4 * i + 2] = (double)(fmaf((float)(fmax - fmin), u2, (float)fmin) * (float)P->airspeed); /* :112 */
        traffic[4 * i + 3] = (double)(360.0f * u3);                                                /* :114 */
    }
    /* the new framework stores a spawned intruder as four float32 values (DESIGN.md "Data layout") */
    for (int q = 0; q < 4 * N; ++q) traffic[q] = (double)(float)traffic[q];
}

void acas2d_oracle_spawn_philox(const acas2d_oracle_params *P, int64_t B, int N,
        uint64_t seed, uint64_t env_id_offset, const uint32_t *episode_idx,
        double *player, double *traffic)
{
    for (int64_t b = 0; b < B; ++b)
        spawn_one(P, N, seed, env_id_offset + (uint64_t)b, episode_idx[b],
                  player + 5 * b, traffic + 4 * (int64_t)N * b);
}

void acas2d_oracle_vec_step(const acas2d_oracle_params *P, int64_t B, int N,
        uint64_t seed, uint64_t env_id_offset, uint32_t *episode_idx,
        double *player, double *traffic, int32_t *steps, double *total_reward,
        double *d_path, double *min_sep,
        const double *actions,
        double *obs, double *reward, uint8_t *flags, uint8_t *outcome,
        double *term_obs, double *ep_return, int32_t *ep_length)
{
    const int L = 5 + 3 * N;
    for (int64_t b = 0; b < B; ++b) {
        double *pl = player + 5 * b, *tr = traffic + 4 * (int64_t)N * b;
        step_one(P, N, pl, tr, steps + b, total_reward + b, 0,
                 d_path ? d_path + b : 0, min_sep ? min_sep + b : 0,
                 actions[b], obs + (int64_t)L * b, reward + b, flags + b, outcome + b);
        if (flags[b] & ACAS2D_ORACLE_DONE) {
            if (term_obs) memcpy(term_obs + (int64_t)L * b, obs + (int64_t)L * b, sizeof(double) * (size_t)L);
            if (ep_return) ep_return[b] = total_reward[b];
            if (ep_length) ep_length[b] = steps[b];
            episode_idx[b] += 1;
            spawn_one(P, N, seed, env_id_offset + (uint64_t)b, episode_idx[b], pl, tr);
            steps[b] = 0;                   /* game.py:30 */
            total_reward[b] = 0.0;          /* game.py:32 */
            if (d_path) d_path[b] = 0.0;    /* game.py:44 */
            game_view g;
            load_player(pl, &g.player);
            g.traffic = tr; g.n = N;
            if (min_sep) min_sep[b] = minimum_separation(&g); /* game.py:141 */
            steps[b] += 1;                  /* environment.py:47 -> game.py:197 */
            observe_row(P, &g, steps[b], obs + (int64_t)L * b);
        }
    }
}
