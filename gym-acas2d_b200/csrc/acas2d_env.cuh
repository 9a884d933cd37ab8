// acas2d_env.cuh -- one environment's step / reset / inject / extract, operating on state
// held in registers and on the SoA arrays of include/acas2d_b200.h.  __host__ __device__
// for the same reason as acas2d_math.cuh: tests/hostcheck compiles this file with g++ to
// check the step orchestration against the CPU oracle; the product only runs it on the GPU.
#pragma once

#include <string.h>

#include "../../include/acas2d_b200.h"
#include "acas2d_math.cuh"

// Store policies (experiment switches: -DACAS2D_OUT_STORE=0|1|2, -DACAS2D_STATE_STORE=0|1|2;
// 0 = default write-back, 1 = .cs streaming / evict-first, 2 = .wt write-through).
#ifndef ACAS2D_OUT_STORE
#define ACAS2D_OUT_STORE 1      /* outputs are not re-read by the step: streaming */
#endif
#ifndef ACAS2D_STORE_256
#define ACAS2D_STORE_256 0      /* experiment: 256-bit st.global.v8.f32 for the 32-byte observation row */
#endif
#ifndef ACAS2D_STATE_STORE
#define ACAS2D_STATE_STORE 1    /* measured: 95.3 -> 91.6 us at 4 Mi envs (state does not fit L2; evict-first keeps the
                                   next step's inputs from being pushed out by lines nobody re-reads soon) */
#endif
#if defined(__CUDA_ARCH__)
#define ACAS_ST_POLICY(sel, ptr, val) \
    do { if ((sel) == 1) __stcs((ptr), (val)); else if ((sel) == 2) __stwt((ptr), (val)); else *(ptr) = (val); } while (0)
#else
#define ACAS_ST_POLICY(sel, ptr, val) (*(ptr) = (val))
#endif
#define ACAS_STCS(ptr, val) ACAS_ST_POLICY(ACAS2D_OUT_STORE, ptr, val)

namespace acas2d {

#if defined(__CUDACC__)
typedef float4 Float4;
#else
struct alignas(16) Float4 { float x, y, z, w; };
#endif

// Cold full-precision complement of a traffic record (see acas2d_b200.h "tres").
struct alignas(32) Residual { double x0, y0, psi, v; };

constexpr int32_t kResidualBit = ACAS2D_STEPS_RESIDUAL_BIT;
constexpr int32_t kCompactBit = ACAS2D_STEPS_COMPACT_BIT;   // the env's fast-path records (tkin / tpsi0) describe it exactly
constexpr int32_t kDownBit = ACAS2D_STEPS_DOWN_BIT;
constexpr int32_t kStepsMask = ACAS2D_STEPS_MASK;

// Kinematic cache of one intruder (acas2d_b200.h "tkin"): origin (float32-exact) and displacement per step.
struct alignas(8) TrafficKin { float x0, y0; double dx, dy; };

struct StatePtrs {
    Vec2d *ppos;
    PlayerAux *paux;
    Float4 *thot;
    Residual *tres;
    uint32_t *episode_idx;
    float *min_sep;
    long long *stats;
    uint64_t seed;
    uint64_t gid0;
    int64_t B;
    TrafficKin *tkin;       // optional (N > 1)
    float *tpsi0;           // optional (N == 1)
    Float4 *pstage;         // optional scratch (N > 1): [7][B] 16-byte words
    float *spawn_sep;       // optional (N > 1): minimum separation of the game at its spawn
};

struct Sinks {
    float *obs;
    float *reward;
    uint8_t *done;
    uint8_t *flags;
    uint8_t *outcome;
    float *term_obs;
    float *ep_return;
    int32_t *ep_length;
};

// ---------------------------------------------------------------- host-side parameter prep
inline DevParams make_dev_params(const acas2d_params &p)
{
    DevParams d;
    memset(&d, 0, sizeof(d));
    const double dt = 1.0 / p.fps;                                  // aircraft.py:18
    d.dt = dt;
    d.airspeed = p.airspeed;
    d.v_dt = p.airspeed * dt;
    d.dpsi_per_action = p.acc_lat_limit / p.airspeed;               // aircraft.py:20-22 (Q1)
    d.acc_lat_limit = p.acc_lat_limit;
    d.lookahead_rad = dt * kDeg2Rad;                                // kinematics.py:57-59 (Q2)
    d.goal_x = p.goal_x; d.goal_y = p.goal_y;
    d.coll_d2 = (2.0 * p.collision_radius) * (2.0 * p.collision_radius);   // game.py:187
    d.goal_r2 = p.goal_radius * p.goal_radius;                      // game.py:192
    d.width = p.width; d.height = p.height;
    d.inv_360 = 1.0 / 360.0;
    d.inv_d_dev_max = 1.0 / p.d_dev_max;
    d.player_x0 = p.player_x0; d.player_y0 = p.player_y0;
    d.player_psi_base = p.player_psi_base;
    d.player_heading_lim = p.player_heading_lim;
    d.traffic_heading_lim = p.traffic_heading_lim;
    d.factor_min = p.airspeed_factor_min;
    d.factor_span = p.airspeed_factor_max - p.airspeed_factor_min;
    d.t0_x = p.width - p.collision_radius;                          // game.py:100
    d.t0_y_up = p.collision_radius;                                 // game.py:101
    d.t0_y_span = p.height - 2.0 * p.collision_radius;
    d.tn_x_span = p.width - p.aircraft_size;                        // game.py:109
    d.tn_y_span = 3.0 * p.height / 5.0;                             // game.py:110
    d.inv_max_steps = (float)(1.0 / p.max_steps);
    d.inv_d_dev_max_f = (float)(1.0 / p.d_dev_max);
    d.inv_d_goal_max = (float)(1.0 / p.d_goal_max);
    d.inv_d_sep_max = (float)(1.0 / p.d_separation_max);
    d.inv_d_cpa_max = (float)(1.0 / p.d_cpa_max);
    d.vc_scale = (float)(p.fps / p.v_closing_max);
    d.fps = (float)p.fps;
    d.inv_safe_distance = (float)(1.0 / p.safe_distance);
    const double d_goal_init = (p.width - p.goal_radius) - 2.0 * p.aircraft_size;   // rewards.py:22,46
    d.rw_dev_max = (float)(d_goal_init / 2.0);
    d.inv_rw_dev_max = (float)(2.0 / d_goal_init);
    d.inv_rw_goal_max = (float)(1.0 / (d_goal_init + (p.airspeed / p.fps) * p.max_steps));
    d.reward_goal = (float)p.reward_goal;
    d.reward_collision = (float)p.reward_collision;
    d.tn_x_span_f = (float)d.tn_x_span; d.tn_y_span_f = (float)d.tn_y_span;
    d.factor_min_f = (float)d.factor_min; d.factor_span_f = (float)d.factor_span; d.airspeed_f = (float)p.airspeed;
    d.n_traffic = p.n_traffic;
    d.max_steps = (int32_t)p.max_steps;
    d.auto_reset = p.auto_reset;
    // spawned speeds are rounded to float32 (records): "exactly AIRSPEED" also needs AIRSPEED to survive that
    d.q3_trivial = (p.airspeed_factor_min == 1.0 && p.airspeed_factor_max == 1.0 &&
                    (double)(float)p.airspeed == p.airspeed) ? 1 : 0;
    const double sure = 2.0 * p.collision_radius - 0.05;
    d.coll_sure_d2 = sure > 0.0 ? (float)(sure * sure) : 0.0f;
    d.dt_f = (float)dt;
    {   // the new game's player-only observation entries (see reset_view)
        Player p0;
        p0.x = p.player_x0; p0.y = p.player_y0; p0.psi = 0.0; p0.c = 1.0; p0.s = 0.0; p0.cl = 1.0; p0.sl = 0.0;
        const PlayerView v0 = player_view(d, p0, 1);
        d.reset_obs0 = v0.obs[0]; d.reset_obs2 = v0.obs[2]; d.reset_obs3 = v0.obs[3]; d.reset_obs4 = v0.obs[4];
    }
    d.c_x0 = (double)(float)d.t0_x;
    d.c_y0_up = (double)(float)d.t0_y_up;
    d.c_y0_down = (double)(float)(d.t0_y_up + 1.0 * d.t0_y_span);
    d.c_v = (double)(float)((d.factor_min + d.factor_span * 0.5) * p.airspeed);      // used only when factor_span == 0
    d.vrel_step = (float)(p.airspeed * dt * (1.0 + p.airspeed_factor_max) * (1.0 + 1e-6));
    d.coll_sure = (float)(2.0 * p.collision_radius - 1e-3);
    return d;
}

inline StatePtrs make_state_ptrs(const acas2d_state &s)
{
    StatePtrs o;
    o.ppos = (Vec2d *)s.ppos;
    o.paux = (PlayerAux *)s.paux;
    o.thot = (Float4 *)s.thot;
    o.tres = (Residual *)s.tres;
    o.episode_idx = s.episode_idx;
    o.min_sep = s.min_sep;
    o.stats = (long long *)s.stats;
    o.seed = s.seed;
    o.gid0 = s.env_id_offset;
    o.B = s.num_envs;
    o.tkin = (TrafficKin *)s.tkin;
    o.tpsi0 = s.tpsi0;
    o.pstage = (Float4 *)s.pstage;
    o.spawn_sep = s.spawn_sep;
    return o;
}

// ---------------------------------------------------------------- episode statistics
struct Tally {
    int episodes, goal, coll, tout;
    long long length, ret_fx, minsep_fx;
};

ACAS_HD void tally_clear(Tally &t)
{
    t.episodes = t.goal = t.coll = t.tout = 0;
    t.length = t.ret_fx = t.minsep_fx = 0;
}

ACAS_HD void tally_add(Tally &t, int outcome, int steps, float ep_return, float minsep, bool has_minsep)
{
    t.episodes += 1;
    t.goal += outcome == ACAS2D_OUTCOME_GOAL;
    t.coll += outcome == ACAS2D_OUTCOME_COLLISION;
    t.tout += outcome == ACAS2D_OUTCOME_TIMEOUT;
    t.length += steps;
    t.ret_fx += llrint((double)ep_return * ACAS2D_STAT_FX_SCALE);
    if (has_minsep) t.minsep_fx += llrint((double)minsep * ACAS2D_STAT_FX_SCALE);
}

// ---------------------------------------------------------------- traffic records
// Hot record: float4 {x0, y0, psi, v}: position at game.steps == 1, heading [deg], speed.
// Spawned intruders are float32-representable by construction (spawn_traffic rounds), so the
// record is exact; injected float64 states keep their remainder in the cold Residual array and
// set kResidualBit in PlayerAux.steps.  The intruder flies a straight line (game.py:243-245,
// a_lat == 0): position after k moves = origin + k * (v cos psi dt, v sin psi dt), in float64.
struct TrafficRec { double x0, y0, psi, v; };

ACAS_HD TrafficRec traffic_load(const StatePtrs &S, int64_t ij, bool residual)
{
    const Float4 h = S.thot[ij];
    TrafficRec t;
    t.x0 = (double)h.x; t.y0 = (double)h.y; t.psi = (double)h.z; t.v = (double)h.w;
    if (residual) {
        const Residual r = S.tres[ij];
        t.x0 += r.x0; t.y0 += r.y0; t.psi += r.psi; t.v += r.v;      // exact: r = full - float(full)
    }
    return t;
}

// Returns true when the record needed a residual (i.e. was not float32-representable).
ACAS_HD bool traffic_store(const StatePtrs &S, int64_t ij, const TrafficRec &t, bool write_residual)
{
    Float4 h;
    h.x = (float)t.x0; h.y = (float)t.y0; h.z = (float)t.psi; h.w = (float)t.v;
    S.thot[ij] = h;
    Residual r;
    r.x0 = t.x0 - (double)h.x; r.y0 = t.y0 - (double)h.y; r.psi = t.psi - (double)h.z; r.v = t.v - (double)h.w;
    const bool need = r.x0 != 0.0 || r.y0 != 0.0 || r.psi != 0.0 || r.v != 0.0;
    if (write_residual) S.tres[ij] = r;
    return need;
}

ACAS_HD Intruder intruder_at(const DevParams &P, const TrafficRec &t, double k)
{
    Intruder it;
    double s, c;
    sincos_deg(t.psi, &s, &c);
    it.dx = (t.v * c) * P.dt;                                           // aircraft.py:25-26 with a_lat = 0
    it.dy = (t.v * s) * P.dt;
    it.x = fma(k, it.dx, t.x0);
    it.y = fma(k, it.dy, t.y0);
    it.dyq = (P.airspeed * s) * P.dt;                                   // Q3 (no division: the player's speed times the sine)
    return it;
}

// The same intruder from its kinematic cache: no sin / cos (the cached dx, dy are the values intruder_at
// derives, bit for bit).  Only for envs carrying kCompactBit: float32-exact origin, speed == AIRSPEED (Q3 trivial).
ACAS_HD Intruder intruder_from_kin(const TrafficKin &q, double k)
{
    Intruder it;
    it.dx = q.dx; it.dy = q.dy; it.dyq = q.dy;
    it.x = fma(k, q.dx, (double)q.x0);
    it.y = fma(k, q.dy, (double)q.y0);
    return it;
}

ACAS_HD void kin_store(const StatePtrs &S, int64_t ij, const TrafficRec &t, const Intruder &it)
{
    if (S.tkin == nullptr) return;
    TrafficKin q;
    q.x0 = (float)t.x0; q.y0 = (float)t.y0; q.dx = it.dx; q.dy = it.dy;
    S.tkin[ij] = q;
}

// May this intruder use the fast-path records?  (float32-exact record AND flying at exactly AIRSPEED.)
ACAS_HD bool traffic_is_plain(const DevParams &P, const TrafficRec &t, bool needs_residual)
{
    return !needs_residual && t.v == P.airspeed;
}

// Flag bits of paux.steps for a freshly spawned env with N > 1 (the N == 1 kernels set theirs in step_env1).
ACAS_HD int32_t spawn_bits(const DevParams &P, const StatePtrs &S)
{
    return (S.tkin != nullptr && P.q3_trivial && P.n_traffic > 1) ? kCompactBit : 0;
}

// Spawn of intruder j of (gid, episode), rounded to float32 (game.py:97-114).
ACAS_HD TrafficRec spawn_traffic(const DevParams &P, uint64_t seed, uint64_t gid, uint32_t episode, int j,
                                 const Spawn0 &sp)
{
    TrafficRec t;
    if (j == 0) { t.x0 = sp.x; t.y0 = sp.y; t.v = sp.v; t.psi = sp.psi; }
    else {
        const SpawnN sn = spawn_slot(P, seed, gid, episode, (uint32_t)j);
        t.x0 = sn.x; t.y0 = sn.y; t.v = sn.v; t.psi = sn.psi;
    }
    t.x0 = (double)(float)t.x0; t.y0 = (double)(float)t.y0;
    t.psi = (double)(float)t.psi; t.v = (double)(float)t.v;
    return t;
}

// Spawn of intruder j with its records written and its reset-observation encounter returned: what
// spawn_traffic + traffic_store + intruder_at(k = 0) + kin_store + encounter do, bit for bit, without the
// float32 <-> float64 round trips in between (each one occupies the conversion pipe for as long as four DFMAs).
ACAS_HD Encounter spawn_intruder(const DevParams &P, const StatePtrs &S, uint64_t gid, uint32_t episode, int j,
                                 const Spawn0 &sp, const Player &rp, int64_t ij)
{
    Float4 h;
    if (j == 0) { h.x = (float)sp.x; h.y = (float)sp.y; h.z = (float)sp.psi; h.w = (float)sp.v; }
    else {
        const U4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), episode, (uint32_t)j,
                                   (uint32_t)S.seed, (uint32_t)(S.seed >> 32));
        h.x = P.tn_x_span_f * u01f(r.x);                                   // game.py:109-114, as spawn_slot
        h.y = P.tn_y_span_f * u01f(r.y);
        h.w = fmaf(P.factor_span_f, u01f(r.z), P.factor_min_f) * P.airspeed_f;
        h.z = 360.0f * u01f(r.w);
    }
    S.thot[ij] = h;
    TrafficRec tr;
    tr.x0 = (double)h.x; tr.y0 = (double)h.y; tr.psi = (double)h.z;
    tr.v = P.q3_trivial ? P.airspeed : (double)h.w;                        // == (double)h.w: the factor is exactly 1
    double sn, cs;
    sincos_deg(tr.psi, &sn, &cs);
    Intruder it;
    it.dx = (tr.v * cs) * P.dt; it.dy = (tr.v * sn) * P.dt; it.dyq = (P.airspeed * sn) * P.dt;
    it.x = tr.x0; it.y = tr.y0;
    if (S.tkin != nullptr) {
        TrafficKin q;
        q.x0 = h.x; q.y0 = h.y; q.dx = it.dx; q.dy = it.dy;
        S.tkin[ij] = q;
    }
    return encounter(P, rp, it);
}

// ---------------------------------------------------------------- N_TRAFFIC == 1
struct Env1 {
    double px, py, psi;
    int32_t steps;          // game.steps, without the flag bits
    int32_t bits;           // the flag bits of paux.steps (kResidualBit | kCompactBit | kDownBit)
    float ret;
    TrafficRec tr;
    float minsep;
    bool respawned;
};

// N == 1, compact form (acas2d_b200.h "tpsi0"): intruder 0 of a SPAWNED env sits in the reference's spawn pattern
// (game.py:97-106) -- fixed x, one of two y, fixed speed when the speed factor is a constant -- so its heading
// and one bit describe it.  The record a spawn writes is rounded to float32 (spawn_traffic); these are the same
// roundings.
ACAS_HD bool compact_ok(const DevParams &P, const StatePtrs &S)
{
    return S.tpsi0 != nullptr && P.factor_span == 0.0 && P.n_traffic == 1;
}

ACAS_HD TrafficRec compact_traffic(const DevParams &P, float psi, bool down)
{
    TrafficRec t;                                       // the float32 roundings are done once, in make_dev_params
    t.x0 = P.c_x0;
    t.y0 = down ? P.c_y0_down : P.c_y0_up;
    t.v = P.c_v;                                        // factor_span == 0 here
    t.psi = (double)psi;
    return t;
}

ACAS_HD void store_obs8(float *row, const PlayerView &v, const Encounter &e, const DevParams &P)
{
    Float4 a; a.x = v.obs[0]; a.y = v.obs[1]; a.z = v.obs[2]; a.w = v.obs[3];
    Float4 b; b.x = v.obs[4]; b.y = e.d * P.inv_d_sep_max; b.z = e.d_cpa * P.inv_d_cpa_max; b.w = e.v_c * P.vc_scale;
#if defined(__CUDA_ARCH__) && ACAS2D_STORE_256
    // one 256-bit streaming store per row (sm_100): a full 32-byte sector per thread in one request
    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(row), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
#else
    ACAS_STCS((Float4 *)row, a);
    ACAS_STCS((Float4 *)row + 1, b);
#endif
}

// New game for a single-intruder env (SB3 DummyVecEnv semantics): Philox spawn, reset
// observation (environment.py:47: steps becomes 1), intruder records.  The caller stores the player state.
template <bool EMIT>
ACAS_HD void respawn_env1(const DevParams &P, const StatePtrs &S, Env1 &e, int64_t i, const Sinks &out)
{
    const uint32_t episode = S.episode_idx[i];
    S.episode_idx[i] = episode + 1u;
    const uint64_t gid = S.gid0 + (uint64_t)i;
    const Spawn0 sp = spawn_slot0(P, S.seed, gid, episode);
    Player p;
    p.x = P.player_x0; p.y = P.player_y0;
    player_set_heading_straight(p, sp.player_psi);                          // a_lat = 0 in a new game
    e.tr = spawn_traffic(P, S.seed, gid, episode, 0, sp);
    const Intruder t = intruder_at(P, e.tr, 0.0);
    const PlayerView v1 = reset_view(P, p.psi);
    const Encounter e1 = encounter(P, p, t);
    if (EMIT) store_obs8(out.obs + 8 * i, v1, e1, P);
    e.px = p.x; e.py = p.y; e.psi = p.psi; e.steps = 1; e.ret = 0.0f;
    e.bits = 0;
    if (compact_ok(P, S)) {
        S.tpsi0[i] = (float)e.tr.psi;
        e.bits = kCompactBit | (sp.y != P.t0_y_up ? kDownBit : 0);
    }
    e.minsep = e1.d;                                                       // game.py:141
    e.respawned = true;
}

// One environment step for a single-intruder env held in registers.
// EMIT = write per-step outputs through `out`; i = local env index.  DEFER: a finished env that must respawn is
// left untouched (its terminal outputs are written) and `true` is returned -- the caller respawns it later with
// respawn_env1 (the persistent kernel queues these so that a lone respawning lane does not hold up its warp).
template <bool MINSEP, bool EMIT, bool DEFER = false>
ACAS_HD bool step_env1(const DevParams &P, const StatePtrs &S, Env1 &e, float action,
                       int64_t i, const Sinks &out, Tally &tally, float *reward_acc)
{
    // ---- game.action (game.py:222-247)
    const double dpsi = (double)action * P.dpsi_per_action;               // game.py:225 + aircraft.py:20-22
    Player p;
    p.x = e.px; p.y = e.py;
    player_set_heading(P, p, wrap360(e.psi + dpsi), dpsi);
    player_advance(P, p);

    const int k = e.steps;                                                 // intruder moves after this step
    const Intruder t = intruder_at(P, e.tr, (double)k);                    // game.py:243-245
    if (MINSEP) {                                                          // game.py:237 (Q10: old traffic)
        const double ox = (t.x - t.dx) - p.x, oy = (t.y - t.dy) - p.y;
        e.minsep = fminf(e.minsep, acas_sqrtf((float)(ox * ox + oy * oy)));
    }

    // ---- game.observe / evaluate / is_done (game.py:194-314)
    const int steps = k + 1;                                               // game.py:197
    const PlayerView v = player_view(P, p, steps);
    const Encounter en = encounter(P, p, t);
    float r = shaped_reward(P, p, v, en, steps);
    const bool coll = en.d2 < P.coll_d2;                                   // game.py:187 (strict, Q8)
    const bool goal = v.dg2 < P.goal_r2;                                   // game.py:192
    const bool tout = steps > P.max_steps;                                 // game.py:183
    // game.py:279-284 (Q9): both bonuses can fall on one step; summed first so that -1000 + 1000 does not
    // round the shaped reward to the float32 spacing at 1000
    r += (coll ? P.reward_collision : 0.0f) + (goal ? P.reward_goal : 0.0f);
    const float ret = e.ret + r;                                           // game.py:287
    const int outcome = tout ? ACAS2D_OUTCOME_TIMEOUT : coll ? ACAS2D_OUTCOME_COLLISION
                        : goal ? ACAS2D_OUTCOME_GOAL : 0;                  // game.py:297-310
    const bool done = outcome != 0;

    if (EMIT) {
        ACAS_STCS(out.reward + i, r);
        out.done[i] = (uint8_t)done;
        if (out.flags) {
            const bool oob = p.x < 0.0 || p.x > P.width || p.y < 0.0 || p.y > P.height;   // aircraft.py:28-29
            out.flags[i] = (uint8_t)((coll ? ACAS2D_FLAG_COLLISION : 0) | (goal ? ACAS2D_FLAG_GOAL : 0) |
                                     (tout ? ACAS2D_FLAG_TIMEOUT : 0) | (done ? ACAS2D_FLAG_DONE : 0) |
                                     (oob ? ACAS2D_FLAG_OOB : 0));
        }
    } else if (reward_acc) {
        *reward_acc += r;
    }

    if (!done || !P.auto_reset) {
        if (EMIT) store_obs8(out.obs + 8 * i, v, en, P);
        if (done) {   // reference ACAS2DEnv semantics: the finished game stays in place until reset()
            if (EMIT) {
                if (out.outcome) out.outcome[i] = (uint8_t)outcome;
                if (out.ep_return) out.ep_return[i] = ret;
                if (out.ep_length) out.ep_length[i] = steps;
            }
            tally_add(tally, outcome, steps, ret, e.minsep, MINSEP);
        }
        e.px = p.x; e.py = p.y; e.psi = p.psi; e.steps = steps; e.ret = ret;
        return false;
    }

    // ---- auto-reset: terminal outputs now, the new game now or (DEFER) later
    if (EMIT) {
        if (out.term_obs) store_obs8(out.term_obs + 8 * i, v, en, P);
        if (out.outcome) out.outcome[i] = (uint8_t)outcome;
        if (out.ep_return) out.ep_return[i] = ret;
        if (out.ep_length) out.ep_length[i] = steps;
    }
    tally_add(tally, outcome, steps, ret, e.minsep, MINSEP);
    if (DEFER) return true;
    respawn_env1<EMIT>(P, S, e, i, out);
    return false;
}

ACAS_HD void load_env1(const StatePtrs &S, int64_t i, Env1 &e, bool minsep)
{
    const Vec2d pp = S.ppos[i];
    const PlayerAux pa = S.paux[i];
    e.px = pp.x; e.py = pp.y; e.psi = pa.psi; e.ret = pa.ep_return;
    e.steps = pa.steps & kStepsMask;
    e.bits = pa.steps & ~kStepsMask;
    e.tr = traffic_load(S, i, (e.bits & kResidualBit) != 0);          // the 16-byte record is valid in compact form too
    e.minsep = minsep ? S.min_sep[i] : 0.0f;
    e.respawned = false;
}

// store_env1 with an L2 cache-hint policy on the two player records (device only): the persistent N == 1 kernel keeps
// a slice of the player state L2-resident across launches (createpolicy evict_last), the rest streams (evict_first).
#if defined(__CUDACC__)
__device__ __forceinline__ void store_env1_hinted(const StatePtrs &S, int64_t i, const Env1 &e, uint64_t policy)
{
    Vec2d pp; pp.x = e.px; pp.y = e.py;
    PlayerAux pa; pa.psi = e.psi; pa.steps = e.steps | e.bits; pa.ep_return = e.ret;
    Float4 a, b;
    a.x = __int_as_float(__double2loint(pp.x)); a.y = __int_as_float(__double2hiint(pp.x));
    a.z = __int_as_float(__double2loint(pp.y)); a.w = __int_as_float(__double2hiint(pp.y));
    b.x = __int_as_float(__double2loint(pa.psi)); b.y = __int_as_float(__double2hiint(pa.psi));
    b.z = __int_as_float(pa.steps); b.w = pa.ep_return;
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(S.ppos + i), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "l"(policy) : "memory");
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(S.paux + i), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "l"(policy) : "memory");
    if (e.respawned) traffic_store(S, i, e.tr, false);
}
#endif

ACAS_HD void store_env1(const StatePtrs &S, int64_t i, const Env1 &e, bool minsep)
{
    Vec2d pp; pp.x = e.px; pp.y = e.py;
    PlayerAux pa; pa.psi = e.psi; pa.steps = e.steps | e.bits; pa.ep_return = e.ret;
    {
        Float4 a, b;                                   // two 16-byte records, stored as 128-bit words
        memcpy(&a, &pp, 16); memcpy(&b, &pa, 16);
        ACAS_ST_POLICY(ACAS2D_STATE_STORE, (Float4 *)(S.ppos + i), a);
        ACAS_ST_POLICY(ACAS2D_STATE_STORE, (Float4 *)(S.paux + i), b);
    }
    if (e.respawned) traffic_store(S, i, e.tr, false);
    if (minsep) S.min_sep[i] = e.minsep;
}

// ---------------------------------------------------------------- any N_TRAFFIC, one thread per env
// Intruders streamed straight from global memory.  Correct for every N; it is the simple
// form the shared-memory tiled kernel is checked against.
template <bool MINSEP>
ACAS_HD void step_env_loop(const DevParams &P, const StatePtrs &S, int64_t i, float action,
                           const Sinks &out, Tally &tally)
{
    const int N = P.n_traffic;
    const int L = 5 + 3 * N;
    const Vec2d pp = S.ppos[i];
    const PlayerAux pa = S.paux[i];
    const bool residual = (pa.steps & kResidualBit) != 0;
    const double dpsi = (double)action * P.dpsi_per_action;
    Player p;
    p.x = pp.x; p.y = pp.y;
    player_set_heading(P, p, wrap360(pa.psi + dpsi), dpsi);
    player_advance(P, p);
    const int k = pa.steps & kStepsMask;
    const int steps = k + 1;
    const PlayerView v = player_view(P, p, steps);
    float *row = out.obs + (int64_t)L * i;
    bool coll = false;
    float minsep = MINSEP ? S.min_sep[i] : 0.0f;
    Encounter e0;
    const double kd = (double)k;
    for (int j = 0; j < N; ++j) {
        const Intruder t = intruder_at(P, traffic_load(S, i * N + j, residual), kd);
        if (MINSEP) {
            const double ox = (t.x - t.dx) - p.x, oy = (t.y - t.dy) - p.y;
            minsep = fminf(minsep, acas_sqrtf((float)(ox * ox + oy * oy)));
        }
        const Encounter en = encounter(P, p, t);
        if (j == 0) e0 = en;
        coll |= en.d2 < P.coll_d2;
        row[5 + 3 * j + 0] = en.d * P.inv_d_sep_max;
        row[5 + 3 * j + 1] = en.d_cpa * P.inv_d_cpa_max;
        row[5 + 3 * j + 2] = en.v_c * P.vc_scale;
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) row[q] = v.obs[q];
    float r = shaped_reward(P, p, v, e0, steps);
    const bool goal = v.dg2 < P.goal_r2;
    const bool tout = steps > P.max_steps;
    // game.py:279-284 (Q9): both bonuses can fall on one step; summed first so that -1000 + 1000 does not
    // round the shaped reward to the float32 spacing at 1000
    r += (coll ? P.reward_collision : 0.0f) + (goal ? P.reward_goal : 0.0f);
    float ret = pa.ep_return + r;
    const int outcome = tout ? ACAS2D_OUTCOME_TIMEOUT : coll ? ACAS2D_OUTCOME_COLLISION
                        : goal ? ACAS2D_OUTCOME_GOAL : 0;
    const bool done = outcome != 0;
    out.reward[i] = r;
    out.done[i] = (uint8_t)done;
    if (out.flags) {
        const bool oob = p.x < 0.0 || p.x > P.width || p.y < 0.0 || p.y > P.height;
        out.flags[i] = (uint8_t)((coll ? ACAS2D_FLAG_COLLISION : 0) | (goal ? ACAS2D_FLAG_GOAL : 0) |
                                 (tout ? ACAS2D_FLAG_TIMEOUT : 0) | (done ? ACAS2D_FLAG_DONE : 0) |
                                 (oob ? ACAS2D_FLAG_OOB : 0));
    }
    int steps_out = steps | (pa.steps & ~kStepsMask);
    if (done) {
        if (out.outcome) out.outcome[i] = (uint8_t)outcome;
        if (out.ep_return) out.ep_return[i] = ret;
        if (out.ep_length) out.ep_length[i] = steps;
        tally_add(tally, outcome, steps, ret, minsep, MINSEP);
        if (P.auto_reset) {
            if (out.term_obs) {
                float *trow = out.term_obs + (int64_t)L * i;
                for (int q = 0; q < L; ++q) trow[q] = row[q];
            }
            const uint32_t episode = S.episode_idx[i];
            S.episode_idx[i] = episode + 1u;
            const uint64_t gid = S.gid0 + (uint64_t)i;
            const Spawn0 sp = spawn_slot0(P, S.seed, gid, episode);
            p.x = P.player_x0; p.y = P.player_y0;
            player_set_heading_straight(p, sp.player_psi);
            const PlayerView v1 = reset_view(P, p.psi);
#pragma unroll
            for (int q = 0; q < 5; ++q) row[q] = v1.obs[q];
            minsep = INFINITY;
            for (int j = 0; j < N; ++j) {
                const TrafficRec tr = spawn_traffic(P, S.seed, gid, episode, j, sp);
                traffic_store(S, i * N + j, tr, false);
                const Intruder it = intruder_at(P, tr, 0.0);
                kin_store(S, i * N + j, tr, it);
                const Encounter en = encounter(P, p, it);
                minsep = fminf(minsep, en.d);
                row[5 + 3 * j + 0] = en.d * P.inv_d_sep_max;
                row[5 + 3 * j + 1] = en.d_cpa * P.inv_d_cpa_max;
                row[5 + 3 * j + 2] = en.v_c * P.vc_scale;
            }
            steps_out = 1 | spawn_bits(P, S);
            ret = 0.0f;
            if (S.spawn_sep) S.spawn_sep[i] = minsep;
        }
    }
    Vec2d np; np.x = p.x; np.y = p.y;
    PlayerAux na; na.psi = p.psi; na.steps = steps_out; na.ep_return = ret;
    S.ppos[i] = np;
    S.paux[i] = na;
    if (MINSEP) S.min_sep[i] = minsep;
}

// ---------------------------------------------------------------- on-device episode records
// The reference appends to per-step Python lists inside action() and evaluate() (game.py:45-75, 231-239,
// 266-276; initial entries game.py:132-160) and its scripts dump them (testing_main.py:113-138).  Here a traced
// env gets one row of ACAS2D_TRACE_DOUBLES (+ 2 per recorded intruder) float64 values per step in a ring buffer
// in HBM.  trace_env runs BEFORE the step of the same actions and computes, without touching the state, the row
// that step is about to produce; a game at steps == 1 (fresh from reset / respawn) first gets its initial row.
struct TraceRing {
    int64_t first_env, num_envs;
    int32_t capacity, n_traffic_rec;
    int32_t *cursor;
    double *rows;
};

ACAS_HD void trace_row(const DevParams &P, const TraceRing &T, int64_t w, const Player &p, int steps, double a_lat,
                       double d_sep, const PlayerView &v, const Encounter &e0, float reward, int flags, bool initial,
                       const double *txy, int nrec)
{
    const int stride = ACAS2D_TRACE_DOUBLES + 2 * T.n_traffic_rec;
    const int c = T.cursor[w];
    T.cursor[w] = c + 1;
    double *r = T.rows + ((int64_t)w * T.capacity + (c % T.capacity)) * stride;
    const RewardTerms t = reward_terms(P, v, e0);
    r[0] = p.x; r[1] = p.y; r[2] = p.psi; r[3] = a_lat; r[4] = d_sep;
    r[5] = (double)v.d_goal; r[6] = (double)t.dh; r[7] = (double)(e0.v_c * P.fps); r[8] = (double)e0.d_cpa;
    r[9] = (double)v.d_dev; r[10] = (double)t.r_goal; r[11] = (double)t.r_head; r[12] = (double)t.r_cpa;
    r[13] = (double)t.r_dev;
    r[14] = initial ? (double)t.r5 : (double)(t.r5 * (1.0f - (float)steps * P.inv_max_steps));   // game.py:159 / 262-263
    r[15] = (double)steps; r[16] = (double)reward; r[17] = (double)flags;
    for (int j = 0; j < nrec; ++j) { r[ACAS2D_TRACE_DOUBLES + 2 * j] = txy[2 * j]; r[ACAS2D_TRACE_DOUBLES + 2 * j + 1] = txy[2 * j + 1]; }
}

ACAS_HD void trace_env(const DevParams &P, const StatePtrs &S, const TraceRing &T, int64_t w, const float *actions)
{
    const int64_t i = T.first_env + w;
    const int N = P.n_traffic;
    const int nrec = T.n_traffic_rec < N ? T.n_traffic_rec : N;
    const Vec2d pp = S.ppos[i];
    const PlayerAux pa = S.paux[i];
    const bool residual = (pa.steps & kResidualBit) != 0;
    const int k = pa.steps & kStepsMask;
    double txy[2 * ACAS2D_TRACE_MAX_TRAFFIC];
    Player p;
    if (k == 1) {
        // initial records of a new game (game.py:132-160): current state, a_lat = 0, no time discount
        p.x = pp.x; p.y = pp.y;
        player_set_heading_straight(p, pa.psi);
        const PlayerView v = player_view(P, p, k);
        Encounter e0;
        double sep2 = INFINITY;
        for (int j = 0; j < N; ++j) {
            const Intruder t = intruder_at(P, traffic_load(S, i * N + j, residual), 0.0);
            if (j < nrec) { txy[2 * j] = t.x; txy[2 * j + 1] = t.y; }
            const Encounter en = encounter(P, p, t);
            if (j == 0) e0 = en;
            sep2 = fmin(sep2, en.d2);
        }
        trace_row(P, T, w, p, k, 0.0, sqrt(sep2), v, e0, 0.0f, 0, true, txy, nrec);
    }
    const float action = actions[i];
    const double dpsi = (double)action * P.dpsi_per_action;
    p.x = pp.x; p.y = pp.y;
    player_set_heading(P, p, wrap360(pa.psi + dpsi), dpsi);
    player_advance(P, p);
    const int steps = k + 1;
    const PlayerView v = player_view(P, p, steps);
    Encounter e0;
    bool coll = false;
    double sep2 = INFINITY;
    for (int j = 0; j < N; ++j) {
        const Intruder t = intruder_at(P, traffic_load(S, i * N + j, residual), (double)k);
        const double ox = (t.x - t.dx) - p.x, oy = (t.y - t.dy) - p.y;       // Q10: records see the OLD traffic
        if (j < nrec) { txy[2 * j] = t.x - t.dx; txy[2 * j + 1] = t.y - t.dy; }
        sep2 = fmin(sep2, ox * ox + oy * oy);
        const Encounter en = encounter(P, p, t);
        if (j == 0) e0 = en;
        coll |= en.d2 < P.coll_d2;
    }
    const bool goal = v.dg2 < P.goal_r2, tout = steps > P.max_steps;
    float r = shaped_reward(P, p, v, e0, steps);
    r += (coll ? P.reward_collision : 0.0f) + (goal ? P.reward_goal : 0.0f);
    const int flags = (coll ? ACAS2D_FLAG_COLLISION : 0) | (goal ? ACAS2D_FLAG_GOAL : 0) | (tout ? ACAS2D_FLAG_TIMEOUT : 0) |
                      ((coll || goal || tout) ? ACAS2D_FLAG_DONE : 0);
    trace_row(P, T, w, p, steps, (double)action * P.acc_lat_limit, sqrt(sep2), v, e0, r, flags, false, txy, nrec);
}

// ---------------------------------------------------------------- reset / inject / extract
// ACAS2DEnv.reset(): new game (game.py:27-160) + observe (game.py:194-220).
ACAS_HD void reset_env(const DevParams &P, const StatePtrs &S, int64_t i, float *obs)
{
    const int N = P.n_traffic;
    const int L = 5 + 3 * N;
    const uint32_t episode = S.episode_idx[i];
    S.episode_idx[i] = episode + 1u;
    const uint64_t gid = S.gid0 + (uint64_t)i;
    const Spawn0 sp = spawn_slot0(P, S.seed, gid, episode);
    Player p;
    p.x = P.player_x0; p.y = P.player_y0;
    player_set_heading_straight(p, sp.player_psi);
    const PlayerView v = reset_view(P, p.psi);
    float *row = obs ? obs + (int64_t)L * i : nullptr;
    if (row) for (int q = 0; q < 5; ++q) row[q] = v.obs[q];
    float minsep = INFINITY;
    for (int j = 0; j < N; ++j) {
        const TrafficRec tr = spawn_traffic(P, S.seed, gid, episode, j, sp);
        traffic_store(S, i * N + j, tr, false);
        const Intruder it = intruder_at(P, tr, 0.0);
        kin_store(S, i * N + j, tr, it);
        const Encounter en = encounter(P, p, it);
        minsep = fminf(minsep, en.d);
        if (row) {
            row[5 + 3 * j + 0] = en.d * P.inv_d_sep_max;
            row[5 + 3 * j + 1] = en.d_cpa * P.inv_d_cpa_max;
            row[5 + 3 * j + 2] = en.v_c * P.vc_scale;
        }
    }
    Vec2d np; np.x = p.x; np.y = p.y;
    PlayerAux na; na.psi = p.psi; na.steps = 1 | spawn_bits(P, S); na.ep_return = 0.0f;
    if (compact_ok(P, S)) {                                               // N == 1: the compact form of intruder 0
        S.tpsi0[i] = (float)sp.psi;
        na.steps |= kCompactBit | (sp.y != P.t0_y_up ? kDownBit : 0);
    }
    S.ppos[i] = np;
    S.paux[i] = na;
    if (S.min_sep) S.min_sep[i] = minsep;
    if (S.spawn_sep) S.spawn_sep[i] = minsep;
}

// game.observe() WITHOUT the steps increment (game.py:199-220): the observation row of the state as it
// stands, with the player's last lateral acceleration taken as 0 (what it is right after a reset or an
// injection; the state does not keep the previous action).
ACAS_HD void observe_env(const DevParams &P, const StatePtrs &S, int64_t i, float *obs)
{
    const int N = P.n_traffic;
    const int L = 5 + 3 * N;
    const Vec2d pp = S.ppos[i];
    const PlayerAux pa = S.paux[i];
    const int st = pa.steps & kStepsMask;
    Player p;
    p.x = pp.x; p.y = pp.y;
    player_set_heading(P, p, pa.psi, 0.0);
    const PlayerView v = player_view(P, p, st);
    float *row = obs + (int64_t)L * i;
    for (int q = 0; q < 5; ++q) row[q] = v.obs[q];
    for (int j = 0; j < N; ++j) {
        const TrafficRec tr = traffic_load(S, i * N + j, (pa.steps & kResidualBit) != 0);
        const Encounter en = encounter(P, p, intruder_at(P, tr, (double)(st - 1)));
        row[5 + 3 * j + 0] = en.d * P.inv_d_sep_max;
        row[5 + 3 * j + 1] = en.d_cpa * P.inv_d_cpa_max;
        row[5 + 3 * j + 2] = en.v_c * P.vc_scale;
    }
}

ACAS_HD void inject_env(const DevParams &P, const StatePtrs &S, int64_t i, const double *player,
                        const double *traffic, const int32_t *steps, const double *total_reward)
{
    const int N = P.n_traffic;
    Vec2d np; np.x = player[3 * i]; np.y = player[3 * i + 1];
    const int st = steps[i] & kStepsMask;
    float minsep = INFINITY;
    const double back = (double)(st - 1);
    bool residual = false, plain = S.tkin != nullptr;
    for (int j = 0; j < N; ++j) {
        const int64_t ij = i * N + j;
        TrafficRec tr;
        const double x = traffic[4 * ij], y = traffic[4 * ij + 1];
        tr.v = traffic[4 * ij + 2]; tr.psi = traffic[4 * ij + 3];
        tr.x0 = x; tr.y0 = y;
        const Intruder it = intruder_at(P, tr, 0.0);                // displacement per step (dx, dy)
        tr.x0 = x - back * it.dx; tr.y0 = y - back * it.dy;         // closed-form origin (steps == 1)
        const bool need = traffic_store(S, ij, tr, true);
        kin_store(S, ij, tr, it);
        residual |= need;
        plain = plain && traffic_is_plain(P, tr, need);
        const double ox = x - np.x, oy = y - np.y;
        minsep = fminf(minsep, sqrtf((float)(ox * ox + oy * oy)));
    }
    PlayerAux na; na.psi = player[3 * i + 2]; na.ep_return = (float)total_reward[i];
    na.steps = st | (residual ? kResidualBit : 0) | ((plain && N > 1) ? kCompactBit : 0);
    S.ppos[i] = np;
    S.paux[i] = na;
    if (S.min_sep) S.min_sep[i] = minsep;
    if (S.spawn_sep) S.spawn_sep[i] = INFINITY;                    // not a spawn: no bound
}

ACAS_HD void extract_env(const DevParams &P, const StatePtrs &S, int64_t i, double *player, double *traffic,
                         int32_t *steps, double *total_reward)
{
    const int N = P.n_traffic;
    const Vec2d pp = S.ppos[i];
    const PlayerAux pa = S.paux[i];
    const int st = pa.steps & kStepsMask;
    if (player) { player[3 * i] = pp.x; player[3 * i + 1] = pp.y; player[3 * i + 2] = pa.psi; }
    if (steps) steps[i] = st;
    if (total_reward) total_reward[i] = (double)pa.ep_return;
    if (traffic) {
        for (int j = 0; j < N; ++j) {
            const int64_t ij = i * N + j;
            const TrafficRec tr = traffic_load(S, ij, (pa.steps & kResidualBit) != 0);
            const Intruder t = intruder_at(P, tr, (double)(st - 1));
            traffic[4 * ij + 0] = t.x;
            traffic[4 * ij + 1] = t.y;
            traffic[4 * ij + 2] = tr.v;
            traffic[4 * ij + 3] = tr.psi;
        }
    }
}

}  // namespace acas2d
