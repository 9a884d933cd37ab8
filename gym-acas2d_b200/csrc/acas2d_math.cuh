// acas2d_math.cuh -- per-environment arithmetic of the ACAS-2D step, written once and
// inlined into every kernel of acas2d_kernels.cu.
//
// Precision plan (DESIGN.md "Numerics"):
//   * the FLAG CHAIN is float64: heading, sin/cos of the heading, player position,
//     intruder position, squared separations and the goal / collision / timeout tests.
//     The reference is float64 Python (aircraft.py:16-26, game.py:182-192); keeping this
//     chain in float64 keeps collision / goal / done flags bit-exact against it.
//   * sign-deciding products (closing-speed dot product, relative-velocity cross product
//     and its x component, Q4/Q12) are formed in float64 from float64 differences, then
//     rounded once to float32.
//   * everything that is only ever emitted as a float32 observation or reward (square
//     roots, atan2, normalisation, reward shaping) is float32.
//
// The functions are __host__ __device__ so tests/hostcheck can compile the very same
// source with g++ and compare it with the CPU oracle without a GPU.  That host build is
// test infrastructure; the product never runs it.
//
// Reference citations: aircraft.py / kinematics.py / rewards.py / game.py =
// gym_ACAS2D/envs/<file>; settings.py = gym_ACAS2D/settings.py.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ACAS_HD __host__ __device__ __forceinline__
#else
#define ACAS_HD static inline
#endif

namespace acas2d {

struct alignas(16) Vec2d { double x, y; };

// 16-byte per-env record: heading [deg], game.steps, game.total_reward so far.
struct alignas(16) PlayerAux { double psi; int32_t steps; float ep_return; };

// Launch-time constants, derived on the host from acas2d_params (see make_dev_params).
struct DevParams {
    // ---- float64 flag chain
    double v_dt;            // AIRSPEED / FPS: player displacement per step (aircraft.py:25-26)
    double dpsi_per_action; // ACC_LAT_LIMIT / AIRSPEED: heading change [deg] per unit action (Q1)
    double acc_lat_limit;   // game.py:225 (episode records only)
    double lookahead_rad;   // (1/FPS) * pi/180: closing_speed's look-ahead turn per degree of dpsi (Q2)
    double goal_x, goal_y;  // game.py:80-81
    double coll_d2;         // (2*COLLISION_RADIUS)^2, game.py:187
    double goal_r2;         // GOAL_RADIUS^2, game.py:192
    double width, height;   // aircraft.py:28-29
    double inv_360;         // obs[1] = psi/360, game.py:200
    double inv_d_dev_max;   // 1/d_dev_max, game.py:201
    // ---- spawn (game.py:85-116)
    double dt, airspeed;
    double player_x0, player_y0, player_psi_base, player_heading_lim;
    double traffic_heading_lim, factor_min, factor_span;
    double t0_x, t0_y_up, t0_y_span;   // intruder 0: (WIDTH-CR, CR + starts_down*(HEIGHT-2CR))
    double tn_x_span, tn_y_span;       // intruders n>0: U(0, WIDTH-SIZE) x U(0, 3*HEIGHT/5)
    // ---- float32 observation / reward block
    float inv_max_steps;    // obs[0], time discount (game.py:199,262)
    float inv_d_dev_max_f;  // obs[2] = d_dev/d_dev_max, game.py:201
    float inv_d_goal_max;   // obs[3], game.py:202
    float inv_d_sep_max;    // game.py:208
    float inv_d_cpa_max;    // game.py:209
    float vc_scale;         // FPS / v_closing_max: displacement-per-step units -> obs (game.py:210)
    float fps;
    float inv_safe_distance;   // rewards.py:16
    float rw_dev_max;          // 704 at defaults, rewards.py:22-23
    float inv_rw_dev_max;
    float inv_rw_goal_max;     // 1/3408 at defaults, rewards.py:46-47
    float reward_goal, reward_collision;
    float tn_x_span_f, tn_y_span_f, factor_min_f, factor_span_f, airspeed_f;   // float32 spawn of intruders n>0
    // ---- integers
    int32_t n_traffic, max_steps, auto_reset;
    int32_t q3_trivial;     // every spawned intruder flies at exactly AIRSPEED (factor min == max == 1): Q3's term == dy
    float coll_sure_d2;     // (2*COLLISION_RADIUS - 0.05)^2: a float32 separation estimate below this IS a collision
    float dt_f;
    float reset_obs0, reset_obs2, reset_obs3, reset_obs4;   // player-only observation entries of a NEW game (all but the heading)
    double c_x0, c_y0_up, c_y0_down, c_v;   // intruder 0's spawn pattern as a spawn stores it: rounded to float32 (compact form)
    float vrel_step;        // upper bound of the player-intruder relative displacement per step (spawned speeds), rounded up
    float coll_sure;        // 2*COLLISION_RADIUS - 1e-3: a separation bound below this proves a collision
};

// forward: used by sincos_deg below
static constexpr double kDeg2Rad = 0.017453292519943295;  // pi/180
static constexpr float kTwoPiF = 6.283185307179586f;
static constexpr float kInvTwoPiF = 0.15915494309189535f;

// float32 special functions on the MUFU unit (one instruction each, <= 2 ulp).  Outputs of the
// observation / reward block are float32 with stated tolerances >= 5e-7, so IEEE-exact sqrtf /
// division (8-10 instructions plus a slow-path branch each) buy nothing here.
ACAS_HD float acas_rsqrtf(float x)
{
#if defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}

ACAS_HD float acas_sqrtf(float x)
{
#if defined(__CUDA_ARCH__)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}

ACAS_HD float acas_rcpf(float x)
{
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}

// atan2(y, x) mod 2*pi expressed in TURNS, [0, 1] (kinematics.py:16-22 divided by 360).
// Octant reduction + degree-7 minimax polynomial in t^2 for atan(t)/(2 pi), t in [0,1]:
// max error 3.2e-8 turns evaluated in float32 (fit: near-minimax Chebyshev, checked on 2e5 points).
// A tiny negative y gives 1.0 exactly like the reference's (atan2 % 2pi) -> 360 degrees.
ACAS_HD float atan2_turns(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = mn * acas_rcpf(mx);
    t = (mx > 0.0f) ? t : 0.0f;                       // atan2(0, 0) == 0
    const float z = t * t;
    float p = -0.0007430583355017006f;
    p = fmaf(p, z, 0.0038461685180664062f);
    p = fmaf(p, z, -0.009448567405343056f);
    p = fmaf(p, z, 0.015766043215990067f);
    p = fmaf(p, z, -0.022308088839054108f);
    p = fmaf(p, z, 0.03178202360868454f);
    p = fmaf(p, z, -0.05304946005344391f);
    p = fmaf(p, z, 0.15915492177009583f);
    float a = p * t;                                  // [0, 1/8]
    a = (ay > ax) ? 0.25f - a : a;
    a = (x < 0.0f) ? 0.5f - a : a;
    a = (y < 0.0f) ? 1.0f - a : a;
    return a;
}

// v with its sign flipped when bit 0 of `flip` is set (integer op on the high word).
ACAS_HD double flip_sign(double v, int flip)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (flip << 31), __double2loint(v));
#else
    return (flip & 1) ? -v : v;
#endif
}

ACAS_HD int acas_rint_i(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2int_rn(x);
#else
    return (int)nearbyint(x);
#endif
}

// Conversions that touch float64 (F2F, F2I, I2F with a 64-bit side) and MUFU share one 16-lane/clk/SM pipe on
// sm_100 (measured, tools/ubench_pipes.cu: DFMA 64, F2F.F64.F32 / I2F.F64 / F2I.F64 / MUFU 16 thread-ops/clk/SM):
// a warp-wide conversion occupies it for 8 cycles -- as long as four DFMAs.  The per-intruder loops are bound by
// that pipe, so the conversions that can be done exactly (or within one float32 ulp) on the integer / FP64
// pipes are done there.

// rint(x) as a double plus the integer in the low word, |x| < 2^51: one DADD instead of F2I.F64 + I2F.F64.
ACAS_HD double acas_rint_magic(double x, int *q)
{
#if defined(__CUDA_ARCH__)
    const double t = x + 6755399441055744.0;            // 2^52 + 2^51: the integer lands in the low mantissa bits
    *q = __double2loint(t);
    return t - 6755399441055744.0;
#else
    const double r = nearbyint(x);
    *q = (int)r;
    return r;
#endif
}

// float32 value of a NON-NEGATIVE double, truncated (result <= v, within one float32 ulp); values below
// 2^-126 give 0..7 denormal ulps (flushed by the .ftz consumers), v must be < 2^128.  Three integer
// instructions instead of one F2F.F32.F64.  Only for quantities that feed sqrt / rsqrt of an output.
#ifndef ACAS2D_ALU_D2F
#define ACAS2D_ALU_D2F 1
#endif
ACAS_HD float acas_d2f_pos(double v)
{
#if defined(__CUDA_ARCH__) && ACAS2D_ALU_D2F
    const int hi = __double2hiint(v) - 0x38000000;      // re-bias the exponent: 1023 -> 127
    const unsigned lo = (unsigned)__double2loint(v);
    return __uint_as_float(__funnelshift_l(lo, (unsigned)max(hi, 0), 3));
#else
    return (float)v;
#endif
}

// sin and cos of an angle given in DEGREES (the unit the reference keeps headings in,
// aircraft.py:22-23).  The argument is reduced in degrees, where the reduction psi - 90*q is
// exact, then one multiply by pi/180 lands in [-pi/4, pi/4] for the fdlibm kernel polynomials.
// |error| < 2.5e-16 (tests/test_hostcheck.py checks 1.5e-16 against 50-digit values), i.e. as close to the
// true value as the reference's own cos((psi/360)*2*pi), whose argument already carries a 4e-16 rounding
// error.  No slow path, no local memory.
// (Tried in round 2: a 1024-entry {sin, cos} table + rotation, 17 float64 operations instead of ~30 plus the
//  constant moves -- fewer instructions, but the 32-address gather per warp sits in the middle of the
//  dependency chain and competes with the streaming traffic for L1: every kernel got 2-20 % slower.)
ACAS_HD void sincos_deg(double deg, double *s, double *c)
{
    // (the coefficients stay 64-bit immediates: moved to the constant bank, ptxas copies them into vector
    //  registers every iteration -- as many instructions, plus a spill in the tiled kernel)
    int q;
    const double qd = acas_rint_magic(deg * (1.0 / 90.0), &q);
    const double r = fma(qd, -90.0, deg);                   // exact
    const double t = r * kDeg2Rad;
    const double z = t * t;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sn = fma(t * z, ps, t);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cs = 1.0 - (0.5 * z - z * z * pc);
    const bool odd = (q & 1) != 0;
    const double a = odd ? cs : sn, b = odd ? sn : cs;
    *s = flip_sign(a, (q >> 1) & 1);
    *c = flip_sign(b, ((q + 1) >> 1) & 1);
}

// Python float `%` with divisor 360 (aircraft.py:22, kinematics.py:58,68, game.py:92,106):
// result in [0, 360], equal to 360.0 only when a tiny negative value rounds up.
ACAS_HD double wrap360_slow(double t)
{
    double r = fmod(t, 360.0);
    if (r != 0.0) { if (r < 0.0) r += 360.0; } else { r = 0.0; }
    return r;
}

ACAS_HD double wrap360(double t)
{
    const bool hi = t >= 360.0, lo = t < 0.0;
    double r = hi ? t - 360.0 : (lo ? t + 360.0 : t);     // exact where fmod is (|t| < 720 resp. >= -360)
    const bool ok = hi ? (r < 360.0) : (lo ? (t >= -360.0) : true);
    if (!ok) r = wrap360_slow(t);                          // |heading step| > 360 deg: unclipped actions only
    return r;
}

// ------------------------------------------------------------------ Philox4x32-10
struct U4 { uint32_t x, y, z, w; };

ACAS_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

ACAS_HD double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

// Action stream of the synthetic rollout: a ~ U(-1,1), float32, one Philox block per
// (global env id, step index); word 0 is used.
ACAS_HD float random_action(uint64_t action_seed, uint64_t gid, uint64_t step_index)
{
    U4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step_index,
                         (uint32_t)(step_index >> 32) ^ 0xAC7105EDu,
                         (uint32_t)action_seed, (uint32_t)(action_seed >> 32));
    return (float)(2.0 * u01(r.x) - 1.0);
}

// ------------------------------------------------------------------ state in registers
struct Player {
    double x, y, psi;   // position [px], heading [deg]
    double c, s;        // cos / sin of the heading
    double cl, sl;      // cos / sin of closing_speed's look-ahead heading (Q2)
};

struct Intruder {
    double x, y;        // CURRENT position
    double dx, dy;      // displacement per step (v cos psi dt, v sin psi dt)
    double dyq;         // Q3: the reference's closing-speed look-ahead multiplies the INTRUDER's sine by the PLAYER's speed
                        // (kinematics.py:71-73): airspeed * sin(psi) * dt; equals dy when the speeds are equal
};

// aircraft.py:16-26 for the player: heading += a_lat/v degrees (Q1), wrap, move one step.
// dpsi is the heading increment in degrees.  Also prepares the look-ahead heading of
// kinematics.py:57-60 (psi + dpsi*dt), a tiny rotation of the new heading.
ACAS_HD void player_set_heading(const DevParams &P, Player &p, double psi, double dpsi)
{
    p.psi = psi;
    sincos_deg(psi, &p.s, &p.c);
    const double d = dpsi * P.lookahead_rad;             // radians, |d| <= 1.7e-4 for |action| <= 1
    double sd, cd;
    if (fabs(d) < 0.0078125) {                           // Taylor: error < 4e-16
        const double d2 = d * d;
        sd = d * (1.0 + d2 * (-1.0 / 6.0 + d2 * (1.0 / 120.0)));
        cd = 1.0 + d2 * (-0.5 + d2 * (1.0 / 24.0 + d2 * (-1.0 / 720.0)));
    } else {                                             // unclipped actions (Q19)
        sincos_deg(dpsi * P.dt, &sd, &cd);
    }
    p.cl = p.c * cd - p.s * sd;
    p.sl = p.s * cd + p.c * sd;
}

// The same for a new game: a_lat = 0, so the look-ahead heading IS the heading (game.py:88-92, kinematics.py:57-60).
ACAS_HD void player_set_heading_straight(Player &p, double psi)
{
    p.psi = psi;
    sincos_deg(psi, &p.s, &p.c);
    p.cl = p.c;
    p.sl = p.s;
}

ACAS_HD void player_advance(const DevParams &P, Player &p)
{
    p.x = p.x + p.c * P.v_dt;
    p.y = p.y + p.s * P.v_dt;
}

// Per-intruder quantities of game.py:205-210: separation, signed distance of closest
// approach (kinematics.py:40-49) and closing speed (kinematics.py:52-79).
struct Encounter {
    double d2;      // squared separation (float64: collision test)
    float d;        // separation
    float d_cpa;    // signed, px
    float v_c;      // px per STEP * FPS is applied by the caller through vc_scale / fps
};

ACAS_HD Encounter encounter(const DevParams &P, const Player &p, const Intruder &t)
{
    Encounter e;
    const double rx = t.x - p.x, ry = t.y - p.y;          // bearing vector player -> intruder
    e.d2 = rx * rx + ry * ry;
    // relative displacement per step at the CURRENT headings (kinematics.py:25-37, times dt)
    const double wx = p.c * P.v_dt - t.dx, wy = p.s * P.v_dt - t.dy;
    const double w2 = wx * wx + wy * wy;
    // d*sin(a_rel - arctan(wy/wx)) == sign(wx) * (ry*wx - rx*wy) / |w|   (Q12 keeps cos(h) >= 0)
    const double cross = ry * wx - rx * wy;
    // one-step look-ahead (kinematics.py:57-77): p' - t' and u' - u_t' with Q3 on the y term
    const double ex = p.cl * P.v_dt - t.dx;
    const double qx = ex - rx;
    const double qy = (p.sl * P.v_dt - t.dy) - ry;
    const double ey = p.sl * P.v_dt - t.dyq;
    const double q2 = qx * qx + qy * qy;
    const double dot = ex * qx + ey * qy;

    e.d = acas_sqrtf(acas_d2f_pos(e.d2));
    const float cr = (float)cross * acas_rsqrtf(acas_d2f_pos(w2));
#if defined(__CUDA_ARCH__)
    // sign(wx) * cr on the integer pipe (wx is a float64 difference: an exact zero is +0, so "wx < 0" is its sign bit)
    e.d_cpa = __int_as_float(__float_as_int(cr) ^ (__double2hiint(wx) & (int)0x80000000));
#else
    e.d_cpa = (wx < 0.0) ? -cr : cr;
#endif
    e.v_c = (float)dot * acas_rsqrtf(acas_d2f_pos(q2));   // displacement units; * FPS = px/s (Q4)
    return e;
}

// Player-only observation terms and the shaped reward (game.py:199-203,249-263,
// rewards.py:5-60).  `steps` is the already incremented counter (Q5, Q6).
struct PlayerView {
    float obs[5];
    double dg2;     // squared goal distance (float64: goal test)
    float d_goal, phi_deg, d_dev;
};

ACAS_HD PlayerView player_view(const DevParams &P, const Player &p, int32_t steps)
{
    PlayerView v;
    const double gx = P.goal_x - p.x, gy = P.goal_y - p.y;
    v.dg2 = gx * gx + gy * gy;
    v.d_goal = acas_sqrtf(acas_d2f_pos(v.dg2));
    // heading_to_goal (game.py:171-173, kinematics.py:16-22): degrees(atan2 mod 2pi) = 360 * turns
    const float gyf = (float)gy;
    const float phi_turns = atan2_turns(gyf, (float)gx);
    v.phi_deg = phi_turns * 360.0f;
    // plan_deviation (game.py:175-180) = d_goal*sin(heading_to_goal) == goal_y - y
    v.d_dev = gyf;
    v.obs[0] = (float)steps * P.inv_max_steps;
    v.obs[1] = (float)(p.psi * P.inv_360);
    v.obs[2] = gyf * P.inv_d_dev_max_f;
    v.obs[3] = v.d_goal * P.inv_d_goal_max;
    v.obs[4] = phi_turns;
    return v;
}

// Player-only observation entries of a freshly spawned game (game.py:199-203 at steps == 1): the player starts at a
// fixed point (game.py:85-86), so everything but obs[1] = psi / 360 is a constant of the parameter set
// (make_dev_params evaluates player_view there once) -- a respawn does not redo the goal distance / bearing math.
ACAS_HD PlayerView reset_view(const DevParams &P, double psi)
{
    PlayerView v;
    v.obs[0] = P.reset_obs0; v.obs[1] = (float)(psi * P.inv_360); v.obs[2] = P.reset_obs2;
    v.obs[3] = P.reset_obs3; v.obs[4] = P.reset_obs4;
    v.dg2 = 0.0; v.d_goal = 0.0f; v.phi_deg = 0.0f; v.d_dev = 0.0f;     // not used by the reset observation
    return v;
}

// The terms of step_reward_5 (rewards.py:5-60) with intruder 0 only (Q7).
struct RewardTerms { float dh, r_head, r_cpa, r_dev, r_goal, r5; };

ACAS_HD RewardTerms reward_terms(const DevParams &P, const PlayerView &v, const Encounter &e0)
{
    RewardTerms t;
    // delta_heading (kinematics.py:82-83) from the float32 heading / bearing (error < 5e-5 deg)
    const float a = fabsf(v.obs[1] * 360.0f - v.phi_deg);
    t.dh = fminf(a, 360.0f - a);
    float h = 1.0f - t.dh * (1.0f / 180.0f);
    h = h * h; h = h * h;                                    // rewards.py:7
    // v_closing <= 0 branch (rewards.py:55-57)
    float c = e0.d_cpa * P.inv_safe_distance;                // rewards.py:16
    c = c * c; c = c * c;
    c = fminf(1.0f, c);                                      // min(1, nan) == 1 in Python, fminf agrees
    const float ad = fabsf(v.d_dev);                         // rewards.py:21-27: 0 beyond d_dev_max
    const float dv = acas_sqrtf(fmaxf(0.0f, 1.0f - ad * P.inv_rw_dev_max));
    // else branch (rewards.py:59-60)
    float g = 1.0f - v.d_goal * P.inv_rw_goal_max;           // rewards.py:48
    g = g * g; g = g * g;
    g = fminf(1.0f, g);
    t.r_head = h; t.r_dev = dv; t.r_goal = g;
    t.r_cpa = (e0.v_c > 0.0f) ? 1.0f : c;                    // rewards.py:13: 1 when separating (Q4)
    t.r5 = (e0.v_c <= 0.0f) ? h * c * dv : h * g;            // rewards.py:54 (NaN -> else, as in Python)
    return t;
}

// step_reward_5 times the time discount (Q6).
ACAS_HD float shaped_reward(const DevParams &P, const Player &p, const PlayerView &v,
                            const Encounter &e0, int32_t steps)
{
    (void)p;
    return reward_terms(P, v, e0).r5 * (1.0f - (float)steps * P.inv_max_steps);      // game.py:262-263
}

// ------------------------------------------------------------------ spawn (game.py:85-116)
// Draw slot 0: player heading jitter, starts_down, intruder-0 speed factor, intruder-0 heading jitter.
struct Spawn0 { double player_psi; double x, y, v, psi; };

ACAS_HD Spawn0 spawn_slot0(const DevParams &P, uint64_t seed, uint64_t gid, uint32_t episode)
{
    const U4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), episode, 0u,
                               (uint32_t)seed, (uint32_t)(seed >> 32));
    Spawn0 o;
    const double hl = P.player_heading_lim, tl = P.traffic_heading_lim;
    o.player_psi = wrap360(P.player_psi_base + (-hl + (2.0 * hl) * u01(r.x)));    // game.py:91-92
    const double sd = (double)(r.y >> 31);                                          // game.py:98
    o.x = P.t0_x;                                                                    // game.py:100
    o.y = P.t0_y_up + sd * P.t0_y_span;                                              // game.py:101
    o.v = (P.factor_min + P.factor_span * u01(r.z)) * P.airspeed;                    // game.py:103
    o.psi = wrap360((145.0 + sd * 70.0) + (-tl + (2.0 * tl) * u01(r.w)));            // game.py:105-106
    return o;
}

// Draw slot i >= 1: intruder i position, speed factor, heading (game.py:109-114).  These values are
// stored as float32 (acas2d_b200.h "thot"), so they are drawn in float32 to begin with: u = 24 random
// bits * 2^-24, one float multiply each -- a respawn at large N_TRAFFIC is as hot as a step.
struct SpawnN { double x, y, v, psi; };

ACAS_HD float u01f(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

ACAS_HD SpawnN spawn_slot(const DevParams &P, uint64_t seed, uint64_t gid, uint32_t episode, uint32_t i)
{
    const U4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), episode, i,
                               (uint32_t)seed, (uint32_t)(seed >> 32));
    SpawnN o;
    o.x = (double)(P.tn_x_span_f * u01f(r.x));
    o.y = (double)(P.tn_y_span_f * u01f(r.y));
    o.v = (double)(fmaf(P.factor_span_f, u01f(r.z), P.factor_min_f) * P.airspeed_f);
    o.psi = (double)(360.0f * u01f(r.w));
    return o;
}

// Per-step displacement of an aircraft flying straight (aircraft.py:23-26 with a_lat = 0).
ACAS_HD void heading_to_velocity(const DevParams &P, double v, double psi, double *dx, double *dy)
{
    double s, c;
    sincos_deg(psi, &s, &c);
    *dx = (v * c) * P.dt;
    *dy = (v * s) * P.dt;
}

}  // namespace acas2d
