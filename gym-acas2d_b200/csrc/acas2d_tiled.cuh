// acas2d_tiled.cuh -- N_TRAFFIC > 1: the shared-memory tiled step kernel.
//
// G lanes cooperate on one env (G = 1, 2, ..., 32; G divides N), so a warp owns E = 32/G consecutive envs
// and every lane handles N/G intruders; a CTA is ONE warp (four when G == 32, see TiledShape).
//   0. one env per warp (N >= 256), lean launch: the game's minimum separation at its spawn (state->spawn_sep)
//      bounds its closest intruder's separation k steps later; when that proves a collision the tile is not
//      fetched and the per-intruder pass is skipped (a respawning env emits none of its results);
//   1. each warp stages its traffic tile -- contiguous in HBM -- into shared memory: ONE TMA bulk copy per warp
//      (cp.async.bulk + the warp's mbarrier) when rows are unpadded (G > 1), cp.async pieces otherwise.  KIN: the
//      tile is the 24-byte kinematic cache {x0, y0 float; dx, dy double} (acas2d_b200.h "tkin"), else the
//      16-byte {x0, y0, psi, v} records;
//   2. PLAYER PHASE, once per env instead of once per lane: with a scratch array (state->pstage) the float64
//      player update (aircraft.py:16-26), the look-ahead heading (kinematics.py:57-60) and the player-only
//      observation terms run as their own one-thread-per-env launch (player_phase_kernel: a ~400-instruction
//      dependent chain at full lane occupancy instead of G-fold redundantly inside every group -- and off the
//      tiled kernel's critical path, which is launched as its programmatic dependent); the lanes read the result
//      while their tile copy flies.  G == 1, or no scratch: each lane does it in registers.
//   3. each lane walks its intruders.  KIN: position = origin + k * cached displacement -- no sin / cos, no
//      float32 -> float64 conversion of a heading; 21 float64 operations, 3 MUFU and 4 conversions per intruder,
//      two intruders in flight (the round-1 loop was bound by the 16-lane conversion / MUFU pipe, see
//      acas2d_math.cuh).  Collision / minimum separation partials stay in registers, the three observation
//      entries go to the warp's observation tile.  One env per warp, lean launch: a cheap first pass (exact
//      cached separations, or a float32 estimate with a 0.05 px margin) may settle the collision and skip this;
//   4. any-collision and min-separation are reduced over the G lanes with xor shuffles;
//   5. lane 0 of each group finishes reward / flags / episode bookkeeping;
//   6. finished envs respawn cooperatively: all 32 lanes take the intruders of each respawning env in turn
//      (one Philox block per intruder, records float32 end to end), write its records and its reset observation;
//   7. the observation tile -- laid out with the HBM span's own phase within a 16-byte word -- is written back as
//      one contiguous span: one TMA bulk store per warp for the aligned body, <= 3 scalar floats at either end.
#pragma once

#include <cuda_pipeline.h>

#include "acas2d_dev.cuh"
#include "acas2d_env.cuh"

namespace acas2d {

// Warps per CTA.  The warps of a CTA share nothing but the launch, and a warp with a respawning env runs about
// twice as long as one without: single-warp CTAs give the SM its resources back warp by warp (measured, N = 64:
// 41.2 -> 37.4 us at 65 536 envs, 136.8 -> 125.4 us at 262 144; N = 8, 16, 32 gain 1-5 %).  With one env per warp
// (G == 32, N >= 256) every warp respawns and four-warp CTAs are better (211 vs 245 us).
template <int G> struct TiledShape { static constexpr int kWarps = (G == 32) ? 4 : 1; };

// What the player phase leaves for the lanes of an env.
struct alignas(16) PlayerStage {
    double x, y;            // new position
    double c, s;            // cos / sin of the new heading
    double cl, sl;          // ... of closing_speed's look-ahead heading
    double psi, dg2;        // new heading, squared goal distance
    float obs[5];           // player-only observation entries
    float d_goal, phi_deg, d_dev, ep_return, minsep;
    int32_t steps_word;     // paux.steps as read (flag bits included)
    float spawn_sep;        // the game's minimum separation at its spawn (+inf: unknown)
};
static_assert(sizeof(PlayerStage) == ACAS2D_PSTAGE_BYTES, "PlayerStage layout");
constexpr int kStageWords = sizeof(PlayerStage) / 16;

__host__ __device__ inline int tiled_row_stride(int N, int G, bool kin)
{
    // G == 1: every lane reads its own row, rows must start in different banks: 16-byte records need an odd
    // stride in records (N + 1 for even N); 24-byte records (read as 64-bit words) too
    if (G == 1) return N | 1;
    (void)kin;
    return N;
}

inline size_t tiled_smem_bytes(int N, int G, bool kin, int warps)
{
    const int E = 32 / G, L = 5 + 3 * N;
    const size_t tile = ((size_t)E * tiled_row_stride(N, G, kin) * (kin ? 24 : 16) + 15) & ~(size_t)15;
    const size_t otile = (((size_t)E * L * 4 + 15) & ~(size_t)15) + 16;      // + one 16-byte word: the span's phase (see 7.)
    return warps * (tile + otile);
}

// float64 player update + player-only observation terms of one env (game.py:222-229, 199-203)
__device__ __forceinline__ void player_phase(const DevParams &P, const StatePtrs &S, const float *__restrict__ actions,
                                             int64_t env, bool minsep, PlayerStage &o)
{
    const Vec2d pp = S.ppos[env];
    const PlayerAux pa = S.paux[env];
    const double dpsi = (double)actions[env] * P.dpsi_per_action;
    Player p;
    p.x = pp.x; p.y = pp.y;
    player_set_heading(P, p, wrap360(pa.psi + dpsi), dpsi);
    player_advance(P, p);
    const int steps = (pa.steps & kStepsMask) + 1;
    const PlayerView v = player_view(P, p, steps);
    o.x = p.x; o.y = p.y; o.c = p.c; o.s = p.s; o.cl = p.cl; o.sl = p.sl; o.psi = p.psi; o.dg2 = v.dg2;
#pragma unroll
    for (int q = 0; q < 5; ++q) o.obs[q] = v.obs[q];
    o.d_goal = v.d_goal; o.phi_deg = v.phi_deg; o.d_dev = v.d_dev;
    o.ep_return = pa.ep_return;
    o.minsep = minsep ? S.min_sep[env] : INFINITY;
    o.steps_word = pa.steps;
    o.spawn_sep = S.spawn_sep ? S.spawn_sep[env] : INFINITY;
}

// PlayerStage <-> seven 16-byte words, through registers (no address of the struct is taken)
__device__ __forceinline__ Float4 pack_dd(double a, double b)
{
    Float4 w;
    w.x = __int_as_float(__double2loint(a)); w.y = __int_as_float(__double2hiint(a));
    w.z = __int_as_float(__double2loint(b)); w.w = __int_as_float(__double2hiint(b));
    return w;
}
__device__ __forceinline__ void unpack_dd(const Float4 &w, double &a, double &b)
{
    a = __hiloint2double(__float_as_int(w.y), __float_as_int(w.x));
    b = __hiloint2double(__float_as_int(w.w), __float_as_int(w.z));
}

__device__ __forceinline__ void stage_store(Float4 *__restrict__ dst, int64_t B, int64_t env, const PlayerStage &o)
{
    Float4 w;
    dst[env] = pack_dd(o.x, o.y);
    dst[B + env] = pack_dd(o.c, o.s);
    dst[2 * B + env] = pack_dd(o.cl, o.sl);
    dst[3 * B + env] = pack_dd(o.psi, o.dg2);
    w.x = o.obs[0]; w.y = o.obs[1]; w.z = o.obs[2]; w.w = o.obs[3];
    dst[4 * B + env] = w;
    w.x = o.obs[4]; w.y = o.d_goal; w.z = o.phi_deg; w.w = o.d_dev;
    dst[5 * B + env] = w;
    w.x = o.ep_return; w.y = o.minsep; w.z = __int_as_float(o.steps_word); w.w = o.spawn_sep;
    dst[6 * B + env] = w;
}

__device__ __forceinline__ void stage_load(const Float4 *__restrict__ src, int64_t B, int64_t env, PlayerStage &o)
{
    const Float4 w0 = src[env], w1 = src[B + env], w2 = src[2 * B + env], w3 = src[3 * B + env];
    const Float4 w4 = src[4 * B + env], w5 = src[5 * B + env], w6 = src[6 * B + env];
    unpack_dd(w0, o.x, o.y); unpack_dd(w1, o.c, o.s); unpack_dd(w2, o.cl, o.sl); unpack_dd(w3, o.psi, o.dg2);
    o.obs[0] = w4.x; o.obs[1] = w4.y; o.obs[2] = w4.z; o.obs[3] = w4.w;
    o.obs[4] = w5.x; o.d_goal = w5.y; o.phi_deg = w5.z; o.d_dev = w5.w;
    o.ep_return = w6.x; o.minsep = w6.y; o.steps_word = __float_as_int(w6.z); o.spawn_sep = w6.w;
}

// The player pre-pass: one thread per env, result as seven 16-byte words per env, structure of arrays.
__global__ void __launch_bounds__(128)
player_phase_kernel(const DevParams P, const StatePtrs S, const float *__restrict__ actions)
{
    // programmatic dependent launch: the tiled kernel that follows may start now -- its tile copies do not depend
    // on this kernel; it waits (griddepcontrol.wait) before it reads the scratch written here
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int64_t env = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (env >= S.B) return;
    PlayerStage ps;
    player_phase(P, S, actions, env, S.min_sep != nullptr, ps);
    stage_store(S.pstage, S.B, env, ps);
}

#ifndef ACAS2D_TILED_MIN_BLOCKS
#define ACAS2D_TILED_MIN_BLOCKS 7        /* 16-byte records: 7 blocks/SM at 72 registers (measured best in round 1) */
#endif
#ifndef ACAS2D_TILED_MIN_BLOCKS_KIN
#define ACAS2D_TILED_MIN_BLOCKS_KIN 7        /* measured: 5 -> 133.3, 6 -> 125.5, 7 -> 123.9, 8 -> 124.5 us (N = 64, 262 144 envs) */
#endif

template <int G, bool MINSEP, bool KIN>
__global__ void __launch_bounds__(TiledShape<G>::kWarps * 32,
                                  (KIN ? ACAS2D_TILED_MIN_BLOCKS_KIN : ACAS2D_TILED_MIN_BLOCKS) * 4 / TiledShape<G>::kWarps)
step_tiled_kernel(const DevParams P, const StatePtrs S, const float *__restrict__ actions, const Sinks out,
                  const uint32_t magic_n)
{
    constexpr int E = 32 / G;
    constexpr int REC = KIN ? 24 : 16;
    constexpr int kTiledWarps = TiledShape<G>::kWarps;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = P.n_traffic;
    const int L = 5 + 3 * N;
    const int TS = tiled_row_stride(N, G, KIN);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tile_bytes = ((size_t)E * TS * REC + 15) & ~(size_t)15;
    const size_t warp_bytes = tile_bytes + (((size_t)E * L * 4 + 15) & ~(size_t)15) + 16;
    unsigned char *tile = smem_raw + warp * warp_bytes;
    __shared__ uint64_t tile_bar[kTiledWarps];                               // one mbarrier per warp (TMA tile copy)

    const int64_t env0 = ((int64_t)blockIdx.x * kTiledWarps + warp) * E;      // first env of this warp
    const int64_t left = S.B - env0;
    const int nvalid = left <= 0 ? 0 : (left < (int64_t)E ? (int)left : E);
    const int e = lane / G, sub = lane % G;
    const bool valid = e < nvalid;
    const int64_t env = env0 + (valid ? e : 0);
    const bool lead = valid && sub == 0;
    // the warp's observation rows are one contiguous span of HBM; the tile is laid out with the span's own phase
    // within a 16-byte word, so that everything but <= 3 floats at either end goes out as one TMA bulk store
    const int64_t span0 = env0 * L;
    float *otile = (float *)(tile + tile_bytes) + (int)(span0 & 3);

    // 0. one env per warp (N >= 256), lean launch: is this game's collision already certain?  Its minimum separation
    //    AT SPAWN (state->spawn_sep) plus the largest distance two aircraft can have closed in k steps bounds the
    //    separation of its closest intruder from above; below 2 * COLLISION_RADIUS the game ends now (game.py:185-189),
    //    respawns, and emits nothing that depends on the other intruders -- the tile is not even fetched.  With the
    //    reference's spawn rule (intruders uniform over the region the player starts in, game.py:109-110) that is
    //    ~99.7 % of the env-steps at N = 256 (203 -> 189 us).
    bool certain = false;
    if (G == 32 && !MINSEP && P.auto_reset && out.term_obs == nullptr && S.pstage && S.spawn_sep && nvalid > 0) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const Float4 w6 = S.pstage[6 * S.B + env0];
        const int k0 = __float_as_int(w6.z) & kStepsMask;
        certain = w6.w + (float)k0 * P.vrel_step < P.coll_sure;
    }

    // 1. stage the traffic tile
    if (nvalid > 0 && !certain) {
        if (G > 1) {                            // unpadded rows: one bulk copy of the warp's contiguous span
            if (lane == 0) {
                mbar_init(&tile_bar[warp], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                const unsigned bytes = (unsigned)(nvalid * N * REC);                    // N is even here: a multiple of 16
                const void *src = KIN ? (const void *)(S.tkin + env0 * N) : (const void *)(S.thot + env0 * N);
                mbar_expect_tx(smem_u32(&tile_bar[warp]), bytes);
                tma_load_1d_plain(smem_u32(tile), src, bytes, smem_u32(&tile_bar[warp]));
            }
        } else if (KIN) {                       // padded rows of 24-byte records: 8-byte pieces, three per record
            const unsigned char *src = (const unsigned char *)(S.tkin + env0 * N);
            const int total = nvalid * N * 3;
            for (int q = lane; q < total; q += 32) {
                const int rec = (int)__umulhi((unsigned)q, 0x55555556u);                // q / 3
                const int row = (int)__umulhi((unsigned)rec, magic_n);                  // rec / N
                __pipeline_memcpy_async(tile + (size_t)(q + 3 * row * (TS - N)) * 8, src + (size_t)q * 8, 8);
            }
            __pipeline_commit();
        } else {
            const Float4 *src = S.thot + env0 * N;
            Float4 *dst4 = (Float4 *)tile;
            const int total = nvalid * N;
            for (int idx = lane; idx < total; idx += 32)
                __pipeline_memcpy_async(dst4 + idx + (int)__umulhi((unsigned)idx, magic_n) * (TS - N), src + idx, 16);   // + row * padding
            __pipeline_commit();
        }
    }

    // 2. player phase: read the pre-pass result (G > 1 with scratch), or do it here in registers
    PlayerStage ps;
    if (valid) {
        if (G > 1 && S.pstage) {
            asm volatile("griddepcontrol.wait;" ::: "memory");      // the pre-pass (launched just before) has completed
            stage_load(S.pstage, S.B, env, ps);
        } else {
            player_phase(P, S, actions, env, MINSEP, ps);
        }
    } else {
        ps.x = ps.y = ps.c = ps.s = ps.cl = ps.sl = ps.psi = ps.dg2 = 0.0;
        ps.obs[0] = ps.obs[1] = ps.obs[2] = ps.obs[3] = ps.obs[4] = 0.0f;
        ps.d_goal = ps.phi_deg = ps.d_dev = ps.ep_return = 0.0f; ps.minsep = INFINITY;
        ps.steps_word = 1; ps.spawn_sep = INFINITY;
    }
    Tally tally;
    tally_clear(tally);
    if (nvalid > 0) {                                                       // warp-uniform
        const PlayerStage *my = &ps;
        Player p;
        p.x = my->x; p.y = my->y; p.c = my->c; p.s = my->s; p.cl = my->cl; p.sl = my->sl; p.psi = 0.0;
        const int steps_word = my->steps_word;
        const int k = steps_word & kStepsMask;
        const int steps = k + 1;
        const bool residual = (steps_word & kResidualBit) != 0;
        float minsep = INFINITY;

        if (G > 1) {
            __syncwarp();                                                   // the barrier was initialised by lane 0
            if (!certain) mbar_wait(smem_u32(&tile_bar[warp]), 0);
        } else {
            __pipeline_wait_prior(0);
            __syncwarp();
        }

        // 3. intruders of this lane: j = sub, sub + G, ... (rotated start for G > 1; G divides N)
        const int per_lane = N / G;
        const int j0 = (G == 1) ? 0 : lane % N;
        int j = j0;
        bool coll = false;
        Encounter e0;
        e0.d2 = 0.0; e0.d = 0.0f; e0.d_cpa = 0.0f; e0.v_c = 0.0f;
        float *orow = otile + e * L;
        const double kd = (double)k;
        auto visit = [&](const Intruder &t, int jj) {
            if (MINSEP) {                                                   // game.py:237 (Q10: old traffic)
                const double ox = (t.x - t.dx) - p.x, oy = (t.y - t.dy) - p.y;
                minsep = fminf(minsep, acas_sqrtf((float)(ox * ox + oy * oy)));
            }
            const Encounter en = encounter(P, p, t);
            if (jj == 0) e0 = en;
            coll |= en.d2 < P.coll_d2;
            orow[5 + 3 * jj + 0] = en.d * P.inv_d_sep_max;
            orow[5 + 3 * jj + 1] = en.d_cpa * P.inv_d_cpa_max;
            orow[5 + 3 * jj + 2] = en.v_c * P.vc_scale;
        };
        // the kinematic cache describes an env exactly only while it carries kCompactBit (spawned at AIRSPEED,
        // or injected so); anything else -- injected float64 states, other speed factors -- takes the records
        const bool fast = KIN && __all_sync(kFull, !valid || (steps_word & kCompactBit) != 0);

        // One env per warp (G == 32, N >= 256): the reference's spawn rule puts intruders on top of the player
        // (game.py:109-110), so nearly every env ends, by collision, on its first step -- and an env that ends and
        // respawns emits none of this step's per-intruder observation entries (SB3 semantics: the row is the reset
        // observation; the terminal row only goes to term_obs).  A cheap first pass settles "is there a collision
        // for sure": exact float64 separations from the cache, or a float32 estimate from the records with a 0.05 px
        // margin (MUFU sin / cos: < 3e-3 px off after 1000 steps).  If so the full pass is skipped; the reward's
        // intruder-0 terms are computed on their own.  Anything short of certain takes the full, exact pass.
        bool skip = certain;
        if (G == 32 && !MINSEP && P.auto_reset && out.term_obs == nullptr && !certain) {
            bool hit = false;
            int jp = j0;
            if (KIN && fast) {
                const TrafficKin *trow = (const TrafficKin *)tile;
                for (int m = 0; m < per_lane; ++m) {
                    const TrafficKin q = trow[jp];
                    const double rx = fma(kd, q.dx, (double)q.x0) - p.x, ry = fma(kd, q.dy, (double)q.y0) - p.y;
                    hit |= fma(rx, rx, ry * ry) < P.coll_d2;
                    jp += G;
                    if (jp >= N) jp -= N;
                }
            } else if (!KIN) {
                const Float4 *trow = (const Float4 *)tile;
                const float pxf = (float)p.x, pyf = (float)p.y, kf = (float)k * P.dt_f;
                for (int m = 0; m < per_lane; ++m) {
                    const Float4 h = trow[jp];
                    float sn, cs;
                    __sincosf(h.z * 0.017453292f, &sn, &cs);
                    const float reach = kf * h.w;
                    const float rx = fmaf(reach, cs, h.x) - pxf, ry = fmaf(reach, sn, h.y) - pyf;
                    hit |= fmaf(rx, rx, ry * ry) < P.coll_sure_d2;
                    jp += G;
                    if (jp >= N) jp -= N;
                }
            }
            skip = __any_sync(kFull, hit);
        }
        if (skip) {
            coll = true;
            if (lane == 0) {                                                // j0 == 0 for lane 0: intruder 0 (Q7)
                if (KIN && fast && !certain) e0 = encounter(P, p, intruder_from_kin(((const TrafficKin *)tile)[0], kd));
                else e0 = encounter(P, p, intruder_at(P, traffic_load(S, env * N, residual), kd));
            }
        } else if (KIN && fast) {
            const TrafficKin *trow = (const TrafficKin *)tile + e * TS;
            int m = 0;
            for (; m + 2 <= per_lane; m += 2) {                              // two independent encounters in flight per lane
                int j1 = j + G;
                if (G > 1 && j1 >= N) j1 -= N;
                const TrafficKin qa = trow[j], qb = trow[j1];
                const Intruder ta = intruder_from_kin(qa, kd), tb = intruder_from_kin(qb, kd);
                visit(ta, j);
                visit(tb, j1);
                j = j1 + G;
                if (G > 1 && j >= N) j -= N;
            }
            if (m < per_lane) visit(intruder_from_kin(trow[j], kd), j);
        } else if (KIN) {
            for (int m = 0; m < per_lane; ++m) {
                if (valid) visit(intruder_at(P, traffic_load(S, env * N + j, residual), kd), j);
                j += G;
                if (G > 1 && j >= N) j -= N;
            }
        } else {
            const Float4 *trow = (const Float4 *)tile + e * TS;
            const bool any_residual = __any_sync(kFull, residual);          // injected float64 states only: keep it a branch
            for (int m = 0; m < per_lane; ++m) {
                const Float4 h = trow[j];
                TrafficRec tr;
                tr.x0 = (double)h.x; tr.y0 = (double)h.y; tr.psi = (double)h.z; tr.v = (double)h.w;
                if (any_residual) {
                    if (residual) {
                        const Residual r = S.tres[env * N + j];
                        tr.x0 += r.x0; tr.y0 += r.y0; tr.psi += r.psi; tr.v += r.v;
                    }
                }
                visit(intruder_at(P, tr, kd), j);
                j += G;
                if (G > 1 && j >= N) j -= N;
            }
        }

        // 4. reductions over the G lanes of the env
        if (!(G == 32 && skip)) {                                           // (a settled collision needs no vote)
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                coll |= (bool)__shfl_xor_sync(kFull, (int)coll, o);
                if (MINSEP) minsep = fminf(minsep, __shfl_xor_sync(kFull, minsep, o));
            }
        }

        // 5. reward, flags, bookkeeping (lane `sub == 0` owns intruder 0, Q7)
        PlayerView v;
#pragma unroll
        for (int q = 0; q < 5; ++q) v.obs[q] = my->obs[q];
        v.dg2 = my->dg2; v.d_goal = my->d_goal; v.phi_deg = my->phi_deg; v.d_dev = my->d_dev;
        p.psi = my->psi;
        if (MINSEP) minsep = fminf(minsep, my->minsep);
        float r = shaped_reward(P, p, v, e0, steps);
        const bool goal = v.dg2 < P.goal_r2;
        const bool tout = steps > P.max_steps;
        // game.py:279-284 (Q9): both bonuses can fall on one step; summed first so that -1000 + 1000 does not
        // round the shaped reward to the float32 spacing at 1000
        r += (coll ? P.reward_collision : 0.0f) + (goal ? P.reward_goal : 0.0f);
        float ret = my->ep_return + r;
        const int outcome = tout ? ACAS2D_OUTCOME_TIMEOUT : coll ? ACAS2D_OUTCOME_COLLISION
                            : goal ? ACAS2D_OUTCOME_GOAL : 0;
        const bool done = outcome != 0;
        int steps_out = steps | (steps_word & ~kStepsMask);
        if (lead) {
#pragma unroll
            for (int q = 0; q < 5; ++q) orow[q] = v.obs[q];
            out.reward[env] = r;
            out.done[env] = (uint8_t)done;
            if (out.flags) {
                const bool oob = p.x < 0.0 || p.x > P.width || p.y < 0.0 || p.y > P.height;
                out.flags[env] = (uint8_t)((coll ? ACAS2D_FLAG_COLLISION : 0) | (goal ? ACAS2D_FLAG_GOAL : 0) |
                                           (tout ? ACAS2D_FLAG_TIMEOUT : 0) | (done ? ACAS2D_FLAG_DONE : 0) |
                                           (oob ? ACAS2D_FLAG_OOB : 0));
            }
            if (done) {
                if (out.outcome) out.outcome[env] = (uint8_t)outcome;
                if (out.ep_return) out.ep_return[env] = ret;
                if (out.ep_length) out.ep_length[env] = steps;
                tally_add(tally, outcome, steps, ret, minsep, MINSEP);
            }
        }

        // 6. auto-reset of the finished envs of this warp
        const bool respawn = valid && done && P.auto_reset;
        const unsigned respawn_mask = __ballot_sync(kFull, respawn);
        if (respawn_mask) {
            __syncwarp();
            if (out.term_obs) {
                for (int row = 0; row < nvalid; ++row) {
                    if (!((respawn_mask >> (row * G)) & 1u)) continue;          // warp-uniform
                    float *dst = out.term_obs + (env0 + row) * L;
                    for (int c = lane; c < L; c += 32) dst[c] = otile[row * L + c];
                }
                __syncwarp();
            }
            // A respawn is a few hundred instructions per intruder; with G lanes per env a lone respawning env
            // would run it at G/32 utilisation.  All 32 lanes take the intruders of each respawning env of the
            // warp in turn instead (warp-uniform loop over the rows).
            for (int row = 0; row < nvalid; ++row) {
                if (!((respawn_mask >> (row * G)) & 1u)) continue;
                const int64_t renv = env0 + row;
                uint32_t episode = 0;
                if (lane == 0) { episode = S.episode_idx[renv]; S.episode_idx[renv] = episode + 1u; }
                episode = __shfl_sync(kFull, episode, 0);
                const uint64_t gid = S.gid0 + (uint64_t)renv;
                const Spawn0 sp = spawn_slot0(P, S.seed, gid, episode);
                Player rp;
                rp.x = P.player_x0; rp.y = P.player_y0;
                player_set_heading_straight(rp, sp.player_psi);
                float ms = INFINITY;
                float *rrow = otile + row * L;
                // one env per warp: eight intruders per lane, two in flight (211 -> 203 us at N = 256); the shorter
                // loops of the smaller groups lose from it (N = 64: 37.4 -> 38.9 us)
#pragma unroll(G == 32 ? 2 : 1)
                for (int jj = lane; jj < N; jj += 32) {
                    const Encounter en = spawn_intruder(P, S, gid, episode, jj, sp, rp, renv * N + jj);
                    ms = fminf(ms, en.d);
                    rrow[5 + 3 * jj + 0] = en.d * P.inv_d_sep_max;
                    rrow[5 + 3 * jj + 1] = en.d_cpa * P.inv_d_cpa_max;
                    rrow[5 + 3 * jj + 2] = en.v_c * P.vc_scale;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ms = fminf(ms, __shfl_xor_sync(kFull, ms, o));
                if (lane == 0) {
                    const PlayerView v1 = reset_view(P, rp.psi);
#pragma unroll
                    for (int q = 0; q < 5; ++q) rrow[q] = v1.obs[q];
                }
                if (lane == 0 && S.spawn_sep) S.spawn_sep[renv] = ms;
                if (e == row) { p.x = rp.x; p.y = rp.y; p.psi = rp.psi; minsep = ms; steps_out = 1 | spawn_bits(P, S); ret = 0.0f; }
            }
        }
        __syncwarp();

        // 7. coalesced write-back of the span: scalar stores up to the first 16-byte boundary, one TMA bulk store for
        //    the aligned body, scalar stores for the rest
        {
            const int span = nvalid * L;
            int head = (4 - (int)(span0 & 3)) & 3;
            if (head > span) head = span;
            const int body = (span - head) & ~3, tail = span - head - body;
            float *dst = out.obs + span0;
            fence_proxy_async_shared();                                     // the lanes' shared writes, for the bulk engine
            __syncwarp();
            if (lane == 0 && body > 0) tma_store_1d_and_wait(dst + head, smem_u32(otile + head), (unsigned)body * 4u);
            if (lane >= 8 && lane < 8 + head) __stcs(dst + (lane - 8), otile[lane - 8]);
            if (lane >= 16 && lane < 16 + tail) __stcs(dst + head + body + (lane - 16), otile[head + body + (lane - 16)]);
        }

        if (lead) {
            Vec2d np; np.x = p.x; np.y = p.y;
            PlayerAux na; na.psi = p.psi; na.steps = steps_out; na.ep_return = ret;
            S.ppos[env] = np;
            S.paux[env] = na;
            if (MINSEP) S.min_sep[env] = minsep;
        }
    }
    tally_flush_warp(S.stats, tally);
}

}  // namespace acas2d
