// acas2d_kernels.cu -- sm_100a kernels and the C ABI (include/acas2d_b200.h) of the
// batched ACAS-2D environment step.
//
// Kernels
//   step_n1_tma_kernel    N_TRAFFIC == 1 (the reference default), persistent: 256-env input tiles
//                         arrive through a TMA bulk-copy ring (cp.async.bulk + mbarrier), threads
//                         compute out of shared memory and stream results with 128-bit stores; compact
//                         4-byte intruder record, respawns deferred to a per-CTA queue, consecutive steps
//                         chained by programmatic dependent launch.
//   step_n1_kernel        same step, one thread per env with plain 128-bit loads (used when the
//                         per-episode minimum separation is tracked; bit-identical results).
//   step_tiled_kernel     (acas2d_tiled.cuh) N_TRAFFIC > 1: G lanes per env (G = 1..32), the warp's traffic
//                         tile -- 24-byte kinematic cache or 16-byte records -- staged in shared memory by one
//                         TMA bulk copy per warp, min-separation / any-collision reduced with warp shuffles,
//                         observation rows assembled in shared memory and written back by one TMA bulk store.
//   player_phase_kernel   (acas2d_tiled.cuh) the float64 player update of a step, one thread per env, for the
//                         tiled kernel's lanes to read.
//   step_loop_kernel      N_TRAFFIC > 1, one thread per env: the simple form the tiled kernel is checked against.
//   rollout_n1_kernel     K fused steps with in-kernel Philox actions (synthetic benchmark, ageing).
//   step_k_n1_kernel      K open-loop steps per launch with every step's outputs.
//   policy_step_n1_kernel / policy_step_n1_tc_kernel (acas2d_policy*.cuh)
//                         the reference agent's actor MLP fused with the env step: float32 on the CUDA cores,
//                         or tcgen05 TF32 MMAs with TMEM accumulators.
//   ppo_* kernels         (acas2d_ppo.cuh) critic forward, GAE, minibatch gradient, fused update.
//   trace_kernel          per-step episode records into ring buffers (acas2d_trace_step).
//   reset / observe / inject / extract / render / random_actions  small utility kernels.
//
// There is no CPU fallback in this file: every entry point launches on the device.
#include <cuda_runtime.h>
#include <cuda_pipeline.h>

#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/acas2d_b200.h"
#include "acas2d_env.cuh"
#include "acas2d_dev.cuh"
#include "acas2d_tiled.cuh"
#include "acas2d_policy_tc.cuh"
#include "acas2d_ppo.cuh"

namespace {

using namespace acas2d;

std::atomic<int64_t> g_launches{0};

constexpr int kBlock = 256;

// SM count of the CURRENT device (cached per device, not per process: one process may drive several GPUs)
int sm_count()
{
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (sms[dev & 63] == 0) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev & 63];
}

int check_args(const acas2d_params *p, const acas2d_state *s)
{
    if (!p || !s) return ACAS2D_E_NULL;
    if (p->n_traffic < 1 || p->n_traffic > ACAS2D_MAX_TRAFFIC) return ACAS2D_E_BAD_TRAFFIC;
    if (s->num_envs < 0 || s->num_envs * (int64_t)(5 + 3 * p->n_traffic) > (int64_t)1 << 40) return ACAS2D_E_BAD_SIZE;
    if (s->num_envs > 0 && (!s->ppos || !s->paux || !s->thot || !s->tres || !s->episode_idx))
        return ACAS2D_E_NULL;                       // an empty batch has nothing to point at
    return 0;
}

int finish_launch(int kernels = 1)
{
    g_launches.fetch_add(kernels, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// Tuning knobs for experiments (read once): ACAS2D_N1_OCC = 1..4 resident CTAs/SM for the N == 1 kernel
// (fewer concurrent DRAM streams win: 2 CTAs x 2 stages = 87.7 us, 3 x 2 = 91.3, 4 x 2 = 93.0, 1 x 2 = 118),
// ACAS2D_FORCE_LOOP = 1 routes N > 1 to the simple per-thread kernel the tiled one is checked against.
struct Tuning { int n1_occupancy; bool force_loop; int n1_tma; int n1_stages; int n1_pdl; int n1_keep_mb; };
Tuning &tuning()
{
    static Tuning t = [] {
        Tuning x;
        const char *o = std::getenv("ACAS2D_N1_OCC");
        x.n1_occupancy = (o && std::atoi(o) >= 1 && std::atoi(o) <= 4) ? std::atoi(o) : 2;   // measured best: 2 CTAs/SM
        const char *f = std::getenv("ACAS2D_FORCE_LOOP");
        x.force_loop = f && std::atoi(f) != 0;
        const char *t = std::getenv("ACAS2D_N1_TMA");
        x.n1_tma = t ? std::atoi(t) : 1;
        const char *g = std::getenv("ACAS2D_N1_STAGES");
        x.n1_stages = g ? std::atoi(g) : 2;                       // ring depth barely matters once loads are off the warps
        const char *d = std::getenv("ACAS2D_N1_PDL");
        x.n1_pdl = d ? std::atoi(d) : 1;                           // programmatic dependent launch of consecutive steps
        const char *kmb = std::getenv("ACAS2D_N1_KEEP_MB");
        x.n1_keep_mb = kmb ? std::atoi(kmb) : 32;                  // MB of player state kept L2-resident across launches
                                                                   // (measured: 0 -> 79.1, 16 -> 77.7, 27 -> 77.4, 40 -> 77.3, 60 -> 77.9 us)
        return x;
    }();
    return t;
}

template <bool MINSEP, int OCC>
__global__ void __launch_bounds__(kBlock, OCC)
step_n1_kernel(const DevParams P, const StatePtrs S, const float *__restrict__ actions, const Sinks out)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    Tally tally;
    tally_clear(tally);
    if (i < S.B) {
        Env1 e;
        load_env1(S, i, e, MINSEP);
        const float a = __ldcs(actions + i);
        step_env1<MINSEP, true>(P, S, e, a, i, out, tally, nullptr);
        store_env1(S, i, e, MINSEP);
    }
    tally_flush_warp(S.stats, tally);
}

// ---------------------------------------------------------------- N_TRAFFIC == 1, persistent + TMA ring
// The direct kernel above is bound by load latency: every warp sits on its five loads
// (long-scoreboard stall, profiles/r01_ncu_step_n1_v1_full.csv) before a ~400-instruction FP64
// dependency chain, and registers cap residency at 8 warps per scheduler.  Here the loads are taken
// off the warps: one elected thread per CTA keeps a STAGES-deep ring of 256-env input tiles in
// flight with TMA bulk copies (cp.async.bulk.shared.global, completion on an mbarrier with
// expect_tx), the CTA is persistent (grid = SMs x resident CTAs) and walks tiles round-robin; all
// 256 threads only compute out of shared memory and stream their results straight from registers
// (128-bit streaming stores).  Bytes in flight per SM = OCC x STAGES x 13 KB, independent of how
// many warps are stalled on arithmetic; the best setting is the one with the FEWEST concurrent
// streams that still keeps the schedulers fed (2 CTAs/SM x 2 stages).
// TILE = envs per tile = threads per CTA (256 or 512); a stage holds ppos | paux | thot | action.

// COMPACT (state->tpsi0 given): the intruder part of a stage is the 4-byte heading array instead of the 16-byte
// records -- 40 instead of 52 bytes read per env-step; envs that are not in the spawn pattern (injected states)
// fetch their record with a plain load.
// Respawns are DEFERRED: a finished env leaves its index in a shared-memory queue and the CTA respawns all of
// them together after its last tile.  A respawn is ~600 instructions on one lane; done in line it holds up
// its warp and, through the per-tile barrier that recycles the stage, the whole CTA -- with ~0.1 % of the
// envs finishing per step, one tile in five has one (measured: 92.6 us per step in steady state against
// 82.4 us with no episode ending).
constexpr int kRespawnQueue = 256;

template <int STAGES, int OCC, int TILE, bool COMPACT>
__global__ void __launch_bounds__(TILE, OCC)
step_n1_tma_kernel(const DevParams P, const StatePtrs S, const float *__restrict__ actions, const Sinks out,
                   const long long full_tiles, const long long keep_tiles)
{
    constexpr int kTileEnvs = TILE, kTraffic = COMPACT ? 4 : 16, kStageBytes = TILE * (16 + 16 + kTraffic + 4);
    constexpr int kOffPaux = TILE * 16, kOffThot = TILE * 32, kOffAct = TILE * (32 + kTraffic);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = (uint64_t *)(smem + STAGES * kStageBytes);
    __shared__ long long queue[kRespawnQueue];
    __shared__ int queue_count;
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        queue_count = 0;
    }
    __syncthreads();

    const uint32_t smem_base = smem_u32(smem), full_base = smem_u32(full);
    const uint64_t pol = ACAS2D_LOAD_HINT ? l2_evict_first_policy() : 0;
    // L2 RESIDENCY for a slice of the state.  Every step sweeps the same player records (36 B per env, rewritten in
    // place) and streams 69 B per env of outputs past them; 4 Mi envs are 151 MB of state against a 126 MB L2, so
    // left alone nothing survives from one launch to the next.  The first `keep_tiles` tiles are therefore loaded and
    // stored with an evict_last policy: ~27 MB of state stay in L2 across launches (what the hint retains on this part,
    // tools/ubench_l2keep.cu), are neither fetched from nor written back to HBM, and everything else keeps streaming
    // with evict_first as before.  Worth 2 % here (79.1 -> 77.3 us; a plain copy kernel gains 12 %): with the HBM time
    // down, the step is held by its own latency -- 63 % issue utilisation at four warps per scheduler.
    const uint64_t pol_keep = l2_evict_last_policy();
    // tile schedule: round-robin over the grid (the CTAs sweep one contiguous window of HBM together), or
    // -DACAS2D_BLOCKED_TILES: each CTA streams through its own contiguous range (experiment)
#ifdef ACAS2D_BLOCKED_TILES
    const long long per_cta = (full_tiles + gridDim.x - 1) / gridDim.x;
    const long long tile_first = (long long)blockIdx.x * per_cta, tile_step = 1;
    const long long tile_end = (tile_first + per_cta < full_tiles) ? tile_first + per_cta : full_tiles;
#else
    const long long tile_first = blockIdx.x, tile_step = gridDim.x, tile_end = full_tiles;
#endif
    auto issue = [&](long long tile, int s) {
        const uint32_t base = smem_base + s * kStageBytes, bar = full_base + 8 * s;
        const long long e0 = tile * kTileEnvs;
        const uint64_t ps = tile < keep_tiles ? pol_keep : pol;
        mbar_expect_tx(bar, kStageBytes);
        tma_load_1d(base, S.ppos + e0, kTileEnvs * 16, bar, ps);
        tma_load_1d(base + kOffPaux, S.paux + e0, kTileEnvs * 16, bar, ps);
        if (COMPACT) tma_load_1d(base + kOffThot, S.tpsi0 + e0, kTileEnvs * 4, bar, ps);
        else tma_load_1d(base + kOffThot, S.thot + e0, kTileEnvs * 16, bar, ps);
        tma_load_1d(base + kOffAct, actions + e0, kTileEnvs * 4, bar, pol);
    };

    // Programmatic dependent launch: consecutive steps are launched so that step k+1's CTAs may take an SM slot as
    // soon as one of step k's CTAs has left it (the launch gap between two dependent kernels is otherwise ~2.4 us
    // of an 80 us step); everything up to here touched no global memory.  Step k's state is complete only once its
    // whole grid has finished: wait for that before the first tile is fetched.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const long long tile = tile_first + (long long)s * tile_step;
            if (tile < tile_end) issue(tile, s);
        }
    }

    Tally tally;
    tally_clear(tally);
    auto finish = [&](Env1 &e, float a, int64_t i, bool keep) {
        e.minsep = 0.0f;
        e.respawned = false;
        if (step_env1<false, true, true>(P, S, e, a, i, out, tally, nullptr)) {
            const int slot = atomicAdd(&queue_count, 1);
            if (slot < kRespawnQueue) { queue[slot] = i; return; }
            respawn_env1<true>(P, S, e, i, out);           // queue full (a whole tile finishing at once): in line
        }
        if (keep) store_env1_hinted(S, i, e, pol_keep);
        else store_env1(S, i, e, false);
    };
    int it = 0;
    for (long long tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
        const int s = it % STAGES;
        mbar_wait(full_base + 8 * s, (unsigned)(it / STAGES) & 1u);
        const unsigned char *base = smem + s * kStageBytes;
        const Vec2d pp = ((const Vec2d *)base)[tid];
        const PlayerAux pa = ((const PlayerAux *)(base + kOffPaux))[tid];
        Float4 h;
        float tpsi = 0.0f;
        if (COMPACT) tpsi = ((const float *)(base + kOffThot))[tid];
        else h = ((const Float4 *)(base + kOffThot))[tid];
        const float a = ((const float *)(base + kOffAct))[tid];
        // WAR across proxies: the generic-proxy shared loads above must have been performed before
        // the async proxy (the TMA refill issued below) may overwrite the stage.  bar.sync alone does
        // not order them -- a warp can pass the barrier with its LDS still queued, and the bulk copy
        // does not go through that queue (observed: whole warps reading the NEXT tile's records).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                   // every thread has drained stage s
        if (tid == 0) {
            const long long next = tile + (long long)STAGES * tile_step;
            if (next < tile_end) issue(next, s);
        }
        const int64_t i = tile * kTileEnvs + tid;
        Env1 e;
        e.px = pp.x; e.py = pp.y; e.psi = pa.psi; e.ret = pa.ep_return;
        e.steps = pa.steps & kStepsMask;
        e.bits = pa.steps & ~kStepsMask;
        if (COMPACT) {
            e.tr = compact_traffic(P, tpsi, (e.bits & kDownBit) != 0);
            if (__any_sync(kFull, (e.bits & kCompactBit) == 0)) {      // injected states only: a real (warp-uniform) branch
                if ((e.bits & kCompactBit) == 0) e.tr = traffic_load(S, i, (e.bits & kResidualBit) != 0);
            }
        } else {
            e.tr.x0 = (double)h.x; e.tr.y0 = (double)h.y; e.tr.psi = (double)h.z; e.tr.v = (double)h.w;
            if (__any_sync(kFull, (e.bits & kResidualBit) != 0)) {     // injected float64 states only
                if (e.bits & kResidualBit) {
                    const Residual r = S.tres[i];
                    e.tr.x0 += r.x0; e.tr.y0 += r.y0; e.tr.psi += r.psi; e.tr.v += r.v;
                }
            }
        }
        finish(e, a, i, tile < keep_tiles);
    }

    // ragged tail (B % 256 envs): one CTA, plain loads
    const int64_t tail0 = (int64_t)full_tiles * kTileEnvs;
    if (tail0 < S.B && blockIdx.x == (unsigned)(full_tiles % gridDim.x)) {
        const int64_t i = tail0 + tid;
        if (i < S.B) {
            Env1 e;
            load_env1(S, i, e, false);
            finish(e, __ldcs(actions + i), i, false);
        }
    }

    // the deferred respawns of this CTA, one lane each
    __syncthreads();
    const int queued = queue_count < kRespawnQueue ? queue_count : kRespawnQueue;
    for (int q = tid; q < queued; q += TILE) {
        const int64_t i = queue[q];
        Env1 e;
        e.minsep = 0.0f;
        respawn_env1<true>(P, S, e, i, out);
        if (i < keep_tiles * kTileEnvs) store_env1_hinted(S, i, e, pol_keep);
        else store_env1(S, i, e, false);
    }
    tally_flush_warp(S.stats, tally);
}

template <int STAGES, int OCC, int TILE = 256>
int launch_n1_tma(const DevParams &P, const StatePtrs &S, const float *actions, const Sinks &out, cudaStream_t st)
{
    constexpr int kTileEnvs = TILE;
    const bool compact = compact_ok(P, S) && (((uintptr_t)S.tpsi0) & 15) == 0;
    const int stage_bytes = TILE * (16 + 16 + (compact ? 4 : 16) + 4);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int sms = sm_count();
    if (!attr_set[dev & 63]) {                      // function attributes are per device
        cudaFuncSetAttribute(step_n1_tma_kernel<STAGES, OCC, TILE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             STAGES * TILE * 52 + 64);
        cudaFuncSetAttribute(step_n1_tma_kernel<STAGES, OCC, TILE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             STAGES * TILE * 40 + 64);
        attr_set[dev & 63] = true;
    }
    const long long full_tiles = S.B / kTileEnvs;
    long long grid = (long long)sms * OCC;
    const long long tiles = full_tiles + ((S.B % kTileEnvs) ? 1 : 0);
    if (grid > tiles) grid = tiles;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(TILE); lc.dynamicSmemBytes = STAGES * stage_bytes + 64; lc.stream = st;
    cudaLaunchAttribute pdl;
    pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl.val.programmaticStreamSerializationAllowed = tuning().n1_pdl ? 1 : 0;
    lc.attrs = &pdl; lc.numAttrs = 1;
    const long long ft = full_tiles;
    // tiles whose player state is kept L2-resident across launches (see the kernel): keep_mb of (36 B per env)
    long long kt = (long long)((double)tuning().n1_keep_mb * 1e6 / (36.0 * kTileEnvs));
    if (kt > full_tiles) kt = full_tiles;
    cudaError_t err = compact ? cudaLaunchKernelEx(&lc, step_n1_tma_kernel<STAGES, OCC, TILE, true>, P, S, actions, out, ft, kt)
                              : cudaLaunchKernelEx(&lc, step_n1_tma_kernel<STAGES, OCC, TILE, false>, P, S, actions, out, ft, kt);
    return err == cudaSuccess ? 0 : (int)err;
}

// ---------------------------------------------------------------- policy + env step (N_TRAFFIC == 1)
// Closed-loop rollout step: actor MLP on the env's current observation row, Gaussian exploration noise
// (optional), clip to the action Box, then the environment step -- one kernel, no host round trip.
// Persistent blocks: the 19 KB weight block is loaded into shared memory once per block.
template <bool STOCHASTIC>
__global__ void __launch_bounds__(kBlock, 2)
policy_step_n1_kernel(const DevParams P, const StatePtrs S, const float *__restrict__ weights,
                      const float *obs_in, float *__restrict__ actions_out, float *__restrict__ logp_out,
                      const Sinks out, const float log_std_value, const uint64_t noise_seed, const uint64_t step_offset,
                      const PolicyDyn dyn)
{
    const float log_std = dyn.log_std ? *dyn.log_std : log_std_value;                 // live values for graph replays
    const uint64_t step_index = step_offset + (dyn.step_base ? *dyn.step_base : 0);
    __shared__ __align__(16) float sw[kPolFloats];
    for (int q = threadIdx.x; q < kPolFloats; q += kBlock) sw[q] = weights[q];
    __syncthreads();
    Tally tally;
    tally_clear(tally);
    const float std_dev = __expf(log_std);
    for (int64_t base = (int64_t)blockIdx.x * kBlock; base < S.B; base += (int64_t)gridDim.x * kBlock) {
        const int64_t i = base + threadIdx.x;
        if (i < S.B) {
            const float4 o0 = ((const float4 *)obs_in)[2 * i], o1 = ((const float4 *)obs_in)[2 * i + 1];
            const float obs[kPolObs] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
            float a = policy_mean(sw, obs);                                  // model.predict(deterministic=True)
            if (STOCHASTIC) {
                const float eps = policy_noise(noise_seed, S.gid0 + (uint64_t)i, step_index);
                a = fmaf(std_dev, eps, a);
                if (logp_out) logp_out[i] = -0.5f * eps * eps - log_std - 0.9189385332046727f;   // log N(a; mean, std)
            }
            if (actions_out) actions_out[i] = a;                             // rollout buffers keep the unclipped sample
            const float clipped = fminf(1.0f, fmaxf(-1.0f, a));              // np.clip to the action Box before env.step
            Env1 e;
            load_env1(S, i, e, false);
            step_env1<false, true>(P, S, e, clipped, i, out, tally, nullptr);
            store_env1(S, i, e, false);
        }
    }
    tally_flush_warp(S.stats, tally);
}

template <bool MINSEP>
__global__ void __launch_bounds__(kBlock)
rollout_n1_kernel(const DevParams P, const StatePtrs S, int num_steps, uint64_t action_seed, uint64_t step0,
                  float *__restrict__ reward_sum)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    Tally tally;
    tally_clear(tally);
    if (i < S.B) {
        Env1 e;
        load_env1(S, i, e, MINSEP);
        float racc = 0.0f;
        Sinks none = {};
        for (int k = 0; k < num_steps; ++k) {
            const float a = random_action(action_seed, S.gid0 + (uint64_t)i, step0 + (uint64_t)k);
            step_env1<MINSEP, false>(P, S, e, a, i, none, tally, &racc);
        }
        store_env1(S, i, e, MINSEP);
        if (reward_sum) reward_sum[i] += racc;
    }
    tally_flush_warp(S.stats, tally);
}

// K consecutive steps per launch for callers that already hold all K actions (open loop: random-action
// rollouts, replays of recorded action sequences).  State stays in registers between the steps -- per
// env-step only the action is read and the step's outputs are written: (80 / K + 41) bytes instead of 121.
__global__ void __launch_bounds__(kBlock)
step_k_n1_kernel(const DevParams P, const StatePtrs S, const int num_steps, const float *__restrict__ actions,
                 const Sinks out)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    Tally tally;
    tally_clear(tally);
    if (i < S.B) {
        Env1 e;
        load_env1(S, i, e, false);
        Sinks o = out;
        float a = __ldcs(actions + i);
        for (int k = 0; k < num_steps; ++k) {
            // the next step's action is in flight while this step's ~400-instruction chain runs
            const float a_next = (k + 1 < num_steps) ? __ldcs(actions + (int64_t)(k + 1) * S.B + i) : 0.0f;
            step_env1<false, true>(P, S, e, a, i, o, tally, nullptr);
            o.obs += 8 * S.B; o.reward += S.B; o.done += S.B;              // next step's [B] slice of the [K][B] outputs
            a = a_next;
        }
        store_env1(S, i, e, false);
    }
    tally_flush_warp(S.stats, tally);
}

// ---------------------------------------------------------------- N_TRAFFIC > 1 (first version)
template <bool MINSEP>
__global__ void __launch_bounds__(kBlock)
step_loop_kernel(const DevParams P, const StatePtrs S, const float *__restrict__ actions, const Sinks out)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    Tally tally;
    tally_clear(tally);
    if (i < S.B) step_env_loop<MINSEP>(P, S, i, actions[i], out, tally);
    tally_flush_warp(S.stats, tally);
}

// Intruders per lane the group size aims at.  Round 1 (player update redone by every lane of a group): 2 per lane
// 0.43 / 0.38 / 0.36 / 0.30 of the roofline at N = 8 / 16 / 32 / 64, 4: 0.60 / 0.53 / 0.50 / 0.42, 8: 0.71 / 0.66 /
// 0.63 / 0.52, 16: 0.70 / 0.51 / 0.49 / 0.40.  Runtime knob (acas2d_set_tiled_tuning / ACAS2D_TILED_PER_LANE).
int &tiled_per_lane()
{
    static int v = [] {
        const char *e = std::getenv("ACAS2D_TILED_PER_LANE");
        const int x = e ? std::atoi(e) : 0;
        return (x >= 1 && x <= 64) ? x : 8;
    }();
    return v;
}

inline int tiled_group(int N)
{
    int want = 1;
    while (want < 32 && want * tiled_per_lane() < N) want <<= 1;
    int g = 1;
    while (g < want && N % (g * 2) == 0) g <<= 1;         // G must divide N
    return g;
}

// Which record the tiled kernel reads (acas2d_set_tuning's third knob / ACAS2D_TILED_KIN: -1 = by N, 0 = always the
// 16-byte records, 1 = the kinematic cache whenever the state carries one).
int &tiled_kin_mode()
{
    static int mode = [] {
        const char *e = std::getenv("ACAS2D_TILED_KIN");
        return e ? std::atoi(e) : -1;
    }();
    return mode;
}

inline bool tiled_use_kin(const StatePtrs &S, int N)
{
    if (!S.tkin || (N & 1)) return false;               // 24-byte rows must stay 16-byte aligned for the bulk copies
    const int mode = tiled_kin_mode();
    if (mode >= 0) return mode != 0;
    // measured (B200, fraction of the HBM roofline, cache vs records): N = 64 at 262 144 envs 0.61 vs 0.53, N = 256 at
    // 65 536 envs 0.35 vs 0.30, N = 8 at 65 536 envs 0.38 vs 0.33 -- but N = 8 at 1 Mi envs 0.67 vs 0.71: with few
    // intruders per env the step is HBM-bound once the cache no longer fits L2, and the cache is 8 bytes more per intruder
    // N >= 256: one env per warp and (reference spawn rule) an episode per step -- a cache entry is written and read
    // once; the records win (203 vs 210 us) and move 0.5 GB less per launch
    if (N >= 256) return false;
    return N >= 16 || (double)S.B * N * 24.0 <= 48e6;
}

template <int G, bool MINSEP, bool KIN>
int launch_tiled_as(const DevParams &P, const StatePtrs &S, const float *actions, const Sinks &out, cudaStream_t st)
{
    constexpr int E = 32 / G, kTiledWarps = TiledShape<G>::kWarps;
    const size_t smem = tiled_smem_bytes(P.n_traffic, G, KIN, kTiledWarps);
    const uint32_t magic_n = (uint32_t)((0x100000000ULL + (uint64_t)P.n_traffic - 1) / (uint64_t)P.n_traffic);   // idx / N for idx < 2^16
    const int64_t warps = (S.B + E - 1) / E;
    const unsigned grid = (unsigned)((warps + kTiledWarps - 1) / kTiledWarps);
    cudaError_t err = cudaFuncSetAttribute(step_tiled_kernel<G, MINSEP, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    if (G > 1 && S.pstage) {                    // the float64 player update, once per env, as its own launch
        player_phase_kernel<<<(unsigned)((S.B + 127) / 128), 128, 0, st>>>(P, S, actions);
        // ... and the tiled kernel as its PROGRAMMATIC dependent: it starts while the pre-pass runs (tile copies,
        // index arithmetic) and waits for it only where it reads the scratch
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid); lc.blockDim = dim3(kTiledWarps * 32); lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute pdl;
        pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl.val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = &pdl; lc.numAttrs = 1;
        err = cudaLaunchKernelEx(&lc, step_tiled_kernel<G, MINSEP, KIN>, P, S, actions, out, magic_n);
        return err == cudaSuccess ? 0 : (int)err;
    }
    step_tiled_kernel<G, MINSEP, KIN><<<grid, kTiledWarps * 32, smem, st>>>(P, S, actions, out, magic_n);
    return 0;
}

template <int G>
int launch_tiled(const DevParams &P, const StatePtrs &S0, const float *actions, const Sinks &out, cudaStream_t st)
{
    // When the cache is not the record of choice for this batch, the step does not maintain it either: respawned
    // envs get no cache entry and lose their "cache valid" bit (a later launch that prefers the cache takes the
    // records for them) -- at N = 256, where nearly every env respawns every step, that is 24 of 52 bytes written.
    const bool kin = tiled_use_kin(S0, P.n_traffic);
    StatePtrs S = S0;
    if (!kin) S.tkin = nullptr;
    if (S.min_sep) return kin ? launch_tiled_as<G, true, true>(P, S, actions, out, st) : launch_tiled_as<G, true, false>(P, S, actions, out, st);
    return kin ? launch_tiled_as<G, false, true>(P, S, actions, out, st) : launch_tiled_as<G, false, false>(P, S, actions, out, st);
}

// ---------------------------------------------------------------- reset / inject / extract
__global__ void __launch_bounds__(kBlock)
reset_kernel(const DevParams P, const StatePtrs S, const uint8_t *__restrict__ mask, float *__restrict__ obs)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= S.B) return;
    if (mask && !mask[i]) return;
    reset_env(P, S, i, obs);
}

__global__ void __launch_bounds__(kBlock)
inject_kernel(const DevParams P, const StatePtrs S, const double *__restrict__ player,
              const double *__restrict__ traffic, const int32_t *__restrict__ steps,
              const double *__restrict__ total_reward)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < S.B) inject_env(P, S, i, player, traffic, steps, total_reward);
}

__global__ void __launch_bounds__(kBlock)
extract_kernel(const DevParams P, const StatePtrs S, double *__restrict__ player, double *__restrict__ traffic,
               int32_t *__restrict__ steps, double *__restrict__ total_reward)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < S.B) extract_env(P, S, i, player, traffic, steps, total_reward);
}

__global__ void __launch_bounds__(kBlock)
observe_kernel(const DevParams P, const StatePtrs S, float *__restrict__ obs)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < S.B) observe_env(P, S, i, obs);
}

// Off-path debug view of ONE env: the scene of game.view() (game.py:323-347) without sprites
// and HUD text -- sky background, discs of AIRCRAFT_SIZE for player (black, with a heading tick), goal
// (green) and intruders (grey), 1-px circles of COLLISION_RADIUS (red) and GOAL_RADIUS (yellow).
__global__ void __launch_bounds__(kBlock)
render_kernel(const DevParams P, const StatePtrs S, const int64_t env, const int W, const int H,
              const float aircraft_r, const float collision_r, const float goal_r, uint8_t *__restrict__ rgb)
{
    const int64_t pix = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (pix >= (int64_t)W * H) return;
    const float x = (float)(pix % W) + 0.5f, y = (float)(pix / W) + 0.5f;
    uint8_t r = 60, g = 150, b = 220;                                         // SKY_RGB
    const Vec2d pp = S.ppos[env];
    const PlayerAux pa = S.paux[env];
    const int st = pa.steps & kStepsMask;
    auto ring = [&](float cx, float cy, float rad) { return fabsf(sqrtf((x - cx) * (x - cx) + (y - cy) * (y - cy)) - rad) < 0.75f; };
    auto disc = [&](float cx, float cy, float rad) { return (x - cx) * (x - cx) + (y - cy) * (y - cy) < rad * rad; };
    const float gx = (float)P.goal_x, gy = (float)P.goal_y, px = (float)pp.x, py = (float)pp.y;
    if (disc(gx, gy, aircraft_r)) { r = 0; g = 255; b = 0; }
    for (int j = 0; j < P.n_traffic; ++j) {
        const TrafficRec tr = traffic_load(S, env * P.n_traffic + j, (pa.steps & kResidualBit) != 0);
        const Intruder t = intruder_at(P, tr, (double)(st - 1));
        if (disc((float)t.x, (float)t.y, aircraft_r)) { r = 90; g = 90; b = 90; }
        if (ring((float)t.x, (float)t.y, collision_r)) { r = 255; g = 0; b = 0; }
    }
    double s, c;
    sincos_deg(pa.psi, &s, &c);
    const float tx = x - px, ty = y - py, along = tx * (float)c + ty * (float)s, across = -tx * (float)s + ty * (float)c;
    if (disc(px, py, aircraft_r) || (along > 0.0f && along < 2.5f * aircraft_r && fabsf(across) < 1.5f)) { r = 0; g = 0; b = 0; }
    if (ring(px, py, collision_r)) { r = 255; g = 0; b = 0; }
    if (ring(gx, gy, goal_r)) { r = 255; g = 255; b = 0; }
    rgb[3 * pix + 0] = r; rgb[3 * pix + 1] = g; rgb[3 * pix + 2] = b;
}

__global__ void __launch_bounds__(128)
trace_kernel(const DevParams P, const StatePtrs S, const TraceRing T, const float *__restrict__ actions)
{
    const int64_t w = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (w < T.num_envs) trace_env(P, S, T, w, actions);
}

__global__ void __launch_bounds__(kBlock)
random_actions_kernel(int64_t B, uint64_t gid0, uint64_t action_seed, uint64_t step_index, float *__restrict__ actions)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < B) actions[i] = random_action(action_seed, gid0 + (uint64_t)i, step_index);
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

Sinks make_sinks(float *obs, float *reward, uint8_t *done, const acas2d_step_aux *aux)
{
    Sinks s;
    s.obs = obs; s.reward = reward; s.done = done;
    s.flags = aux ? aux->flags : nullptr;
    s.outcome = aux ? aux->outcome : nullptr;
    s.term_obs = aux ? aux->term_obs : nullptr;
    s.ep_return = aux ? aux->ep_return : nullptr;
    s.ep_length = aux ? aux->ep_length : nullptr;
    return s;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

int acas2d_abi_version(void) { return ACAS2D_ABI_VERSION; }

int acas2d_params_default(acas2d_params *p, int32_t n_traffic)
{
    if (!p) return ACAS2D_E_NULL;
    if (n_traffic < 1 || n_traffic > ACAS2D_MAX_TRAFFIC) return ACAS2D_E_BAD_TRAFFIC;
    std::memset(p, 0, sizeof(*p));
    p->width = 1600; p->height = 1000; p->fps = 100; p->max_steps = 1000;          // settings.py:9,15-17
    p->aircraft_size = 24;                                                           // settings.py:33
    p->collision_radius = 2 * p->aircraft_size;                                      // settings.py:34
    p->goal_radius = 6 * p->aircraft_size;                                           // settings.py:35
    p->safe_distance = 4 * p->collision_radius;                                      // settings.py:36
    p->airspeed = 200; p->airspeed_factor_min = 1; p->airspeed_factor_max = 1;       // settings.py:39-41
    p->acc_lat_limit = 20 * 9.80665;                                                 // settings.py:42
    p->player_heading_lim = 3; p->traffic_heading_lim = 15;                          // settings.py:43-44
    p->reward_goal = 1000; p->reward_collision = -1000;                              // settings.py:47-48
    p->goal_x = p->width - p->goal_radius; p->goal_y = p->height / 2;                // game.py:80-81
    p->player_x0 = p->collision_radius; p->player_y0 = p->height / 2;                // game.py:85-86
    double b = std::fmod(std::atan2(p->goal_y - p->player_y0, p->goal_x - p->player_x0), 2 * 3.141592653589793);
    if (b < 0) b += 2 * 3.141592653589793;
    p->player_psi_base = b * (180.0 / 3.141592653589793);                            // game.py:91
    const double reach = (p->airspeed / p->fps) * p->max_steps;
    const double dxg = p->player_x0 - p->goal_x, dyg = p->player_y0 - p->goal_y;
    p->d_goal_max = std::sqrt(dxg * dxg + dyg * dyg) + reach;                        // game.py:120
    p->d_dev_max = reach;                                                            // game.py:122
    const double diag = std::sqrt(p->width * p->width + p->height * p->height);
    p->d_separation_max = diag + 2 * reach;                                          // game.py:124
    p->d_cpa_max = diag;                                                             // game.py:126
    p->v_closing_max = 2 * (p->airspeed_factor_max * p->airspeed);                   // game.py:128
    p->n_traffic = n_traffic;
    p->auto_reset = 0;
    return 0;
}

int acas2d_reset(const acas2d_params *params, const acas2d_state *state, const uint8_t *mask, float *obs, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (state->num_envs == 0) return 0;
    reset_kernel<<<grid_for(state->num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        make_dev_params(*params), make_state_ptrs(*state), mask, obs);
    return finish_launch();
}

int acas2d_step(const acas2d_params *params, const acas2d_state *state, const float *actions, float *obs,
                float *reward, uint8_t *done, const acas2d_step_aux *aux, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (state->num_envs == 0) return 0;
    if (!actions || !obs || !reward || !done) return ACAS2D_E_NULL;
    const DevParams P = make_dev_params(*params);
    const StatePtrs S = make_state_ptrs(*state);
    const Sinks out = make_sinks(obs, reward, done, aux);
    const unsigned grid = grid_for(state->num_envs);
    cudaStream_t st = (cudaStream_t)stream;
    if (P.n_traffic == 1) {
        const bool occ4 = tuning().n1_occupancy == 4;
        const bool aligned = (((uintptr_t)actions | (uintptr_t)S.ppos | (uintptr_t)S.paux | (uintptr_t)S.thot) & 15) == 0;
        if (!S.min_sep && tuning().n1_tma && aligned) {
            const int stages = tuning().n1_stages;
            const int occ = tuning().n1_occupancy;
            if (occ == 4) {
                if (stages <= 2) launch_n1_tma<2, 4>(P, S, actions, out, st);
                else if (stages == 3) launch_n1_tma<3, 4>(P, S, actions, out, st);
                else launch_n1_tma<4, 4>(P, S, actions, out, st);
            } else if (occ == 2) {
                if (stages <= 2) launch_n1_tma<2, 2>(P, S, actions, out, st);
                else if (stages == 3) launch_n1_tma<3, 2>(P, S, actions, out, st);
                else launch_n1_tma<4, 2>(P, S, actions, out, st);
            } else if (occ == 1) {
                if (stages <= 2) launch_n1_tma<2, 1>(P, S, actions, out, st);
                else if (stages == 3) launch_n1_tma<3, 1>(P, S, actions, out, st);
                else launch_n1_tma<4, 1>(P, S, actions, out, st);
            } else {
                if (stages <= 2) launch_n1_tma<2, 3>(P, S, actions, out, st);
                else if (stages == 3) launch_n1_tma<3, 3>(P, S, actions, out, st);
                else if (stages == 4) launch_n1_tma<4, 3>(P, S, actions, out, st);
                else launch_n1_tma<5, 3>(P, S, actions, out, st);
            }
        } else if (S.min_sep) {
            if (occ4) step_n1_kernel<true, 4><<<grid, kBlock, 0, st>>>(P, S, actions, out);
            else step_n1_kernel<true, 3><<<grid, kBlock, 0, st>>>(P, S, actions, out);
        } else {
            if (occ4) step_n1_kernel<false, 4><<<grid, kBlock, 0, st>>>(P, S, actions, out);
            else step_n1_kernel<false, 3><<<grid, kBlock, 0, st>>>(P, S, actions, out);
        }
    } else if (tuning().force_loop) {
        StatePtrs Sl = S;
        if (!tiled_use_kin(S, P.n_traffic)) Sl.tkin = nullptr;      // same record policy as the tiled kernel (below)
        if (S.min_sep) step_loop_kernel<true><<<grid, kBlock, 0, st>>>(P, Sl, actions, out);
        else step_loop_kernel<false><<<grid, kBlock, 0, st>>>(P, Sl, actions, out);
    } else {
        int rc = 0;
        switch (tiled_group(P.n_traffic)) {
            case 1: rc = launch_tiled<1>(P, S, actions, out, st); break;
            case 2: rc = launch_tiled<2>(P, S, actions, out, st); break;
            case 4: rc = launch_tiled<4>(P, S, actions, out, st); break;
            case 8: rc = launch_tiled<8>(P, S, actions, out, st); break;
            case 16: rc = launch_tiled<16>(P, S, actions, out, st); break;
            default: rc = launch_tiled<32>(P, S, actions, out, st); break;
        }
        if (rc) return rc;
    }
    return finish_launch();
}

// Host-buffer step.  Large batches are cut into chunks that rotate over three internal streams so
// that the PCIe H2D copy of chunk c+1, the kernel of chunk c and the D2H copies of chunk c-1
// overlap (the D2H leg, 4L+5 bytes per env, is the bound); small batches use the caller's stream.
namespace {
struct HostPipe {
    cudaStream_t s[3];
    cudaEvent_t start, done[3];
    bool ok = false;
};
// One pipe per device (streams and events belong to the device that was current when they were created), and one
// caller at a time per device: the pipe's streams are shared state.
std::mutex g_pipe_mutex[64];
HostPipe &host_pipe(int dev)
{
    static HostPipe pipes[64];
    static bool made[64] = {};
    HostPipe &q = pipes[dev & 63];
    if (!made[dev & 63]) {
        q.ok = true;
        for (int i = 0; i < 3; ++i) {
            q.ok = q.ok && cudaStreamCreateWithFlags(&q.s[i], cudaStreamNonBlocking) == cudaSuccess;
            q.ok = q.ok && cudaEventCreateWithFlags(&q.done[i], cudaEventDisableTiming) == cudaSuccess;
        }
        q.ok = q.ok && cudaEventCreateWithFlags(&q.start, cudaEventDisableTiming) == cudaSuccess;
        made[dev & 63] = true;
    }
    return q;
}


constexpr int64_t kHostChunk = 256 * 1024;     // envs per chunk (multiple of the 256-env TMA tile)
}  // namespace

int acas2d_step_host(const acas2d_params *params, const acas2d_state *state, const float *h_actions, float *h_obs,
                     float *h_reward, uint8_t *h_done, float *d_actions, float *d_obs, float *d_reward,
                     uint8_t *d_done, const acas2d_step_aux *aux, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (state->num_envs == 0) return 0;
    if (!h_actions || !h_obs || !h_reward || !h_done || !d_actions || !d_obs || !d_reward || !d_done)
        return ACAS2D_E_NULL;
    const int64_t B = state->num_envs;
    const int N = params->n_traffic, L = 5 + 3 * N;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err;
#define ACAS_TRY(call) do { err = (call); if (err != cudaSuccess) return (int)err; } while (0)
    int dev = 0;
    cudaGetDevice(&dev);
    std::unique_lock<std::mutex> pipe_lock(g_pipe_mutex[dev & 63], std::defer_lock);
    HostPipe *pipe = nullptr;
    if (B >= 2 * kHostChunk) {
        pipe_lock.lock();
        pipe = &host_pipe(dev);
    }
    if (!pipe || !pipe->ok) {
        ACAS_TRY(cudaMemcpyAsync(d_actions, h_actions, sizeof(float) * B, cudaMemcpyHostToDevice, st));
        if (int e = acas2d_step(params, state, d_actions, d_obs, d_reward, d_done, aux, stream)) return e;
        ACAS_TRY(cudaMemcpyAsync(h_obs, d_obs, sizeof(float) * B * L, cudaMemcpyDeviceToHost, st));
        ACAS_TRY(cudaMemcpyAsync(h_reward, d_reward, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
        ACAS_TRY(cudaMemcpyAsync(h_done, d_done, sizeof(uint8_t) * B, cudaMemcpyDeviceToHost, st));
        return (int)cudaStreamSynchronize(st);
    }
    ACAS_TRY(cudaEventRecord(pipe->start, st));
    for (int i = 0; i < 3; ++i) ACAS_TRY(cudaStreamWaitEvent(pipe->s[i], pipe->start, 0));
    int c = 0;
    for (int64_t off = 0; off < B; off += kHostChunk, ++c) {
        const int64_t n = (B - off) < kHostChunk ? (B - off) : kHostChunk;
        cudaStream_t cs = pipe->s[c % 3];
        acas2d_state sub = *state;
        sub.num_envs = n;
        sub.ppos = (char *)state->ppos + 16 * off;
        sub.paux = (char *)state->paux + 16 * off;
        sub.thot = (char *)state->thot + 16 * N * off;
        sub.tres = (char *)state->tres + 32 * N * off;
        sub.episode_idx = state->episode_idx + off;
        sub.min_sep = state->min_sep ? state->min_sep + off : nullptr;
        sub.tkin = state->tkin ? (char *)state->tkin + 24 * N * off : nullptr;
        sub.tpsi0 = state->tpsi0 ? state->tpsi0 + off : nullptr;
        sub.spawn_sep = state->spawn_sep ? state->spawn_sep + off : nullptr;
        sub.pstage = state->pstage ? (char *)state->pstage + ACAS2D_PSTAGE_BYTES * off : nullptr;   // [7][n] block of this chunk
        sub.env_id_offset = state->env_id_offset + (uint64_t)off;
        acas2d_step_aux sa = {};
        if (aux) {
            sa.flags = aux->flags ? aux->flags + off : nullptr;
            sa.outcome = aux->outcome ? aux->outcome + off : nullptr;
            sa.term_obs = aux->term_obs ? aux->term_obs + L * off : nullptr;
            sa.ep_return = aux->ep_return ? aux->ep_return + off : nullptr;
            sa.ep_length = aux->ep_length ? aux->ep_length + off : nullptr;
        }
        ACAS_TRY(cudaMemcpyAsync(d_actions + off, h_actions + off, sizeof(float) * n, cudaMemcpyHostToDevice, cs));
        if (int e = acas2d_step(params, &sub, d_actions + off, d_obs + L * off, d_reward + off, d_done + off,
                                aux ? &sa : nullptr, cs))
            return e;
        ACAS_TRY(cudaMemcpyAsync(h_obs + L * off, d_obs + L * off, sizeof(float) * n * L, cudaMemcpyDeviceToHost, cs));
        ACAS_TRY(cudaMemcpyAsync(h_reward + off, d_reward + off, sizeof(float) * n, cudaMemcpyDeviceToHost, cs));
        ACAS_TRY(cudaMemcpyAsync(h_done + off, d_done + off, sizeof(uint8_t) * n, cudaMemcpyDeviceToHost, cs));
    }
    for (int i = 0; i < 3 && i < c; ++i) {
        ACAS_TRY(cudaEventRecord(pipe->done[i], pipe->s[i]));
        ACAS_TRY(cudaStreamWaitEvent(st, pipe->done[i], 0));
    }
#undef ACAS_TRY
    return (int)cudaStreamSynchronize(st);
}

int acas2d_step_mapped(const acas2d_params *params, const acas2d_state *state, const float *actions, float *obs,
                       float *reward, uint8_t *done, const acas2d_step_aux *aux, void *stream)
{
    if (int e = acas2d_step(params, state, actions, obs, reward, done, aux, stream)) return e;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

int acas2d_step_host_packed(const acas2d_params *params, const acas2d_state *state, const float *h_actions,
                            float *d_actions, float *obs, float *reward, uint8_t *done, const acas2d_step_aux *aux,
                            const void *d_packed, void *h_packed, int64_t packed_bytes, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (state->num_envs == 0) return 0;
    if (!h_actions || !d_actions || !obs || !reward || !done || !d_packed || !h_packed) return ACAS2D_E_NULL;
    if (packed_bytes <= 0) return ACAS2D_E_BAD_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err = cudaMemcpyAsync(d_actions, h_actions, sizeof(float) * state->num_envs, cudaMemcpyHostToDevice, st);
    if (err != cudaSuccess) return (int)err;
    if (int e = acas2d_step(params, state, d_actions, obs, reward, done, aux, stream)) return e;
    err = cudaMemcpyAsync(h_packed, d_packed, (size_t)packed_bytes, cudaMemcpyDeviceToHost, st);
    if (err != cudaSuccess) return (int)err;
    return (int)cudaStreamSynchronize(st);
}

int acas2d_trace_step(const acas2d_params *params, const acas2d_state *state, const float *actions,
                      const acas2d_trace *trace, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (!trace || !actions) return ACAS2D_E_NULL;
    if (trace->num_envs == 0) return 0;
    if (!trace->cursor || !trace->rows) return ACAS2D_E_NULL;
    if (trace->first_env < 0 || trace->num_envs < 0 || trace->first_env + trace->num_envs > state->num_envs ||
        trace->capacity < 2 || trace->n_traffic_rec < 0 || trace->n_traffic_rec > ACAS2D_TRACE_MAX_TRAFFIC ||
        trace->n_traffic_rec > params->n_traffic)
        return ACAS2D_E_BAD_SIZE;
    TraceRing T;
    T.first_env = trace->first_env; T.num_envs = trace->num_envs; T.capacity = trace->capacity;
    T.n_traffic_rec = trace->n_traffic_rec; T.cursor = trace->cursor; T.rows = trace->rows;
    trace_kernel<<<(unsigned)((T.num_envs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        make_dev_params(*params), make_state_ptrs(*state), T, actions);
    return finish_launch();
}

int acas2d_inject_state(const acas2d_params *params, const acas2d_state *state, const double *player,
                        const double *traffic, const int32_t *steps, const double *total_reward, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (state->num_envs == 0) return 0;
    if (!player || !traffic || !steps || !total_reward) return ACAS2D_E_NULL;
    inject_kernel<<<grid_for(state->num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        make_dev_params(*params), make_state_ptrs(*state), player, traffic, steps, total_reward);
    return finish_launch();
}

int acas2d_extract_state(const acas2d_params *params, const acas2d_state *state, double *player, double *traffic,
                         int32_t *steps, double *total_reward, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (state->num_envs == 0) return 0;
    extract_kernel<<<grid_for(state->num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        make_dev_params(*params), make_state_ptrs(*state), player, traffic, steps, total_reward);
    return finish_launch();
}

int acas2d_observe(const acas2d_params *params, const acas2d_state *state, float *obs, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (state->num_envs == 0) return 0;
    if (!obs) return ACAS2D_E_NULL;
    observe_kernel<<<grid_for(state->num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        make_dev_params(*params), make_state_ptrs(*state), obs);
    return finish_launch();
}

int acas2d_render(const acas2d_params *params, const acas2d_state *state, int64_t env_index, uint8_t *rgb, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (!rgb) return ACAS2D_E_NULL;
    if (env_index < 0 || env_index >= state->num_envs) return ACAS2D_E_BAD_SIZE;
    const int W = (int)params->width, H = (int)params->height;
    render_kernel<<<grid_for((int64_t)W * H), kBlock, 0, (cudaStream_t)stream>>>(
        make_dev_params(*params), make_state_ptrs(*state), env_index, W, H, (float)(params->aircraft_size / 2),
        (float)params->collision_radius, (float)params->goal_radius, rgb);
    return finish_launch();
}

int acas2d_rollout_random(const acas2d_params *params, const acas2d_state *state, int32_t num_steps,
                          uint64_t action_seed, uint64_t step0, float *reward_sum, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (params->n_traffic != 1) return ACAS2D_E_BAD_TRAFFIC;
    if (state->num_envs == 0 || num_steps <= 0) return 0;
    DevParams P = make_dev_params(*params);
    P.auto_reset = 1;
    const StatePtrs S = make_state_ptrs(*state);
    const unsigned grid = grid_for(state->num_envs);
    if (S.min_sep) rollout_n1_kernel<true><<<grid, kBlock, 0, (cudaStream_t)stream>>>(P, S, num_steps, action_seed, step0, reward_sum);
    else rollout_n1_kernel<false><<<grid, kBlock, 0, (cudaStream_t)stream>>>(P, S, num_steps, action_seed, step0, reward_sum);
    return finish_launch();
}

namespace {
// Function attributes (per device) of the policy kernels; cudaFuncGetAttributes also forces the lazily loaded
// kernels in, so that a first launch may happen inside a CUDA-graph capture.
int policy_prepare_device()
{
    static bool ready[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (ready[dev & 63]) return 0;
    cudaError_t err = cudaFuncSetAttribute(policy_step_n1_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (err != cudaSuccess) return (int)err;
    err = cudaFuncSetAttribute(policy_step_n1_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (err != cudaSuccess) return (int)err;
    cudaFuncAttributes fa;
    if ((err = cudaFuncGetAttributes(&fa, policy_step_n1_kernel<true>)) != cudaSuccess) return (int)err;
    if ((err = cudaFuncGetAttributes(&fa, policy_step_n1_kernel<false>)) != cudaSuccess) return (int)err;
    ready[dev & 63] = true;
    return 0;
}
}  // namespace

int acas2d_step_k(const acas2d_params *params, const acas2d_state *state, int32_t num_steps, const float *actions,
                  float *obs, float *reward, uint8_t *done, const acas2d_step_aux *aux, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (params->n_traffic != 1 || state->min_sep) return ACAS2D_E_BAD_TRAFFIC;
    if (num_steps < 0) return ACAS2D_E_BAD_SIZE;
    if (state->num_envs == 0 || num_steps == 0) return 0;
    if (!actions || !obs || !reward || !done) return ACAS2D_E_NULL;
    step_k_n1_kernel<<<grid_for(state->num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        make_dev_params(*params), make_state_ptrs(*state), num_steps, actions, make_sinks(obs, reward, done, aux));
    return finish_launch();
}

int acas2d_policy_step(const acas2d_params *params, const acas2d_state *state, const float *weights,
                       float log_std, const float *obs_in, float *actions_out, float *logp_out, float *obs_out,
                       float *reward, uint8_t *done, const acas2d_step_aux *aux, int32_t stochastic,
                       uint64_t noise_seed, uint64_t step_index, int32_t tensor_cores, void *stream)
{
    return acas2d_policy_step_dyn(params, state, weights, log_std, obs_in, actions_out, logp_out, obs_out, reward, done,
                                  aux, stochastic, noise_seed, step_index, tensor_cores, nullptr, nullptr, stream);
}

int acas2d_policy_step_dyn(const acas2d_params *params, const acas2d_state *state, const float *weights,
                           float log_std, const float *obs_in, float *actions_out, float *logp_out, float *obs_out,
                           float *reward, uint8_t *done, const acas2d_step_aux *aux, int32_t stochastic,
                           uint64_t noise_seed, uint64_t step_index, int32_t tensor_cores,
                           const float *log_std_dev, const uint64_t *step_base_dev, void *stream)
{
    PolicyDyn dyn;
    dyn.log_std = log_std_dev; dyn.step_base = step_base_dev;
    if (int e = check_args(params, state)) return e;
    if (params->n_traffic != 1 || state->min_sep) return ACAS2D_E_BAD_TRAFFIC;    // the trained actor takes 8 inputs
    if (state->num_envs == 0) return 0;
    if (!weights || !obs_in || !obs_out || !reward || !done) return ACAS2D_E_NULL;
    const DevParams P = make_dev_params(*params);
    const StatePtrs S = make_state_ptrs(*state);
    const Sinks out = make_sinks(obs_out, reward, done, aux);
    const int sms = sm_count();
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = policy_prepare_device()) return e;
    if (tensor_cores) {
        long long g = (long long)sms * 4;                   // 4 CTAs/SM: 52 KB smem + 64 TMEM columns each
        const long long t = (S.B + kTcTile - 1) / kTcTile;
        if (g > t) g = t;
        if (stochastic)
            policy_step_n1_tc_kernel<true><<<(unsigned)g, kTcTile, kTcSmemBytes, st>>>(
                P, S, weights, obs_in, actions_out, logp_out, out, log_std, noise_seed, step_index, dyn);
        else
            policy_step_n1_tc_kernel<false><<<(unsigned)g, kTcTile, kTcSmemBytes, st>>>(
                P, S, weights, obs_in, actions_out, logp_out, out, log_std, noise_seed, step_index, dyn);
        return finish_launch();
    }
    long long grid = (long long)sms * 2;
    const long long tiles = (S.B + kBlock - 1) / kBlock;
    if (grid > tiles) grid = tiles;
    if (stochastic)
        policy_step_n1_kernel<true><<<(unsigned)grid, kBlock, 0, st>>>(P, S, weights, obs_in, actions_out, logp_out, out,
                                                                       log_std, noise_seed, step_index, dyn);
    else
        policy_step_n1_kernel<false><<<(unsigned)grid, kBlock, 0, st>>>(P, S, weights, obs_in, actions_out, logp_out, out,
                                                                        log_std, noise_seed, step_index, dyn);
    return finish_launch();
}

int acas2d_random_actions(const acas2d_state *state, uint64_t action_seed, uint64_t step_index, float *actions, void *stream)
{
    if (!state || !actions) return ACAS2D_E_NULL;
    if (state->num_envs <= 0) return state->num_envs < 0 ? ACAS2D_E_BAD_SIZE : 0;
    random_actions_kernel<<<grid_for(state->num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        state->num_envs, state->env_id_offset, action_seed, step_index, actions);
    return finish_launch();
}

// ---------------------------------------------------------------- PPO learner (acas2d_ppo.cuh)
namespace { int ppo_prepare_device(); }
int acas2d_ppo_values(const float *params, const float *obs, int64_t n, float *values, void *stream)
{
    if (n < 0) return ACAS2D_E_BAD_SIZE;
    if (n == 0) return 0;
    if (!params || !obs || !values) return ACAS2D_E_NULL;
    const int sms = sm_count();
    if (int e = ppo_prepare_device()) return e;
    long long grid = (long long)sms * 2;
    const long long tiles = (n + kPpoValTile - 1) / kPpoValTile;
    if (grid > tiles) grid = tiles;
    ppo_values_kernel<<<(unsigned)grid, kPpoThreads, kPpoValSmemBytes, (cudaStream_t)stream>>>(params, obs, n, values);
    return finish_launch();
}

int acas2d_ppo_gae(const acas2d_ppo_config *cfg, const float *rewards, const uint8_t *dones, const float *values,
                   int32_t n_steps, int64_t num_envs, float *advantages, float *returns, void *stream)
{
    if (n_steps < 0 || num_envs < 0) return ACAS2D_E_BAD_SIZE;
    if (n_steps == 0 || num_envs == 0) return 0;
    if (!cfg || !rewards || !dones || !values || !advantages || !returns) return ACAS2D_E_NULL;
    ppo_gae_kernel<<<grid_for(num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        rewards, dones, values, n_steps, num_envs, cfg->gamma, cfg->gae_lambda, advantages, returns);
    return finish_launch();
}

namespace {
// Function attributes are per device; cudaFuncGetAttributes also forces the (lazily loaded) kernels in, so
// that a first launch may happen inside a CUDA-graph capture.
int ppo_prepare_device()
{
    static bool ready[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (ready[dev & 63]) return 0;
    cudaError_t err = cudaFuncSetAttribute(ppo_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPpoSmemBytes);
    if (err != cudaSuccess) return (int)err;
    err = cudaFuncSetAttribute(ppo_values_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPpoValSmemBytes);
    if (err != cudaSuccess) return (int)err;
    cudaFuncAttributes fa;
    if ((err = cudaFuncGetAttributes(&fa, ppo_reduce_kernel)) != cudaSuccess) return (int)err;
    if ((err = cudaFuncGetAttributes(&fa, ppo_adam_kernel)) != cudaSuccess) return (int)err;
    if ((err = cudaFuncGetAttributes(&fa, ppo_update_kernel)) != cudaSuccess) return (int)err;
    if ((err = cudaFuncGetAttributes(&fa, ppo_gae_kernel)) != cudaSuccess) return (int)err;
    ready[dev & 63] = true;
    return 0;
}

PpoAdam make_adam(const acas2d_ppo_config &c)
{
    PpoAdam h;
    h.lr = c.lr; h.beta1 = c.beta1; h.beta2 = c.beta2; h.eps = c.adam_eps; h.max_grad_norm = c.max_grad_norm;
    return h;
}

int launch_ppo_grad(const acas2d_ppo_config *cfg, const float *params, const float *obs, const float *actions,
                    const float *old_logp, const float *advantages, const float *returns, const int64_t *indices,
                    int64_t minibatch, float *workspace, int32_t *adam_step, cudaStream_t st)
{
    const int64_t tiles = (minibatch + kPpoTile - 1) / kPpoTile;
    const int ctas = (int)(tiles < kPpoMaxCtas ? tiles : kPpoMaxCtas);
    PpoBatch b;
    b.obs = obs; b.actions = actions; b.old_logp = old_logp; b.adv = advantages; b.ret = returns;
    b.idx = indices; b.mb = minibatch;
    // programmatic dependent of whatever precedes it in the stream (the previous step's update kernel signals early)
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)ctas, 2); lc.blockDim = dim3(kPpoThreads); lc.dynamicSmemBytes = kPpoSmemBytes; lc.stream = st;
    cudaLaunchAttribute pdl;
    pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl.val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = &pdl; lc.numAttrs = 1;
    cudaLaunchKernelEx(&lc, ppo_grad_kernel, params, b, (int)cfg->normalize_advantage, cfg->clip_range, cfg->vf_coef,
                       workspace + ACAS2D_PPO_WORKSPACE_HEAD, adam_step);
    return ctas;
}
}  // namespace

int acas2d_ppo_prepare(void)
{
    if (int e = policy_prepare_device()) return e;
    return ppo_prepare_device();
}

int acas2d_ppo_grad(const acas2d_ppo_config *cfg, const float *params, const float *obs, const float *actions,
                    const float *old_logp, const float *advantages, const float *returns, const int64_t *indices,
                    int64_t minibatch, float *workspace, float *grad, float *loss_stats, int32_t *adam_step,
                    void *stream)
{
    if (minibatch <= 0) return ACAS2D_E_BAD_SIZE;
    if (!cfg || !params || !obs || !actions || !old_logp || !advantages || !returns || !workspace || !grad)
        return ACAS2D_E_NULL;
    if (int e = ppo_prepare_device()) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = launch_ppo_grad(cfg, params, obs, actions, old_logp, advantages, returns, indices, minibatch,
                                     workspace, adam_step, st);
    ppo_reduce_kernel<<<kPpoUpdateCtas, 256, 0, st>>>(workspace + ACAS2D_PPO_WORKSPACE_HEAD, ctas, cfg->ent_coef,
                                                      1.0f / (float)minibatch, grad, loss_stats);
    return finish_launch(2);
}

int acas2d_ppo_adam(const acas2d_ppo_config *cfg, float *params, const float *grad, float grad_scale,
                    float *adam_m, float *adam_v, const int32_t *adam_step, float *loss_stats, void *stream)
{
    if (!cfg || !params || !grad || !adam_m || !adam_v || !adam_step) return ACAS2D_E_NULL;
    if (int e = ppo_prepare_device()) return e;
    ppo_adam_kernel<<<kPpoUpdateCtas, 256, 0, (cudaStream_t)stream>>>(
        params, grad, grad_scale, adam_m, adam_v, adam_step, make_adam(*cfg), loss_stats);
    return finish_launch();
}

int acas2d_ppo_step(const acas2d_ppo_config *cfg, float *params, const float *obs, const float *actions,
                    const float *old_logp, const float *advantages, const float *returns, const int64_t *indices,
                    int64_t minibatch, float *workspace, float *adam_m, float *adam_v, int32_t *sync,
                    float *loss_stats, float *grad_out, int32_t rank, int32_t world, void *const *peer_exchange,
                    void *stream)
{
    if (minibatch <= 0 || world < 1 || world > kPpoMaxRanks || rank < 0 || rank >= world) return ACAS2D_E_BAD_SIZE;
    if (!cfg || !params || !obs || !actions || !old_logp || !advantages || !returns || !workspace || !adam_m ||
        !adam_v || !sync || (world > 1 && !peer_exchange))
        return ACAS2D_E_NULL;
    PpoPeers peers;
    for (int r = 0; r < kPpoMaxRanks; ++r) peers.block[r] = nullptr;
    for (int r = 0; r < world && world > 1; ++r) {
        if (!peer_exchange[r]) return ACAS2D_E_NULL;
        peers.block[r] = (float *)peer_exchange[r];
    }
    if (int e = ppo_prepare_device()) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = launch_ppo_grad(cfg, params, obs, actions, old_logp, advantages, returns, indices, minibatch,
                                     workspace, sync, st);
    // cooperative: the update kernel's CTAs meet at grid barriers, so they must all be resident together -- the
    // driver guarantees it or refuses the launch (also inside a stream capture: the graph node keeps the attribute)
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(kPpoUpdateCtas); lc.blockDim = dim3(kPpoUpdateThreads); lc.dynamicSmemBytes = 0; lc.stream = st;
    cudaLaunchAttribute coop;
    coop.id = cudaLaunchAttributeCooperative;
    coop.val.cooperative = 1;
    lc.attrs = &coop; lc.numAttrs = 1;
    const cudaError_t lerr = cudaLaunchKernelEx(&lc, ppo_update_kernel, params, (const float *)(workspace + ACAS2D_PPO_WORKSPACE_HEAD),
                                                ctas, cfg->ent_coef, 1.0f / (float)minibatch, workspace, peers, (int)rank, (int)world,
                                                adam_m, adam_v, sync, make_adam(*cfg), loss_stats, grad_out);
    if (lerr != cudaSuccess) return (int)lerr;
    return finish_launch(2);
}

int64_t acas2d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int acas2d_set_tuning(int32_t n1_occupancy, int32_t force_loop)
{
    if (n1_occupancy >= 1 && n1_occupancy <= 4) tuning().n1_occupancy = n1_occupancy;
    if (force_loop >= 0) tuning().force_loop = force_loop != 0;
    return 0;
}

int acas2d_set_tiled_tuning(int32_t kin_mode, int32_t per_lane)
{
    if (kin_mode >= -1 && kin_mode <= 1) tiled_kin_mode() = kin_mode;
    if (per_lane >= 1 && per_lane <= 64) tiled_per_lane() = per_lane;
    return 0;
}

int acas2d_set_n1_kernel(int32_t use_tma, int32_t stages)
{
    if (use_tma >= 0) tuning().n1_tma = use_tma != 0;
    if (stages >= 2 && stages <= 5) tuning().n1_stages = stages;
    return 0;
}

}  // extern "C"
