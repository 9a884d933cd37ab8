// acas2d_kernels.cu -- sm_100a kernels and the C ABI (include/acas2d_b200.h) of the
// batched ACAS-2D environment step.
//
// Kernels
//   step_n1_kernel        thread per env, N_TRAFFIC == 1 (the reference default): five
//                         128-bit loads + one 32-bit load per env, four 128-bit stores +
//                         reward + done.  HBM-bound streaming kernel.
//   step_tiled_kernel     N_TRAFFIC > 1: G lanes per env (G = 1..32), the warp's traffic
//                         tile staged in shared memory with cp.async, min-separation /
//                         any-collision reduced with warp shuffles, observation rows
//                         assembled in shared memory and written back coalesced.
//   rollout_n1_kernel     K fused steps with in-kernel Philox actions (synthetic benchmark).
//   reset / inject / extract / random_actions  small utility kernels.
//
// There is no CPU fallback in this file: every entry point launches on the device.
#include <cuda_runtime.h>
#include <cuda_pipeline.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/acas2d_b200.h"
#include "acas2d_env.cuh"

namespace {

using namespace acas2d;

std::atomic<int64_t> g_launches{0};

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;

int check_args(const acas2d_params *p, const acas2d_state *s)
{
    if (!p || !s) return ACAS2D_E_NULL;
    if (p->n_traffic < 1 || p->n_traffic > ACAS2D_MAX_TRAFFIC) return ACAS2D_E_BAD_TRAFFIC;
    if (s->num_envs < 0 || s->num_envs * (int64_t)(5 + 3 * p->n_traffic) > (int64_t)1 << 40) return ACAS2D_E_BAD_SIZE;
    if (!s->ppos || !s->paux || !s->tpos0 || !s->tvel || !s->tpsi || !s->tvair || !s->episode_idx)
        return ACAS2D_E_NULL;
    return 0;
}

int finish_launch()
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- warp-level statistics flush
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// All 32 lanes must call.  One set of atomics per warp that saw a finished episode, spread
// over ACAS2D_STAT_SLOTS 128-byte slots so that same-address serialisation stays negligible.
__device__ __forceinline__ void tally_flush_warp(long long *stats, const Tally &t)
{
    if (stats == nullptr) return;
    if (!__any_sync(kFull, t.episodes != 0)) return;
    const int episodes = __reduce_add_sync(kFull, t.episodes);
    const int goal = __reduce_add_sync(kFull, t.goal);
    const int coll = __reduce_add_sync(kFull, t.coll);
    const int tout = __reduce_add_sync(kFull, t.tout);
    const long long length = warp_sum_ll(t.length);
    const long long ret_fx = warp_sum_ll(t.ret_fx);
    const long long minsep_fx = warp_sum_ll(t.minsep_fx);
    if ((threadIdx.x & 31) == 0) {
        const unsigned warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        unsigned long long *slot = (unsigned long long *)stats +
                                   (size_t)(warp_global % ACAS2D_STAT_SLOTS) * ACAS2D_STAT_FIELDS;
        atomicAdd(slot + ACAS2D_STAT_EPISODES, (unsigned long long)episodes);
        if (goal) atomicAdd(slot + ACAS2D_STAT_GOAL, (unsigned long long)goal);
        if (coll) atomicAdd(slot + ACAS2D_STAT_COLLISION, (unsigned long long)coll);
        if (tout) atomicAdd(slot + ACAS2D_STAT_TIMEOUT, (unsigned long long)tout);
        atomicAdd(slot + ACAS2D_STAT_LENGTH, (unsigned long long)length);
        atomicAdd(slot + ACAS2D_STAT_RETURN_FX, (unsigned long long)ret_fx);
        if (minsep_fx) atomicAdd(slot + ACAS2D_STAT_MINSEP_FX, (unsigned long long)minsep_fx);
    }
}

template <bool MINSEP>
__global__ void __launch_bounds__(kBlock)
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
    if (!actions || !obs || !reward || !done) return ACAS2D_E_NULL;
    if (state->num_envs == 0) return 0;
    const DevParams P = make_dev_params(*params);
    const StatePtrs S = make_state_ptrs(*state);
    const Sinks out = make_sinks(obs, reward, done, aux);
    const unsigned grid = grid_for(state->num_envs);
    cudaStream_t st = (cudaStream_t)stream;
    if (P.n_traffic == 1) {
        if (S.min_sep) step_n1_kernel<true><<<grid, kBlock, 0, st>>>(P, S, actions, out);
        else step_n1_kernel<false><<<grid, kBlock, 0, st>>>(P, S, actions, out);
    } else {
        if (S.min_sep) step_loop_kernel<true><<<grid, kBlock, 0, st>>>(P, S, actions, out);
        else step_loop_kernel<false><<<grid, kBlock, 0, st>>>(P, S, actions, out);
    }
    return finish_launch();
}

int acas2d_step_host(const acas2d_params *params, const acas2d_state *state, const float *h_actions, float *h_obs,
                     float *h_reward, uint8_t *h_done, float *d_actions, float *d_obs, float *d_reward,
                     uint8_t *d_done, const acas2d_step_aux *aux, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (!h_actions || !h_obs || !h_reward || !h_done || !d_actions || !d_obs || !d_reward || !d_done)
        return ACAS2D_E_NULL;
    const int64_t B = state->num_envs;
    const int L = 5 + 3 * params->n_traffic;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err = cudaMemcpyAsync(d_actions, h_actions, sizeof(float) * B, cudaMemcpyHostToDevice, st);
    if (err != cudaSuccess) return (int)err;
    if (int e = acas2d_step(params, state, d_actions, d_obs, d_reward, d_done, aux, stream)) return e;
    err = cudaMemcpyAsync(h_obs, d_obs, sizeof(float) * B * L, cudaMemcpyDeviceToHost, st);
    if (err != cudaSuccess) return (int)err;
    err = cudaMemcpyAsync(h_reward, d_reward, sizeof(float) * B, cudaMemcpyDeviceToHost, st);
    if (err != cudaSuccess) return (int)err;
    err = cudaMemcpyAsync(h_done, d_done, sizeof(uint8_t) * B, cudaMemcpyDeviceToHost, st);
    if (err != cudaSuccess) return (int)err;
    return (int)cudaStreamSynchronize(st);
}

int acas2d_inject_state(const acas2d_params *params, const acas2d_state *state, const double *player,
                        const double *traffic, const int32_t *steps, const double *total_reward, void *stream)
{
    if (int e = check_args(params, state)) return e;
    if (!player || !traffic || !steps || !total_reward) return ACAS2D_E_NULL;
    if (state->num_envs == 0) return 0;
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

int acas2d_random_actions(const acas2d_state *state, uint64_t action_seed, uint64_t step_index, float *actions, void *stream)
{
    if (!state || !actions) return ACAS2D_E_NULL;
    if (state->num_envs <= 0) return state->num_envs < 0 ? ACAS2D_E_BAD_SIZE : 0;
    random_actions_kernel<<<grid_for(state->num_envs), kBlock, 0, (cudaStream_t)stream>>>(
        state->num_envs, state->env_id_offset, action_seed, step_index, actions);
    return finish_launch();
}

int64_t acas2d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
