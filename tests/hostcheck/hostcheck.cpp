// TEST INFRASTRUCTURE.  Compiles the product's per-environment source
// (gym-acas2d_b200/csrc/acas2d_env.cuh, acas2d_math.cuh -- the __host__ __device__ bodies the
// CUDA kernels inline) with g++ so the step logic can be compared with the CPU oracle on a
// machine without a GPU.  It shares the C ABI's structs but takes HOST pointers.  Nothing in
// the product loads this library; GPU parity is established separately by the -m gpu tests.
#include <cstdint>
#include <cmath>

#include "../../gym-acas2d_b200/csrc/acas2d_env.cuh"
#include "../../gym-acas2d_b200/csrc/acas2d_policy.cuh"

using namespace acas2d;

static Sinks sinks_of(float *obs, float *reward, uint8_t *done, const acas2d_step_aux *aux)
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

static void flush(const StatePtrs &S, const Tally &t)
{
    if (!S.stats || !t.episodes) return;
    long long *slot = S.stats;   // slot 0
    slot[ACAS2D_STAT_EPISODES] += t.episodes;
    slot[ACAS2D_STAT_GOAL] += t.goal;
    slot[ACAS2D_STAT_COLLISION] += t.coll;
    slot[ACAS2D_STAT_TIMEOUT] += t.tout;
    slot[ACAS2D_STAT_LENGTH] += t.length;
    slot[ACAS2D_STAT_RETURN_FX] += t.ret_fx;
    slot[ACAS2D_STAT_MINSEP_FX] += t.minsep_fx;
}

extern "C" {

int hostcheck_reset(const acas2d_params *p, const acas2d_state *s, const uint8_t *mask, float *obs)
{
    const DevParams P = make_dev_params(*p);
    const StatePtrs S = make_state_ptrs(*s);
    for (int64_t i = 0; i < S.B; ++i)
        if (!mask || mask[i]) reset_env(P, S, i, obs);
    return 0;
}

// variant 0: the N == 1 register path (step_env1); variant 1: the per-env loop path.
int hostcheck_step(const acas2d_params *p, const acas2d_state *s, const float *actions, float *obs,
                   float *reward, uint8_t *done, const acas2d_step_aux *aux, int variant)
{
    const DevParams P = make_dev_params(*p);
    const StatePtrs S = make_state_ptrs(*s);
    const Sinks out = sinks_of(obs, reward, done, aux);
    Tally tally;
    tally_clear(tally);
    for (int64_t i = 0; i < S.B; ++i) {
        if (variant == 0) {
            if (P.n_traffic != 1) return ACAS2D_E_BAD_TRAFFIC;
            Env1 e;
            if (S.min_sep) {
                load_env1(S, i, e, true);
                step_env1<true, true>(P, S, e, actions[i], i, out, tally, nullptr);
                store_env1(S, i, e, true);
            } else {
                load_env1(S, i, e, false);
                step_env1<false, true>(P, S, e, actions[i], i, out, tally, nullptr);
                store_env1(S, i, e, false);
            }
        } else {
            if (S.min_sep) step_env_loop<true>(P, S, i, actions[i], out, tally);
            else step_env_loop<false>(P, S, i, actions[i], out, tally);
        }
    }
    flush(S, tally);
    return 0;
}

int hostcheck_rollout_random(const acas2d_params *p, const acas2d_state *s, int32_t num_steps,
                             uint64_t action_seed, uint64_t step0, float *reward_sum)
{
    DevParams P = make_dev_params(*p);
    P.auto_reset = 1;
    const StatePtrs S = make_state_ptrs(*s);
    Tally tally;
    tally_clear(tally);
    Sinks none = {};
    for (int64_t i = 0; i < S.B; ++i) {
        Env1 e;
        load_env1(S, i, e, S.min_sep != nullptr);
        float racc = 0.0f;
        for (int k = 0; k < num_steps; ++k) {
            const float a = random_action(action_seed, S.gid0 + (uint64_t)i, step0 + (uint64_t)k);
            if (S.min_sep) step_env1<true, false>(P, S, e, a, i, none, tally, &racc);
            else step_env1<false, false>(P, S, e, a, i, none, tally, &racc);
        }
        store_env1(S, i, e, S.min_sep != nullptr);
        if (reward_sum) reward_sum[i] += racc;
    }
    flush(S, tally);
    return 0;
}

int hostcheck_random_actions(const acas2d_state *s, uint64_t action_seed, uint64_t step_index, float *actions)
{
    for (int64_t i = 0; i < s->num_envs; ++i)
        actions[i] = random_action(action_seed, s->env_id_offset + (uint64_t)i, step_index);
    return 0;
}

int hostcheck_observe(const acas2d_params *p, const acas2d_state *s, float *obs)
{
    const DevParams P = make_dev_params(*p);
    const StatePtrs S = make_state_ptrs(*s);
    for (int64_t i = 0; i < S.B; ++i) observe_env(P, S, i, obs);
    return 0;
}

int hostcheck_inject(const acas2d_params *p, const acas2d_state *s, const double *player, const double *traffic,
                     const int32_t *steps, const double *total_reward)
{
    const DevParams P = make_dev_params(*p);
    const StatePtrs S = make_state_ptrs(*s);
    for (int64_t i = 0; i < S.B; ++i) inject_env(P, S, i, player, traffic, steps, total_reward);
    return 0;
}

int hostcheck_extract(const acas2d_params *p, const acas2d_state *s, double *player, double *traffic,
                      int32_t *steps, double *total_reward)
{
    const DevParams P = make_dev_params(*p);
    const StatePtrs S = make_state_ptrs(*s);
    for (int64_t i = 0; i < S.B; ++i) extract_env(P, S, i, player, traffic, steps, total_reward);
    return 0;
}

void hostcheck_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const U4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

double hostcheck_wrap360(double t) { return wrap360(t); }

void hostcheck_sincos_deg(double deg, double *s, double *c) { sincos_deg(deg, s, c); }

void hostcheck_policy_mean(const float *weights, const float *obs, int64_t B, float *mean)
{
    for (int64_t i = 0; i < B; ++i) mean[i] = policy_mean(weights, obs + 8 * i);
}

void hostcheck_policy_noise(uint64_t seed, uint64_t gid0, uint64_t step, int64_t B, float *eps)
{
    for (int64_t i = 0; i < B; ++i) eps[i] = policy_noise(seed, gid0 + (uint64_t)i, step);
}

}  // extern "C"
