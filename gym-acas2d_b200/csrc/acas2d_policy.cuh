// acas2d_policy.cuh -- the actor of the reference's trained agent, fused with the environment step.
//
// The reference trains stable-baselines3 PPO('MlpPolicy') on ACAS2D-v0 (gym_ACAS2D/training_main.py:44-52)
// and evaluates it with model.predict(obs, deterministic=True) (testing_main.py:74).  The actor saved in
// models/best_model_1048576_11/best_model.zip/policy.pth is
//     latent = tanh(W2 . tanh(W1 . obs + b1) + b2)        mlp_extractor.policy_net.{0,2}   8 -> 64 -> 64
//     mean   = W3 . latent + b3                            action_net                        64 -> 1
//     action ~ N(mean, exp(log_std)^2)  (deterministic: action = mean), clipped to the Box [-1, 1]
// all float32.  One thread evaluates it for one env straight from the env's own observation row and
// feeds the clipped action into step_env1 -- a closed-loop rollout never leaves the GPU.
//
// This first version is float32 on the CUDA cores (4.7 k FMA per env-step, weights broadcast from
// shared memory): exact against a torch float32 reference.  It is FMA/LDS-issue-bound, ~7x the cost of
// the env step itself; a tcgen05 version is the next step.
#pragma once

#include "acas2d_env.cuh"

namespace acas2d {

constexpr int kPolObs = 8, kPolHidden = 64;
// packed weight block: W1[64][8] | b1[64] | W2[64][64] | b2[64] | W3[64] | b3[1] (+3 pad)
constexpr int kPolW1 = 0, kPolB1 = kPolW1 + kPolHidden * kPolObs, kPolW2 = kPolB1 + kPolHidden,
              kPolB2 = kPolW2 + kPolHidden * kPolHidden, kPolW3 = kPolB2 + kPolHidden,
              kPolB3 = kPolW3 + kPolHidden, kPolFloats = kPolB3 + 4;

ACAS_HD float acas_tanhf(float x)
{
#if defined(__CUDA_ARCH__)
    // 1 - 2/(e^{2x}+1): two MUFU ops, |error| < 3e-7; saturates correctly for large |x|
    const float t = __expf(2.0f * x);
    return 1.0f - 2.0f * acas_rcpf(t + 1.0f);
#else
    return tanhf(x);
#endif
}

// Mean action of the actor for one observation row.  `w` is the packed block above (shared memory on
// the device: every lane reads the same address, a broadcast).
ACAS_HD float policy_mean(const float *w, const float *obs)
{
    float h1[kPolHidden];
#pragma unroll
    for (int j = 0; j < kPolHidden; ++j) {
        float acc = w[kPolB1 + j];
#pragma unroll
        for (int i = 0; i < kPolObs; ++i) acc = fmaf(w[kPolW1 + j * kPolObs + i], obs[i], acc);
        h1[j] = acas_tanhf(acc);
    }
    float mean = w[kPolB3];
#pragma unroll 1
    for (int j = 0; j < kPolHidden; ++j) {
        float acc = w[kPolB2 + j];
        const float *row = w + kPolW2 + j * kPolHidden;
#pragma unroll
        for (int i = 0; i < kPolHidden; ++i) acc = fmaf(row[i], h1[i], acc);
        mean = fmaf(w[kPolW3 + j], acas_tanhf(acc), mean);
    }
    return mean;
}

// Optional device-resident overrides of two by-value launch arguments, so that a captured rollout graph
// (T policy-step launches) can be replayed while the learner keeps changing log_std and the noise counter
// advances: log_std is read from *log_std (the learner's parameter block), the noise counter of the launch is
// step_offset + *step_base.  Null pointers: the by-value arguments are used as they are.
struct PolicyDyn { const float *log_std; const uint64_t *step_base; };

// Standard normal from one Philox block (Box-Muller), keyed like the other streams:
// key = noise_seed, counter = (global env id, step index, tag).
ACAS_HD float policy_noise(uint64_t noise_seed, uint64_t gid, uint64_t step_index)
{
    const U4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step_index,
                               (uint32_t)(step_index >> 32) ^ 0x9A055E5Du,
                               (uint32_t)noise_seed, (uint32_t)(noise_seed >> 32));
    const float u1 = ((float)(r.x >> 8) + 0.5f) * 5.9604644775390625e-08f;      // (0, 1)
    const float u2 = (float)(r.y >> 8) * 5.9604644775390625e-08f;               // [0, 1)
#if defined(__CUDA_ARCH__)
    return acas_sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
#else
    return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
#endif
}

}  // namespace acas2d
