// acas2d_ppo.cuh -- PPO learner kernels for the reference's agent (BASELINE config 5).
//
// The reference trains stable_baselines3.PPO('MlpPolicy', env, seed=13) (gym_ACAS2D/training_main.py:44-52);
// the hyper-parameters stored in models/best_model_1048576_11/best_model.zip/data are the SB3 1.1.0 defaults:
// separate 8 -> 64 -> 64 tanh actor and critic, state-independent log_std, clipped surrogate (clip 0.2) on
// minibatch-normalised advantages, 0.5 * MSE value loss, no entropy bonus, global grad-norm clip 0.5, Adam.
// SB3 is not vendored by the reference; what is restated here is SB3 1.1.0's published PPO.train() arithmetic
// (ppo.py there), and the numerics reference in tests/ is the same loss written in plain torch float32
// (gym_ACAS2D/ppo.py: reference_loss) differentiated by autograd.
//
// One gradient step = two kernels:
//   ppo_grad_kernel       forward + backward of BOTH networks over the minibatch: grid (G, 2), blockIdx.y
//                         picks actor or critic (their losses do not interact), each CTA walks 32-sample
//                         tiles, everything in shared memory, weight gradients accumulated in registers
//                         across its tiles, one partial-gradient row per CTA; the actor CTAs first take the
//                         mean / unbiased std of the minibatch's advantages (two passes, every CTA the same)
//   ppo_update_kernel     fixed-order sum of the partial rows, the data-parallel gradient exchange over
//                         NVLink peer memory (each rank publishes its 38 KB gradient in a peer-mapped,
//                         double-buffered block, signals its peers with release stores and sums all ranks'
//                         blocks in rank order -- bit-identical on every rank), global norm, clip, Adam:
//                         38 co-resident CTAs with a counter barrier between the phases
// plus, for callers that exchange gradients themselves (torch.distributed all-reduce), the same work split
// as ppo_grad_kernel -> ppo_reduce_kernel (-> all-reduce) -> ppo_adam_kernel.
// The matrices are 64-wide: a minibatch step is ~0.1 GFLOP, launch/latency-bound, so this is float32 on
// the CUDA cores (within 2e-4 of the torch float32 reference's gradients) rather than TF32 tensor-core code.
// Everything is deterministic: no floating-point atomics.
#pragma once

#include "acas2d_policy.cuh"
#include "acas2d_dev.cuh"

namespace acas2d {

constexpr int kPpoLogStd = 2 * kPolFloats;          // parameter block: actor | critic | log_std | 3 pad
constexpr int kPpoParams = ACAS2D_PPO_PARAM_FLOATS;
constexpr int kPpoPartial = ACAS2D_PPO_PARTIAL_FLOATS;
constexpr int kPpoMaxCtas = ACAS2D_PPO_MAX_CTAS;
constexpr int kPpoTile = 32, kPpoThreads = 256;     // gradient kernel: 32-sample tiles -> two CTAs per SM (16 warps) at
                                                    // a 4096-row minibatch.  Measured equal to 64-sample tiles (one
                                                    // CTA per SM, 12 % occupancy): the serial chain of phases bounds
                                                    // the kernel, not occupancy; kept for small minibatches
constexpr int kPpoTq = kPpoTile / 16;               // sample rows per thread in the 16 x 16 thread grid
constexpr int kPpoLps = kPpoThreads / kPpoTile;     // lanes per sample in the output-unit phase
constexpr int kPpoValTile = 64;                     // critic-forward kernel: 64-row tiles
constexpr int kPpoLd = 68;                          // shared-memory row stride of the 64-wide matrices (16-byte rows, bank-skewed)
constexpr int kPpoLdW1 = 12;                        // same for W1 rows (8 wide)
// partial row tail: dlog_std | sum pg term | sum value term | sum approx-kl term | clipped count
constexpr int kPpoStatBase = kPolFloats;
static_assert(kPpoParams == 2 * kPolFloats + 4 && kPpoPartial >= kPolFloats + 8, "PPO block sizes");

struct PpoBatch {
    const float *obs;        // [n][8]
    const float *actions;    // [n]   unclipped samples
    const float *old_logp;   // [n]
    const float *adv;        // [n]
    const float *ret;        // [n]
    const int64_t *idx;      // [mb] rows of this minibatch (nullptr: 0..mb-1)
    int64_t mb;
};

// shared-memory plan of ppo_grad_kernel (floats)
constexpr int kPpoSmW2 = 0, kPpoSmW2T = kPpoSmW2 + 64 * kPpoLd, kPpoSmW1 = kPpoSmW2T + 64 * kPpoLd,
              kPpoSmB1 = kPpoSmW1 + 64 * kPpoLdW1, kPpoSmB2 = kPpoSmB1 + 64, kPpoSmW3 = kPpoSmB2 + 64,
              kPpoSmX = kPpoSmW3 + 64, kPpoSmH1 = kPpoSmX + kPpoTile * 8, kPpoSmC = kPpoSmH1 + kPpoTile * kPpoLd,
              kPpoSmDy = kPpoSmC + kPpoTile * kPpoLd, kPpoSmRed = kPpoSmDy + kPpoTile,
              kPpoSmFloats = kPpoSmRed + 2 * kPpoThreads;
constexpr int kPpoSmemBytes = kPpoSmFloats * 4;

#if defined(__CUDACC__)

// acc[q][r] += sum_{k<K} A[(mg + 16 q) * lda + k] * B[(ng + 16 r) * ldb + k], q < TQ, r < 4: both operands are read as
// 128-bit words along k.  Rows are dealt to the 16 x 16 thread grid interleaved (row = group + 16 q), so
// the 8 threads of a quarter warp read 8 consecutive rows of B -- with a row stride of 68 (or 12) floats
// those fall in 8 distinct 4-bank groups -- while A is a 2-address broadcast.
template <int K, int TQ>
__device__ __forceinline__ void ppo_gemm_nt(const float *A, int lda, const float *B, int ldb, int mg, int ng,
                                            float (&acc)[TQ][4])
{
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 a[TQ], b[4];
#pragma unroll
        for (int q = 0; q < TQ; ++q) a[q] = *(const float4 *)(A + (mg + 16 * q) * lda + k);
#pragma unroll
        for (int r = 0; r < 4; ++r) b[r] = *(const float4 *)(B + (ng + 16 * r) * ldb + k);
#pragma unroll
        for (int q = 0; q < TQ; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float s = acc[q][r];
                s = fmaf(a[q].x, b[r].x, s); s = fmaf(a[q].y, b[r].y, s);
                s = fmaf(a[q].z, b[r].z, s); s = fmaf(a[q].w, b[r].w, s);
                acc[q][r] = s;
            }
    }
}

__device__ __forceinline__ float ppo_block_sum(float v, float *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.0f;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];            // same order in every thread
    return s;
}

__device__ __forceinline__ double ppo_block_sum_d(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(kPpoThreads, 2)
ppo_grad_kernel(const float *__restrict__ params, const PpoBatch b, const int normalize_advantage,
                const float clip_range, const float vf_coef, float *__restrict__ partials, int32_t *adam_step)
{
    extern __shared__ __align__(16) float sm[];
    float *sW2 = sm + kPpoSmW2, *sW2T = sm + kPpoSmW2T, *sW1 = sm + kPpoSmW1, *sb1 = sm + kPpoSmB1,
          *sb2 = sm + kPpoSmB2, *sw3 = sm + kPpoSmW3, *sX = sm + kPpoSmX, *sH1 = sm + kPpoSmH1, *sC = sm + kPpoSmC,
          *sdy = sm + kPpoSmDy, *sred = sm + kPpoSmRed;
    const int t = threadIdx.x;
    const int net = blockIdx.y;                                   // 0 = actor, 1 = critic
    const float *w = params + net * kPolFloats;
    const int mg = t >> 4, ng = t & 15;                           // 16 x 16 thread grid of the 64 x 64 products

    // Gathers run one tile ahead of the arithmetic (and the first tile's ahead of the weight staging and the
    // advantage statistics): observation-row halves by threads 0 .. 2 * tile - 1, the per-sample scalars by lane 0
    // of each sample's lane group.  index -> row is two dependent trips to L2 / HBM otherwise on the critical path.
    float4 pre_obs = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float pre_act = 0.0f, pre_logp = 0.0f, pre_adv = 0.0f, pre_ret = 0.0f;
    auto prefetch = [&](const int64_t tile) {
        if (t < 2 * kPpoTile) {
            const int64_t k = tile * kPpoTile + (t >> 1);
            pre_obs = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (k < b.mb) pre_obs = ((const float4 *)b.obs)[2 * (b.idx ? b.idx[k] : k) + (t & 1)];
        }
        if (t % kPpoLps == 0) {
            const int64_t k = tile * kPpoTile + t / kPpoLps;
            if (k < b.mb) {
                const int64_t row = b.idx ? b.idx[k] : k;
                if (net == 0) { pre_act = b.actions[row]; pre_logp = b.old_logp[row]; pre_adv = b.adv[row]; }
                else pre_ret = b.ret[row];
            }
        }
    };
    prefetch(blockIdx.x);

    // SB3 ppo.py: advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8) over the minibatch
    // (std with Bessel's correction).  Every actor CTA computes it itself, in the same order.
    float adv_mean = 0.0f, adv_scale = 1.0f;
    if (normalize_advantage && net == 0) {                                       // block-uniform
        double *dred = (double *)sred;
        constexpr int kKeep = 16;                          // minibatches up to 4096 rows are gathered exactly once
        float a[kKeep];
#pragma unroll
        for (int u = 0; u < kKeep; ++u) {                  // all gathers (index, then value) in flight together
            const int64_t k = u * kPpoThreads + t;
            a[u] = (k < b.mb) ? b.adv[b.idx ? b.idx[k] : k] : 0.0f;
        }
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < kKeep; ++u) s += (double)a[u];
        for (int64_t k = (int64_t)kKeep * kPpoThreads + t; k < b.mb; k += kPpoThreads) s += (double)b.adv[b.idx ? b.idx[k] : k];
        const double mean = ppo_block_sum_d(s, dred) / (double)b.mb;
        double q = 0.0;
#pragma unroll
        for (int u = 0; u < kKeep; ++u) {
            const double d = (double)a[u] - mean;
            q += (u * kPpoThreads + t < b.mb) ? d * d : 0.0;
        }
        for (int64_t k = (int64_t)kKeep * kPpoThreads + t; k < b.mb; k += kPpoThreads) {
            const double d = (double)b.adv[b.idx ? b.idx[k] : k] - mean;
            q += d * d;
        }
        const double var = ppo_block_sum_d(q, dred) / (double)(b.mb > 1 ? b.mb - 1 : 1);
        adv_mean = (float)mean;
        adv_scale = 1.0f / ((float)sqrt(var) + 1e-8f);
    }

    // Programmatic dependent launch: everything above reads rollout data only and may overlap the tail of the update
    // kernel of the previous gradient step (38 CTAs: most SMs are free while it runs).  The parameters it writes, the
    // partial rows it reads and the step counter are touched only from here on.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    {   // W2[j][i], rows as SB3 stores them, and its transpose; W1: 128-bit loads, all in flight
        float4 v2[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) v2[it] = ((const float4 *)(w + kPolW2))[t + kPpoThreads * it];
        float4 v1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (t < 128) v1 = ((const float4 *)(w + kPolW1))[t];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int e = 4 * (t + kPpoThreads * it), j = e >> 6, i = e & 63;
            *(float4 *)(sW2 + j * kPpoLd + i) = v2[it];
            sW2T[(i + 0) * kPpoLd + j] = v2[it].x; sW2T[(i + 1) * kPpoLd + j] = v2[it].y;
            sW2T[(i + 2) * kPpoLd + j] = v2[it].z; sW2T[(i + 3) * kPpoLd + j] = v2[it].w;
        }
        if (t < 128) *(float4 *)(sW1 + (t >> 1) * kPpoLdW1 + 4 * (t & 1)) = v1;
    }
    if (t < 64) { sb1[t] = w[kPolB1 + t]; sb2[t] = w[kPolB2 + t]; sw3[t] = w[kPolW3 + t]; }
    const float b3 = w[kPolB3];
    const float log_std = params[kPpoLogStd];
    const float inv_var = __expf(-2.0f * log_std);
    const float inv_mb = 1.0f / (float)b.mb;
    if (adam_step && blockIdx.x == 0 && net == 0 && t == 0) *adam_step += 1;     // nobody reads it in this kernel

    // gradient accumulators, kept in registers across this CTA's tiles
    float gW2[4][4] = {};                    // dW2[4 jg + q][4 ig + r]   (jg = t >> 4, ig = t & 15)
    float gW1[2] = {0.0f, 0.0f};             // dW1[t >> 2][2 (t & 3) + {0, 1}]
    float gb1 = 0.0f, gb2 = 0.0f, gw3 = 0.0f, gb3 = 0.0f;        // t < 64: unit t; gb3: t == 0
    float st_dls = 0.0f, st_pg = 0.0f, st_v = 0.0f, st_kl = 0.0f, st_clip = 0.0f;
    __syncthreads();

    const int64_t ntiles = (b.mb + kPpoTile - 1) / kPpoTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- this tile's gathered rows (zero rows past the end of the minibatch); next tile's gathers take off
        if (t < 2 * kPpoTile) *(float4 *)(sX + (t >> 1) * 8 + 4 * (t & 1)) = pre_obs;
        const float cur_act = pre_act, cur_logp = pre_logp, cur_adv = pre_adv, cur_ret = pre_ret;
        if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x);
        __syncthreads();

        // ---- layer 1: H1[s][j] = tanh(b1[j] + sum_c X[s][c] W1[j][c])
        {
            float acc[kPpoTq][4] = {};
            ppo_gemm_nt<8, kPpoTq>(sX, 8, sW1, kPpoLdW1, mg, ng, acc);
#pragma unroll
            for (int q = 0; q < kPpoTq; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    sH1[(mg + 16 * q) * kPpoLd + ng + 16 * r] = acas_tanhf(acc[q][r] + sb1[ng + 16 * r]);
        }
        __syncthreads();

        // ---- layer 2: H2[s][j] = tanh(b2[j] + sum_i H1[s][i] W2[j][i])  -> sC
        {
            float acc[kPpoTq][4] = {};
            ppo_gemm_nt<64, kPpoTq>(sH1, kPpoLd, sW2, kPpoLd, mg, ng, acc);
#pragma unroll
            for (int q = 0; q < kPpoTq; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    sC[(mg + 16 * q) * kPpoLd + ng + 16 * r] = acas_tanhf(acc[q][r] + sb2[ng + 16 * r]);
        }
        __syncthreads();

        // ---- output unit and the loss derivative of each sample (kPpoLps lanes per sample)
        {
            constexpr int kJ = 64 / kPpoLps;
            const int s = t / kPpoLps, part = t % kPpoLps;
            float y = 0.0f;
#pragma unroll
            for (int jj = 0; jj < kJ; ++jj) y = fmaf(sw3[part * kJ + jj], sC[s * kPpoLd + part * kJ + jj], y);
#pragma unroll
            for (int o = 1; o < kPpoLps; o <<= 1) y += __shfl_xor_sync(kFull, y, o);
            y += b3;
            if (part == 0) {
                const int64_t k = tile * kPpoTile + s;
                float dy = 0.0f;
                if (k < b.mb) {
                    if (net == 0) {
                        // SB3 ppo.py train(): ratio = exp(log_prob - old_log_prob); policy_loss =
                        // -mean(min(adv * ratio, adv * clamp(ratio, 1 - clip, 1 + clip)))
                        const float an = (cur_adv - adv_mean) * adv_scale;
                        const float diff = cur_act - y;
                        const float z2 = diff * diff * inv_var;
                        const float logp = -0.5f * z2 - log_std - 0.9189385332046727f;
                        const float lr = logp - cur_logp;
                        const float ratio = __expf(lr);
                        const float rc = fminf(fmaxf(ratio, 1.0f - clip_range), 1.0f + clip_range);
                        st_pg += -fminf(an * ratio, an * rc);
                        st_kl += (ratio - 1.0f) - lr;
                        st_clip += (fabsf(ratio - 1.0f) > clip_range) ? 1.0f : 0.0f;
                        const bool cut = (an > 0.0f && ratio > 1.0f + clip_range) || (an < 0.0f && ratio < 1.0f - clip_range);
                        const float g = cut ? 0.0f : -an * ratio * inv_mb;        // d loss / d logp
                        dy = g * diff * inv_var;                                    // d logp / d mean
                        st_dls += g * (z2 - 1.0f);                                  // d logp / d log_std
                    } else {
                        // value_loss = F.mse_loss(returns, values), weighted by vf_coef
                        const float d = y - cur_ret;
                        st_v += d * d;
                        dy = vf_coef * 2.0f * d * inv_mb;
                    }
                }
                sdy[s] = dy;
            }
        }
        __syncthreads();

        // ---- dw3, db3; dZ2 = dy * w3 * (1 - H2^2) in place of H2; db2 (column sums: 4 x kPpoTile/4 samples per unit)
        {
            const int j = t & 63, part = t >> 6;
            const float w3j = sw3[j];
            float pw3 = 0.0f, pb2 = 0.0f;
#pragma unroll 4
            for (int s = part * (kPpoTile / 4); s < (part + 1) * (kPpoTile / 4); ++s) {
                const float h = sC[s * kPpoLd + j], dyv = sdy[s];
                pw3 = fmaf(dyv, h, pw3);
                const float dz = dyv * w3j * (1.0f - h * h);
                sC[s * kPpoLd + j] = dz;
                pb2 += dz;
            }
            sred[t] = pw3;
            sred[kPpoThreads + t] = pb2;
        }
        __syncthreads();
        if (t < 64) {
            gw3 += (sred[t] + sred[64 + t]) + (sred[128 + t] + sred[192 + t]);
            gb2 += (sred[kPpoThreads + t] + sred[kPpoThreads + 64 + t]) +
                   (sred[kPpoThreads + 128 + t] + sred[kPpoThreads + 192 + t]);
        }
        if (t < 32) {
            float v = sdy[t] + (kPpoTile > 32 ? sdy[(t + 32) % kPpoTile] : 0.0f);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (t == 0) gb3 += v;
        }

        // ---- dW2[j][i] += sum_s dZ2[s][j] H1[s][i]: outer products, contiguous 4 x 4 ownership, two
        //      128-bit loads per sample (dZ2 row: broadcast; H1 row: consecutive words)
        {
            const int j0 = 4 * mg, i0 = 4 * ng;
#pragma unroll 4
            for (int s = 0; s < kPpoTile; ++s) {
                const float4 a = *(const float4 *)(sC + s * kPpoLd + j0);
                const float4 h = *(const float4 *)(sH1 + s * kPpoLd + i0);
                gW2[0][0] = fmaf(a.x, h.x, gW2[0][0]); gW2[0][1] = fmaf(a.x, h.y, gW2[0][1]);
                gW2[0][2] = fmaf(a.x, h.z, gW2[0][2]); gW2[0][3] = fmaf(a.x, h.w, gW2[0][3]);
                gW2[1][0] = fmaf(a.y, h.x, gW2[1][0]); gW2[1][1] = fmaf(a.y, h.y, gW2[1][1]);
                gW2[1][2] = fmaf(a.y, h.z, gW2[1][2]); gW2[1][3] = fmaf(a.y, h.w, gW2[1][3]);
                gW2[2][0] = fmaf(a.z, h.x, gW2[2][0]); gW2[2][1] = fmaf(a.z, h.y, gW2[2][1]);
                gW2[2][2] = fmaf(a.z, h.z, gW2[2][2]); gW2[2][3] = fmaf(a.z, h.w, gW2[2][3]);
                gW2[3][0] = fmaf(a.w, h.x, gW2[3][0]); gW2[3][1] = fmaf(a.w, h.y, gW2[3][1]);
                gW2[3][2] = fmaf(a.w, h.z, gW2[3][2]); gW2[3][3] = fmaf(a.w, h.w, gW2[3][3]);
            }
        }
        __syncthreads();                                   // all reads of H1 done before it is overwritten

        // ---- dZ1[s][i] = (sum_j dZ2[s][j] W2[j][i]) * (1 - H1[s][i]^2), in place of H1
        {
            float acc[kPpoTq][4] = {};
            ppo_gemm_nt<64, kPpoTq>(sC, kPpoLd, sW2T, kPpoLd, mg, ng, acc);
#pragma unroll
            for (int q = 0; q < kPpoTq; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float *p = sH1 + (mg + 16 * q) * kPpoLd + ng + 16 * r;
                    const float h = *p;
                    *p = acc[q][r] * (1.0f - h * h);
                }
        }
        __syncthreads();

        // ---- db1 (column sums of dZ1) and dW1[j][c] += sum_s dZ1[s][j] X[s][c]
        {
            const int j = t & 63, part = t >> 6;
            float pb1 = 0.0f;
#pragma unroll 4
            for (int s = part * (kPpoTile / 4); s < (part + 1) * (kPpoTile / 4); ++s) pb1 += sH1[s * kPpoLd + j];
            sred[t] = pb1;
            const int jw = t >> 2, c0 = 2 * (t & 3);
#pragma unroll 4
            for (int s = 0; s < kPpoTile; ++s) {
                const float d = sH1[s * kPpoLd + jw];
                const float2 x = *(const float2 *)(sX + s * 8 + c0);
                gW1[0] = fmaf(d, x.x, gW1[0]);
                gW1[1] = fmaf(d, x.y, gW1[1]);
            }
        }
        __syncthreads();
        if (t < 64) gb1 += (sred[t] + sred[64 + t]) + (sred[128 + t] + sred[192 + t]);
        __syncthreads();                                   // sred / sX / sH1 are rewritten by the next tile
    }

    // ---- this CTA's partial row
    float *row = partials + ((size_t)net * gridDim.x + blockIdx.x) * kPpoPartial;
    {
        const int j0 = 4 * mg, i0 = 4 * ng;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *(float4 *)(row + kPolW2 + (j0 + q) * 64 + i0) = make_float4(gW2[q][0], gW2[q][1], gW2[q][2], gW2[q][3]);
        *(float2 *)(row + kPolW1 + (t >> 2) * 8 + 2 * (t & 3)) = make_float2(gW1[0], gW1[1]);
    }
    if (t < 64) { row[kPolB1 + t] = gb1; row[kPolB2 + t] = gb2; row[kPolW3 + t] = gw3; }
    if (t == 0) { row[kPolB3] = gb3; row[kPolB3 + 1] = 0.0f; row[kPolB3 + 2] = 0.0f; row[kPolB3 + 3] = 0.0f; }
    {   // the five per-sample statistics: warp shuffles, then lane 0 of every warp -> 8 x 5 words, summed in warp order
        float st[5] = {st_dls, st_pg, st_v, st_kl, st_clip};
#pragma unroll
        for (int f = 0; f < 5; ++f)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) st[f] += __shfl_xor_sync(kFull, st[f], o);
        if ((t & 31) == 0)
#pragma unroll
            for (int f = 0; f < 5; ++f) sred[(t >> 5) * 5 + f] = st[f];
        __syncthreads();
        if (t < 5) {
            float v = 0.0f;
            for (int wp = 0; wp < kPpoThreads / 32; ++wp) v += sred[wp * 5 + t];
            row[kPpoStatBase + t] = v;
        }
    }
}

// Sum of the CTAs' partial rows c = first, first + stride, ... for parameter p (fixed order; 8 loads in flight).
__device__ __forceinline__ float ppo_reduce_param(const float *__restrict__ partials, const int ctas, const int p,
                                                  const int first, const int stride)
{
    if (p > kPpoLogStd) return 0.0f;
    const int net = (p < 2 * kPolFloats && p >= kPolFloats) ? 1 : 0;
    const int q = (p == kPpoLogStd) ? kPpoStatBase : p - net * kPolFloats;
    const float *src = partials + (size_t)net * ctas * kPpoPartial + q;
    float s = 0.0f;
    for (int c0 = first; c0 < ctas; c0 += 8 * stride) {
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * stride;
            a[u] = (c < ctas) ? __ldcg(src + (size_t)c * kPpoPartial) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s += a[u];
    }
    return s;
}

// loss_stats: 0 policy loss, 1 value loss (MSE), 2 approx KL, 3 clip fraction (SB3's logged quantities);
// 4 = gradient norm before clipping.  Called by all 32 lanes of one warp for statistic f = 1..4
// (pg / value / kl / clip): lanes take rows lane, lane + 32, ...; xor-shuffle tree (fixed order).
__device__ __forceinline__ void ppo_write_loss_stat(const float *__restrict__ partials, const int ctas, const int f,
                                                    const float inv_mb, float *__restrict__ loss_stats)
{
    const int net = (f == 2) ? 1 : 0, lane = threadIdx.x & 31;
    float s = 0.0f;
    for (int c = lane; c < ctas; c += 32) s += __ldcg(partials + ((size_t)net * ctas + c) * kPpoPartial + kPpoStatBase + f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (lane == 0) loss_stats[f - 1] = s * inv_mb;
}

__global__ void __launch_bounds__(256)
ppo_reduce_kernel(const float *__restrict__ partials, const int ctas, const float ent_coef, const float inv_mb,
                  float *__restrict__ grad, float *__restrict__ loss_stats)
{
    const int p = blockIdx.x * 256 + threadIdx.x;
    // entropy of N(., sigma) = log_std + const; the loss has -ent_coef * entropy
    if (p < kPpoParams) grad[p] = ppo_reduce_param(partials, ctas, p, 0, 1) - (p == kPpoLogStd ? ent_coef : 0.0f);
    if (blockIdx.x == 0 && threadIdx.x < 128 && loss_stats)
        ppo_write_loss_stat(partials, ctas, 1 + (threadIdx.x >> 5), inv_mb, loss_stats);
}

struct PpoAdam { float lr, beta1, beta2, eps, max_grad_norm; };

// torch.optim.Adam(lr, betas, eps).step() for one parameter, g already clipped.
__device__ __forceinline__ void ppo_adam_param(float *params, float *m, float *v, const int p, const float g,
                                               const int step, const PpoAdam &h)
{
    const float mn = h.beta1 * m[p] + (1.0f - h.beta1) * g;
    const float vn = h.beta2 * v[p] + (1.0f - h.beta2) * g * g;
    m[p] = mn; v[p] = vn;
    const float bc1 = 1.0f - powf(h.beta1, (float)step), bc2 = 1.0f - powf(h.beta2, (float)step);
    const float denom = sqrtf(vn) / sqrtf(bc2) + h.eps;
    params[p] -= (h.lr / bc1) * (mn / denom);
}

// torch.nn.utils.clip_grad_norm_(max_norm) followed by torch.optim.Adam(lr, betas, eps).step(), on
// grad * grad_scale (1/world after a SUM all-reduce).  Every CTA recomputes the global norm in the same
// order, so all of them apply the same clip coefficient without a grid-wide barrier.
__global__ void __launch_bounds__(256)
ppo_adam_kernel(float *__restrict__ params, const float *__restrict__ grad, const float grad_scale,
                float *__restrict__ m, float *__restrict__ v, const int32_t *__restrict__ adam_step,
                const PpoAdam h, float *__restrict__ loss_stats)
{
    __shared__ float red[8];
    float q = 0.0f;
    for (int p = threadIdx.x; p < kPpoParams; p += 256) {
        const float g = grad[p] * grad_scale;
        q = fmaf(g, g, q);
    }
    const float norm = sqrtf(ppo_block_sum(q, red));
    float coef = (h.max_grad_norm > 0.0f) ? h.max_grad_norm / (norm + 1e-6f) : 1.0f;
    coef = fminf(coef, 1.0f);
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss_stats) loss_stats[4] = norm;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p > kPpoLogStd) return;
    ppo_adam_param(params, m, v, p, grad[p] * grad_scale * coef, *adam_step, h);
}

// ---------------------------------------------------------------- fused update (+ NVLink gradient exchange)
constexpr int kPpoUpdateCtas = (kPpoParams + 255) / 256;          // 38: co-resident on any B200
constexpr int kPpoMaxRanks = ACAS2D_PPO_MAX_RANKS;
constexpr int kPpoXFlags = 2 * kPpoParams;                        // exchange block: grad[2][9612] | uint32 flags[16]

struct PpoPeers { float *block[kPpoMaxRanks]; };                  // every rank's exchange block, peer-mapped

__device__ __forceinline__ void st_release_sys_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float *p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Barrier over the CTAs of ONE launch (all co-resident: 38 CTAs): arrivals are counted on a monotonic 64-bit
// counter, `target` = arrivals expected once every CTA has reached this barrier instance.
// SYSTEM: the fence that publishes this CTA's stores (ordered before it by the bar.sync) is system-scope, so
// that a peer GPU which later observes this rank's release flag also observes them.
// The launch is COOPERATIVE (cudaLaunchAttributeCooperative: the driver either makes all CTAs resident together or
// fails the launch), and every spin is BOUNDED: a barrier or a peer flag that does not come within ~10 s -- a dead
// or late peer rank -- records ACAS2D_PPO_ERR_* in sync[1] and lets the kernel finish instead of hanging the GPU;
// the host reads the word back (FusedLearner.check()).
constexpr unsigned kPpoSpinLimit = 1u << 25;
template <bool SYSTEM = false>
__device__ __forceinline__ void ppo_grid_barrier(unsigned long long *counter, const unsigned long long target, int32_t *err)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        if (SYSTEM) __threadfence_system(); else __threadfence();
        atomicAdd(counter, 1ull);
        unsigned spins = 0;
        while (ld_acquire_gpu_u64(counter) < target) {
            if (++spins > kPpoSpinLimit) { atomicExch(err, ACAS2D_PPO_ERR_BARRIER); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// sync: int32[4] = { Adam step (incremented by ppo_grad_kernel), error word (0 = ok), 64-bit barrier-arrival counter }.  The
// counter is only ever advanced by this kernel, by exactly gridDim.x * barriers_per_step per launch, so its value
// at launch (rounded down: early CTAs of the same launch may already have arrived) numbers the launch -- the
// barrier targets, the exchange-buffer parity and the peer flags all derive from that sequence number, not from
// the Adam step count, and stay consistent whatever else the caller does with `sync[0]`.
constexpr int kPpoUpdateThreads = 1024;

__global__ void __launch_bounds__(kPpoUpdateThreads)
ppo_update_kernel(float *__restrict__ params, const float *__restrict__ partials, const int ctas, const float ent_coef,
                  const float inv_mb, float *__restrict__ norm_parts, const __grid_constant__ PpoPeers peers, const int rank, const int world,
                  float *__restrict__ m, float *__restrict__ v, int32_t *sync, const PpoAdam h,
                  float *__restrict__ loss_stats, float *__restrict__ grad_out)
{
    __shared__ float red[32];
    __shared__ unsigned long long s_before;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");       // the next step's gradient kernel may begin its gathers
    const int t = threadIdx.x, sub = t & 3, p = blockIdx.x * 256 + (t >> 2);
    const bool owner = sub == 0 && p < kPpoParams;
    const int step = sync[0];
    unsigned long long *arrivals = (unsigned long long *)(sync + 2);
    const unsigned per_launch = gridDim.x * (world > 1 ? 2u : 1u);
    if (t == 0) s_before = (ld_acquire_gpu_u64(arrivals) / per_launch) * per_launch;
    __syncthreads();
    const unsigned long long arrivals_before = s_before;
    const unsigned seq = (unsigned)(arrivals_before / per_launch) + 1u;       // 1, 2, ...: this launch's number

    // 1. this rank's gradient: fixed-order sum of the partial rows (all four lanes end with the same value)
    float g = ppo_reduce_param(partials, ctas, p, sub, 4);
    g += __shfl_xor_sync(kFull, g, 1);
    g += __shfl_xor_sync(kFull, g, 2);
    if (p == kPpoLogStd) g -= ent_coef;            // entropy of N(., sigma) = log_std + const; loss has -ent_coef * entropy
    if (blockIdx.x == gridDim.x - 1 && t >= kPpoUpdateThreads - 128 && loss_stats)
        ppo_write_loss_stat(partials, ctas, 1 + ((t >> 5) & 3), inv_mb, loss_stats);

    // 2. data-parallel exchange over peer memory: publish, signal, wait, sum in rank order
    if (world > 1) {
        const int half = (int)(seq & 1u) * kPpoParams;      // double-buffered: a fast rank's next step cannot overwrite
        float *mine = peers.block[rank];                   // what a slow peer is still reading
        if (owner) mine[half + p] = g;
        ppo_grid_barrier<true>(arrivals, arrivals_before + gridDim.x, sync + 1);    // the whole gradient of this rank is published
        if (blockIdx.x == 0 && t < world)
            st_release_sys_u32((unsigned *)(peers.block[t] + kPpoXFlags) + rank, seq);
        if (t < world) {
            const unsigned *flag = (const unsigned *)(mine + kPpoXFlags) + t;
            unsigned spins = 0;
            while ((int)(ld_acquire_sys_u32(flag) - seq) < 0) {
                if (++spins > kPpoSpinLimit) { atomicExch(sync + 1, ACAS2D_PPO_ERR_PEER); break; }
                __nanosleep(64);
            }
        }
        __syncthreads();
        if (owner) {
            float a[kPpoMaxRanks];                         // all ranks' words in flight at once (one NVLink round trip),
#pragma unroll
            for (int r = 0; r < kPpoMaxRanks; ++r)         // then summed in rank order
                a[r] = (r < world) ? ld_relaxed_sys_f32(peers.block[r] + half + p) : 0.0f;
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < kPpoMaxRanks; ++r) s += a[r];
            g = s * (1.0f / (float)world);
        }
    }
    if (grad_out && owner) grad_out[p] = g;

    // 3. global norm: one partial per CTA, barrier, every CTA sums the 38 partials in the same order
    const float q = ppo_block_sum(owner ? g * g : 0.0f, red);
    if (t == 0) norm_parts[blockIdx.x] = q;
    ppo_grid_barrier(arrivals, arrivals_before + per_launch, sync + 1);
    if (t < 32) {                                          // 38 partials: lanes take c and c + 32, fixed shuffle tree
        float x = 0.0f;
        for (int c = t; c < (int)gridDim.x; c += 32) x += __ldcg(norm_parts + c);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
        if (t == 0) red[0] = x;
    }
    __syncthreads();
    const float total = red[0];
    if (!owner || p > kPpoLogStd) return;
    const float norm = sqrtf(total);
    float coef = (h.max_grad_norm > 0.0f) ? h.max_grad_norm / (norm + 1e-6f) : 1.0f;
    coef = fminf(coef, 1.0f);
    if (p == 0 && loss_stats) loss_stats[4] = norm;

    // 4. Adam
    ppo_adam_param(params, m, v, p, g * coef, step, h);
}

// Critic forward over n observation rows (the value head of SB3's MlpPolicy: mlp_extractor.value_net +
// value_net): the forward half of ppo_grad_kernel, 64-row tiles, persistent CTAs.
constexpr int kPpoValSmFloats = 64 * kPpoLd + 64 * kPpoLdW1 + 3 * 64 + kPpoValTile * 8 + 2 * kPpoValTile * kPpoLd;
constexpr int kPpoValSmemBytes = kPpoValSmFloats * 4;

__global__ void __launch_bounds__(kPpoThreads, 2)
ppo_values_kernel(const float *__restrict__ params, const float *__restrict__ obs, const int64_t n, float *__restrict__ values)
{
    extern __shared__ __align__(16) float sm[];
    float *sW2 = sm, *sW1 = sW2 + 64 * kPpoLd, *sb1 = sW1 + 64 * kPpoLdW1, *sb2 = sb1 + 64, *sw3 = sb2 + 64,
          *sX = sw3 + 64, *sH1 = sX + kPpoValTile * 8, *sC = sH1 + kPpoValTile * kPpoLd;
    const int t = threadIdx.x, mg = t >> 4, ng = t & 15;
    const float *w = params + kPolFloats;
    for (int e = t; e < 64 * 64; e += kPpoThreads) sW2[(e >> 6) * kPpoLd + (e & 63)] = w[kPolW2 + e];
    for (int e = t; e < 64 * 8; e += kPpoThreads) sW1[(e >> 3) * kPpoLdW1 + (e & 7)] = w[kPolW1 + e];
    if (t < 64) { sb1[t] = w[kPolB1 + t]; sb2[t] = w[kPolB2 + t]; sw3[t] = w[kPolW3 + t]; }
    const float b3 = w[kPolB3];
    __syncthreads();
    const int64_t ntiles = (n + kPpoValTile - 1) / kPpoValTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (t < 2 * kPpoValTile) {
            const int64_t k = tile * kPpoValTile + (t >> 1);
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (k < n) v = __ldcs((const float4 *)obs + 2 * k + (t & 1));
            *(float4 *)(sX + (t >> 1) * 8 + 4 * (t & 1)) = v;
        }
        __syncthreads();
        {
            float acc[4][4] = {};
            ppo_gemm_nt<8, 4>(sX, 8, sW1, kPpoLdW1, mg, ng, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    sH1[(mg + 16 * q) * kPpoLd + ng + 16 * r] = acas_tanhf(acc[q][r] + sb1[ng + 16 * r]);
        }
        __syncthreads();
        {
            float acc[4][4] = {};
            ppo_gemm_nt<64, 4>(sH1, kPpoLd, sW2, kPpoLd, mg, ng, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    sC[(mg + 16 * q) * kPpoLd + ng + 16 * r] = acas_tanhf(acc[q][r] + sb2[ng + 16 * r]);
        }
        __syncthreads();
        {
            const int s = t >> 2, part = t & 3;
            float y = 0.0f;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) y = fmaf(sw3[part * 16 + jj], sC[s * kPpoLd + part * 16 + jj], y);
            y += __shfl_xor_sync(kFull, y, 1);
            y += __shfl_xor_sync(kFull, y, 2);
            const int64_t k = tile * kPpoValTile + s;
            if (part == 0 && k < n) values[k] = y + b3;
        }
        __syncthreads();
    }
}

// Generalised advantage estimation over a [T][B] rollout (SB3 RolloutBuffer.compute_returns_and_advantage):
// values[T + 1][B] (row T bootstraps), dones[t][b] = the step t ended its episode.  One thread per env.
__global__ void __launch_bounds__(256)
ppo_gae_kernel(const float *__restrict__ rewards, const uint8_t *__restrict__ dones, const float *__restrict__ values,
               const int T, const int64_t B, const float gamma, const float lam, float *__restrict__ adv,
               float *__restrict__ ret)
{
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= B) return;
    float last = 0.0f, vnext = values[(int64_t)T * B + e];
    for (int t = T - 1; t >= 0; --t) {
        const int64_t k = (int64_t)t * B + e;
        const float nonterminal = dones[k] ? 0.0f : 1.0f;
        const float vt = values[k];
        const float delta = rewards[k] + gamma * vnext * nonterminal - vt;
        last = delta + gamma * lam * nonterminal * last;
        adv[k] = last;
        ret[k] = last + vt;
        vnext = vt;
    }
}

#endif  // __CUDACC__

}  // namespace acas2d
