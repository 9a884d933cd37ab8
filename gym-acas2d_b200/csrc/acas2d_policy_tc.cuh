// acas2d_policy_tc.cuh -- the actor MLP on the 5th-generation tensor cores (tcgen05 + TMEM), fused
// with the environment step.  Device-only (included by acas2d_kernels.cu).
//
// Work per CTA tile: 128 envs (M = 128 = one TMEM lane per env).
//   layer 1   D1[128 x 64] = OBS[128 x 8]  . W1^T     one  tcgen05.mma kind::tf32  (K = 8)
//   layer 2   D2[128 x 64] = H1 [128 x 64] . W2^T     eight tcgen05.mma kind::tf32 (K = 8 each)
//   layer 3   mean = w3 . tanh(D2 + b2) + b3          64 FMAs per thread in the epilogue
// Operands live in shared memory in the canonical K-major, SWIZZLE_NONE ("interleaved") UMMA layout:
// 16-byte chunks along K, chunk c of row r at  c * LBO + (r / 8) * SBO + (r % 8) * 16  with
// SBO = 128 and LBO = rows * 16 -- i.e. chunk c of row r sits at c * (rows*16) + r * 16, so thread r
// writes its own row with conflict-free 16-byte stores.  Accumulators live in TMEM (128 columns per
// CTA: D1 | D2) and come back with tcgen05.ld.32x32b, which hands thread r exactly row r: the
// bias + tanh epilogue, the action sample and the env step of env r all stay in thread r.
// TF32 keeps 10 mantissa bits of the activations / weights (fp32 accumulate); together with
// tanh.approx the action mean agrees with the fp32 reference to ~2e-3 (tests/test_policy.py), far
// inside the exploration noise exp(log_std) ~ 0.2 of the agent.
#pragma once

#include "acas2d_dev.cuh"
#include "acas2d_policy.cuh"

namespace acas2d {

constexpr int kTcTile = 128;                        // envs per tile == threads per CTA
constexpr int kTcA2 = 0;                            // H1   : 128 rows x 64 tf32 = 16 chunks x 2048 B
constexpr int kTcA1 = kTcA2;                        // OBS  : 128 rows x 8  tf32 = 2 chunks x 2048 B; aliases H1's
                                                    //        first two chunks (dead once layer 1 has completed)
constexpr int kTcB1 = kTcA2 + 16 * 2048;            // W1   :  64 rows x 8  tf32 = 2 chunks x 1024 B
constexpr int kTcB2 = kTcB1 + 2 * 1024;             // W2   :  64 rows x 64 tf32 = 16 chunks x 1024 B
constexpr int kTcVec = kTcB2 + 16 * 1024;           // b1[64] | b2[64] | w3[64] | b3
constexpr int kTcBar = kTcVec + 200 * 4;            // mbarrier (8 B) + TMEM base address (4 B)
constexpr int kTcSmemBytes = kTcBar + 16;
constexpr uint32_t kTcTmemCols = 64;                // D1 and D2 share the columns: D1 is drained before layer 2 is issued

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading /
// stride byte offsets in 16-byte units, version 1 (Blackwell), SWIZZLE_NONE.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tanh_mufu(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 32 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v)
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
}

template <bool STOCHASTIC>
__global__ void __launch_bounds__(kTcTile, 4)
policy_step_n1_tc_kernel(const DevParams P, const StatePtrs S, const float *__restrict__ weights,
                         const float *obs_in, float *__restrict__ actions_out, float *__restrict__ logp_out,
                         const Sinks out, const float log_std_value, const uint64_t noise_seed, const uint64_t step_offset,
                         const PolicyDyn dyn)
{
    const float log_std = dyn.log_std ? *dyn.log_std : log_std_value;                 // live values for graph replays
    const uint64_t step_index = step_offset + (dyn.step_base ? *dyn.step_base : 0);
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    float *sB1 = (float *)(smem + kTcB1), *sB2 = (float *)(smem + kTcB2), *sVec = (float *)(smem + kTcVec);
    uint32_t *tmem_slot = (uint32_t *)(smem + kTcBar + 8);
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar_addr = smem_base + kTcBar;

    // ---- one-time setup: weights into the UMMA B layout, barrier, TMEM columns
    for (int idx = tid; idx < kPolHidden * kPolObs; idx += kTcTile) {            // W1[n][k], k < 8
        const int n = idx >> 3, k = idx & 7;
        sB1[(k >> 2) * 256 + n * 4 + (k & 3)] = weights[kPolW1 + idx];
    }
    for (int idx = tid; idx < kPolHidden * kPolHidden; idx += kTcTile) {          // W2[n][k], k < 64
        const int n = idx >> 6, k = idx & 63;
        sB2[(k >> 2) * 256 + n * 4 + (k & 3)] = weights[kPolW2 + idx];
    }
    if (tid < kPolHidden) {
        sVec[tid] = weights[kPolB1 + tid];
        sVec[64 + tid] = weights[kPolB2 + tid];
        sVec[128 + tid] = weights[kPolW3 + tid];
    }
    if (tid == 0) {
        sVec[192] = weights[kPolB3];
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_base + kTcBar + 8), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                  // weights: generic -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);         // this warp's 32 TMEM lanes

    constexpr uint32_t idesc = umma_idesc_tf32(128, 64);
    const uint64_t a1_desc = umma_desc(smem_base + kTcA1, 2048, 128), b1_desc = umma_desc(smem_base + kTcB1, 1024, 128);
    const float std_dev = __expf(log_std);
    uint32_t phase = 0;
    Tally tally;
    tally_clear(tally);

    // Software prefetch: the records of the NEXT tile are requested before this tile's MMAs and
    // epilogues, so the two global-load latencies of a tile are off its critical path.
    struct Rec { float4 o0, o1; Vec2d pp; PlayerAux pa; Float4 h; };
    auto fetch = [&](int64_t i, Rec &r) {
        if (i < S.B) {
            r.o0 = ((const float4 *)obs_in)[2 * i]; r.o1 = ((const float4 *)obs_in)[2 * i + 1];
            r.pp = S.ppos[i]; r.pa = S.paux[i]; r.h = S.thot[i];
        } else {
            r.o0 = make_float4(0.f, 0.f, 0.f, 0.f); r.o1 = r.o0;
        }
    };
    const int64_t stride = (int64_t)gridDim.x * kTcTile;
    auto stage_obs = [&](const Rec &r) {                 // A1 <- this env's observation row (two 16-byte chunks)
        *(float4 *)(smem + kTcA1 + tid * 16) = r.o0;
        *(float4 *)(smem + kTcA1 + 2048 + tid * 16) = r.o1;
    };
    auto sync_for_mma = [&]() {                          // smem writes -> async proxy; orders prior tcgen05.ld
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    };
    auto issue_layer1 = [&]() {
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_tf32(tmem_base, a1_desc, b1_desc, idesc, 0u);
            umma_commit(bar_addr);
        }
    };

    // Prologue: first tile's records, its layer 1 in flight, second tile's records requested.
    int64_t base = (int64_t)blockIdx.x * kTcTile;
    bool have = base < S.B;
    Rec cur, nxt;
    fetch(base + tid, cur);
    fetch(base + stride + tid, nxt);
    if (have) {
        stage_obs(cur);
        sync_for_mma();
        issue_layer1();
    }

    while (have) {
        const int64_t i = base + tid;
        const bool valid = i < S.B;

        // ---- layer 1 was issued one stage ago (behind the previous tile's env step)
        mbar_wait(bar_addr, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue 1: H1 = tanh(D1 + b1) -> A2 (sixteen 16-byte chunks of this row)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld32(tmem_lane + half * 32, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 b = *(const float4 *)(sVec + half * 32 + c * 4);       // b1, broadcast 16-byte read
                float4 h;
                h.x = tanh_mufu(v[c * 4 + 0] + b.x);
                h.y = tanh_mufu(v[c * 4 + 1] + b.y);
                h.z = tanh_mufu(v[c * 4 + 2] + b.z);
                h.w = tanh_mufu(v[c * 4 + 3] + b.w);
                *(float4 *)(smem + kTcA2 + (half * 8 + c) * 2048 + tid * 16) = h;
            }
        }
        sync_for_mma();

        // ---- layer 2 on the tensor cores: eight K = 8 slices accumulate into D2
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_tf32(tmem_base, umma_desc(smem_base + kTcA2 + k * 2 * 2048, 2048, 128),
                          umma_desc(smem_base + kTcB2 + k * 2 * 1024, 1024, 128), idesc, k > 0 ? 1u : 0u);
            umma_commit(bar_addr);
        }
        mbar_wait(bar_addr, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue 2: mean = w3 . tanh(D2 + b2) + b3
        float mean = sVec[192];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld32(tmem_lane + half * 32, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 b = *(const float4 *)(sVec + 64 + half * 32 + c * 4);   // b2
                const float4 w = *(const float4 *)(sVec + 128 + half * 32 + c * 4);  // w3
                mean = fmaf(w.x, tanh_mufu(v[c * 4 + 0] + b.x), mean);
                mean = fmaf(w.y, tanh_mufu(v[c * 4 + 1] + b.y), mean);
                mean = fmaf(w.z, tanh_mufu(v[c * 4 + 2] + b.z), mean);
                mean = fmaf(w.w, tanh_mufu(v[c * 4 + 3] + b.w), mean);
            }
        }

        // ---- next tile: its observation rows go to A1 (D2 and A2 are drained) and its layer 1 is issued
        //      NOW, so that the tensor-core round trip runs behind this tile's env step
        const int64_t next_base = base + stride;
        const bool have_next = next_base < S.B;
        const Vec2d pp = cur.pp; const PlayerAux pa = cur.pa; const Float4 h = cur.h;
        cur = nxt;
        if (have_next) {
            fetch(next_base + stride + tid, nxt);
            stage_obs(cur);
            sync_for_mma();
            issue_layer1();
        }

        // ---- sample, clip, env step of this tile
        if (valid) {
            float a = mean;
            if (STOCHASTIC) {
                const float eps = policy_noise(noise_seed, S.gid0 + (uint64_t)i, step_index);
                a = fmaf(std_dev, eps, a);
                if (logp_out) logp_out[i] = -0.5f * eps * eps - log_std - 0.9189385332046727f;
            }
            if (actions_out) actions_out[i] = a;
            const float clipped = fminf(1.0f, fmaxf(-1.0f, a));
            Env1 e;
            e.px = pp.x; e.py = pp.y; e.psi = pa.psi; e.ret = pa.ep_return;
            e.steps = pa.steps & kStepsMask;
            e.bits = pa.steps & ~kStepsMask;
            e.tr.x0 = (double)h.x; e.tr.y0 = (double)h.y; e.tr.psi = (double)h.z; e.tr.v = (double)h.w;
            if (__builtin_expect((e.bits & kResidualBit) != 0, 0)) {
                const Residual r = S.tres[i];
                e.tr.x0 += r.x0; e.tr.y0 += r.y0; e.tr.psi += r.psi; e.tr.v += r.v;
            }
            e.minsep = 0.0f;
            e.respawned = false;
            step_env1<false, true>(P, S, e, clipped, i, out, tally, nullptr);
            store_env1(S, i, e, false);
        }
        base = next_base;
        have = have_next;
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
    tally_flush_warp(S.stats, tally);
}

}  // namespace acas2d
