// acas2d_dev.cuh -- device-only helpers shared by the kernels: warp-level episode-statistics flush,
// mbarrier / TMA bulk-copy wrappers (inline PTX).
#pragma once

#include "acas2d_env.cuh"

namespace acas2d {

constexpr unsigned kFull = 0xffffffffu;

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

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t addr, unsigned parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}

#ifndef ACAS2D_LOAD_HINT
#define ACAS2D_LOAD_HINT 1      /* L2 evict-first cache hint on the TMA loads: the records are rewritten, not re-read (90.8 vs 91.6 us) */
#endif

__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t l2_evict_last_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 1-D TMA bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(uint32_t smem_dst, const void *gmem_src, unsigned bytes, uint32_t bar,
                                            uint64_t policy = 0)
{
#if ACAS2D_LOAD_HINT
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
#else
    (void)policy;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
#endif
}

// Same without a cache hint (data that is re-read by the next step and fits L2).
__device__ __forceinline__ void tma_load_1d_plain(uint32_t smem_dst, const void *gmem_src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}

// 1-D TMA bulk copy shared -> global (bytes % 16 == 0, both addresses 16-byte aligned), then wait until the
// shared source has been read.  The caller orders its generic-proxy shared writes before this with
// fence.proxy.async + a barrier.
__device__ __forceinline__ void tma_store_1d_and_wait(void *gmem_dst, uint32_t smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_shared()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace acas2d
