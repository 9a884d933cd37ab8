// Micro-benchmark (development tool): can a PART of a cyclically swept, read-modify-write working set that is larger
// than L2 be kept L2-resident across launches with evict_last cache hints (createpolicy), while the rest and a
// streaming output pass through with evict_first?  Working set like the N == 1 step: 36 B/env state rewritten in
// place + 37 B/env of streamed outputs, 4 Mi envs.
//   nvcc -arch=sm_100a -O3 -o ubench_l2keep ubench_l2keep.cu && ./ubench_l2keep
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr long long B = 4ll << 20;

__device__ __forceinline__ uint64_t policy(bool keep)
{
    uint64_t p;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_hint(const float4 *p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint(float4 *p, float4 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// state: 2 x float4 per env (ppos, paux) rewritten in place; out: 2 x float4 per env streamed
__global__ void step_like(float4 *ppos, float4 *paux, float4 *obs, long long keep_envs)
{
    const uint64_t pk = policy(true), ps = policy(false);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
        const uint64_t pol = i < keep_envs ? pk : ps;
        float4 a = ld_hint(ppos + i, pol), b = ld_hint(paux + i, pol);
        a.x += 1.0f; b.y += a.x;
        st_hint(ppos + i, a, pol); st_hint(paux + i, b, pol);
        __stcs(obs + 2 * i, a); __stcs(obs + 2 * i + 1, b);
    }
}

int main()
{
    int dev = 0, l2 = 0, maxp = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
    cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
    printf("L2 %d MB, max persisting %d MB\n", l2 >> 20, maxp >> 20);
    float4 *ppos, *paux, *obs;
    cudaMalloc(&ppos, B * 16); cudaMalloc(&paux, B * 16); cudaMalloc(&obs, B * 32);
    cudaMemset(ppos, 0, B * 16); cudaMemset(paux, 0, B * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int set_limit = 0; set_limit < 2; ++set_limit) {
        if (set_limit) { cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp); printf("set persisting limit: %s\n", cudaGetErrorString(e)); }
        for (double frac : {0.0, 0.2, 0.3, 0.4, 0.5, 0.6, 0.8}) {
            const long long keep = (long long)(frac * B);
            for (int i = 0; i < 5; ++i) step_like<<<148 * 8, 256>>>(ppos, paux, obs, keep);
            cudaEventRecord(e0);
            for (int i = 0; i < 20; ++i) step_like<<<148 * 8, 256>>>(ppos, paux, obs, keep);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("keep %.0f %% of the state (%.0f MB): %.1f us per sweep  (%.0f GB/s nominal on %d B/env)\n", 100 * frac, keep * 32e-6, 1e3 * ms / 20,
                   96.0 * B / (1e3 * ms / 20) * 1e-3, 96);
        }
    }
    return 0;
}
