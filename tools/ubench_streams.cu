// Micro-benchmark (development tool): does the NUMBER of concurrent DRAM streams matter for the N == 1 step's
// traffic (40 B read + 69 B written per env, 4 Mi envs)?  Same bytes, different splits into arrays.
//   nvcc -arch=sm_100a -O3 -o ubench_streams ubench_streams.cu && ./ubench_streams
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

constexpr long long B = 4ll << 20;

// pattern A: the shipped layout.  reads ppos(16) paux(16) tpsi(4) act(4); writes ppos(16) paux(16) obs(32) rew(4) done(1)
__global__ void patA(const float4 *ppos, const float4 *paux, const float *tpsi, const float *act, float4 *oppos, float4 *opaux,
                     float4 *obs, float *rew, uint8_t *done)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
        float4 a = __ldcs(ppos + i), b = __ldcs(paux + i);
        float t = __ldcs(tpsi + i), c = __ldcs(act + i);
        a.x += t; b.y += c;
        __stcs(oppos + i, a); __stcs(opaux + i, b);
        __stcs(obs + 2 * i, a); __stcs(obs + 2 * i + 1, b);
        __stcs(rew + i, t + c); done[i] = (uint8_t)(c > 0.5f);
    }
}
// pattern B: player state as ONE 32-byte record (2 streams fewer)
__global__ void patB(const float4 *pst, const float *tpsi, const float *act, float4 *opst, float4 *obs, float *rew, uint8_t *done)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
        float4 a = __ldcs(pst + 2 * i), b = __ldcs(pst + 2 * i + 1);
        float t = __ldcs(tpsi + i), c = __ldcs(act + i);
        a.x += t; b.y += c;
        __stcs(opst + 2 * i, a); __stcs(opst + 2 * i + 1, b);
        __stcs(obs + 2 * i, a); __stcs(obs + 2 * i + 1, b);
        __stcs(rew + i, t + c); done[i] = (uint8_t)(c > 0.5f);
    }
}
// pattern C: + reward and done folded into a 40-byte output record... approximated by 48 B (3 x float4) output, 48 B input
__global__ void patC(const float4 *in, float4 *out_state, float4 *out_obs)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
        float4 a = __ldcs(in + 3 * i), b = __ldcs(in + 3 * i + 1), c = __ldcs(in + 3 * i + 2);
        a.x += c.x;
        __stcs(out_state + 2 * i, a); __stcs(out_state + 2 * i + 1, b);
        __stcs(out_obs + 3 * i, a); __stcs(out_obs + 3 * i + 1, b); __stcs(out_obs + 3 * i + 2, c);
    }
}

template <class F> float timeit(F f, int iters = 20)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return 1e3f * ms / iters;
}

int main()
{
    void *buf[12];
    for (int i = 0; i < 12; ++i) { cudaMalloc(&buf[i], B * 48); cudaMemset(buf[i], 0, B * 48); }
    const int grid = 148 * 8, block = 256;
    float a = timeit([&] { patA<<<grid, block>>>((float4 *)buf[0], (float4 *)buf[1], (float *)buf[2], (float *)buf[3], (float4 *)buf[0],
                                                 (float4 *)buf[1], (float4 *)buf[6], (float *)buf[7], (uint8_t *)buf[8]); });
    float b = timeit([&] { patB<<<grid, block>>>((float4 *)buf[0], (float *)buf[2], (float *)buf[3], (float4 *)buf[0], (float4 *)buf[6],
                                                 (float *)buf[7], (uint8_t *)buf[8]); });
    float c = timeit([&] { patC<<<grid, block>>>((float4 *)buf[9], (float4 *)buf[0], (float4 *)buf[10]); });
    printf("A shipped layout (4 read + 5 write streams, 109 B/env): %.1f us = %.0f GB/s\n", a, 109.0 * B / a * 1e-3);
    printf("B merged player record (3 read + 4 write streams, 109 B/env): %.1f us = %.0f GB/s\n", b, 109.0 * B / b * 1e-3);
    printf("C two fat streams (1 read + 2 write, 48 + 80 = 128 B/env): %.1f us = %.0f GB/s\n", c, 128.0 * B / c * 1e-3);
    return 0;
}
