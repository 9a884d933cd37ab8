// Micro-benchmark (development tool, not product code): per-SM throughput of the instruction kinds the tiled
// kernel's inner loop is made of -- DFMA, F2F f32<->f64, I2F/F2I f64, MUFU, FFMA, IMAD, LDS.128 -- to decide
// what the loop should be made of.   nvcc -arch=sm_100a -O3 -o ubench_pipes ubench_pipes.cu && ./ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void k(double *out, float seed, long long *cycles)
{
    double d0 = seed + threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;
    float f0 = seed + threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    __shared__ float4 sm[256];
    sm[threadIdx.x & 255] = make_float4(f0, f1, f2, f3);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
        if (OP == 0) { d0 = fma(d0, 1.0000001, 0.5); d1 = fma(d1, 1.0000001, 0.5); d2 = fma(d2, 1.0000001, 0.5); d3 = fma(d3, 1.0000001, 0.5); }
        if (OP == 1) { d0 = (double)f0; d1 = (double)f1; d2 = (double)f2; d3 = (double)f3;      // F2F.F64.F32 + F2F.F32.F64 + FADD round trip
                       f0 = (float)d0 + 1.0f; f1 = (float)d1 + 1.0f; f2 = (float)d2 + 1.0f; f3 = (float)d3 + 1.0f; }
        if (OP == 2) { d0 = (double)f0 + d0; d1 = (double)f1 + d1; d2 = (double)f2 + d2; d3 = (double)f3 + d3;   // F2F.F64.F32 + DADD
                       f0 += 1.0f; f1 += 1.0f; f2 += 1.0f; f3 += 1.0f; }
        if (OP == 3) { asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(f0) : "f"(f0)); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(f1) : "f"(f1));
                       asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(f2) : "f"(f2)); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(f3) : "f"(f3)); }
        if (OP == 4) { f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f); }
        if (OP == 5) { i0 = i0 * 3 + it; i1 = i1 * 3 + it; i2 = i2 * 3 + it; i3 = i3 * 3 + it; }
        if (OP == 6) { d0 = (double)i0 + d0; d1 = (double)i1 + d1; d2 = (double)i2 + d2; d3 = (double)i3 + d3; i0 += it; i1 += it; i2 += it; i3 += it; }   // I2F.F64 + DADD
        if (OP == 7) { float4 a = sm[(threadIdx.x + it) & 255], b = sm[(threadIdx.x + it + 64) & 255], c = sm[(threadIdx.x + it + 128) & 255], d = sm[(threadIdx.x + it + 192) & 255];
                       f0 += a.x + b.y; f1 += c.z + d.w; f2 += a.w; f3 += c.x; }
        if (OP == 8) { d0 = d0 + 1.5; d1 = d1 + 1.5; d2 = d2 + 1.5; d3 = d3 + 1.5; }
        if (OP == 9) { i0 += __double2int_rn(d0); i1 += __double2int_rn(d1); i2 += __double2int_rn(d2); i3 += __double2int_rn(d3);
                       d0 += 1.25; d1 += 1.25; d2 += 1.25; d3 += 1.25; }                                       // F2I.F64 + DADD
        if (OP == 11) { f0 = (float)d0; f1 = (float)d1; f2 = (float)d2; f3 = (float)d3; d0 += f0; d1 += f1; d2 += f2; d3 += f3; }   // F2F.F32.F64 + F2F.F64.F32 + DADD
        if (OP == 10) { bool p0 = d0 < d1, p1 = d2 < d3; i0 += p0; i1 += p1; d0 += 1e-9 * i0; d2 += 1e-9 * i1; }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3 + f0 + f1 + f2 + f3 + i0 + i1 + i2 + i3;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int ops_per_iter)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, blocks = sms;          // one full CTA per SM: 32 warps, 8 per scheduler
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * threads * blocks); cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<OP><<<blocks, threads>>>(out, 1.0f, cyc);
    k<OP><<<blocks, threads>>>(out, 1.0f, cyc);
    cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    double per_clk = (double)ITERS * ops_per_iter * threads / avg;
    printf("%-28s %8.1f thread-ops/clk/SM  (%.2f warp-instr/clk/SM)\n", name, per_clk, per_clk / 32);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("DFMA", 4); run<8>("DADD", 4); run<1>("F2F up+down+FADD (x4)", 4); run<2>("F2F.F64.F32+DADD (x4)", 4); run<6>("I2F.F64+DADD (x4)", 4);
    run<9>("F2I.F64+DADD (x4)", 4); run<11>("F2F down+up+DADD (x4)", 4); run<3>("MUFU.RSQ", 4); run<4>("FFMA", 4); run<5>("IMAD", 4); run<7>("LDS.128", 4); run<10>("DSETP+misc (2 setp)", 2);
    return 0;
}
