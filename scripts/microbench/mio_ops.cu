// Micro-benchmark 3: MIO-side cost of SHFL vs LDS.32 (conflict-free / 4-way) vs LDS.128, per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define NACC 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int stride)
{
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = i * 1e-3f;
    __syncthreads();
    float r[NACC];
    for (int i = 0; i < NACC; i++) r[i] = threadIdx.x + i;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (MODE == 0) r[i] += __shfl_down_sync(0xffffffffu, r[(i + 1) % NACC], 1);
            if (MODE == 1) r[i] += sm[(threadIdx.x + i * 32 + it) & 4095];                       // LDS.32 conflict free
            if (MODE == 2) r[i] += sm[(4 * threadIdx.x + i * 32 + it) & 4095];                   // LDS.32 stride 4 (4-way)
            if (MODE == 3) { float4 v = *reinterpret_cast<const float4*>(&sm[(4 * (threadIdx.x + i * 32 + it)) & 4092]); r[i] += v.x + v.y + v.z + v.w; }  // LDS.128
            if (MODE == 4) { float2 v = *reinterpret_cast<const float2*>(&sm[(2 * (threadIdx.x + i * 32 + it)) & 4094]); r[i] += v.x + v.y; }  // LDS.64
            if (MODE == 5) r[i] += __shfl_xor_sync(0xffffffffu, r[(i + 1) % NACC], 1);
        }
    }
    float s = 0; for (int i = 0; i < NACC; i++) s += r[i];
    if (s == 123.456f) out[0] = s + lane;
}
template <int MODE> void run(const char* name, float* d)
{
    int sms = 148, blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winstr = (double)blocks * 8 * ITERS * NACC;
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-36s %8.3f ms  warp-ops/cycle/SM = %.3f\n", name, ms, winstr / cyc / sms);
}
int main()
{
    float* d; cudaMalloc(&d, 4);
    run<0>("SHFL.DOWN + FADD", d);
    run<5>("SHFL.BFLY + FADD", d);
    run<1>("LDS.32 conflict-free + FADD", d);
    run<2>("LDS.32 stride-4 (4-way) + FADD", d);
    run<4>("LDS.64 + 2 FADD", d);
    run<3>("LDS.128 + 4 FADD", d);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
