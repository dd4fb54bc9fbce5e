// Micro-benchmark 2: can the issue slots freed by packed f32x2 arithmetic be used by other pipes?
#include <cstdio>
#include <cuda_runtime.h>
#define NACC 8
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int ia)
{
    __shared__ float sm[256 * 4];
    float r[NACC]; float2 r2[NACC]; int q[NACC];
    for (int i = 0; i < NACC; i++) { r[i] = threadIdx.x * 1e-3f + i; r2[i] = make_float2(r[i], r[i] + 0.5f); q[i] = threadIdx.x + i; }
    float2 a2 = make_float2(a, a * 1.5f);
    sm[threadIdx.x] = a; sm[threadIdx.x + 256] = b; __syncthreads();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (MODE == 0) { r2[i] = __fadd2_rn(r2[i], a2); r[i] = fminf(r[i], a); }                  // FADD2 + FMNMX
            if (MODE == 1) { r2[i] = __fadd2_rn(r2[i], a2); q[i] = (q[i] ^ ia) + it; }               // FADD2 + 2 int ops
            if (MODE == 2) { r2[i] = __fadd2_rn(r2[i], a2); r[i] += sm[(threadIdx.x + i * 32 + it) & 1023]; } // FADD2 + LDS + FADD
            if (MODE == 3) { r2[i] = __ffma2_rn(r2[i], a2, a2); r[i] = fminf(r[i], a); q[i] = (q[i] ^ ia) + it; } // FFMA2 + FMNMX + 2 int
            if (MODE == 4) { r[i] = __fadd_rn(r[i], a); r[(i + 4) % NACC] = __fmul_rn(r[(i + 4) % NACC], b); q[i] = (q[i] ^ ia) + it; } // 2 scalar FP + 2 int
            if (MODE == 5) { r2[i] = __fadd2_rn(r2[i], make_float2(r[i], r[i])); r[i] = fminf(r[i], a); }   // FADD2 with broadcast operand
        }
    }
    float s = 0; for (int i = 0; i < NACC; i++) s += r[i] + r2[i].x + r2[i].y + q[i];
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name, int ops, float* d)
{
    int sms = 148, blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winstr = (double)blocks * 8 * ITERS * NACC * ops;
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-44s %8.3f ms  source-ops/cycle/SM = %.3f\n", name, ms, winstr / cyc / sms);
}
int main()
{
    float* d; cudaMalloc(&d, 4);
    run<0>("FADD2 + FMNMX (2 ops)", 2, d);
    run<1>("FADD2 + XOR + IADD (3 ops)", 3, d);
    run<2>("FADD2 + LDS + FADD (3 ops)", 3, d);
    run<3>("FFMA2 + FMNMX + XOR + IADD (4 ops)", 4, d);
    run<4>("FADD + FMUL + XOR + IADD (4 ops)", 4, d);
    run<5>("FADD2(bcast operand) + FMNMX (2 ops)", 2, d);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
