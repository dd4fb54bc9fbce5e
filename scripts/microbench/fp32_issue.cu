// Micro-benchmark: issue rate of scalar vs packed (f32x2) FP32 arithmetic on sm_100a.
// Decides whether the photometric stage should be written with add/mul/fma.rn.f32x2.
#include <cstdio>
#include <cuda_runtime.h>
#define NACC 8
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b)
{
    float r[NACC]; float2 r2[NACC];
    for (int i = 0; i < NACC; i++) { r[i] = threadIdx.x * 1e-3f + i; r2[i] = make_float2(r[i], r[i] + 0.5f); }
    float2 a2 = make_float2(a, a * 1.5f), b2 = make_float2(b, b * 0.5f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (MODE == 0) r[i] = __fadd_rn(r[i], a);
            if (MODE == 1) r[i] = __fmaf_rn(r[i], a, b);
            if (MODE == 2) r[i] = __fmul_rn(r[i], a);
            if (MODE == 3) r2[i] = __fadd2_rn(r2[i], a2);
            if (MODE == 4) r2[i] = __ffma2_rn(r2[i], a2, b2);
            if (MODE == 5) r2[i] = __fmul2_rn(r2[i], a2);
            if (MODE == 6) { r[i] = __fadd_rn(r[i], a); r2[i] = __fadd2_rn(r2[i], a2); }       // mix
            if (MODE == 7) { r[i] = __fadd_rn(r[i], r[(i + 1) % NACC]); }                      // 2 distinct regs
            if (MODE == 8) { r[i] = __fmaf_rn(r[i], r[(i + 1) % NACC], r[(i + 3) % NACC]); }  // 3 distinct regs
            if (MODE == 9) { r[i] = __fdiv_rn(r[i], a); }                                     // IEEE division
            if (MODE == 10) { r[i] = fminf(r[i], a); }                                        // alu pipe
            if (MODE == 11) { r[i] = __fadd_rn(r[i], a); r[(i + 1) % NACC] = fminf(r[(i + 1) % NACC], b); } // fma+alu mix
        }
    }
    float s = 0; for (int i = 0; i < NACC; i++) s += r[i] + r2[i].x + r2[i].y;
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name, int ops_per_iter_per_acc, float* d)
{
    int sms = 148, blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winstr = (double)blocks * 8 /*warps*/ * ITERS * NACC * ops_per_iter_per_acc;
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  warp-instr/cycle/SM = %.3f (at nominal %d MHz)\n", name, ms, winstr / cyc / sms, clk / 1000);
}
int main()
{
    float* d; cudaMalloc(&d, 4);
    run<0>("FADD scalar (imm-like)", 1, d); run<1>("FFMA scalar", 1, d); run<2>("FMUL scalar", 1, d);
    run<3>("FADD2 packed", 1, d); run<4>("FFMA2 packed", 1, d); run<5>("FMUL2 packed", 1, d);
    run<6>("FADD + FADD2 mix", 2, d); run<7>("FADD 2 distinct regs", 1, d); run<8>("FFMA 3 distinct regs", 1, d);
    run<9>("FDIV IEEE (per div)", 1, d); run<10>("FMNMX", 1, d); run<11>("FADD+FMNMX mix", 2, d);
    cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
