// FP64 operand-pattern microbenchmark (sm_100a): how many cycles per warp-instruction per SM sub-partition does the FP64 pipe
// take for DFMA/DADD/DMUL with 1, 2, 3 distinct register operands, 20 warps/SM, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(160) k(double *out, int iters, double s0, double s1, double s2) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = s0 + i + threadIdx.x; b[i] = s1 + i * 0.5; c[i] = s2 + i * 0.25; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fma(a[i], 1.0000001, 0.5);              // 1 register operand
                if (MODE == 1) a[i] = fma(a[i], b[i], 0.5);                   // 2 register operands
                if (MODE == 2) a[i] = fma(a[i], b[i], c[i]);                  // 3 register operands
                if (MODE == 3) a[i] = fma(a[i], b[(i + r) & 7], c[(i + 3 * r) & 7]);   // 3 operands, varying registers
                if (MODE == 4) a[i] = a[i] + b[i];                            // DADD 2 regs
                if (MODE == 5) a[i] = a[i] * b[i];                            // DMUL 2 regs
                if (MODE == 6) a[i] = fma(a[i], a[i], a[i]);                  // same register three times
                if (MODE == 7) { a[i] = fma(a[i], b[i], c[i]); b[i] = fma(b[i], c[i], a[i]); }  // 2 dependent-ish, 3 operands
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, double per, int ctas) {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * ctas * 160);
    k<MODE><<<sms * ctas, 160>>>(out, 10, 1.5, 0.999, 0.001);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas, 160>>>(out, iters, 1.5, 0.999, 0.001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)sms * ctas * 5 * iters * 64 * per, cycles = ms * 1e-3 * clk * 1e3;
    printf("%-44s %2d warps/SM %8.3f ms  %.2f cycles per warp-inst per SMSP\n", name, ctas * 5, ms, cycles * sms * 4 / winst);
    cudaFree(out);
}
int main() {
    run<0>("DFMA r, imm, imm", 1, 4);
    run<1>("DFMA r, r, imm", 1, 4);
    run<2>("DFMA r, r, r", 1, 4);
    run<3>("DFMA r, r', r'' (rotating registers)", 1, 4);
    run<4>("DADD r, r", 1, 4);
    run<5>("DMUL r, r", 1, 4);
    run<6>("DFMA a, a, a", 1, 4);
    run<7>("2 x DFMA r, r, r (cross-dependent)", 2, 4);
    run<2>("DFMA r, r, r", 1, 2);
    run<2>("DFMA r, r, r", 1, 1);
    return 0;
}
