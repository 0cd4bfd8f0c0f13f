// Pipe-throughput microbenchmark for the instruction mix of the window loop (sm_100a):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each test runs `warps` warps per SM (5 warps x CTAs) of independent chains and reports warp-instructions per cycle per
// SM sub-partition for: DFMA, I2F.F64, MUFU.RCP64H, IMAD, PRMT and the loop's mix (DFMA : I2F : MUFU : ALU).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double s2d(int d) { return __hiloint2double(0x43300000, (int)((uint32_t)d ^ 0x80000000u)) - (4503599627370496.0 + 2147483648.0); }
template <int MODE, int CVT>
__global__ void __launch_bounds__(160) k(double *out, int iters, double seed, int iseed) {
    __shared__ double lut[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) lut[i] = (double)(i - 510);
    __syncthreads();
    double a[8];
    int n[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; n[i] = iseed + i * 7 + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fma(a[i], 1.0000001, 0.5);                                  // DFMA
                if (MODE == 1) { a[i] += (double)n[i]; n[i] += 3; }                               // I2F.F64 + DADD + IADD
                if (MODE == 2) { double rr; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(a[i])); a[i] = rr; }   // MUFU.RCP64H (+MOV)
                if (MODE == 3) n[i] = n[i] * 3 + i;                                               // IMAD
                if (MODE == 4) n[i] = __byte_perm(n[i], n[(i + 1) & 7], 0x3210 + i);              // PRMT
                if (MODE == 5) {   // the loop's mix per sample: 15 DFMA-class, 4 I2F, 1 MUFU, ~10 ALU
                    double rr;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(a[i]));
                    double e = fma(-a[i], rr, 1.0); e = fma(e, e, e); rr = fma(rr, e, rr);
                    double x = fma(a[i], 0.5, 3.0) * rr, y = fma(a[i], 0.25, 2.0) * rr;
                    double tx = __dadd_rd(x, 6755399441055744.0), ty = __dadd_rd(y, 6755399441055744.0);
                    int px = __double2loint(tx), py = __double2loint(ty);
                    double fy = y - (ty - 6755399441055744.0);
                    int q = n[i] + py * 1600 + px;
                    int g00 = __byte_perm(q, 0, 0x4440), g01 = __byte_perm(q, 0, 0x4441), g10 = __byte_perm(q, 0, 0x4442), g11 = __byte_perm(q, 0, 0x4443);
                    int ndx = g00 - g01, idy = g10 - g00, ndxy = g10 - g11 - ndx;
                    int k0 = px * ndx + g00, k1 = px * ndxy + idy;
                    double c;
                    if (CVT == 0) c = fma(fy, fma(-x, (double)ndxy, (double)k1), fma(-x, (double)ndx, (double)k0));
                    else if (CVT == 1) c = fma(fy, fma(-x, (double)ndxy, s2d(k1)), fma(-x, (double)ndx, s2d(k0)));
                    else if (CVT == 2) c = fma(fy, fma(-x, s2d(ndxy), s2d(k1)), fma(-x, s2d(ndx), s2d(k0)));
                    else c = fma(fy, fma(-x, lut[ndxy + 510], (double)k1), fma(-x, lut[ndx + 510], (double)k0));   // small differences from a shared-memory table
                    a[i] = a[i] + c * 1e-9 + 1.0;
                    n[i] += 12345;
                }
            }
        }
    }
    double s = 0;
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; t += n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int MODE, int CVT = 0>
void run(const char *name, double instPerInner, int ctasPerSm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * ctasPerSm * 160);
    const int iters = 2000;
    k<MODE, CVT><<<sms * ctasPerSm, 160>>>(out, 10, 1.5, 3);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, CVT><<<sms * ctasPerSm, 160>>>(out, iters, 1.5, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)sms * ctasPerSm * 5, inner = (double)iters * 64;
    const double winst = warps * inner * instPerInner;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %2d warps/SM  %8.3f ms  %.4f counted warp-inst/cycle/SMSP  (%.1f cycles per counted inst per SMSP)\n", name, ctasPerSm * 5, ms,
           winst / cycles / (sms * 4), cycles * sms * 4 / winst);
    cudaFree(out);
}

int main() {
    for (int c = 12; c >= 2; c -= 2) {
        if (c == 4) {
            run<0>("DFMA", 1, c);
            run<1>("I2F.F64 (+DADD+IADD)", 1, c);
            run<2>("MUFU.RCP64H", 1, c);
            run<4>("PRMT", 1, c);
        }
        run<5, 0>("loop mix, 4 I2F", 1, c);
        run<5, 1>("loop mix, 2 I2F + 2 magic", 1, c);
        run<5, 2>("loop mix, 4 magic", 1, c);
        run<5, 3>("loop mix, 2 I2F + 2 LDS table", 1, c);
    }
    return 0;
}
