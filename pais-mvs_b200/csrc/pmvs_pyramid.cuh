/*
 * pmvs_pyramid.cuh — the camera pyramids built on the device (reference: Camera ctor, TMVS/mvs/camera.cpp:63-92).
 *
 *   level i grey  = cv::resize(level 0, Size(), r^i, r^i, INTER_AREA)            camera.cpp:81-85
 *   level i edge  = sqrt(gx^2 + gy^2), gx/gy = Sobel ksize 1 = [-1,0,1] central difference with reflect-101 borders,
 *                   min-max normalised to [0,1] per level                        camera.cpp:71-78, :87-91
 *
 * OpenCV 2.4's INTER_AREA for a non-integer scale (imgproc resize.cpp: computeResizeAreaTab + ResizeArea_Invoker) is
 * restated: per destination index a run of source indices with float weights; a destination pixel is the weighted sum
 * over its source rows of the weighted horizontal sums, all in float, accumulated in source order, then
 * saturate_cast<uchar> (round half to even). The host builds the two 1-D weight tables; the kernels are pure streaming
 * passes: resize reads ~(1/r^i)^2 source bytes per destination byte (L2/L1-served overlap), the edge pass reads the grey
 * level twice (min/max, then normalise+store) and writes 8 B/pixel once — 10 B/pixel of HBM traffic instead of the 25
 * a store-then-normalise formulation would move.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

struct AreaTab {           /* one axis */
    std::vector<int> start, count;
    std::vector<float> w;  /* dsize * maxTaps */
    int maxTaps = 0, dsize = 0;
};

/* computeResizeAreaTab (OpenCV 2.4 imgproc/src/imgwarp.cpp) for one axis */
static inline void build_area_tab(int ssize, int dsize, double scale, AreaTab &t) {
    t.dsize = dsize;
    t.maxTaps = (int)std::ceil(scale) + 2;
    t.start.assign(dsize, 0);
    t.count.assign(dsize, 0);
    t.w.assign((size_t)dsize * t.maxTaps, 0.f);
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        float *w = &t.w[(size_t)dx * t.maxTaps];
        int n = 0, first = sx1;
        if (sx1 - fsx1 > 1e-3) { first = sx1 - 1; w[n++] = (float)((sx1 - fsx1) / cell); }
        for (int sx = sx1; sx < sx2; ++sx) w[n++] = (float)(1.0 / cell);
        if (fsx2 - sx2 > 1e-3) w[n++] = (float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell);
        t.start[dx] = first;
        t.count[dx] = n;
    }
}

__global__ void resize_area_kernel(const uint8_t *__restrict__ src, size_t spitch, int scols, int srows, uint8_t *__restrict__ dst,
                                   size_t dpitch, int dcols, int drows, const int *__restrict__ xs, const int *__restrict__ xn,
                                   const float *__restrict__ xw, int xTaps, const int *__restrict__ ys, const int *__restrict__ yn,
                                   const float *__restrict__ yw, int yTaps) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dcols || dy >= drows) return;
    const int x0 = xs[dx], nx = xn[dx], y0 = ys[dy], ny = yn[dy];
    const float *wx = xw + (size_t)dx * xTaps, *wy = yw + (size_t)dy * yTaps;
    float acc = 0.f;
    for (int ky = 0; ky < ny; ++ky) {
        const uint8_t *row = src + (size_t)(y0 + ky) * spitch + x0;
        float h = 0.f;
        for (int kx = 0; kx < nx; ++kx) h += wx[kx] * (float)row[kx];
        acc += wy[ky] * h;
    }
    int v = __float2int_rn(acc);            /* saturate_cast<uchar>(float): cvRound, then clamp */
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    dst[(size_t)dy * dpitch + dx] = (uint8_t)v;
}

/* INTER_AREA for an integer scale (OpenCV resize.cpp ResizeAreaFast): integer block sums; 2 x 2 blocks (sum + 2) >> 2,
 * n x n blocks saturate_cast<uchar>(sum * (1.f / area)), blocks cut by the right / bottom border (float)sum / count */
__global__ void resize_area_fast_kernel(const uint8_t *__restrict__ src, size_t spitch, int scols, int srows, uint8_t *__restrict__ dst,
                                        size_t dpitch, int dcols, int drows, int iscale) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dcols || dy >= drows) return;
    const int x0 = dx * iscale, y0 = dy * iscale;
    const int x1 = min(x0 + iscale, scols), y1 = min(y0 + iscale, srows);
    int sum = 0;
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) sum += src[(size_t)y * spitch + x];
    const int count = (x1 - x0) * (y1 - y0);
    int v;
    if (count == iscale * iscale) v = iscale == 2 ? (sum + 2) >> 2 : __float2int_rn((float)sum * (1.f / (float)(iscale * iscale)));
    else v = count > 0 ? __float2int_rn((float)sum / (float)count) : 0;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    dst[(size_t)dy * dpitch + dx] = (uint8_t)v;
}

__device__ __forceinline__ double edge_magnitude(const uint8_t *__restrict__ g, size_t pitch, int cols, int rows, int x, int y) {
    const int xl = x == 0 ? (cols > 1 ? 1 : 0) : x - 1, xr = x == cols - 1 ? (cols > 1 ? cols - 2 : 0) : x + 1;
    const int yu = y == 0 ? (rows > 1 ? 1 : 0) : y - 1, yd = y == rows - 1 ? (rows > 1 ? rows - 2 : 0) : y + 1;
    const double gx = (double)g[(size_t)y * pitch + xr] - (double)g[(size_t)y * pitch + xl];
    const double gy = (double)g[(size_t)yd * pitch + x] - (double)g[(size_t)yu * pitch + x];
    return sqrt(gx * gx + gy * gy);
}

/* pass 1: min / max of the gradient magnitude (non-negative doubles order like their bit patterns) */
__global__ void edge_minmax_kernel(const uint8_t *__restrict__ g, size_t pitch, int cols, int rows, unsigned long long *__restrict__ mm) {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int y = blockIdx.y; y < rows; y += gridDim.y)
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < cols; x += gridDim.x * blockDim.x) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(edge_magnitude(g, pitch, cols, rows, x, y));
            lo = b < lo ? b : lo;
            hi = b > hi ? b : hi;
        }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    /* one atomic pair per CTA, not per warp: 12 MP / 32 same-address atomics serialised in L2 (178 us at 4000x3000, 69 GB/s) */
    __shared__ unsigned long long sLo[8], sHi[8];
    const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { sLo[warp] = lo; sHi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; ++k) {
            lo = sLo[k] < lo ? sLo[k] : lo;
            hi = sHi[k] > hi ? sHi[k] : hi;
        }
        atomicMin(mm, lo);
        atomicMax(mm + 1, hi);
    }
}

/* pass 2: (e - min) / (max - min), camera.cpp:77 */
__global__ void edge_normalise_kernel(const uint8_t *__restrict__ g, size_t pitch, int cols, int rows,
                                      const unsigned long long *__restrict__ mm, double *__restrict__ edge) {
    const double mn = __longlong_as_double((long long)mm[0]), mx = __longlong_as_double((long long)mm[1]);
    const double range = mx - mn;
    for (int y = blockIdx.y; y < rows; y += gridDim.y)
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < cols; x += gridDim.x * blockDim.x)
            edge[(size_t)y * cols + x] = (edge_magnitude(g, pitch, cols, rows, x, y) - mn) / range;
}
