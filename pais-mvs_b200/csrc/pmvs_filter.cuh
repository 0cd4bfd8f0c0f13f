/*
 * pmvs_filter.cuh — the PCMVS neighbour filter's pair scan (MVS::neighborPatchFiltering, TMVS/mvs/mvs.cpp:448-525).
 *
 * The reference, for every patch, computes the distance to every other patch, sorts the list and counts the entries
 * with dist <= neighborRadius (:470-499, the `break` at :496). Only the COUNT is used (:507-521), so the sort is
 * dropped: counts[i] = #{ j != i : norm(c_i - c_j) <= radius }, an exact integer.
 *
 * cv::norm(Vec3d) = sqrt(((dx*dx + dy*dy) + dz*dz)) with separate roundings; the kernel forms the same sum with
 * __dmul_rn / __dadd_rn (no contraction) and decides s <= r^2 outside a 2^-50 relative guard band, falling back to
 * the reference's sqrt-then-compare inside it, so every decision is the reference's bit for bit.
 *
 * One thread owns one patch i and walks all patches through 256-point shared-memory tiles (HBM traffic 24 B per point
 * per CTA; the scan is FP64-pipe bound: 3 subtractions, 3 products, 2 sums and 2 compares per pair).
 */
#pragma once
#include <cuda_runtime.h>

#define PMVS_NB_TILE 256

__global__ void __launch_bounds__(PMVS_NB_TILE) neighbor_count_kernel(int n, const double *__restrict__ centers, double radius,
                                                                     int first, int count, int *__restrict__ counts) {
    __shared__ double sx[PMVS_NB_TILE], sy[PMVS_NB_TILE], sz[PMVS_NB_TILE];
    const int i = first + blockIdx.x * PMVS_NB_TILE + threadIdx.x;
    const bool live = i < first + count;
    double cx = 0, cy = 0, cz = 0;
    if (live) { cx = centers[3 * (size_t)i]; cy = centers[3 * (size_t)i + 1]; cz = centers[3 * (size_t)i + 2]; }
    const double r2 = __dmul_rn(radius, radius);
    const double lo = __dmul_rn(r2, 1.0 - 0x1p-50), hi = __dmul_rn(r2, 1.0 + 0x1p-50);
    int c = 0;
    for (int t0 = 0; t0 < n; t0 += PMVS_NB_TILE) {
        const int j = t0 + threadIdx.x;
        __syncthreads();
        if (j < n) { sx[threadIdx.x] = centers[3 * (size_t)j]; sy[threadIdx.x] = centers[3 * (size_t)j + 1]; sz[threadIdx.x] = centers[3 * (size_t)j + 2]; }
        __syncthreads();
        const int m = n - t0 < PMVS_NB_TILE ? n - t0 : PMVS_NB_TILE;
#pragma unroll 4
        for (int k = 0; k < m; ++k) {
            const double dx = cx - sx[k], dy = cy - sy[k], dz = cz - sz[k];
            const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            bool in = s < lo;
            if (!in && !(s > hi)) in = __dsqrt_rn(s) <= radius;      /* guard band (and NaN): the reference's own test */
            c += (in && t0 + k != i) ? 1 : 0;
        }
    }
    if (live) counts[i - first] = c;
}
