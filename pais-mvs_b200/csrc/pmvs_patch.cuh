/*
 * pmvs_patch.cuh — Patch::refine() and everything it calls, CTA-cooperative (one CTA = one patch at a time).
 *
 * Reference functions restated here (TMVS/mvs/patch.cpp unless noted):
 *   :114-176 refine            :180-219 psoOptimization      :221-267 setCorrelationTable
 *   :269-288 getHomographyRegionRatio (+ OpenCV 2.4 fitEllipse, restated from the published algorithm)
 *   :332-386 getHomographyPatch :415-445 setReferenceCameraIndex  :447-461 setDepthAndRay
 *   :463-509 setDepthRange      :511-610 setLOD   :612-625 setPriority   :627-653 setImagePoint
 *   :655-721 removeInvisibleCamera   :723-761 expandVisibleCamera
 * Convention: every cta_* function is entered and left with the CTA synchronised; patch state lives in shared
 * memory; scalar bookkeeping the reference does serially is done by thread 0 in the reference's order.
 */
#pragma once
#include "pmvs_device.cuh"

struct PatchS {
    double center[3], normal[3], normalS[2], ray[3], depth, depthRange[2], fitness, priority, correlation;
    double pt[2];                       /* scratch: reference-image point at LOD */
    double imgPoint[PMVS_MAX_VIEWS][2];
    double ratio[PMVS_MAX_VIEWS];       /* scratch: homography region ratios */
    int LOD, refCamIdx, type, id, drop, nCam, psoRuns, psoIterations, nImgPoint;
    unsigned evals, status, windowEvals;
    int flag, nx, ny;
    uint16_t camIdx[PMVS_MAX_VIEWS];
    /* memo of the last removeInvisibleCamera() that removed nothing: its inputs (see cta_remove_invisible) */
    double visCenter[3], visNormal[3];
    int visValid, visRef, visLOD, visN;
    uint16_t visCam[PMVS_MAX_VIEWS];
};

/* everything one CTA owns in shared memory */
struct CtaS {
    PatchS p;
    EvalCtx E;
    PsoS pso;
    ParticleS *part;                    /* 2*particleNum entries (seeds run 2P particles, patch.cpp:192) */
    MoveS mv;
    PmvsPatchOut out;
    RefWin rw;                          /* reference window of the current swarm run (tables carved by the kernel) */
    int nextIdx;
};

/* -------------------------------------------------------------------------------------------------------
 * OpenCV 2.4 fitEllipse (imgproc shapedescr.cpp cvFitEllipse2) with cvSolve(CV_SVD) = one-sided Jacobi SVD
 * + truncated back-substitution, restated from the published algorithm (OpenCV is not under the reference tree).
 * ----------------------------------------------------------------------------------------------------- */
__device__ inline void svd_solve(const double *A, const double *b, int m, int n, double *x) {
    double At[5][8], Vt[5][5], W[5];
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < m; ++k) At[i][k] = A[k * n + i];
        for (int k = 0; k < n; ++k) Vt[i][k] = (i == k) ? 1.0 : 0.0;
    }
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < m; ++k) sd += At[i][k] * At[i][k];
        W[i] = sd;
    }
    const double eps = DBL_EPSILON * 10;
    const int max_iter = m > 30 ? m : 30;
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
        for (int i = 0; i < n - 1; ++i)
            for (int j = i + 1; j < n; ++j) {
                double a = W[i], p = 0, bb = W[j];
                for (int k = 0; k < m; ++k) p += At[i][k] * At[j][k];
                if (fabs(p) <= eps * sqrt(a * bb)) continue;
                p *= 2;
                double beta = a - bb, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = bb = 0;
                for (int k = 0; k < m; ++k) {
                    double t0 = c * At[i][k] + s * At[j][k];
                    double t1 = -s * At[i][k] + c * At[j][k];
                    At[i][k] = t0;
                    At[j][k] = t1;
                    a += t0 * t0;
                    bb += t1 * t1;
                }
                W[i] = a;
                W[j] = bb;
                changed = true;
                for (int k = 0; k < n; ++k) {
                    double t0 = c * Vt[i][k] + s * Vt[j][k];
                    double t1 = -s * Vt[i][k] + c * Vt[j][k];
                    Vt[i][k] = t0;
                    Vt[j][k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < m; ++k) sd += At[i][k] * At[i][k];
        W[i] = sqrt(sd);
    }
    double threshold = 0;
    for (int i = 0; i < n; ++i) threshold += W[i];
    threshold *= DBL_EPSILON * 2;
    for (int k = 0; k < n; ++k) x[k] = 0;
    for (int i = 0; i < n; ++i) {
        if (W[i] <= threshold) continue;
        double ub = 0;
        for (int k = 0; k < m; ++k) ub += At[i][k] * b[k];
        double coef = ub / (W[i] * W[i]);
        for (int k = 0; k < n; ++k) x[k] += coef * Vt[i][k];
    }
}

__device__ inline double fit_ellipse_ratio(const float *px, const float *py, int n) {
    const double min_eps = 1e-8;
    double gfp[5], rp[5], t;
    double Ad[8 * 5], bd[8];
    float cx = 0, cy = 0;
    for (int i = 0; i < n; ++i) { cx += px[i]; cy += py[i]; }
    cx /= n;
    cy /= n;
    for (int i = 0; i < n; ++i) {
        float x = px[i] - cx, y = py[i] - cy;
        bd[i] = 10000.0;
        Ad[i * 5] = -(double)x * x;
        Ad[i * 5 + 1] = -(double)y * y;
        Ad[i * 5 + 2] = -(double)x * y;
        Ad[i * 5 + 3] = x;
        Ad[i * 5 + 4] = y;
    }
    svd_solve(Ad, bd, n, 5, gfp);
    double A2[4] = {2 * gfp[0], gfp[2], gfp[2], 2 * gfp[1]}, b2[2] = {gfp[3], gfp[4]};
    svd_solve(A2, b2, 2, 2, rp);
    for (int i = 0; i < n; ++i) {
        float x = px[i] - cx, y = py[i] - cy;
        bd[i] = 1.0;
        Ad[i * 3] = (x - rp[0]) * (x - rp[0]);
        Ad[i * 3 + 1] = (y - rp[1]) * (y - rp[1]);
        Ad[i * 3 + 2] = (x - rp[0]) * (y - rp[1]);
    }
    svd_solve(Ad, bd, n, 3, gfp);
    rp[4] = -0.5 * atan2(gfp[2], gfp[1] - gfp[0]);
    t = sin(-2.0 * rp[4]);
    if (fabs(t) > fabs(gfp[2]) * min_eps) t = gfp[2] / t;
    else t = gfp[1] - gfp[0];
    rp[2] = fabs(gfp[0] + gfp[1] - t);
    if (rp[2] > min_eps) rp[2] = sqrt(2.0 / rp[2]);
    rp[3] = fabs(gfp[0] + gfp[1] + t);
    if (rp[3] > min_eps) rp[3] = sqrt(2.0 / rp[3]);
    const float w = (float)(rp[2] * 2), h = (float)(rp[3] * 2);
    const float lo = h < w ? h : w, hi = w < h ? h : w;      /* std::min / std::max */
    return (double)(lo / hi);
}

/* Patch::getHomographyRegionRatio, patch.cpp:269-288 */
__device__ __noinline__ double region_ratio(int r, const double *pt, const double *H) {
    const double x[8] = {pt[0] - r, pt[0] - r, pt[0] + r, pt[0] + r, pt[0] - r, pt[0], pt[0] + r, pt[0]};
    const double y[8] = {pt[1] - r, pt[1] + r, pt[1] + r, pt[1] - r, pt[1], pt[1] + r, pt[1], pt[1] - r};
    float fx[8], fy[8];
    for (int i = 0; i < 8; ++i) {
        const double w = H[6] * x[i] + H[7] * y[i] + H[8];
        fx[i] = (float)((H[0] * x[i] + H[1] * y[i] + H[2]) / w);
        fy[i] = (float)((H[3] * x[i] + H[4] * y[i] + H[5]) / w);
    }
    return fit_ellipse_ratio(fx, fy, 8);
}

/* ------------------------------------------------------------------------------------------------------- */
__device__ inline void set_normalS(PatchS &p, double theta, double phi) {   /* abstractpatch.cpp:47-50 */
    p.normalS[0] = theta;
    p.normalS[1] = phi;
    spherical2Normal(theta, phi, p.normal);
}

/* thread 0: setReferenceCameraIndex (:415-445) + setDepthAndRay (:447-461) */
__device__ inline void t0_set_ref_depth_ray(const DevScene &S, PatchS &p) {
    if (!p.drop) {
        if (p.nCam < S.cfg.minCamNum) p.drop = 1;
        else {
            p.refCamIdx = -1;
            double maxCorr = -DBL_MAX;
            for (int i = 0; i < p.nCam; i++) {
                const double *on = S.cams[p.camIdx[i]].optN;
                const double neg[3] = {-on[0], -on[1], -on[2]};
                const double corr = dot3(p.normal, neg);
                if (corr > maxCorr) { maxCorr = corr; p.refCamIdx = p.camIdx[i]; }
            }
            if (p.refCamIdx < 0) { p.refCamIdx = p.camIdx[0]; p.drop = 1; }
        }
    }
    if (!p.drop) {
        if (p.refCamIdx < 0) p.drop = 1;
        else {
            const double *C = S.cams[p.refCamIdx].center;
            for (int k = 0; k < 3; ++k) p.ray[k] = p.center[k] - C[k];
            p.depth = sqrt(p.ray[0] * p.ray[0] + p.ray[1] * p.ray[1] + p.ray[2] * p.ray[2]);
            const double inv = 1.0 / p.depth;
            for (int k = 0; k < 3; ++k) p.ray[k] = p.ray[k] * inv;
        }
    }
}

/* warp 0: setDepthRange (:463-509) — lanes over views, exact max */
__device__ inline void w0_set_depth_range(const DevScene &S, PatchS &p) {
    const int lane = threadIdx.x & 31;
    if (p.drop) return;
    if (p.nCam < S.cfg.minCamNum) {
        __syncwarp();
        if (lane == 0) p.drop = 1;
        __syncwarp();
        return;
    }
    const DevCamera &rc = S.cams[p.refCamIdx];
    double c2[3], c1[3];
    for (int k = 0; k < 3; ++k) {
        c2[k] = p.ray[k] * (p.depth + 1.0) + rc.center[k];
        c1[k] = p.center[k];
    }
    double best = -DBL_MAX;
    for (int i = lane; i < p.nCam; i += 32) {
        if (p.camIdx[i] == p.refCamIdx) continue;
        const DevCamera &cam = S.cams[p.camIdx[i]];
        double p1[2], p2[2];
        project_pt(cam.R, cam.t, cam.focal, cam.pp, S.lodScale[0], c1, p1);
        project_pt(cam.R, cam.t, cam.focal, cam.pp, S.lodScale[0], c2, p2);
        const double dx = p1[0] - p2[0], dy = p1[1] - p2[1];
        const double imgDist = sqrt(dx * dx + dy * dy);
        const double worldDist = 1.0 / imgDist;
        if (worldDist > best && imgDist >= 0.01) best = worldDist;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double b2 = __shfl_xor_sync(PMVS_FULL, best, o);
        if (b2 > best) best = b2;
    }
    __syncwarp();
    if (lane == 0) {
        if (best == -DBL_MAX) p.drop = 1;
        else {
            const double a = p.depth - best * S.cfg.depthRangeScalar;
            p.depthRange[0] = (a < 0.0) ? 0.0 : a;                                  /* std::max(a, 0.0) */
            const double m1 = best * S.cfg.depthRangeScalar, m2 = S.cfg.neighborRadius * 100;
            p.depthRange[1] = p.depth + ((m2 < m1) ? m2 : m1);                      /* std::min(m1, m2) */
        }
    }
    __syncwarp();
}

/* warp 0: setLOD (:511-610) — lanes over the window */
__device__ inline void w0_set_lod(const DevScene &S, PatchS &p) {
    const int lane = threadIdx.x & 31;
    if (p.drop) return;
    if (p.refCamIdx < 0) {
        __syncwarp();
        if (lane == 0) p.drop = 1;
        __syncwarp();
        return;
    }
    const int r = S.cfg.patchRadius;
    const DevCamera &rc = S.cams[p.refCamIdx];
    const double center[3] = {p.center[0], p.center[1], p.center[2]};
    double variance = 0;
    int LOD = S.cfg.minLOD - 1;
    while (variance < S.cfg.textureVariation) {
        LOD++;
        if (LOD >= rc.maxLOD) { LOD = rc.maxLOD; break; }
        double pt[2];
        project_pt(rc.R, rc.t, rc.focal, rc.pp, S.lodScale[LOD], center, pt);
        const DevLevel &L = rc.level[LOD];
        if (!in_image(pt[0], pt[1], L.cols, L.rows)) { LOD = (LOD - 1 > 0) ? LOD - 1 : 0; break; }
        const int cx = __double2int_rn(pt[0]), cy = __double2int_rn(pt[1]);
        if (cx - r < 0 || cx + r >= L.cols || cy - r < 0 || cy + r >= L.rows) { LOD = (LOD - 1 > 0) ? LOD - 1 : 0; break; }
        const int ps = 2 * r + 1, count = ps * ps;
        int isum = 0;
        for (int s = lane; s < count; s += 32) {
            const int x = cx - r + s % ps, y = cy - r + s / ps;
            isum += (int)(__ldg(L.quad + (size_t)y * L.cols + x) & 0xff);
        }
        isum = __reduce_add_sync(PMVS_FULL, isum);
        const double mean = (double)isum / count;
        double var = 0;
        for (int s = lane; s < count; s += 32) {
            const int x = cx - r + s % ps, y = cy - r + s / ps;
            const double t = (double)(__ldg(L.quad + (size_t)y * L.cols + x) & 0xff) - mean;
            var += t * t;
        }
        variance = warp_sum(var) / count;
    }
    __syncwarp();
    if (lane == 0) p.LOD = LOD;
    __syncwarp();
}

/* prologue block used at refine() entry and after every swarm: :125-128 / :163-166 */
__device__ inline void cta_update_information(const DevScene &S, PatchS &p) {
    const int tid = threadIdx.x;
    if (tid == 0) t0_set_ref_depth_ray(S, p);
    __syncthreads();
    if (tid < 32) {
        w0_set_depth_range(S, p);
        w0_set_lod(S, p);
    }
    __syncthreads();
}

/* Patch::expandVisibleCamera (:723-761); p.camIdx holds the PARENT's cameras on entry */
__device__ inline void cta_expand_visible(const DevScene &S, PatchS &p) {
    if (threadIdx.x == 0 && !p.drop) {
        int exp[PMVS_MAX_VIEWS + PMVS_MAX_VIEWS];
        int n = 0;
        bool over = false;
        for (int i = 0; i < S.nCams; ++i) {
            const double *on = S.cams[i].optN;
            const double neg[3] = {-on[0], -on[1], -on[2]};
            if (dot3(p.normal, neg) >= S.cfg.visibleCorrelation) {
                if (n < PMVS_MAX_VIEWS) exp[n++] = i;
                else over = true;
            }
        }
        if (!over && n < S.cfg.minCamNum) {
            for (int i = 0; i < p.nCam; ++i) {
                const double *on = S.cams[p.camIdx[i]].optN;
                const double neg[3] = {-on[0], -on[1], -on[2]};
                if (dot3(p.normal, neg) >= S.cfg.visibleCorrelation / 2.0) exp[n++] = p.camIdx[i];
            }
            for (int a = 1; a < n; ++a) {                      /* sort + unique */
                const int v = exp[a];
                int b = a - 1;
                while (b >= 0 && exp[b] > v) { exp[b + 1] = exp[b]; --b; }
                exp[b + 1] = v;
            }
            int m = 0;
            for (int a = 0; a < n; ++a)
                if (m == 0 || exp[m - 1] != exp[a]) exp[m++] = exp[a];
            n = m;
            if (n > PMVS_MAX_VIEWS) over = true;
        }
        if (over) {                                            /* capacity of the C-ABI records, not in the reference */
            p.status |= PMVS_S_TOO_MANY_VIEWS;
            p.nCam = 0;
            p.drop = 1;
        } else {
            for (int a = 0; a < n; ++a) p.camIdx[a] = (uint16_t)exp[a];
            p.nCam = n;
            if (p.nCam < S.cfg.minCamNum) p.drop = 1;
        }
    }
    __syncthreads();
}

/* (re)build the evaluation context from the patch state */
__device__ inline void cta_build_ctx(const DevScene &S, CtaS &c) {
    build_eval_ctx(S, c.E, c.p.ray, c.p.refCamIdx, c.p.LOD, c.p.nCam, c.p.camIdx, threadIdx.x, blockDim.x);
    __syncthreads();
    finish_eval_ctx(c.E, threadIdx.x);
    __syncthreads();
}

struct PatchEval {
    const DevScene &S;
    const EvalCtx &E;
    const double *sDistW;
    WarpWork W;
    unsigned *windowEvals;
    __device__ __forceinline__ double operator()(const double *pos) const {
        const double f = warp_fitness_any(S, E, sDistW, W, pos[0], pos[1], pos[2]);
        if ((threadIdx.x & 31) == 0 && f != DBL_MAX) atomicAdd(windowEvals, 1u);
        return f;
    }
    /* m <= PMVS_EVAL_BATCH particles part[p0 + k stride] of this patch's swarm: fitness stored in the particles */
    __device__ __forceinline__ void batch(int m, ParticleS *part, int p0, int stride) const {
        const unsigned ran = warp_fitness_batch(S, E, sDistW, W, m, part, p0, stride);
        if ((threadIdx.x & 31) == 0 && ran) atomicAdd(windowEvals, ran);
    }
};

/* Patch::psoOptimization, :180-219 */
__device__ inline void cta_pso_optimization(const DevScene &S, CtaS &c, const double *sDistW, const WarpWork &W) {
    PatchS &p = c.p;
    const int tid = threadIdx.x;
    cta_build_ctx(S, c);
    if (W.rw) {                                           /* uniform: the launch either carries reference windows or not */
        /* canonical hypothesis centre = the initial particle's (patch.cpp:204, :944) */
        const double ctr[3] = {c.E.ray[0] * p.depth + c.E.refC[0], c.E.ray[1] * p.depth + c.E.refC[1], c.E.ray[2] * p.depth + c.E.refC[2]};
        build_ref_win<false>(S, c.E, c.rw, ctr, tid, blockDim.x);
#if PMVS_TILE
        stage_tiles(S, c.E, c.rw, ctr, p.normal, W.H);      /* only warp 0 touches it: its own homography area, free before the swarm starts */
#endif
#if PMVS_FOOTPRINT
        if (tid < 8) { c.rw.foot[tid][0] = c.rw.foot[tid][1] = 1 << 30; c.rw.foot[tid][2] = c.rw.foot[tid][3] = -(1 << 30); c.rw.footn[tid] = 0; }
        __syncthreads();
#endif
    }
    __shared__ double sInit[3];
    if (tid == 0) {
        const double PI = 3.14159265358979323846;
        double L[3] = {0.0, p.normalS[1] - PI / 2.0, p.depthRange[0]};
        double U[3] = {PI, p.normalS[1] + PI / 2.0, p.depthRange[1]};
        sInit[0] = p.normalS[0];
        sInit[1] = p.normalS[1];
        sInit[2] = p.depth;
        int maxIter, P;
        if (p.type == PMVS_TYPE_SEED) {
            maxIter = S.cfg.maxIteration * 2;
            P = S.cfg.particleNum * 2;
        } else {
            const double lo = p.normalS[0] - PI / S.cfg.reduceNormalRange, hi = p.normalS[0] + PI / S.cfg.reduceNormalRange;
            L[0] = (0.0 < lo) ? lo : 0.0;      /* std::max(0.0, lo) */
            U[0] = (hi < PI) ? hi : PI;        /* std::min(PI, hi)  */
            L[1] = p.normalS[1] - PI / S.cfg.reduceNormalRange;
            U[1] = p.normalS[1] + PI / S.cfg.reduceNormalRange;
            maxIter = S.cfg.maxIteration;
            P = S.cfg.particleNum;
        }
        pso_setup(c.pso, L, U, maxIter, P, pmvs_stream_key(S.seed, p.id, p.psoRuns));
    }
    __syncthreads();
    PatchEval ev = {S, c.E, sDistW, W, &p.windowEvals};
    const unsigned evals = pso_run(c.pso, c.part, c.mv, ev, sInit, true);
    if (tid == 0) {
        const ParticleS &g = c.part[c.pso.gBestIdx];
        p.fitness = c.pso.gBestFitness;
        set_normalS(p, g.pBest[0], g.pBest[1]);
        p.depth = g.pBest[2];
        const double *C = S.cams[p.refCamIdx].center;
        for (int k = 0; k < 3; ++k) p.center[k] = p.ray[k] * p.depth + C[k];
        p.psoIterations = c.pso.iteration;
        p.psoRuns++;
        p.evals += evals;
#if PMVS_FOOTPRINT
        if (S.dbgFoot && W.rw && p.psoRuns == 1 && c.nextIdx < 65536)
            for (int v = 0; v < 8; ++v)
                for (int k = 0; k < 4; ++k) S.dbgFoot[(c.nextIdx * 8 + v) * 4 + k] = v < c.E.V ? c.rw.foot[v][k] : (v == 7 ? c.rw.footn[k] : (v == 6 && k == 0 ? c.rw.footn[4] : 0));
#endif
    }
    __syncthreads();
}

/*
 * Patch::removeInvisibleCamera (:655-721) with setCorrelationTable (:221-267) and getHomographyPatch (:332-386).
 * Hc: V*9 doubles, xs/ys: window axes, corr: V*V doubles (all shared); hp: this CTA's scratch in HBM (V*ps*ps).
 */
__device__ __noinline__ void cta_remove_invisible(const DevScene &S, CtaS &c, double *Hc, double *xs, double *ys, double *corr,
                                            double *hp) {
    PatchS &p = c.p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    if (p.drop) return;                                   /* uniform: p.drop was written before the last barrier */
    const int V = p.nCam, r = S.cfg.patchRadius, ps = S.cfg.patchSize, SS = ps * ps;
    /* removeInvisibleCamera is a pure function of (center, normal, reference camera, LOD, camera list). The caller's
     * trailing call (mvs.cpp:215, :574) very often repeats the call made inside refine() (:162) with identical inputs —
     * nothing was removed and the reference camera / LOD were re-selected to the same values — and then it would
     * reproduce the same correlation and remove nothing again: skip it. */
    if (tid == 0) {
        bool same = p.visValid && p.visRef == p.refCamIdx && p.visLOD == p.LOD && p.visN == V;
        for (int k = 0; same && k < 3; ++k) same = p.visCenter[k] == p.center[k] && p.visNormal[k] == p.normal[k];
        for (int k = 0; same && k < V; ++k) same = p.visCam[k] == p.camIdx[k];
        p.flag = same ? 2 : 0;
    }
    __syncthreads();
    if (p.flag == 2) return;
    cta_build_ctx(S, c);
    const EvalCtx &E = c.E;
    if (!E.valid) {                                       /* missing pyramid level: cannot be evaluated */
        if (tid == 0) { p.drop = 1; p.correlation = 0; }
        __syncthreads();
        return;
    }
    if (warp == 0) {
        const double ctr[3] = {p.center[0], p.center[1], p.center[2]}, nrm[3] = {p.normal[0], p.normal[1], p.normal[2]};
        warp_homographies(E, ctr, nrm, Hc);                                           /* :663 */
        double pt[2];
        project_pt(E.refR, E.refT, E.refFocal, E.refPP, E.sc, ctr, pt);               /* :233, :683 */
        const int nxy = warp_window_axes(pt, r, ps, xs, ys);
        if (lane == 0) {
            p.pt[0] = pt[0];
            p.pt[1] = pt[1];
            p.nx = nxy & 0xffff;
            p.ny = nxy >> 16;
            p.flag = 0;
        }
    }
    __syncthreads();
    const int nx = p.nx, ny = p.ny, total = nx * ny;
    /* normalised homography patches, one warp per view */
    for (int v = warp; v < V; v += NW) {
        const double *H = Hc + 9 * v;
        const ViewS &vw = E.view[v];
        double *dst = hp + (size_t)v * SS;
        double sum = 0;
        bool oob = false;
        for (int s = lane; s < total; s += 32) {
            const double x = xs[s % nx], y = ys[s / nx];
            const double w = (H[6] * x + H[7] * y + H[8]);
            const double ix = (H[0] * x + H[1] * y + H[2]) / w, iy = (H[3] * x + H[4] * y + H[5]) / w;
            if (!(ix >= 0.0 && ix < (double)(vw.cols - 1) && iy >= 0.0 && iy < (double)(vw.rows - 1)) || w == 0.0) {   /* :355 */
                oob = true;
                dst[s] = 0;
                continue;
            }
            const double val = quad_bilinear(vw.quad, vw.cols, ix, iy);
            dst[s] = val;
            sum += val * val;
        }
        if (__any_sync(PMVS_FULL, oob)) {
            if (lane == 0) atomicOr(&p.flag, 1);
        }
        const double rn = 1.0 / sqrt(warp_sum(sum));                                   /* :384; Mat /= s multiplies by 1./s */
        for (int s = lane; s < total; s += 32) dst[s] = dst[s] * rn;
    }
    __syncthreads();
    if (p.flag) {                                                                     /* :244-247 */
        for (int k = tid; k < V * V; k += blockDim.x) corr[k] = 0;
        if (tid == 0) { p.drop = 1; p.correlation = 0; }
        __syncthreads();
    } else {
        __threadfence_block();
        const int nPairs = V * (V - 1) / 2;
        for (int k = tid; k < V; k += blockDim.x) corr[k * V + k] = 0;
        for (int pi = warp; pi < nPairs; pi += NW) {
            int i = 0, rem = pi;
            while (rem >= V - 1 - i) { rem -= V - 1 - i; ++i; }
            const int j = i + 1 + rem;
            const double *a = hp + (size_t)i * SS, *b = hp + (size_t)j * SS;
            double acc = 0;
            for (int s = lane; s < total; s += 32) acc += a[s] * b[s];
            acc = warp_sum(acc);
            if (lane == 0) { corr[i * V + j] = acc; corr[j * V + i] = acc; }
        }
        __syncthreads();
        if (tid == 0) {                                                               /* :259-266 */
            double cs = 0;
            for (int i = 0; i < V; ++i)
                for (int j = 0; j < V; ++j) cs += corr[i * V + j];
            p.correlation = cs / (V * V - V);
        }
    }
    /* region ratios: lane 0 of each warp takes views round-robin (:690, :269-288) */
    if (lane == 0)
        for (int v = warp; v < V; v += NW) p.ratio[v] = region_ratio(r, p.pt, Hc + 9 * v);
    __syncthreads();
    if (tid == 0) {
        double maxCorr = -DBL_MAX;
        int maxIdx = 0;
        for (int i = 0; i < V; ++i) {                                                 /* :666-680 */
            double corrSum = 0;
            for (int j = 0; j < V; ++j) corrSum += corr[i * V + j];
            if (corrSum >= maxCorr) { maxIdx = i; maxCorr = corrSum; }
        }
        int removeIdx[PMVS_MAX_VIEWS], nRemove = 0;
        for (int i = 0; i < V; ++i) {                                                 /* :686-707 */
            if (p.ratio[i] < S.cfg.minRegionRatio) { removeIdx[nRemove++] = p.camIdx[i]; continue; }
            const double *on = S.cams[p.camIdx[i]].optN;
            const double neg[3] = {-on[0], -on[1], -on[2]};
            if (dot3(p.normal, neg) < 0) { removeIdx[nRemove++] = p.camIdx[i]; continue; }
            if (i == maxIdx) continue;
            if (corr[maxIdx * V + i] < S.cfg.minCorrelation) { removeIdx[nRemove++] = p.camIdx[i]; continue; }
        }
        int n = V;
        for (int k = 0; k < nRemove; ++k)                                             /* :709-716 */
            for (int a = 0; a < n; ++a)
                if (p.camIdx[a] == removeIdx[k]) {
                    for (int b = a; b + 1 < n; ++b) p.camIdx[b] = p.camIdx[b + 1];
                    --n;
                    break;
                }
        p.nCam = n;
        if (p.nCam < S.cfg.minCamNum) p.drop = 1;
        p.visValid = (nRemove == 0 && !p.drop) ? 1 : 0;
        if (p.visValid) {
            p.visRef = p.refCamIdx;
            p.visLOD = p.LOD;
            p.visN = V;
            for (int k = 0; k < 3; ++k) { p.visCenter[k] = p.center[k]; p.visNormal[k] = p.normal[k]; }
            for (int k = 0; k < V; ++k) p.visCam[k] = p.camIdx[k];
        }
    }
    __syncthreads();
}

/* thread 0: setPriority (:612-625) + setImagePoint (:627-653; colour lookup is outside the hot path) */
__device__ inline void t0_priority_imgpoint(const DevScene &S, PatchS &p) {
    if (p.drop) return;
    const double camRatio = ((double)p.nCam) / ((double)S.nCams);
    p.priority = p.fitness * exp(-p.correlation / 1.0 - camRatio / 1.0) * (p.LOD + 1.0);
    if (p.nCam == 0) return;
    p.nImgPoint = p.nCam;
    for (int i = 0; i < p.nCam; ++i) {
        const DevCamera &cam = S.cams[p.camIdx[i]];
        project_pt(cam.R, cam.t, cam.focal, cam.pp, S.lodScale[0], p.center, p.imgPoint[i]);
    }
}

/* Patch::refine, :114-176 */
__device__ inline void cta_refine(const DevScene &S, CtaS &c, const double *sDistW, const WarpWork &W, double *Hc, double *xs,
                                  double *ys, double *corr, double *hp) {
    PatchS &p = c.p;
    const int tid = threadIdx.x;
    if (p.nCam < S.cfg.minCamNum) {
        if (tid == 0) { p.fitness = DBL_MAX; p.priority = DBL_MAX; p.drop = 1; }
        __syncthreads();
        return;
    }
    cta_update_information(S, p);                                                     /* :125-128 */
    if (p.drop) return;
    int beforeRef = p.refCamIdx, afterRef = -1, beforeNum = p.nCam, afterNum = -1, count = 0;
    const int totalCamNum = beforeNum;
    while ((beforeRef != afterRef || beforeNum != afterNum) && count++ <= totalCamNum) {
        if (p.nCam < S.cfg.minCamNum) {
            if (tid == 0) { p.fitness = DBL_MAX; p.priority = DBL_MAX; p.drop = 1; }
            __syncthreads();
            return;
        }
        beforeRef = p.refCamIdx;
        beforeNum = p.nCam;
        cta_pso_optimization(S, c, sDistW, W);                                        /* :153 */
        if (p.fitness > S.cfg.maxFitness) {                                           /* :156-159 */
            __syncthreads();
            if (tid == 0) p.drop = 1;
            __syncthreads();
            return;
        }
        cta_remove_invisible(S, c, Hc, xs, ys, corr, hp);                             /* :162 */
        cta_update_information(S, p);                                                 /* :163-166 */
        if (p.type == PMVS_TYPE_EXPAND) break;
        afterRef = p.refCamIdx;
        afterNum = p.nCam;
        __syncthreads();
    }
    if (tid == 0) t0_priority_imgpoint(S, p);                                         /* :174-175 */
    __syncthreads();
}
