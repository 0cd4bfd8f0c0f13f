/*
 * pmvs_b200.cu — kernels and the C-ABI (include/pmvs_b200.h) of the B200-native patch-refinement path.
 *
 * Kernels:
 *   pack_quad_kernel      u8 level image -> 4-taps-per-word "quad" layout (see pmvs_device.cuh)
 *   fitness_batch_kernel  seam 1: one warp = one PAIS::getFitness call (patch.cpp:914-1047)
 *   refine_kernel         seam 2: persistent CTAs, one CTA = one Patch::refine() (+ trailing removeInvisibleCamera)
 *   pso_test_kernel       the swarm alone on analytic functions (known-answer tests against the unmodified
 *                         reference solver)
 * There is no CPU fallback: every compute entry point needs an sm_100 device.
 */
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "pmvs_patch.cuh"
#include "pmvs_pyramid.cuh"
#include "pmvs_filter.cuh"

#define PMVS_VERSION "pmvs_b200 0.1 (sm_100a)"
/* the alternative register budget of refine_kernel: __launch_bounds__(PMVS_ALT_T, PMVS_ALT_B) (160 x 4 = 96 registers) */
#ifndef PMVS_ALT_T
#define PMVS_ALT_T 160
#define PMVS_ALT_B 4
#endif
#ifndef PMVS_DEFAULT_LEAN
#define PMVS_DEFAULT_LEAN true
#endif

/* ======================================================================================================= */
__global__ void pack_quad_kernel(const uint8_t *__restrict__ grey, size_t pitch, int cols, int rows, uint32_t *__restrict__ quad) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols || y >= rows) return;
    const int x1 = min(x + 1, cols - 1), y1 = min(y + 1, rows - 1);
    const uint32_t g00 = grey[(size_t)y * pitch + x], g01 = grey[(size_t)y * pitch + x1];
    const uint32_t g10 = grey[(size_t)y1 * pitch + x], g11 = grey[(size_t)y1 * pitch + x1];
    quad[(size_t)y * cols + x] = g00 | (g01 << 8) | (g10 << 16) | (g11 << 24);
}

/* carve the dynamic shared memory of one CTA */
struct SmemPlan {
    size_t ctaOff, viewOff, distOff, warpOff, corrOff, refWinOff, total;
    size_t perWarp;      /* doubles per warp: H + xs + ys + column constants + colour stash */
    int vcap, ps, nWarps, slotViews, nRefWin, useVL, grad, stash;
};
/* nRefWin: reference windows (RefWin tables) the CTA holds — one per CTA (refine), one per warp (fitness), 0 = none */
static SmemPlan plan_smem(int vcap, int ps, int nWarps, bool withCorr, size_t headBytes, bool useVL, int nRefWin, bool grad) {
    SmemPlan pl;
    pl.vcap = vcap;
    pl.ps = ps;
    pl.nWarps = nWarps;
    pl.useVL = useVL ? 1 : 0;
    pl.grad = grad ? 1 : 0;
    pl.slotViews = PMVS_SLOT_VIEWS_OF(vcap, useVL);
    pl.nRefWin = nRefWin;
    pl.stash = nRefWin ? PMVS_STASH_DOUBLES(vcap, useVL, grad) : 0;
    size_t off = 0;
    pl.ctaOff = off;
    off += (headBytes + 15) & ~(size_t)15;
    pl.viewOff = off;
    off += sizeof(ViewS) * (size_t)vcap;
    off = (off + 15) & ~(size_t)15;
    pl.distOff = off;
    off += sizeof(double) * ((size_t)PMVS_DIST_PAD(ps) + 64);     /* distance weights + exp table */
    pl.warpOff = off;
    pl.perWarp = (((size_t)PMVS_HCAP(vcap) * 9 + 2 * (size_t)ps + 1) & ~(size_t)1) + PMVS_HYP_DOUBLES +
                 PMVS_COLV_DOUBLES_N(vcap, ps, pl.slotViews, pl.useVL && nRefWin) + pl.stash;
    off += sizeof(double) * pl.perWarp * nWarps;
    pl.corrOff = off;
    if (withCorr && !PMVS_CORR_GLOBAL(vcap)) off += sizeof(double) * (size_t)vcap * vcap;
    off = (off + 15) & ~(size_t)15;
    pl.refWinOff = off;
    off += sizeof(double) * PMVS_REFWIN_DOUBLES(ps, grad) * (size_t)nRefWin;
    pl.total = off;
    return pl;
}
struct SmemArgs {
    unsigned ctaOff, viewOff, distOff, warpOff, corrOff, perWarp, refWinOff;
    int vcap, ps, slotViews, nRefWin, useVL, grad, stash;
};
static SmemArgs to_args(const SmemPlan &pl) {
    SmemArgs a;
    a.ctaOff = (unsigned)pl.ctaOff;
    a.viewOff = (unsigned)pl.viewOff;
    a.distOff = (unsigned)pl.distOff;
    a.warpOff = (unsigned)pl.warpOff;
    a.corrOff = (unsigned)pl.corrOff;
    a.perWarp = (unsigned)pl.perWarp;
    a.refWinOff = (unsigned)pl.refWinOff;
    a.vcap = pl.vcap;
    a.ps = pl.ps;
    a.slotViews = pl.slotViews;
    a.nRefWin = pl.nRefWin;
    a.useVL = pl.useVL && pl.nRefWin;
    a.grad = pl.grad;
    a.stash = pl.stash;
    return a;
}
/* per-warp area: H | xs | ys | hyp | [slots | view table | rowf | rowi] (column-lane loop only) | colour stash */
__device__ __forceinline__ WarpWork warp_work(unsigned char *smem, const SmemArgs &a, int warp) {
    double *base = (double *)(smem + a.warpOff) + (size_t)a.perWarp * warp;
    WarpWork W;
    W.H = base;
    W.xs = base + (size_t)PMVS_HCAP(a.vcap) * 9;
    W.ys = W.xs + a.ps;
    W.colv = base + a.perWarp - a.stash - PMVS_COLV_DOUBLES_N(a.vcap, a.ps, a.slotViews, a.useVL);
    W.hyp = W.colv - PMVS_HYP_DOUBLES;
    W.slotViews = a.slotViews;
    W._padw = 0;
    W.gv = a.useVL ? nullptr : W.colv + PMVS_COLV_SLOTS_N(a.slotViews);
    W.rowf = a.useVL ? nullptr : W.gv + PMVS_GV_DOUBLES_N(a.vcap, a.slotViews, a.useVL);
    W.rowi = a.useVL ? nullptr : (int2 *)(W.rowf + PMVS_PS_PAD(a.ps));
    W.rw = nullptr;
    W.stash = a.stash ? base + a.perWarp - a.stash : nullptr;
    return W;
}

/* ---- seam 1 ------------------------------------------------------------------------------------------- */
struct FitWarpS {
    EvalCtx E;
    RefWin rw;
};
__global__ void __launch_bounds__(128) fitness_batch_kernel(const __grid_constant__ DevScene S, const SmemArgs a, int n,
                                                            const PmvsHypothesis *__restrict__ in, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    double *sDistW = (double *)(smem + a.distOff);
    for (int k = tid; k < a.ps * a.ps && !PMVS_DIST_GLOBAL; k += blockDim.x) sDistW[k] = S.distW[k];
    load_exp_table(sDistW, a.ps, tid, blockDim.x);
    /* one EvalCtx + view table per warp: ctaOff holds NW contexts, viewOff NW*vcap views */
    EvalCtx &E = ((FitWarpS *)(smem + a.ctaOff))[warp].E;
    RefWin &R = ((FitWarpS *)(smem + a.ctaOff))[warp].rw;
    if (lane == 0) {
        E.view = (ViewS *)(smem + a.viewOff) + (size_t)warp * a.vcap;
        if (a.nRefWin) carve_ref_win(R, (double *)(smem + a.refWinOff) + PMVS_REFWIN_DOUBLES(a.ps, a.grad) * (size_t)warp, a.ps, a.grad != 0);
    }
    __syncthreads();
    WarpWork W = warp_work(smem, a, warp);
    if (a.nRefWin) W.rw = &R;
    for (int h = blockIdx.x * NW + warp; h < n; h += gridDim.x * NW) {
        const PmvsHypothesis &hy = in[h];
        const double ray[3] = {hy.ray[0], hy.ray[1], hy.ray[2]};
        __syncwarp();
        build_eval_ctx(S, E, ray, hy.refCamIdx, hy.LOD, hy.nCam, hy.camIdx, lane, 32);
        __syncwarp();
        finish_eval_ctx(E, lane);
        __syncwarp();
        if (a.nRefWin) {       /* every hypothesis is its own patch here: its reference window is built for it alone */
            double ctr[3];
            for (int k = 0; k < 3; ++k) ctr[k] = E.ray[k] * hy.depth + E.refC[k];
            build_ref_win<true>(S, E, R, ctr, lane, 32);
        }
        const double f = warp_fitness_any(S, E, sDistW, W, hy.theta, hy.phi, hy.depth);
        if (lane == 0) out[h] = f;
    }
}

/* ---- seam 2 ------------------------------------------------------------------------------------------- */
__device__ inline void patch_from_in(const PmvsPatchIn &in, PatchS &p) {   /* AbstractPatch::init, abstractpatch.cpp:25-40 */
    for (int k = 0; k < 3; ++k) {
        p.center[k] = in.center[k];
        p.normal[k] = in.normal[k];
        p.ray[k] = 0;
    }
    p.normalS[0] = in.normalS[0];
    p.normalS[1] = in.normalS[1];
    p.depth = 0;
    p.depthRange[0] = p.depthRange[1] = 0;
    p.fitness = DBL_MAX;
    p.priority = DBL_MAX;
    p.correlation = 0;
    p.LOD = -1;
    p.refCamIdx = -1;
    p.type = in.type;
    p.id = in.id;
    p.drop = 0;
    int n = in.nCam;
    n = n < 0 ? 0 : (n > PMVS_MAX_VIEWS ? PMVS_MAX_VIEWS : n);
    p.nCam = n;
    for (int i = 0; i < n; ++i) p.camIdx[i] = in.camIdx[i];
    p.psoRuns = 0;
    p.psoIterations = 0;
    p.nImgPoint = 0;
    p.evals = 0;
    p.status = 0;
    p.windowEvals = 0;
    p.flag = 0;
    p.visValid = 0;
}
__device__ inline void patch_to_out(const PatchS &p, PmvsPatchOut &o) {
    for (int k = 0; k < 3; ++k) {
        o.center[k] = p.center[k];
        o.normal[k] = p.normal[k];
        o.ray[k] = p.ray[k];
    }
    o.normalS[0] = p.normalS[0];
    o.normalS[1] = p.normalS[1];
    o.depth = p.depth;
    o.depthRange[0] = p.depthRange[0];
    o.depthRange[1] = p.depthRange[1];
    o.fitness = p.fitness;
    o.priority = p.priority;
    o.correlation = p.correlation;
    o.LOD = p.LOD;
    o.refCamIdx = p.refCamIdx;
    o.nCam = p.nCam;
    o.drop = p.drop ? 1 : 0;
    o.psoRuns = p.psoRuns;
    o.psoIterations = p.psoIterations;
    o.evaluations = p.evals;
    o.status = p.status;
    o.windowEvaluations = p.windowEvals;
    for (int i = 0; i < p.nCam; ++i) o.camIdx[i] = p.camIdx[i];
    o.nImgPoint = p.nImgPoint;
    for (int i = 0; i < p.nImgPoint; ++i) {
        o.imgPoint[i][0] = p.imgPoint[i][0];
        o.imgPoint[i][1] = p.imgPoint[i][1];
    }
}

/*
 * Work distribution of the persistent CTAs. The batch is cut into nChunks contiguous index ranges, one per SM; the
 * CTAs resident on an SM pull consecutive indices from their SM's range, so the patches an SM works on at the same time
 * are neighbours in batch order. Callers hand candidates over in spatial order (grid scan / sorted by reference cell),
 * so neighbouring patches share most of their image footprint and the SM's L1 holds one footprint instead of one
 * per CTA. A CTA whose home range is exhausted steals from the following ranges; counters: counter[0..nChunks).
 * nChunks = 1 is the plain shared counter.
 */
#define PMVS_MAX_CHUNKS 1024
__device__ __forceinline__ int next_patch(int *counter, int n, int nChunks, int home) {
    for (int t = 0; t < nChunks; ++t) {
        int k = home + t;
        if (k >= nChunks) k -= nChunks;
        const int lo = (int)((long long)n * k / nChunks), len = (int)((long long)n * (k + 1) / nChunks) - lo;
        if (len > 0 && *(volatile int *)(counter + k) < len) {
            const int i = atomicAdd(counter + k, 1);
            if (i < len) return lo + i;
        }
    }
    return n;
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) refine_kernel(const __grid_constant__ DevScene S, const SmemArgs a, int n,
                                                        const PmvsPatchIn *__restrict__ in, PmvsPatchOut *__restrict__ out,
                                                        uint32_t flags, int *__restrict__ counter, int nChunks) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    CtaS &c = *(CtaS *)(smem + a.ctaOff);
    double *sDistW = (double *)(smem + a.distOff);
    double *hp = S.scratch + (size_t)S.scratchStride * blockIdx.x;
    double *corr = PMVS_CORR_GLOBAL(a.vcap) ? hp + (size_t)a.vcap * a.ps * a.ps : (double *)(smem + a.corrOff);
    for (int k = tid; k < a.ps * a.ps && !PMVS_DIST_GLOBAL; k += blockDim.x) sDistW[k] = S.distW[k];
    load_exp_table(sDistW, a.ps, tid, blockDim.x);
    if (tid == 0) {
        c.E.view = (ViewS *)(smem + a.viewOff);
        c.part = (ParticleS *)(smem + a.ctaOff + ((sizeof(CtaS) + 15) & ~(size_t)15));
    }
    WarpWork W = warp_work(smem, a, warp);
    const WarpWork W0 = warp_work(smem, a, 0);
    if (a.nRefWin) {
        if (tid == 0) carve_ref_win(c.rw, (double *)(smem + a.refWinOff), a.ps, a.grad != 0);
        W.rw = &c.rw;
    }
    int home = 0;
    if (nChunks > 1) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        home = (int)(smid % (unsigned)nChunks);
    }
    __syncthreads();
    for (;;) {
        if (tid == 0) c.nextIdx = next_patch(counter, n, nChunks, home);
        __syncthreads();
        const int idx = c.nextIdx;
        if (idx >= n) break;
        if (tid == 0) {
            patch_from_in(in[idx], c.p);
            /* records this launch cannot hold are dropped in-band with a status bit instead of being evaluated:
             * more camera entries than the scene's view tables were carved for (duplicate indices in a caller's list),
             * or an index that is not a camera of the scene (corrupt .mvs / raw C-ABI caller) */
            if (in[idx].nCam > a.vcap || in[idx].nCam < 0) {
                c.p.status |= PMVS_S_TOO_MANY_VIEWS;
                c.p.drop = 1;
                c.p.nCam = 0;
            }
            for (int i = 0; i < c.p.nCam; ++i)
                if ((int)c.p.camIdx[i] >= S.nCams) {
                    c.p.status |= PMVS_S_BAD_CAMERA;
                    c.p.drop = 1;
                }
            if (c.p.status & PMVS_S_BAD_CAMERA) c.p.nCam = 0;
        }
        __syncthreads();
        if (!c.p.drop) {
            if ((flags & PMVS_F_EXPAND_VISIBLE) && c.p.type == PMVS_TYPE_EXPAND) cta_expand_visible(S, c.p);
            cta_refine(S, c, sDistW, W, W0.H, W0.xs, W0.ys, corr, hp);
            __syncthreads();
            if (flags & PMVS_F_POST_REMOVE_INVISIBLE) cta_remove_invisible(S, c, W0.H, W0.xs, W0.ys, corr, hp);   /* mvs.cpp:215, :574 */
        }
        __syncthreads();
        /* assemble the record in shared memory, then store it with coalesced 32-bit words */
        uint32_t *ow = (uint32_t *)&c.out;
        for (int k = tid; k < (int)(sizeof(PmvsPatchOut) / 4); k += blockDim.x) ow[k] = 0;
        __syncthreads();
        if (tid == 0) patch_to_out(c.p, c.out);
        __syncthreads();
        uint32_t *gw = (uint32_t *)(out + idx);
        for (int k = tid; k < (int)(sizeof(PmvsPatchOut) / 4); k += blockDim.x) gw[k] = ow[k];
        __syncthreads();
    }
}

/* ---- exchange record of the multi-GPU driver: {centre 3, normal 3, fitness, drop} as 8 doubles per patch ---------- */
__global__ void pack_records_kernel(int n, const PmvsPatchOut *__restrict__ out, double *__restrict__ rec, unsigned long long *__restrict__ counters) {
    unsigned long long ev = 0, wev = 0, kept = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const PmvsPatchOut &o = out[i];
        double *r = rec + 8 * (size_t)i;
        r[0] = o.center[0]; r[1] = o.center[1]; r[2] = o.center[2];
        r[3] = o.normal[0]; r[4] = o.normal[1]; r[5] = o.normal[2];
        r[6] = o.fitness;
        r[7] = (double)o.drop;
        ev += o.evaluations;
        wev += o.windowEvaluations;
        kept += o.drop ? 0 : 1;
    }
    if (counters) {
        for (int off = 16; off > 0; off >>= 1) {
            ev += __shfl_xor_sync(0xffffffffu, ev, off);
            wev += __shfl_xor_sync(0xffffffffu, wev, off);
            kept += __shfl_xor_sync(0xffffffffu, kept, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters, ev);
            atomicAdd(counters + 1, wev);
            atomicAdd(counters + 2, kept);
        }
    }
}

/* ---- swarm on analytic functions (test support) ------------------------------------------------------- */
struct TestEval {
    int fn;
    __device__ __forceinline__ double operator()(const double *x) const {
        switch (fn) {
        default:
        case 0: return (x[0] - 0.3) * (x[0] - 0.3) + (x[1] + 0.2) * (x[1] + 0.2) + (x[2] - 1.5) * (x[2] - 1.5);
        case 1: {
            double aa = x[1] - x[0] * x[0], b = 1 - x[0], cc = x[2] - x[1] * x[1], d = 1 - x[1];
            return 100 * aa * aa + b * b + 100 * cc * cc + d * d;
        }
        case 3: return (x[0] > 0.5) ? DBL_MAX : fabs(x[0]) + fabs(x[1]) + fabs(x[2]);
        case 4: return floor(4 * fabs(x[0])) + floor(4 * fabs(x[1])) + floor(4 * fabs(x[2]));
        }
    }
    __device__ __forceinline__ void batch(int m, ParticleS *part, int p0, int stride) const {
        for (int k = 0; k < m; ++k) {
            const double f = (*this)(part[p0 + k * stride].pos);
            if ((threadIdx.x & 31) == 0) part[p0 + k * stride].fitness = f;
        }
        __syncwarp();
    }
};
#define PMVS_PSO_TEST_MAX 64          /* particle rows of the test ABI's `particles` output */
struct PsoTestS {
    PsoS pso;
    ParticleS part[PMVS_PSO_TEST_MAX];
    MoveS mv;
    double init[3];
};
/* one CTA per problem; io = per problem {L3,U3,init3,hasInit,maxIter,P,fn,key(as 2 doubles via bits)} */
struct PsoTestProblem {
    double L[3], U[3], init[3];
    uint64_t key;
    int hasInit, maxIter, P, fn;
};
struct PsoTestResult {
    double gbest[3], gbestFitness;
    int iterations, _pad;
    double particles[PMVS_PSO_TEST_MAX][8];
};
__global__ void __launch_bounds__(512) pso_test_kernel(int n, const PsoTestProblem *__restrict__ prob, PsoTestResult *__restrict__ res) {
    __shared__ PsoTestS s;
    const int tid = threadIdx.x;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const PsoTestProblem &pr = prob[b];
        if (tid == 0) {
            pso_setup(s.pso, pr.L, pr.U, pr.maxIter, pr.P, pr.key);
            for (int d = 0; d < 3; ++d) s.init[d] = pr.init[d];
        }
        __syncthreads();
        TestEval ev = {pr.fn};
        pso_run(s.pso, s.part, s.mv, ev, s.init, pr.hasInit != 0);
        if (tid == 0) {
            PsoTestResult &r = res[b];
            const ParticleS &g = s.part[s.pso.gBestIdx];
            for (int d = 0; d < 3; ++d) r.gbest[d] = g.pBest[d];
            r.gbestFitness = s.pso.gBestFitness;
            r.iterations = s.pso.iteration;
            for (int i = 0; i < pr.P; ++i) {
                const ParticleS &q = s.part[i];
                for (int d = 0; d < 3; ++d) {
                    r.particles[i][d] = q.pos[d];
                    r.particles[i][3 + d] = q.vec[d];
                }
                r.particles[i][6] = q.fitness;
                r.particles[i][7] = q.pbf;
            }
        }
        __syncthreads();
    }
}

/* =======================================================================================================
 * host side: context + C-ABI
 * ===================================================================================================== */
struct pmvs_ctx {
    int device = -1;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    PmvsConfig cfg;
    int nCams = 0;
    int vcap = 0;
    uint64_t seed = 0;
    std::vector<void *> allocs;          /* pyramid storage */
    DevCamera *dCams = nullptr;
    double *dDistW = nullptr;
    double *dScratch = nullptr;
    size_t scratchStride = 0;
    int scratchCtas = 0;
    int *dCounter = nullptr;
    int *dDbgFoot = nullptr;             /* experiment builds only (PMVS_FOOTPRINT) */
    cudaEvent_t lastLaunch = nullptr;    /* refine launches of one context share its work counters and scratch slabs: they are serialised */
    void *dIn = nullptr, *dOut = nullptr;
    size_t dInBytes = 0, dOutBytes = 0;
    int64_t launches = 0;
    std::string err;
    DevScene scene;
};

static int fail(pmvs_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? PMVS_E_NOMEM : PMVS_E_CUDA,                \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                              \
    } while (0)

/* MVS::initPatchDistanceWeighting, mvs.cpp:97-114 */
static std::vector<double> dist_weight(const PmvsConfig &cfg) {
    const int ps = cfg.patchSize, r = cfg.patchRadius;
    std::vector<double> w((size_t)ps * ps);
    const double sigma = cfg.distWeighting;
    const double s2 = 1.0 / (2.0 * sigma * sigma);
    const double s = 1.0 / (2.0 * M_PI * sigma * sigma);
    for (int x = 0; x < ps; ++x)
        for (int y = 0; y < ps; ++y) {
            const double e = -(pow((double)(x - r), 2) + pow((double)(y - r), 2)) * s2;
            w[(size_t)x * ps + y] = s * exp(e);
        }
    /* cv::sum on CV_64F adds four elements at a time; `Mat / n` multiplies by 1./n (OpenCV 2.4 stat.cpp, matop.cpp) */
    double n = 0;
    size_t i = 0;
    for (; i + 4 <= w.size(); i += 4) n += w[i] + w[i + 1] + w[i + 2] + w[i + 3];
    for (; i < w.size(); ++i) n += w[i];
    const double rn = 1. / n;
    for (i = 0; i < w.size(); ++i) w[i] = w[i] * rn;
    return w;
}

static int check_config(pmvs_ctx *ctx, const PmvsConfig &cfg) {
    if (cfg.patchRadius < 0 || cfg.patchRadius > PMVS_MAX_RADIUS)
        return fail(ctx, PMVS_E_UNSUPPORTED, "patchRadius must be in [0, " + std::to_string(PMVS_MAX_RADIUS) + "]");
    if (cfg.particleNum < 1 || cfg.particleNum > PMVS_MAX_PARTICLE_NUM)
        return fail(ctx, PMVS_E_UNSUPPORTED, "particleNum must be in [1, " + std::to_string(PMVS_MAX_PARTICLE_NUM) + "]");
    if (cfg.maxIteration < 0) return fail(ctx, PMVS_E_ARG, "maxIteration < 0");
    if (!(cfg.lodRatio > 0.0 && cfg.lodRatio <= 1.0)) return fail(ctx, PMVS_E_ARG, "lodRatio must be in (0,1]");
    return PMVS_OK;
}

static int apply_config(pmvs_ctx *ctx, const PmvsConfig *cfg) {
    int rc = check_config(ctx, *cfg);
    if (rc) return rc;
    ctx->cfg = *cfg;
    ctx->cfg.patchSize = (cfg->patchRadius << 1) + 1;                               /* mvs.cpp:67 */
    const std::vector<double> w = dist_weight(ctx->cfg);
    if (ctx->dDistW) { cudaFree(ctx->dDistW); ctx->dDistW = nullptr; }
    /* the table, followed by its separable factor: w[x][y] = g[x] g[y] / (sum g)^2 (ones when the weight is disabled) */
    std::vector<double> wg(w);
    {
        const int ps = ctx->cfg.patchSize, r = ctx->cfg.patchRadius;
        const double s2 = 1.0 / (2.0 * ctx->cfg.distWeighting * ctx->cfg.distWeighting);
        std::vector<double> g((size_t)ps, 1.0);
        if (ctx->cfg.adaptiveDistanceEnable) {
            double sum = 0;
            for (int k = 0; k < ps; ++k) { g[k] = exp(-pow((double)(k - r), 2) * s2); sum += g[k]; }
            for (int k = 0; k < ps; ++k) g[k] = g[k] / sum;
        }
        wg.insert(wg.end(), g.begin(), g.end());
    }
    CK(cudaMalloc(&ctx->dDistW, wg.size() * sizeof(double)));
    CK(cudaMemcpy(ctx->dDistW, wg.data(), wg.size() * sizeof(double), cudaMemcpyHostToDevice));
    DevScene &s = ctx->scene;
    s.cfg = ctx->cfg;
    s.cams = ctx->dCams;
    s.distW = ctx->dDistW;
    s.distG = ctx->dDistW + w.size();
    s.dbgFoot = nullptr;
#if PMVS_FOOTPRINT
    if (!ctx->dDbgFoot) CK(cudaMalloc(&ctx->dDbgFoot, sizeof(int) * 65536 * 32));
    CK(cudaMemset(ctx->dDbgFoot, 0, sizeof(int) * 65536 * 32));
    s.dbgFoot = ctx->dDbgFoot;
#endif
    s.nCams = ctx->nCams;
    s.seed = ctx->seed;
    {
        const char *envVL = getenv("PMVS_VL");          /* tuning / A-B: PMVS_VL=0 keeps the column-lane loop */
        s.useVL = (envVL && atoi(envVL) == 0) ? 0 : 1;
        if (const char *envG = getenv("PMVS_ROWS_GLN")) s.useVL |= (atoi(envG) & 15) << 4;      /* tuning: lanes per pixel of fitness_vl_rows */
    }
    for (int l = 0; l < PMVS_MAX_LEVELS; ++l) s.lodScale[l] = pow(ctx->cfg.lodRatio, l);
    /* correlation scratch: one slab per resident CTA */
    const size_t stride = (size_t)ctx->vcap * ctx->cfg.patchSize * ctx->cfg.patchSize + (size_t)ctx->vcap * ctx->vcap + ctx->vcap;   /* windows + (global) correlation table + ratios */
    const int ctas = ctx->smCount * 4;
    if (stride > ctx->scratchStride || ctas > ctx->scratchCtas || !ctx->dScratch) {
        if (ctx->dScratch) cudaFree(ctx->dScratch);
        ctx->dScratch = nullptr;
        CK(cudaMalloc(&ctx->dScratch, stride * ctas * sizeof(double)));
        ctx->scratchStride = stride;
        ctx->scratchCtas = ctas;
    }
    s.scratch = ctx->dScratch;
    s.scratchStride = ctx->scratchStride;
    return PMVS_OK;
}

extern "C" {

const char *pmvs_version(void) { return PMVS_VERSION; }
const char *pmvs_last_error(const pmvs_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
int64_t pmvs_launch_count(const pmvs_ctx *ctx) { return ctx ? ctx->launches : 0; }

void pmvs_destroy(pmvs_ctx *ctx) {
    if (!ctx) return;
    if (ctx->device >= 0) cudaSetDevice(ctx->device);
    for (void *p : ctx->allocs) cudaFree(p);
    if (ctx->dCams) cudaFree(ctx->dCams);
    if (ctx->dDistW) cudaFree(ctx->dDistW);
    if (ctx->dScratch) cudaFree(ctx->dScratch);
    if (ctx->dCounter) cudaFree(ctx->dCounter);
    if (ctx->lastLaunch) cudaEventDestroy(ctx->lastLaunch);
    if (ctx->dIn) cudaFree(ctx->dIn);
    if (ctx->dOut) cudaFree(ctx->dOut);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}


/* ---- device pyramid builder (camera.cpp:63-92), shared by pmvs_create and pmvs_build_pyramid ----------------- */
static inline int cv_round(double v) { return (int)std::nearbyint(v); }

struct DevU8 {
    uint8_t *p = nullptr;
    size_t pitch = 0;
    int cols = 0, rows = 0;
};

static cudaError_t upload_u8(const uint8_t *host, int64_t hpitch, int cols, int rows, DevU8 &d, cudaStream_t st) {
    cudaError_t e = cudaMallocPitch(&d.p, &d.pitch, (size_t)cols, (size_t)rows);
    if (e != cudaSuccess) return e;
    d.cols = cols;
    d.rows = rows;
    return cudaMemcpy2DAsync(d.p, d.pitch, host, (size_t)hpitch, (size_t)cols, (size_t)rows, cudaMemcpyHostToDevice, st);
}

/* level l (cols x rows) = INTER_AREA resize of `src` by factor f; launches counted in *launches */
static cudaError_t resize_level(const DevU8 &src, double f, int dcols, int drows, DevU8 &dst, cudaStream_t st, int64_t *launches) {
    AreaTab tx, ty;
    const double scale = 1.0 / f;
    const int iscale = cv_round(scale);
    if (std::fabs(scale - iscale) < DBL_EPSILON) {            /* is_area_fast (OpenCV resize.cpp) */
        cudaError_t e = cudaMallocPitch(&dst.p, &dst.pitch, (size_t)dcols, (size_t)drows);
        if (e != cudaSuccess) return e;
        dst.cols = dcols;
        dst.rows = drows;
        dim3 block(32, 8), grid((dcols + 31) / 32, (drows + 7) / 8);
        resize_area_fast_kernel<<<grid, block, 0, st>>>(src.p, src.pitch, src.cols, src.rows, dst.p, dst.pitch, dcols, drows, iscale);
        ++*launches;
        e = cudaGetLastError();
        return e == cudaSuccess ? cudaStreamSynchronize(st) : e;
    }
    build_area_tab(src.cols, dcols, scale, tx);
    build_area_tab(src.rows, drows, scale, ty);
    cudaError_t e = cudaMallocPitch(&dst.p, &dst.pitch, (size_t)dcols, (size_t)drows);
    if (e != cudaSuccess) return e;
    dst.cols = dcols;
    dst.rows = drows;
    const size_t nI = (size_t)2 * (dcols + drows), nW = tx.w.size() + ty.w.size();
    int *dI = nullptr;
    float *dW = nullptr;
    if ((e = cudaMalloc(&dI, nI * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&dW, nW * sizeof(float))) != cudaSuccess) { cudaFree(dI); return e; }
    std::vector<int> hI;
    hI.insert(hI.end(), tx.start.begin(), tx.start.end());
    hI.insert(hI.end(), tx.count.begin(), tx.count.end());
    hI.insert(hI.end(), ty.start.begin(), ty.start.end());
    hI.insert(hI.end(), ty.count.begin(), ty.count.end());
    std::vector<float> hW(tx.w);
    hW.insert(hW.end(), ty.w.begin(), ty.w.end());
    e = cudaMemcpyAsync(dI, hI.data(), nI * sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dW, hW.data(), nW * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        dim3 block(32, 8), grid((dcols + 31) / 32, (drows + 7) / 8);
        resize_area_kernel<<<grid, block, 0, st>>>(src.p, src.pitch, src.cols, src.rows, dst.p, dst.pitch, dcols, drows, dI, dI + dcols,
                                                    dW, tx.maxTaps, dI + 2 * dcols, dI + 2 * dcols + drows, dW + tx.w.size(), ty.maxTaps);
        ++*launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);     /* the host tables must outlive the copies */
    cudaFree(dI);
    cudaFree(dW);
    return e;
}

/* min-max normalised gradient magnitude of one level into `edge` (cols*rows doubles on the device) */
static cudaError_t edge_level(const DevU8 &g, double *edge, unsigned long long *dMinMax, cudaStream_t st, int64_t *launches) {
    const unsigned long long init[2] = {~0ull, 0ull};
    cudaError_t e = cudaMemcpyAsync(dMinMax, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    dim3 grid((g.cols + 255) / 256, g.rows < 1024 ? g.rows : 1024);
    {   /* the reduction pass: a few CTAs per SM striding the image, so the same-address atomics stay in the hundreds */
        int gx = (g.cols + 255) / 256, gy = 1184 / gx;
        gy = gy < 1 ? 1 : (gy > g.rows ? g.rows : gy);
        edge_minmax_kernel<<<dim3(gx, gy), 256, 0, st>>>(g.p, g.pitch, g.cols, g.rows, dMinMax);
    }
    edge_normalise_kernel<<<grid, 256, 0, st>>>(g.p, g.pitch, g.cols, g.rows, dMinMax, edge);
    *launches += 2;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);     /* `init` is a stack buffer */
    return e;
}

static void level_dims(int cols0, int rows0, double lodRatio, int l, int &c, int &r) {   /* cv::resize rounds the size */
    const double f = pow(lodRatio, l);
    c = l == 0 ? cols0 : cv_round(cols0 * f);
    r = l == 0 ? rows0 : cv_round(rows0 * f);
}

static int create_impl(pmvs_ctx *ctx, const PmvsConfig *cfg, int nCams, const PmvsCamera *cams, int device, uint64_t rngSeed) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(ctx, PMVS_E_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                                          " (this library has no CPU path)");
    if (device < 0 || device >= count) return fail(ctx, PMVS_E_ARG, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(ctx, PMVS_E_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                          std::to_string(prop.minor) + "; this library is built for sm_100a only");
    ctx->device = device;
    ctx->smCount = prop.multiProcessorCount;
    ctx->seed = rngSeed;
    ctx->nCams = nCams;
    ctx->vcap = nCams < PMVS_MAX_VIEWS ? nCams : PMVS_MAX_VIEWS;
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&ctx->dCounter, sizeof(int) * PMVS_MAX_CHUNKS));

    std::vector<DevCamera> hc(nCams);
    for (int i = 0; i < nCams; ++i) {
        const PmvsCamera &c = cams[i];
        DevCamera &d = hc[i];
        memset(&d, 0, sizeof(d));
        memcpy(d.center, c.center, sizeof(d.center));
        memcpy(d.R, c.R, sizeof(d.R));
        memcpy(d.t, c.t, sizeof(d.t));
        memcpy(d.KR, c.KR, sizeof(d.KR));
        memcpy(d.KT, c.KT, sizeof(d.KT));
        memcpy(d.optN, c.opticalNormal, sizeof(d.optN));
        memcpy(d.focal, c.focal, sizeof(d.focal));
        memcpy(d.pp, c.principal, sizeof(d.pp));
        if (c.maxLOD < 0 || c.maxLOD >= PMVS_MAX_LEVELS) return fail(ctx, PMVS_E_ARG, "camera maxLOD out of range");
        d.maxLOD = c.maxLOD;
        /* level 0 must come from the host; any further level whose grey pointer is NULL is built on the device
         * (INTER_AREA from level 0, camera.cpp:81-85); edges likewise when the config needs them (camera.cpp:71-92) */
        const PmvsLevel &L0 = c.level[0];
        if (L0.cols <= 0 || L0.rows <= 0 || !L0.grey || L0.pitch < L0.cols) return fail(ctx, PMVS_E_ARG, "bad pyramid level 0");
        DevU8 g0;
        e = upload_u8(L0.grey, L0.pitch, L0.cols, L0.rows, g0, ctx->stream);
        if (e != cudaSuccess) { cudaFree(g0.p); return fail(ctx, PMVS_E_CUDA, std::string("level upload: ") + cudaGetErrorString(e)); }
        unsigned long long *dMinMax = nullptr;
        for (int l = 0; l <= c.maxLOD && e == cudaSuccess; ++l) {
            const PmvsLevel &L = c.level[l];
            DevU8 gl;
            if (l == 0) gl = g0;
            else if (L.grey) {
                if (L.cols <= 0 || L.rows <= 0 || L.pitch < L.cols) { cudaFree(g0.p); return fail(ctx, PMVS_E_ARG, "bad pyramid level"); }
                e = upload_u8(L.grey, L.pitch, L.cols, L.rows, gl, ctx->stream);
            } else {
                int lc, lr;
                level_dims(L0.cols, L0.rows, cfg->lodRatio, l, lc, lr);
                if ((L.cols > 0 && L.cols != lc) || (L.rows > 0 && L.rows != lr) || lc <= 0 || lr <= 0) {
                    cudaFree(g0.p);
                    return fail(ctx, PMVS_E_ARG, "pyramid level size does not match round(size0 * lodRatio^l)");
                }
                e = resize_level(g0, pow(cfg->lodRatio, l), lc, lr, gl, ctx->stream, &ctx->launches);
            }
            uint32_t *quad = nullptr;
            if (e == cudaSuccess) e = cudaMalloc(&quad, (size_t)gl.cols * gl.rows * sizeof(uint32_t));
            if (e == cudaSuccess) {
                ctx->allocs.push_back(quad);
                dim3 grid((gl.cols + 255) / 256, gl.rows);
                pack_quad_kernel<<<grid, 256, 0, ctx->stream>>>(gl.p, gl.pitch, gl.cols, gl.rows, quad);
                ctx->launches++;
                e = cudaGetLastError();
            }
            d.level[l].quad = quad;
            d.level[l].cols = gl.cols;
            d.level[l].rows = gl.rows;
            d.level[l].edge = nullptr;
            if (e == cudaSuccess && (L.edge || cfg->adaptiveGradientEnable)) {
                double *edge = nullptr;
                e = cudaMalloc(&edge, (size_t)gl.cols * gl.rows * sizeof(double));
                if (e == cudaSuccess) {
                    ctx->allocs.push_back(edge);
                    d.level[l].edge = edge;
                    if (L.edge) e = cudaMemcpyAsync(edge, L.edge, (size_t)gl.cols * gl.rows * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
                    else {
                        if (!dMinMax) e = cudaMalloc(&dMinMax, 2 * sizeof(unsigned long long));
                        if (e == cudaSuccess) e = edge_level(gl, edge, dMinMax, ctx->stream, &ctx->launches);
                    }
                }
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (l > 0) cudaFree(gl.p);
        }
        cudaFree(g0.p);
        if (dMinMax) cudaFree(dMinMax);
        if (e != cudaSuccess) return fail(ctx, e == cudaErrorMemoryAllocation ? PMVS_E_NOMEM : PMVS_E_CUDA, std::string("pyramid build: ") + cudaGetErrorString(e));
    }
    CK(cudaMalloc(&ctx->dCams, sizeof(DevCamera) * (size_t)nCams));
    CK(cudaMemcpy(ctx->dCams, hc.data(), sizeof(DevCamera) * (size_t)nCams, cudaMemcpyHostToDevice));
    return apply_config(ctx, cfg);
}

int pmvs_create(pmvs_ctx **out, const PmvsConfig *cfg, int nCams, const PmvsCamera *cams, int device, uint64_t rngSeed) {
    if (!out) return PMVS_E_ARG;
    *out = nullptr;
    pmvs_ctx *ctx = new (std::nothrow) pmvs_ctx();
    if (!ctx) return PMVS_E_NOMEM;
    *out = ctx;      /* returned even on failure so pmvs_last_error() can explain; caller destroys it */
    if (!cfg || !cams || nCams <= 0 || nCams > 65535) return fail(ctx, PMVS_E_ARG, "bad arguments to pmvs_create");
    return create_impl(ctx, cfg, nCams, cams, device, rngSeed);
}

int pmvs_set_neighbor_radius(pmvs_ctx *ctx, double neighborRadius) {
    if (!ctx || ctx->device < 0) return PMVS_E_ARG;
    ctx->cfg.neighborRadius = neighborRadius;
    ctx->scene.cfg.neighborRadius = neighborRadius;
    return PMVS_OK;
}

int pmvs_set_config(pmvs_ctx *ctx, const PmvsConfig *cfg) {
    if (!ctx || !cfg || ctx->device < 0) return PMVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    return apply_config(ctx, cfg);
}

static int ensure_buffers(pmvs_ctx *ctx, size_t inBytes, size_t outBytes) {
    if (inBytes > ctx->dInBytes) {
        if (ctx->dIn) cudaFree(ctx->dIn);
        ctx->dIn = nullptr;
        ctx->dInBytes = 0;
        CK(cudaMalloc(&ctx->dIn, inBytes));
        ctx->dInBytes = inBytes;
    }
    if (outBytes > ctx->dOutBytes) {
        if (ctx->dOut) cudaFree(ctx->dOut);
        ctx->dOut = nullptr;
        ctx->dOutBytes = 0;
        CK(cudaMalloc(&ctx->dOut, outBytes));
        ctx->dOutBytes = outBytes;
    }
    return PMVS_OK;
}

int pmvs_fitness_batch(pmvs_ctx *ctx, int n, const PmvsHypothesis *in, double *outFitness) {
    if (!ctx || ctx->device < 0) return PMVS_E_ARG;
    if (n < 0 || (n > 0 && (!in || !outFitness))) return fail(ctx, PMVS_E_ARG, "bad arguments to pmvs_fitness_batch");
    if (n == 0) return PMVS_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_buffers(ctx, sizeof(PmvsHypothesis) * (size_t)n, sizeof(double) * (size_t)n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->dIn, in, sizeof(PmvsHypothesis) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    const bool useVL = ctx->scene.useVL != 0, grad = ctx->cfg.adaptiveGradientEnable != 0;
    /* per-warp contexts: NW EvalCtx + NW x MAX_VIEWS view records in the head / view areas, the per-warp area planned for
     * MAX_VIEWS views, one reference window per warp */
    int NW = 4;
    SmemPlan pl;
    for (;; NW >>= 1) {
        pl = plan_smem(PMVS_MAX_VIEWS, ctx->cfg.patchSize, NW, false, sizeof(FitWarpS) * NW, useVL, useVL ? NW : 0, grad);
        /* the view area must hold NW x MAX_VIEWS records: re-carve with the larger view table */
        const size_t extraViews = sizeof(ViewS) * (size_t)PMVS_MAX_VIEWS * (NW - 1);
        pl.distOff += extraViews;
        pl.warpOff += extraViews;
        pl.corrOff += extraViews;
        pl.refWinOff += extraViews;
        pl.total += extraViews;
        if (pl.total <= 227 * 1024 || NW == 1) break;
    }
    if (pl.total > 227 * 1024) return fail(ctx, PMVS_E_UNSUPPORTED, "fitness kernel does not fit on an SM with this patch size");
    SmemArgs a = to_args(pl);
    CK(cudaFuncSetAttribute(fitness_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
    int grid = (n + NW - 1) / NW;
    if (grid > ctx->smCount * 8) grid = ctx->smCount * 8;
    fitness_batch_kernel<<<grid, NW * 32, pl.total, ctx->stream>>>(ctx->scene, a, n, (const PmvsHypothesis *)ctx->dIn, (double *)ctx->dOut);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(outFitness, ctx->dOut, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PMVS_OK;
}

/* one launch configuration of refine_kernel: NW warps per CTA on one of the register budgets */
typedef void (*RefineFn)(const DevScene, const SmemArgs, int, const PmvsPatchIn *, PmvsPatchOut *, uint32_t, int *, int);
struct RefineCfg {
    int NW, perSm;
    RefineFn fn;
    SmemPlan pl;
};
static int refine_config(pmvs_ctx *ctx, int NW, RefineCfg &c) {
    int nPart = ctx->cfg.particleNum * 2;
    if (nPart > PMVS_MAX_PARTICLES) nPart = PMVS_MAX_PARTICLES;
    const size_t ctaBytes = ((sizeof(CtaS) + 15) & ~(size_t)15);
    c.NW = NW;
    const bool useVL = ctx->scene.useVL != 0;
    c.pl = plan_smem(ctx->vcap, ctx->cfg.patchSize, NW, true, ctaBytes + sizeof(ParticleS) * (size_t)nPart, useVL,
                     (useVL && ctx->vcap >= 2) ? 1 : 0, ctx->cfg.adaptiveGradientEnable != 0);
    /* three register budgets of the same kernel: 96 registers (20 warps/SM, NW <= 5), 128 (16 warps/SM, NW <= 8,
     * or one 16-warp CTA). Measured (8192 patches, P = 15): 3 views 217k vs 199k patches/s and 5 views 142k vs 137k in
     * favour of 96 registers, 8 views 54.8k vs 57.3k in favour of 128 (the wider view loops spill), 12 views equal. */
    const char *envRegs = getenv("PMVS_REGS");
    const bool lean = NW <= PMVS_ALT_T / 32 && (envRegs ? atoi(envRegs) != 128 : (PMVS_DEFAULT_LEAN && ctx->vcap <= 6));
    c.fn = lean ? (RefineFn)refine_kernel<PMVS_ALT_T, PMVS_ALT_B> : (NW > 8 ? (RefineFn)refine_kernel<512, 1> : (RefineFn)refine_kernel<256, 2>);
    const int warpsPerSm = lean ? PMVS_ALT_T * PMVS_ALT_B / 32 : 16;
    c.perSm = 0;
    if (c.pl.total > 227 * 1024) return PMVS_OK;          /* does not fit: perSm = 0 */
    CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.pl.total));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.perSm, c.fn, NW * 32, c.pl.total));
    if (c.perSm > warpsPerSm / NW) c.perSm = warpsPerSm / NW;
    return PMVS_OK;
}

static int refine_launch(pmvs_ctx *ctx, int n, const PmvsPatchIn *d_in, PmvsPatchOut *d_out, uint32_t flags, cudaStream_t st) {
    /* Warps per CTA: the swarm's P particles (2P for seeds) are evaluated in rounds of NW, one warp per particle.
     * Throughput configuration: small CTAs (4..8 warps) leave room for several patches per SM, so one patch's serial
     * phases (swarm bookkeeping, visibility) overlap another's evaluations; the NW that wastes the fewest warp slots
     * in the last round. Latency configurations: a batch that cannot fill the GPU anyway (the host driver's expansion
     * rounds are bounded by the reconstruction's frontier: a few hundred candidates per call) finishes sooner with
     * more warps per patch — 8 (two CTAs per SM) or 16 (one) — because a generation then takes fewer evaluation
     * rounds. Picked per launch by waves x rounds; results do not depend on the choice (tests). */
    const int P = ctx->cfg.particleNum;
    RefineCfg cfg;
    int NW = 0;
    if (const char *envNw = getenv("PMVS_NW")) {       /* tuning override, 4..16 */
        NW = atoi(envNw);
        if (NW < 4 || NW > 16) NW = 5;
        const int rc = refine_config(ctx, NW, cfg);
        if (rc) return rc;
    } else {
        /* candidates: useful warps per SM = CTAs/SM x NW x (P / (rounds x NW)) for throughput; waves x rounds for latency */
        static const int cand[] = {4, 5, 6, 7, 8, 16};
        RefineCfg all[6];
        double bestThr = -1;
        int bestT = -1;
        for (int k = 0; k < 6; ++k) {
            all[k].perSm = 0;
            if (cand[k] == 16 && P <= 8) continue;
            const int rc = refine_config(ctx, cand[k], all[k]);
            if (rc) return rc;
            if (all[k].perSm < 1) continue;
            const int rounds = (P + cand[k] - 1) / cand[k];
            /* one CTA per SM leaves a patch's serial phases (swarm bookkeeping, visibility) uncovered — except where the
             * scene's tables are so large (more than 48 views: ~100 KB per 8-warp CTA) that two CTAs leave the tap stream
             * 60 KB of L1: there one 16-warp CTA (92 KB of L1) wins (config 5: 3.56 vs 3.47 k patches/s, r2_ab_runs.txt run 10) */
            const double solo = ctx->vcap > 48 ? 1.05 : 0.85;
            const double thr = (double)all[k].perSm * P / rounds * (all[k].perSm == 1 ? solo : 1.0);
            if (thr > bestThr + 1e-9 || (thr > bestThr - 1e-9 && bestT >= 0 && rounds < (P + cand[bestT] - 1) / cand[bestT])) { bestThr = thr; bestT = k; }
        }
        if (bestT < 0) return fail(ctx, PMVS_E_UNSUPPORTED, "refine kernel does not fit on an SM with this configuration");
        int pick = bestT;
        if ((long)n < 2L * ctx->smCount * all[bestT].perSm) {          /* cannot fill the GPU twice: finish sooner instead */
            long bestScore = 1L << 60;
            for (int k = 0; k < 6; ++k) {
                if (all[k].perSm < 1) continue;
                const long cap = (long)ctx->smCount * all[k].perSm;
                const long score = ((n + cap - 1) / cap) * ((P + cand[k] - 1) / cand[k]);
                if (score < bestScore || (score == bestScore && k == bestT)) { bestScore = score; pick = k; }
            }
        }
        cfg = all[pick];
        NW = cfg.NW;
    }
    const SmemPlan &pl = cfg.pl;
    RefineFn fn = cfg.fn;
    int perSm = cfg.perSm;
    if (perSm < 1) return fail(ctx, PMVS_E_UNSUPPORTED, "refine kernel does not fit on an SM with this configuration");
    int grid = ctx->smCount * perSm;
    if (grid > ctx->scratchCtas) grid = ctx->scratchCtas;
    if (grid > n) grid = n;
    {   /* leave everything the CTAs do not need to L1: the tap stream lives there */
        /* the carve-out comes in steps; ask for the smallest step that holds the CTAs (the driver rounds a percentage
         * UP to the next step, so the request is rounded down). Measured at config 2, 4 CTAs/SM: 164 KB (L1 92 KB)
         * 141.5 k patches/s, 196 KB 133.4 k, 228 KB 124.4 k. */
        static const int steps[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
        const size_t need = (size_t)perSm * (pl.total + 1024);
        int kb = 228;
        for (int k = 0; k < 9; ++k)
            if ((size_t)steps[k] * 1024 >= need) { kb = steps[k]; break; }
        int pct = (int)((100 * (size_t)kb * 1024) / 233472);
        if (pct > 100) pct = 100;
        if (getenv("PMVS_DEBUG")) fprintf(stderr, "refine_launch: n %d NW %d, %zu B of shared memory per CTA, %d CTAs/SM -> carve-out %d KB (%d %%)\n", n, NW, pl.total, perSm, kb, pct);
        if (const char *envC = getenv("PMVS_CARVEOUT")) pct = atoi(envC);      /* tuning: percent of 228 KB given to shared memory */
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    /* one index range per SM when the batch gives every SM a few patches per CTA slot; else one shared counter */
    int nChunks = ctx->smCount < PMVS_MAX_CHUNKS ? ctx->smCount : PMVS_MAX_CHUNKS;
    if (n < 4 * grid) nChunks = 1;
    if (const char *envL = getenv("PMVS_LOCAL")) { if (atoi(envL) == 0) nChunks = 1; }      /* tuning / A-B */
    /* a launch on another stream must not overlap the previous one of this context (shared counters / scratch) */
    if (!ctx->lastLaunch) CK(cudaEventCreateWithFlags(&ctx->lastLaunch, cudaEventDisableTiming));
    else CK(cudaStreamWaitEvent(st, ctx->lastLaunch, 0));
    CK(cudaMemsetAsync(ctx->dCounter, 0, sizeof(int) * PMVS_MAX_CHUNKS, st));
    fn<<<grid, NW * 32, pl.total, st>>>(ctx->scene, to_args(pl), n, d_in, d_out, flags, ctx->dCounter, nChunks);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->lastLaunch, st));
    return PMVS_OK;
}

int pmvs_refine_batch_device(pmvs_ctx *ctx, int n, const PmvsPatchIn *d_in, PmvsPatchOut *d_out, uint32_t flags, void *cudaStream) {
    if (!ctx || ctx->device < 0) return PMVS_E_ARG;
    if (n < 0 || (n > 0 && (!d_in || !d_out))) return fail(ctx, PMVS_E_ARG, "bad arguments to pmvs_refine_batch_device");
    if (n == 0) return PMVS_OK;
    CK(cudaSetDevice(ctx->device));
    return refine_launch(ctx, n, d_in, d_out, flags, cudaStream ? (cudaStream_t)cudaStream : ctx->stream);
}

#if PMVS_FOOTPRINT
/* experiment builds only: bounding boxes {x0, y0, x1, y1} per patch and view (8 views) of the first swarm's evaluated windows */
int pmvs_debug_footprints(pmvs_ctx *ctx, int n, int *out) {
    if (!ctx || !ctx->dDbgFoot || n > 65536) return PMVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, ctx->dDbgFoot, sizeof(int) * (size_t)n * 32, cudaMemcpyDeviceToHost));
    return PMVS_OK;
}
#endif

int pmvs_pack_records_device(pmvs_ctx *ctx, int n, const PmvsPatchOut *d_out, double *d_records, uint64_t *d_counters, void *cudaStream) {
    if (!ctx || ctx->device < 0) return PMVS_E_ARG;
    if (n < 0 || (n > 0 && (!d_out || !d_records))) return fail(ctx, PMVS_E_ARG, "bad arguments to pmvs_pack_records_device");
    if (n == 0) return PMVS_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cudaStream ? (cudaStream_t)cudaStream : ctx->stream;
    int grid = (n + 255) / 256;
    if (grid > ctx->smCount * 8) grid = ctx->smCount * 8;
    pack_records_kernel<<<grid, 256, 0, st>>>(n, d_out, d_records, (unsigned long long *)d_counters);
    ctx->launches++;
    CK(cudaGetLastError());
    return PMVS_OK;
}

int pmvs_refine_batch(pmvs_ctx *ctx, int n, const PmvsPatchIn *in, PmvsPatchOut *out, uint32_t flags) {
    if (!ctx || ctx->device < 0) return PMVS_E_ARG;
    if (n < 0 || (n > 0 && (!in || !out))) return fail(ctx, PMVS_E_ARG, "bad arguments to pmvs_refine_batch");
    if (n == 0) return PMVS_OK;
    for (int i = 0; i < n; ++i) {       /* host buffers can be validated up front (the device-resident call flags such records in-band) */
        if (in[i].nCam < 0 || in[i].nCam > PMVS_MAX_VIEWS) return fail(ctx, PMVS_E_ARG, "patch " + std::to_string(i) + ": nCam out of range");
        for (int k = 0; k < in[i].nCam; ++k)
            if ((int)in[i].camIdx[k] >= ctx->nCams)
                return fail(ctx, PMVS_E_ARG, "patch " + std::to_string(i) + ": camera index " + std::to_string(in[i].camIdx[k]) + " is not a camera of this scene");
    }
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_buffers(ctx, sizeof(PmvsPatchIn) * (size_t)n, sizeof(PmvsPatchOut) * (size_t)n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->dIn, in, sizeof(PmvsPatchIn) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = refine_launch(ctx, n, (const PmvsPatchIn *)ctx->dIn, (PmvsPatchOut *)ctx->dOut, flags, ctx->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->dOut, sizeof(PmvsPatchOut) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PMVS_OK;
}

/* camera.cpp:63-64 + cv::resize size rounding: host-only helper, no device needed */
int pmvs_pyramid_levels(int cols, int rows, double lodRatio, int cfgMaxLOD, int *maxLOD, int32_t *levelCols, int32_t *levelRows) {
    if (cols <= 0 || rows <= 0 || !(lodRatio > 0.0 && lodRatio < 1.0) || !maxLOD) return PMVS_E_ARG;
    int m = (int)(log((double)(cols > rows ? cols : rows)) / log(1.0 / lodRatio));
    if (m > cfgMaxLOD) m = cfgMaxLOD;
    if (m >= PMVS_MAX_LEVELS) m = PMVS_MAX_LEVELS - 1;
    if (m < 0) m = 0;
    *maxLOD = m;
    for (int l = 0; l <= m; ++l) {
        int c, r;
        level_dims(cols, rows, lodRatio, l, c, r);
        if (levelCols) levelCols[l] = c;
        if (levelRows) levelRows[l] = r;
    }
    return PMVS_OK;
}

/* The pyramid of one camera built on `device` and copied back to caller-allocated host levels (levels[0].grey may be
 * NULL: level 0 is the input). Used by tests and by hosts that want the arrays; pmvs_create builds missing levels
 * itself without a host round trip. */
int pmvs_build_pyramid(int device, const uint8_t *grey0, int cols, int rows, int64_t pitch, double lodRatio, int maxLOD, int withEdge,
                       PmvsLevelOut *levels) {
    if (!grey0 || cols <= 0 || rows <= 0 || pitch < cols || maxLOD < 0 || maxLOD >= PMVS_MAX_LEVELS || !levels) return PMVS_E_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return PMVS_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PMVS_E_CUDA;
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return PMVS_E_CUDA;
    int64_t launches = 0;
    DevU8 g0;
    unsigned long long *dMinMax = nullptr;
    double *dEdge = nullptr;
    cudaError_t e = upload_u8(grey0, pitch, cols, rows, g0, st);
    if (e == cudaSuccess && withEdge) e = cudaMalloc(&dMinMax, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess && withEdge) e = cudaMalloc(&dEdge, (size_t)cols * rows * sizeof(double));
    int rc = PMVS_OK;
    for (int l = 0; l <= maxLOD && e == cudaSuccess; ++l) {
        int lc, lr;
        level_dims(cols, rows, lodRatio, l, lc, lr);
        PmvsLevelOut &L = levels[l];
        if (L.cols != lc || L.rows != lr || (l > 0 && (!L.grey || L.pitch < lc)) || (withEdge && !L.edge)) { rc = PMVS_E_ARG; break; }
        DevU8 gl = g0;
        if (l > 0) e = resize_level(g0, pow(lodRatio, l), lc, lr, gl, st, &launches);
        if (e == cudaSuccess && l > 0) e = cudaMemcpy2DAsync(L.grey, (size_t)L.pitch, gl.p, gl.pitch, (size_t)lc, (size_t)lr, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && withEdge) {
            e = edge_level(gl, dEdge, dMinMax, st, &launches);
            if (e == cudaSuccess) e = cudaMemcpyAsync(L.edge, dEdge, (size_t)lc * lr * sizeof(double), cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (l > 0) cudaFree(gl.p);
    }
    cudaFree(g0.p);
    cudaFree(dMinMax);
    cudaFree(dEdge);
    cudaStreamDestroy(st);
    if (rc != PMVS_OK) return rc;
    return e == cudaSuccess ? PMVS_OK : (e == cudaErrorMemoryAllocation ? PMVS_E_NOMEM : PMVS_E_CUDA);
}

int pmvs_neighbor_counts(int device, int n, const double *centers, double radius, int first, int count, int *counts) {
    if (n < 0 || first < 0 || count < 0 || first + (long long)count > n || (n > 0 && !centers) || (count > 0 && !counts)) return PMVS_E_ARG;
    if (count == 0) return PMVS_OK;
    int devs = 0;
    if (cudaGetDeviceCount(&devs) != cudaSuccess || device < 0 || device >= devs) return PMVS_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PMVS_E_CUDA;
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return PMVS_E_CUDA;
    double *dC = nullptr;
    int *dN = nullptr;
    cudaError_t e = cudaMalloc(&dC, sizeof(double) * 3 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&dN, sizeof(int) * (size_t)count);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dC, centers, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        neighbor_count_kernel<<<(count + PMVS_NB_TILE - 1) / PMVS_NB_TILE, PMVS_NB_TILE, 0, st>>>(n, dC, radius, first, count, dN);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, dN, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dC);
    cudaFree(dN);
    cudaStreamDestroy(st);
    return e == cudaSuccess ? PMVS_OK : (e == cudaErrorMemoryAllocation ? PMVS_E_NOMEM : PMVS_E_CUDA);
}

/* test support: the swarm alone on analytic functions. L,U,init: n*3; keys,maxIter,P,fn,hasInit: n. Outputs:
 * gbest n*3, gbestFitness n, iterations n, particles n*64*8 (may be NULL). */
int pmvs_pso_test(pmvs_ctx *ctx, int n, const double *L, const double *U, const double *init, const int *hasInit,
                  const int *maxIter, const int *P, const int *fn, const uint64_t *keys, double *gbest, double *gbestFitness,
                  int *iterations, double *particles) {
    if (!ctx || ctx->device < 0 || n <= 0) return PMVS_E_ARG;
    CK(cudaSetDevice(ctx->device));
    std::vector<PsoTestProblem> pr(n);
    for (int i = 0; i < n; ++i) {
        if (P[i] < 1 || P[i] > PMVS_PSO_TEST_MAX) return fail(ctx, PMVS_E_UNSUPPORTED, "P out of range");
        for (int d = 0; d < 3; ++d) {
            pr[i].L[d] = L[3 * i + d];
            pr[i].U[d] = U[3 * i + d];
            pr[i].init[d] = init ? init[3 * i + d] : 0.0;
        }
        pr[i].key = keys[i];
        pr[i].hasInit = hasInit ? hasInit[i] : 0;
        pr[i].maxIter = maxIter[i];
        pr[i].P = P[i];
        pr[i].fn = fn[i];
    }
    PsoTestProblem *dp = nullptr;
    PsoTestResult *dr = nullptr;
    CK(cudaMalloc(&dp, sizeof(PsoTestProblem) * n));
    cudaError_t e = cudaMalloc(&dr, sizeof(PsoTestResult) * n);
    if (e != cudaSuccess) { cudaFree(dp); return fail(ctx, PMVS_E_NOMEM, "cudaMalloc failed"); }
    cudaMemsetAsync(dr, 0, sizeof(PsoTestResult) * n, ctx->stream);      /* particle slots beyond P are copied back too */
    std::vector<PsoTestResult> hr(n);
    e = cudaMemcpy(dp, pr.data(), sizeof(PsoTestProblem) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        pso_test_kernel<<<n < 1024 ? n : 1024, 512, 0, ctx->stream>>>(n, dp, dr);
        ctx->launches++;
        e = cudaStreamSynchronize(ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpy(hr.data(), dr, sizeof(PsoTestResult) * n, cudaMemcpyDeviceToHost);
    cudaFree(dp);
    cudaFree(dr);
    if (e != cudaSuccess) return fail(ctx, PMVS_E_CUDA, std::string("pso_test: ") + cudaGetErrorString(e));
    for (int i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) gbest[3 * i + d] = hr[i].gbest[d];
        gbestFitness[i] = hr[i].gbestFitness;
        iterations[i] = hr[i].iterations;
        if (particles) memcpy(particles + (size_t)i * PMVS_PSO_TEST_MAX * 8, hr[i].particles, sizeof(hr[i].particles));
    }
    return PMVS_OK;
}

}   // extern "C"
