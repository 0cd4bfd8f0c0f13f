/*
 * pmvs_device.cuh — device-side data layout and the warp-/CTA-cooperative building blocks of the patch
 * refinement path (sm_100a). Compiled with -fmad=false: every a*b+c below is two roundings, like the reference
 * (MSVC /fp:precise) — fused multiply-adds appear only where fma() is written out (the sample loop).
 *
 * Reference functions restated here (paths relative to the reference tree, TMVS/):
 *   mvs/camera.cpp:138-160  Camera::project            -> project_pt
 *   mvs/patch.cpp:290-330   Patch::getHomographies     -> warp_homographies
 *   mvs/patch.cpp:914-1047  PAIS::getFitness           -> warp_fitness
 *   pso/psosolver.cpp       PsoSolver (whole file)     -> pso_run / pso_scans / pso_apply_moves
 *
 * HBM layout. Every pyramid level of every camera is stored as a "quad" image: one 32-bit word per pixel (x,y)
 * holding the four bilinear taps g(y,x) | g(y,x+1)<<8 | g(y+1,x)<<16 | g(y+1,x+1)<<24 (edge-replicated). The
 * reference issues four byte loads per sample and view (patch.cpp:1014-1017); here that is ONE aligned 32-bit
 * load, coalesced across the warp because lanes walk the window along x. It costs 4 B/pixel of the 180 GB HBM
 * (config 5: 64 views x 12 MP x 2.8 levels x 4 B = 8.6 GB) and needs no per-patch staging pass.
 */
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/pmvs_b200.h"
#include "pmvs_rng.h"

#define PMVS_MAX_PARTICLES 128        /* swarm size: seeds run 2 * particleNum particles (patch.cpp:192) */
#define PMVS_MAX_PARTICLE_NUM 64     /* MvsConfig.particleNum accepted by pmvs_create */
#define PMVS_MAX_RADIUS 31
#define PMVS_MAX_PS (2 * PMVS_MAX_RADIUS + 1)
#define PMVS_FULL 0xffffffffu
/* build switches of the evaluation path (shared memory is taken from L1, which the tap stream lives in: keep it small) */
#ifndef PMVS_EVAL_BATCH
#define PMVS_EVAL_BATCH 4       /* hypotheses whose scalar part (homographies ...) one warp computes together; 1 = off */
#endif
#ifndef PMVS_DIST_GLOBAL
#define PMVS_DIST_GLOBAL 1      /* distance weights read through L1 from one global table instead of a copy per CTA */
#endif
/* non-reference views with per-lane slots in the column loop (V = 2..16) */
#ifndef PMVS_SLOT_VIEWS
#define PMVS_SLOT_VIEWS 10       /* up to this many non-reference views keep per-lane slots; more form the x-part inline (L1 over slots) */
#endif
/* scenes with more than 16 cameras keep no slots at all (their shared memory goes to occupancy: the tables that scale
 * with the camera count are large already); every view count forms the x-parts inline there */
#define PMVS_COLV_VIEWS(vcap) ((vcap) > 16 ? 0 : ((vcap) < 2 ? 1 : ((vcap) - 1 > PMVS_SLOT_VIEWS ? PMVS_SLOT_VIEWS : (vcap) - 1)))
#define PMVS_CORR_GLOBAL(vcap) ((vcap) > 16)   /* the V x V correlation table (+ V region ratios) in the CTA's global scratch */
/* scenes the view-lane loop covers (<= 9 cameras) keep no slots either: only its rare fallbacks reach the old loop */
#define PMVS_VL_VIEWS 9                 /* views the register form of the view-lane loop covers (fitness_vl) */
#define PMVS_SLOT_VIEWS_OF(vcap, useVL) ((useVL) ? 0 : PMVS_COLV_VIEWS(vcap))
/* per-warp colour stash of the many-view / gradient form (fitness_vl_many): 4 doubles per lane and chunk of 4 views */
#ifndef PMVS_VL_MANY_STASH
#define PMVS_VL_MANY_STASH 0            /* 1: the stash form (fitness_vl_many); 0: the register form (fitness_vl_rows) — no stash */
#endif
#define PMVS_STASH_DOUBLES(vcap, useVL, grad) ((PMVS_VL_MANY_STASH && (useVL) && ((vcap) > PMVS_VL_VIEWS || (grad))) ? (((vcap) - 1 + 3) / 4) * 128 : 0)
#define PMVS_COLV_SLOTS_N(slotViews) (3 * (slotViews) * 32)
/* useVL: no view table at all — the rare hypotheses the view-lane loop does not take go to the lane-strided loops */
#define PMVS_GV_DOUBLES_N(vcap, slotViews, useVL) ((useVL) ? 0 : (((slotViews) > 0 && (vcap) - 1 <= (slotViews)) ? 6 * ((vcap) < 2 ? 1 : (vcap) - 1) : 12 * ((vcap) < 2 ? 1 : (vcap) - 1)))
#define PMVS_COLV_DOUBLES_N(vcap, ps, slotViews, useVL) (PMVS_COLV_SLOTS_N(slotViews) + PMVS_GV_DOUBLES_N(vcap, slotViews, useVL) + ((useVL) ? 0 : 2 * PMVS_PS_PAD(ps)))
#define PMVS_COLV_SLOTS(vcap) (3 * PMVS_COLV_VIEWS(vcap) * 32)
/* compact table of the non-reference views. Slot mode (<= PMVS_SLOT_VIEWS of them): 6 doubles each — h1, h4, h7, quad
 * pointer, cols, spare; inline mode: 12 — h0..h8, quad pointer, cols, spare */
#define PMVS_GV_DOUBLES_TOTAL(vcap) ((vcap) - 1 <= PMVS_SLOT_VIEWS ? 6 * ((vcap) < 2 ? 1 : (vcap) - 1) : 12 * ((vcap) - 1))
#define PMVS_PS_PAD(ps) (((ps) + 1) & ~1)
/* per-warp area of the column loop: per-lane slots | compact non-reference view table | per-row {fy} | per-row {py*cols, sel} */
#define PMVS_COLV_DOUBLES(vcap, ps) (PMVS_COLV_SLOTS(vcap) + PMVS_GV_DOUBLES_TOTAL(vcap) + 2 * PMVS_PS_PAD(ps))
/* homographies kept per warp: one hypothesis of vcap views, or a batch of up to 4 hypotheses with batch * V <= 32 */
#define PMVS_HCAP(vcap) ((PMVS_EVAL_BATCH < 2 || (vcap) > 32) ? (vcap) : (PMVS_EVAL_BATCH * (vcap) < 32 ? PMVS_EVAL_BATCH * (vcap) : 32))
#define PMVS_HYP_DOUBLES 12

struct DevLevel {
    const uint32_t *quad;
    const double *edge;
    int cols, rows;
};
struct DevCamera {
    double center[3], R[9], t[3], KR[9], KT[3], optN[3], focal[2], pp[2];
    int maxLOD, _pad;
    DevLevel level[PMVS_MAX_LEVELS];
};
struct DevScene {
    PmvsConfig cfg;
    const DevCamera *cams;
    const double *distW;       /* patchSize^2, index x*patchSize+y (mvs.cpp:104-109) */
    const double *distG;       /* patchSize: distW[x][y] = distG[x] * distG[y] up to rounding (the Gaussian is separable) */
    int *dbgFoot;              /* experiment builds (PMVS_FOOTPRINT): per patch 8 views x 4 ints, else nullptr */
    double *scratch;           /* per-CTA correlation windows: scratchStride doubles per CTA */
    unsigned long long scratchStride;
    int nCams, useVL;          /* useVL: the view-lane window loop (fitness_vl) is enabled */
    uint64_t seed;
    double lodScale[PMVS_MAX_LEVELS];
};

/* ---- per-patch evaluation context (shared memory, read-only while the swarm runs) ---------------------- */
struct ViewS {
    double KR[9], KT[3];
    const uint32_t *quad;
    int cols, rows, isRef, _pad;
};
struct EvalCtx {
    double ray[3], refC[3], refOptN[3], refR[9], refT[3], refKR[9], refKT[3], refFocal[2], refPP[2], sc;
    const uint32_t *refQuad;
    const double *refEdge;
    ViewS *view;
    int refCols, refRows, refCam, LOD, V, valid;
    int refView, _pad;   /* index of the reference camera in view[] (-1: not among the views) */
};
/*
 * Per-patch constants of the reference view (shared memory; built once per swarm run by build_ref_win).
 * center = ray * depth + C_ref (patch.cpp:944) lies on the reference camera's viewing ray for every depth, so its
 * reference-image position pt (patch.cpp:951), the window axes (:979-980), the reference view's samples (H = I,
 * :317-319), the background-mask pixels (:986) and the distance weights (:1030-1032) are the same for every
 * hypothesis of a swarm run up to the rounding of pt (~1e-13 pixel; hypotheses further than PMVS_PT_TOL from the
 * canonical pt take the per-hypothesis path).
 */
#define PMVS_PT_TOL 1e-9
#ifndef PMVS_TILE
#define PMVS_TILE 0             /* T > 0: experiment build that stages a T x T quad tile per non-reference view in shared memory with
                                   cp.async.bulk + mbarrier (the TMA unit) once per swarm and reads the taps of every evaluation that stays
                                   inside the tiles from it (profiles/r2_tile_staging.md) */
#endif
#ifndef PMVS_FOOTPRINT
#define PMVS_FOOTPRINT 0        /* 1: experiment build that records the image footprint of every swarm (tools/footprints.py) */
#endif
#define PMVS_PS_PAD8(ps) (((ps) + 7) & ~7)          /* window rows incl. the padding rows of the view-lane loop's trips */
#define PMVS_NYP(ps) (((PMVS_PS_PAD8(ps) + 11) & ~15) + 4)     /* row pitch of the colour table: = 4 (mod 16) doubles: conflict-free quads */
struct RefWin {
    double pt[2];
    double *xs;                  /* nx sample columns (canonical window axes) */
    double2 *ysg;                /* PAD8(ny) x {y, row factor of the separable distance weight}; padding rows repeat the last row */
    double *gx;                  /* nx column factors of the distance weight */
    unsigned long long *mask;    /* nx: bit (64/gl) (j % gl) + j / gl = reference pixel of (column, row j) is not background */
    double *refc;                /* nx x nyp: the reference view's sample of every window position */
    double *pw;                  /* gradient weighting on: nx x nyp per-pixel weights = distance weight x gradient weight
                                    (patch.cpp:1030-1032, :1036-1038) x background mask; nullptr otherwise */
    int nx, ny, nyp, ok, gl, _pad;
#if PMVS_TILE
    uint32_t *tile;              /* experiment build: 4 quad tiles of PMVS_TILE x PMVS_TILE words, one per non-reference view */
    int tox[4], toy[4];          /* tile origins in the views' level images */
    unsigned long long mbar;     /* mbarrier of the bulk copies */
    int tileOk, tilePhase;
#endif
#if PMVS_FOOTPRINT
    int foot[8][4];              /* experiment build: per view, bounding box {x0, y0, x1, y1} of every window the swarm evaluated */
    int footc[8][2];             /* per view: centre of the first window evaluated (the tile a per-patch staging would be centred on) */
    int footn[8];                /* evaluations: all, inside a 48 / 64 / 96 / 128 pixel tile around footc in EVERY view; [5] = centre set */
#endif
};
/* doubles one RefWin's tables take for patch size ps: xs | gx | ysg | mask | refc | pw */
#define PMVS_REFWIN_DOUBLES(ps, grad) (5 * (size_t)PMVS_PS_PAD8(ps) + (size_t)(ps) * PMVS_NYP(ps) * ((grad) ? 2 : 1) + (size_t)(4 * PMVS_TILE * PMVS_TILE / 2))
__device__ __forceinline__ void carve_ref_win(RefWin &R, double *base, int ps, bool grad) {
    const int pp = PMVS_PS_PAD8(ps);
    R.xs = base;
    R.gx = base + pp;
    R.ysg = (double2 *)(base + 2 * pp);
    R.mask = (unsigned long long *)(base + 4 * pp);
    R.refc = base + 5 * pp;
    R.pw = grad ? R.refc + (size_t)ps * PMVS_NYP(ps) : nullptr;
#if PMVS_TILE
    R.tile = (uint32_t *)(R.refc + (size_t)ps * PMVS_NYP(ps) * (grad ? 2 : 1));     /* 16-byte aligned: every table before it is a multiple of 16 bytes */
    R.tileOk = 0;
    R.tilePhase = -1;
#endif
    R.nyp = PMVS_NYP(ps);
    R.nx = R.ny = 0;
    R.ok = 0;
    R.gl = 4;
}

/* per-warp scratch (shared memory) */
struct WarpWork {
    double *H;      /* V*9 */
    double *xs;     /* patchSize */
    double *ys;     /* patchSize */
    double *hyp;    /* PMVS_EVAL_BATCH HypoS records: per-hypothesis results of the batched scalar part */
    double *colv;   /* PMVS_COLV_SLOTS(vcap): per-lane column constants of the unchecked loop */
    int slotViews, _padw;   /* non-reference views the slot area holds */
    double *gv;     /* PMVS_GV_DOUBLES_TOTAL(vcap): compact table of the non-reference views */
    double *rowf;   /* PMVS_PS_PAD(ps): fractional part of each window row in the reference view */
    int2 *rowi;     /* PMVS_PS_PAD(ps): {floor(y) * refCols, 2 * (cvRound(y) - floor(y))} per window row */
    const RefWin *rw;   /* the patch's reference window (nullptr: none built) */
    double *stash;      /* PMVS_STASH_DOUBLES: per-lane colours of the many-view loop (nullptr: none) */
};

__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* Utility::spherical2Normal, mvs/utility.h:25-29 */
__device__ __forceinline__ void spherical2Normal(double theta, double phi, double *n) {
    double st, ct, sp, cp;
    sincos(theta, &st, &ct);
    sincos(phi, &sp, &cp);
    n[0] = st * cp;
    n[1] = st * sp;
    n[2] = ct;
}

/* Camera::project without distortion (camera.cpp:138-160); the in-image test is the caller's. */
__device__ __forceinline__ void project_pt(const double *R, const double *t, const double *focal, const double *pp, double sc,
                                           const double *X, double *out) {
    double x2 = (R[0] * X[0] + R[1] * X[1] + R[2] * X[2]) + t[0];
    double y2 = (R[3] * X[0] + R[4] * X[1] + R[5] * X[2]) + t[1];
    double z2 = (R[6] * X[0] + R[7] * X[1] + R[8] * X[2]) + t[2];
    out[0] = focal[0] * (x2 / z2);
    out[1] = focal[1] * (y2 / z2);
    out[0] += pp[0];
    out[1] += pp[1];
    out[0] *= sc;
    out[1] *= sc;
}
/* Camera::inImage(Vec2d), camera.h:116-131 */
__device__ __forceinline__ bool in_image(double x, double y, int cols, int rows) {
    if (isnan(x) || isnan(y)) return false;
    return !(x < 0 || x >= cols || y < 0 || y >= rows);
}

/* OpenCV 2.4 cv::invert, 3x3 CV_64F closed form (used by Mat_::inv() at patch.cpp:314) */
__device__ __forceinline__ void inv3(const double *S, double *D) {
#define Sd(r, c) S[(r)*3 + (c)]
    double d = Sd(0, 0) * (Sd(1, 1) * Sd(2, 2) - Sd(1, 2) * Sd(2, 1)) - Sd(0, 1) * (Sd(1, 0) * Sd(2, 2) - Sd(1, 2) * Sd(2, 0)) +
               Sd(0, 2) * (Sd(1, 0) * Sd(2, 1) - Sd(1, 1) * Sd(2, 0));
    if (d != 0.) {
        d = 1. / d;
        D[0] = (Sd(1, 1) * Sd(2, 2) - Sd(1, 2) * Sd(2, 1)) * d;
        D[1] = (Sd(0, 2) * Sd(2, 1) - Sd(0, 1) * Sd(2, 2)) * d;
        D[2] = (Sd(0, 1) * Sd(1, 2) - Sd(0, 2) * Sd(1, 1)) * d;
        D[3] = (Sd(1, 2) * Sd(2, 0) - Sd(1, 0) * Sd(2, 2)) * d;
        D[4] = (Sd(0, 0) * Sd(2, 2) - Sd(0, 2) * Sd(2, 0)) * d;
        D[5] = (Sd(0, 2) * Sd(1, 0) - Sd(0, 0) * Sd(1, 2)) * d;
        D[6] = (Sd(1, 0) * Sd(2, 1) - Sd(1, 1) * Sd(2, 0)) * d;
        D[7] = (Sd(0, 1) * Sd(2, 0) - Sd(0, 0) * Sd(2, 1)) * d;
        D[8] = (Sd(0, 0) * Sd(1, 1) - Sd(0, 1) * Sd(1, 0)) * d;
    } else {
        for (int i = 0; i < 9; ++i) D[i] = 0;
    }
#undef Sd
}

/* d*L*KR - L*KT*n^T, L = diag(s,s,1): the bracket of patch.cpp:314 / :328 */
__device__ __forceinline__ void plane_matrix(const double *KR, const double *KT, const double *n, double d, double sc, double *M) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double L = (r < 2) ? sc : 1.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) M[r * 3 + c] = d * (L * KR[r * 3 + c]) - (L * KT[r]) * n[c];
    }
}

/* Patch::getHomographies (patch.cpp:290-330): lane v builds H_v (ref -> view v) in f64 and stores it in Hw[9v..]. */
__device__ __forceinline__ void warp_homographies(const EvalCtx &E, const double *center, const double *normal, double *Hw) {
    const int lane = threadIdx.x & 31;
    if (lane < E.V || E.V > 32) {
        const double d = -dot3(center, normal);
        double Mref[9], inv[9];
        plane_matrix(E.refKR, E.refKT, normal, d, E.sc, Mref);
        inv3(Mref, inv);
        for (int v = lane; v < E.V; v += 32) {
            const ViewS &vw = E.view[v];
            double *Hi = Hw + 9 * v;
            if (vw.isRef) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Hi[k] = (k % 4 == 0) ? 1.0 : 0.0;
                continue;
            }
            double M[9];
            plane_matrix(vw.KR, vw.KT, normal, d, E.sc, M);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double acc = M[r * 3] * inv[c];
                    acc += M[r * 3 + 1] * inv[3 + c];
                    acc += M[r * 3 + 2] * inv[6 + c];
                    Hi[r * 3 + c] = acc;
                }
        }
    }
    __syncwarp();
}

/* The reference walks the window with `for (double x = pt-r; x <= pt+r; ++x)` (patch.cpp:979-980, :347-348):
 * repeated +1.0 in f64. Lane 0 / lane 1 replay that recurrence so sample positions (and, in the rare case a
 * rounding step overshoots pt+r, the sample COUNT) are bit-identical. Returns nx | ny<<16 to every lane. */
__device__ __forceinline__ int warp_window_axes(const double *pt, int radius, int ps, double *xs, double *ys) {
    const int lane = threadIdx.x & 31;
    /* Fast path: while x0 = pt-r and x0+2r share a binade (and x0 >= 1), every x0+k is exactly representable, so the
     * recurrence never rounds and x_k == x0 + k bit for bit; the count is #{k < ps : x0+k <= pt+r}. */
    const double x0 = pt[0] - radius, y0 = pt[1] - radius, xh = pt[0] + radius, yh = pt[1] + radius;
    const bool exact = x0 >= 1.0 && y0 >= 1.0 && (__double2hiint(x0) >> 20) == (__double2hiint(x0 + 2.0 * radius) >> 20) &&
                       (__double2hiint(y0) >> 20) == (__double2hiint(y0 + 2.0 * radius) >> 20);
    if (exact) {
        int nx = 0, ny = 0;
        for (int k = lane; k < ps; k += 32) {
            const double x = x0 + (double)k, y = y0 + (double)k;
            if (x <= xh) { xs[k] = x; ++nx; }
            if (y <= yh) { ys[k] = y; ++ny; }
        }
        nx = __reduce_add_sync(PMVS_FULL, nx);
        ny = __reduce_add_sync(PMVS_FULL, ny);
        __syncwarp();
        return nx | (ny << 16);
    }
    int n = 0;
    if (lane < 2) {
        double *dst = lane ? ys : xs;
        const double hi = pt[lane] + radius;
        for (double x = pt[lane] - radius; x <= hi && n < ps; x = x + 1.0) dst[n++] = x;
    }
    const int nx = __shfl_sync(PMVS_FULL, n, 0), ny = __shfl_sync(PMVS_FULL, n, 1);
    __syncwarp();
    return nx | (ny << 16);
}

#define PMVS_MAGIC_FLOOR 6755399441055744.0 /* 1.5 * 2^52: low word of (x + magic, rounded down) = floor(x) */
__device__ __forceinline__ double u2d(uint32_t g) { return __hiloint2double(0x43300000, (int)g) - 4503599627370496.0; }
__device__ __forceinline__ double s2d(int d) {
    return __hiloint2double(0x43300000, (int)((uint32_t)d ^ 0x80000000u)) - (4503599627370496.0 + 2147483648.0);
}

/* bilinear sample of one view (patch.cpp:1005-1017) from the quad layout; ix,iy known in-bounds and >= 0 */
__device__ __forceinline__ double quad_bilinear(const uint32_t *__restrict__ quad, int cols, double ix, double iy) {
    const double tx = __dadd_rd(ix, PMVS_MAGIC_FLOOR), ty = __dadd_rd(iy, PMVS_MAGIC_FLOOR);
    const int px = __double2loint(tx), py = __double2loint(ty);
    const double fx = ix - (tx - PMVS_MAGIC_FLOOR), fy = iy - (ty - PMVS_MAGIC_FLOOR);
    const uint32_t q = __ldg(quad + (size_t)py * cols + px);
    const int g00 = q & 0xff, g01 = (q >> 8) & 0xff, g10 = (q >> 16) & 0xff, g11 = q >> 24;
    const double top = fma(s2d(g01 - g00), fx, u2d(g00));
    const double bot = fma(s2d(g11 - g10), fx, u2d(g10));
    return fma(bot - top, fy, top);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PMVS_FULL, v, o);
    return v;
}

/* One sample of one view: homography + bounds (CHECK) + bilinear. Returns false when out of bounds. */
template <bool CHECK>
__device__ __forceinline__ bool sample_view(const double *__restrict__ H, const ViewS &vw, double x, double y, double &c) {
    const double w = fma(H[6], x, fma(H[7], y, H[8]));
    const double nx = fma(H[0], x, fma(H[1], y, H[2]));
    const double ny = fma(H[3], x, fma(H[4], y, H[5]));
    double ix, iy;
    if (CHECK) {
        ix = nx / w;
        iy = ny / w;
        /* patch.cpp:999; NaN coordinates count as out of bounds (the reference would index with (int)NaN) */
        if (!(ix >= 2.0 && ix < (double)(vw.cols - 3) && iy >= 2.0 && iy < (double)(vw.rows - 3)) || w == 0.0) return false;
    } else {
        const double rw = 1.0 / w;
        ix = nx * rw;
        iy = ny * rw;
    }
    c = quad_bilinear(vw.quad, vw.cols, ix, iy);
    return true;
}

/*
 * The sample loop of getFitness (patch.cpp:979-1042). Lanes stride the window along x; each lane accumulates its
 * samples in index order, then a fixed xor-tree combines lanes (deterministic). VCAP > 0: the per-view colours of
 * one sample live in registers (V <= VCAP); VCAP == 0: two passes per sample (V up to 64, no storage).
 * Returns false when an unmasked sample left a view (only possible with CHECK).
 */
template <int VCAP, bool CHECK>
__device__ __noinline__ bool fitness_samples(const DevScene &S, const EvalCtx &E, const double *__restrict__ sDistW,
                                                const double *__restrict__ Hw, const double *__restrict__ xs,
                                                const double *__restrict__ ys, int nx, int ny, double &fitOut, double &swOut) {
    const int lane = threadIdx.x & 31;
    const int V = E.V;
    const double invV = 1.0 / (double)V;
    const bool useDist = S.cfg.adaptiveDistanceEnable, useDiff = S.cfg.adaptiveDifferenceEnable, useGrad = S.cfg.adaptiveGradientEnable;
    const double invDiffW = 1.0 / S.cfg.diffWeighting, gradW = S.cfg.gradientWeighting;
    double fit = 0, sw = 0;
    bool oob = false;
    const int total = nx * ny;
    int i = lane % nx, j = lane / nx;
    for (int s = lane; s < total; s += 32) {
        const double x = xs[i], y = ys[j];
        const int idx = i * ny + j;          /* position of the reference's distance-table iterator (patch.cpp:975,980) */
        i += 32;
        while (i >= nx) { i -= nx; ++j; }
        const int rx = __double2int_rn(x), ry = __double2int_rn(y);                    /* cvRound */
        const size_t rofs = (size_t)ry * E.refCols + rx;
        if ((__ldg(E.refQuad + rofs) & 0xff) == 0) continue;                          /* patch.cpp:986 */
        double mean = 0, sad = 0;
        if (VCAP > 0) {
            double c[VCAP > 0 ? VCAP : 1];
#pragma unroll
            for (int v = 0; v < VCAP; ++v) {
                if (v < V) {
                    if (!sample_view<CHECK>(Hw + 9 * v, E.view[v], x, y, c[v])) oob = true;
                    if (CHECK && oob) break;
                    mean += c[v];
                }
            }
            if (CHECK && oob) break;
            mean *= invV;
#pragma unroll
            for (int v = 0; v < VCAP; ++v)
                if (v < V) sad += fabs(c[v] - mean);
        } else {
            for (int v = 0; v < V; ++v) {
                double c;
                if (!sample_view<CHECK>(Hw + 9 * v, E.view[v], x, y, c)) oob = true;
                if (CHECK && oob) break;
                mean += c;
            }
            if (CHECK && oob) break;
            mean *= invV;
            for (int v = 0; v < V; ++v) {
                double c = 0;
                sample_view<CHECK>(Hw + 9 * v, E.view[v], x, y, c);   /* same arithmetic as pass one */
                sad += fabs(c - mean);
            }
        }
        const double avgSad = sad * invV;
        double weight = 1.0;
        if (useDist) weight *= PMVS_DIST_GLOBAL ? __ldg(S.distW + idx) : sDistW[idx];                                           /* patch.cpp:1030-1032 */
        if (useDiff) weight *= exp(-avgSad * avgSad * invDiffW);                      /* patch.cpp:1033-1035 */
        if (useGrad) weight *= exp(-1.0 / (__ldg(E.refEdge + rofs) * gradW));         /* patch.cpp:1036-1038 */
        sw += weight;
        fit = fma(weight, avgSad, fit);
    }
    if (CHECK && __any_sync(PMVS_FULL, oob)) return false;
    fitOut = warp_sum(fit);
    swOut = warp_sum(sw);
    return true;
}

/* exp(x) for the weights of the unchecked loop (exp(-avgSad^2/diffWeighting), exp(-1/(edge*gradientWeighting)),
 * patch.cpp:1034,1037; x <= 0, valid up to x < 709): Cody-Waite reduction
 * x = k ln2 + r, |r| <= ln2/2, degree-11 polynomial (Chebyshev-node interpolant of exp on that interval, relative error
 * 1.6e-17 before rounding, coefficients from tools/exp_poly.py), scaling by exponent arithmetic. The coefficients sit in
 * the constant bank so every fma takes its constant as an operand: no register or uniform-register traffic. Results
 * below 2^-1020 are flushed to zero (such weights cannot matter). */
__constant__ double kExpPoly[12] = {0x1.0000000000000p+0,  0x1.0000000000000p+0,  0x1.0000000000011p-1,  0x1.555555555555ap-3,
                                    0x1.555555554f0bap-5,  0x1.111111110f21ep-7,  0x1.6c16c1880029fp-10, 0x1.a01a01b1461c5p-13,
                                    0x1.a01991a10d9aep-16, 0x1.71ddf56d8deb5p-19, 0x1.28b4101c77212p-22, 0x1.af632a0f7e2cep-26};
__device__ __forceinline__ double exp_nonpos(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);      /* low word = round-to-nearest integer k */
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = kExpPoly[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, r, kExpPoly[i]);
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return (x >= -707.0) ? res : ((x != x) ? x : 0.0);
}

/* 1/w for the unchecked path: MUFU.RCP64H seed r0 (~2^-20), then r0*(1+e+e^2) with e = 1-w*r0 (residual e^3); w is
 * finite, normal and > 0 there. */
__device__ __forceinline__ double rcp_fast(double w) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
    double e = fma(-w, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}

/* explicit shared-memory loads: these helpers run inside non-inlined functions where the compiler only sees
 * generic pointers; `a` is a 32-bit shared-window address from smem_addr() */
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned long long lds_u64(unsigned a) {
    unsigned long long v;
    asm("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(unsigned a) {
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

/* bilinear taps of the unchecked path (patch.cpp:1005-1017) in difference form,
 *   c = g00 + fx*(g01-g00) + fy*((g10-g00) + fx*(g11-g10-g01+g00)),
 * tap differences formed in integers, converted by the 2^52+2^31 bias trick (lo word = value + 2^31). */
#define PMVS_BIAS_S32 (4503599627370496.0 + 2147483648.0)
__device__ __forceinline__ double quad_bilinear_fast(const uint32_t *__restrict__ quad, int cols, double ix, double iy) {
    const double tx = __dadd_rd(ix, PMVS_MAGIC_FLOOR), ty = __dadd_rd(iy, PMVS_MAGIC_FLOOR);
    const int px = __double2loint(tx), py = __double2loint(ty);
    const double fx = ix - (tx - PMVS_MAGIC_FLOOR), fy = iy - (ty - PMVS_MAGIC_FLOOR);
    const uint32_t q = __ldg(quad + (py * cols + px));
    /* unsigned (wrapping) arithmetic on purpose: d + 2^31 mod 2^32 is the biased low word for either sign of d */
    const uint32_t g00 = q & 0xffu, g01 = (q >> 8) & 0xffu, g10 = (q >> 16) & 0xffu, g11 = q >> 24;
    const double c00 = __hiloint2double(0x43300000, (int)g00) - 4503599627370496.0;
    const double dx = __hiloint2double(0x43300000, (int)((g01 - g00) ^ 0x80000000u)) - PMVS_BIAS_S32;
    const double dy = __hiloint2double(0x43300000, (int)((g10 - g00) ^ 0x80000000u)) - PMVS_BIAS_S32;
    const double dxy = __hiloint2double(0x43300000, (int)((g11 - g10 - g01 + g00) ^ 0x80000000u)) - PMVS_BIAS_S32;
    return fma(fy, fma(fx, dxy, dy), fma(fx, dx, c00));
}

/*
 * Unchecked sample loop for exactly V views (compile-time, 2..16; more views: fitness_columns_many below), one of which is
 * the reference view.
 *
 * Lane = one window column (several row groups when the window is narrow, several column passes when it is wider
 * than 32), so the x-dependent part of every homography row, A = H0*x+H2, B = H3*x+H5, C = H6*x+H8, is computed once
 * per evaluation and parked in a per-lane shared-memory slot (registers are the scarce resource); a sample of a
 * non-reference view then costs, on the FP64 pipe:
 *   3 fma   projective coordinates  X = h1*y+A, Y = h4*y+B, w = h7*y+C
 *   3 fma   reciprocal: MUFU.RCP64H seed r0, r = r0*(1+e+e^2), e = 1-w*r0
 *   2 mul   ix = X*r, iy = Y*r
 *   2 add   round-down magic adds: low words = px, py (the tap address)
 *   2 add   fy = iy - floor(iy)
 *   3 fma   bilinear blend (patch.cpp:1005-1017) without forming fx: with integer tap differences dx, dy, dxy,
 *             c = g00 + fx*dx + fy*(dy + fx*dxy),  fx = ix - px
 *               = fma(fy, fma(ix, dxy, dy - px*dxy), fma(ix, dx, g00 - px*dx))
 *           the integer parts dy - px*dxy and g00 - px*dx are exact (IMAD, |.| < 2^22) and each fma rounds once, so
 *           this is as accurate as the fx form and saves the two adds that would form fx.
 *   3 add   cross-view mean and absolute deviation
 * The reference view has H = I (patch.cpp:317-319): its sample is (x, y) bit for bit, so its pixel index and
 * fractions are per-column / per-row constants (no homography, reciprocal or floor), and the background-mask pixel
 * img(cvRound(y), cvRound(x)) (patch.cpp:986) is one of the four bytes of the tap word it loads anyway.
 * Above PMVS_SLOT_VIEWS non-reference views the slots would take so much shared memory (768 B per view and warp) that
 * occupancy and the L1 the taps live in suffer: those instantiations form the x-part inline from a compact homography
 * table, X = fma(h1, y, fma(h0, x, h2)) — the same expression, three more fma per sample.
 * The body is branch-free; each lane sums its rows in ascending order, then the fixed xor-tree.
 */
__device__ __forceinline__ double lds_f64_v(unsigned a) {   /* ordered against the volatile stores below */
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64_v(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ int2 lds_s32x2(unsigned a) {
    int2 v;
    asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

/* small signed integer -> f64. PMVS_CVT_MIX picks the pipe: the conversion unit (I2F.F64, 16 lanes/SM) or one FP64 add
 * on the biased-exponent form (2^52 + 2^31 + n) - (2^52 + 2^31); both exact. */
#ifndef PMVS_CVT_MIX
#define PMVS_CVT_MIX 0
#endif
#ifndef PMVS_TWO_ROW_VIEWS
#define PMVS_TWO_ROW_VIEWS 4    /* up to this many non-reference views, both rows of a trip are staged together */
#endif
#ifndef PMVS_CHUNK
#define PMVS_CHUNK 4            /* non-reference views staged together in the chunked row */
#endif
__device__ __forceinline__ double cvt_a(int n) { return PMVS_CVT_MIX >= 2 ? s2d(n) : (double)n; }
__device__ __forceinline__ double cvt_b(int n) { return PMVS_CVT_MIX >= 1 ? s2d(n) : (double)n; }

/* exp(x) for the difference weight exp(-avgSad^2/diffWeighting) (patch.cpp:1034) when -700 <= x <= 0 is guaranteed:
 * x = (k/64) ln2 + r, |r| <= ln2/128; exp = 2^(k>>6) * T[k&63] * (1 + r + ... + r^5/120) (remainder < 4e-17),
 * T[i] = 2^(i/64) from shared memory. */
__constant__ double kExpTab64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
#define PMVS_DIST_PAD(ps) (PMVS_DIST_GLOBAL ? 0 : (((ps) * (ps) + 1) & ~1))   /* the table sits behind the distance weights in shared memory */
__device__ __forceinline__ void load_exp_table(double *sDistW, int ps, int tid, int nthreads) {
    for (int k = tid; k < 64; k += nthreads) sDistW[PMVS_DIST_PAD(ps) + k] = kExpTab64[k];
}
__device__ __forceinline__ double exp_table(double x, unsigned tabA) {
    const double t = fma(x, 92.33248261689366, 6755399441055744.0);     /* 64/ln2; low word = round-to-nearest k */
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -0.01083042469326756, x);                              /* ln2/64 split: hi has 32 significant bits */
    r = fma(kd, -2.9815858269852933e-12, r);
    const double T = lds_f64(tabA + 8u * (unsigned)(k & 63));
    double q = fma(r, 8.33333333333333321769e-03, 4.16666666666666643537e-02);
    q = fma(r, q, 1.66666666666666657415e-01);
    q = fma(r, q, 0.5);
    q = fma(r, q, 1.0);
    const double p = fma(T * r, q, T);
    return __hiloint2double(__double2hiint(p) + ((k >> 6) << 20), __double2loint(p));
}

/* compact table of the non-reference views (skipping the reference view's index): homography, quad pointer, cols */
__device__ __forceinline__ void fill_view_table(const EvalCtx &E, const WarpWork &W, int NG, int refV) {
    for (int k = threadIdx.x & 31; k < NG; k += 32) {
        const int v = k + (k >= refV ? 1 : 0);
        const ViewS &vw = E.view[v];
        double *g = W.gv + 12 * k;
#pragma unroll
        for (int q = 0; q < 9; ++q) g[q] = W.H[9 * v + q];
        *(const uint32_t **)(g + 9) = vw.quad;
        *(int *)(g + 10) = vw.cols;
    }
}

template <int N>
struct ColumnTaps {
    double ix[N], fy[N];
    uint32_t q[N];
    int px[N];
};

/* projective coordinates, floors and tap loads of the non-reference views [k0, k0+N) for the row y of this lane's
 * column; gvA = compact view table (h1, h4, h7, quad, cols per view), cvA = this lane's slot base: A,B,C of view k at
 * cvA + 256*(3k + {0,1,2}) */
template <int N>
__device__ __forceinline__ void column_coords(unsigned gvA, unsigned cvA, int k0, double y, ColumnTaps<N> &t) {
    double w[N], r[N], e[N];
#pragma unroll
    for (int k = 0; k < N; ++k) w[k] = fma(lds_f64(gvA + 48u * (k0 + k) + 16u), y, lds_f64_v(cvA + 256u * (3 * (k0 + k) + 2)));
#pragma unroll
    for (int k = 0; k < N; ++k) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[k]) : "d"(w[k]));
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = fma(-w[k], r[k], 1.0);
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = fma(e[k], e[k], e[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = fma(r[k], e[k], r[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        t.ix[k] = fma(lds_f64(gvA + 48u * (k0 + k)), y, lds_f64_v(cvA + 256u * (3 * (k0 + k)))) * r[k];
        t.fy[k] = fma(lds_f64(gvA + 48u * (k0 + k) + 8u), y, lds_f64_v(cvA + 256u * (3 * (k0 + k) + 1))) * r[k];     /* iy */
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        w[k] = __dadd_rd(t.ix[k], PMVS_MAGIC_FLOOR);
        r[k] = __dadd_rd(t.fy[k], PMVS_MAGIC_FLOOR);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const uint32_t *__restrict__ quad = (const uint32_t *)lds_u64(gvA + 48u * (k0 + k) + 24u);
        const int cols = lds_s32(gvA + 48u * (k0 + k) + 32u);
        t.px[k] = __double2loint(w[k]);
        t.q[k] = __ldg(quad + (__double2loint(r[k]) * cols + t.px[k]));
    }
#pragma unroll
    for (int k = 0; k < N; ++k) t.fy[k] = t.fy[k] - (r[k] - PMVS_MAGIC_FLOOR);
}

template <int N>
__device__ __forceinline__ void column_blend(const ColumnTaps<N> &t, double *c) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int g00 = (int)__byte_perm(t.q[k], 0, 0x4440), g01 = (int)__byte_perm(t.q[k], 0, 0x4441);
        const int g10 = (int)__byte_perm(t.q[k], 0, 0x4442), g11 = (int)__byte_perm(t.q[k], 0, 0x4443);
        /* negated differences, so the integer parts are plain multiply-adds and the sign rides on the fma operand */
        const int ndx = g00 - g01, idy = g10 - g00, ndxy = g10 - g11 - ndx;
        const int k0 = t.px[k] * ndx + g00, k1 = t.px[k] * ndxy + idy;
        c[k] = fma(t.fy[k], fma(-t.ix[k], cvt_a(ndxy), cvt_b(k1)), fma(-t.ix[k], cvt_a(ndx), cvt_b(k0)));
    }
}

/* the reference view's sample of one row: tap word, mask bit, colour */
struct RefColumn {
    const uint32_t *__restrict__ quad;   /* refQuad + floor(x) */
    double fx;
    int selx;                            /* cvRound(x) - floor(x) */
};
__device__ __forceinline__ double ref_blend(const RefColumn &rc, uint32_t q, double fy) {
    const int g00 = (int)__byte_perm(q, 0, 0x4440), g01 = (int)__byte_perm(q, 0, 0x4441);
    const int g10 = (int)__byte_perm(q, 0, 0x4442), g11 = (int)__byte_perm(q, 0, 0x4443);
    const int idx = g01 - g00, idy = g10 - g00, idxy = g11 - g10 - idx;
    return fma(fy, fma(rc.fx, cvt_a(idxy), cvt_b(idy)), fma(rc.fx, cvt_a(idx), cvt_b(g00)));
}
__device__ __forceinline__ double ref_sample(const RefColumn &rc, int2 ri, double fy, bool &keep) {
    const uint32_t q = __ldg(rc.quad + ri.x);
    keep = __byte_perm(q, 0, 0x4440 + ri.y + rc.selx) != 0;                            /* patch.cpp:986 */
    return ref_blend(rc, q, fy);
}

/* cross-view mean and summed absolute deviation (patch.cpp:1019-1027; the division by V is folded into the callers'
 * constants) */
template <int N>
__device__ __forceinline__ double tree_sum(const double *c) {       /* pairwise: depth log2(N) instead of N */
    if constexpr (N == 1) return c[0];
    else return tree_sum<N / 2>(c) + tree_sum<N - N / 2>(c + N / 2);
}
template <int V>
__device__ __forceinline__ double sum_abs_dev(const double *c, double invV) {
    const double mean = tree_sum<V>(c) * invV;
    double d[V];
#pragma unroll
    for (int v = 0; v < V; ++v) d[v] = fabs(c[v] - mean);
    return tree_sum<V>(d);
}

template <int N>
__device__ __forceinline__ void direct_coords(unsigned gvK, double x, double y, ColumnTaps<N> &t);

/* the non-reference views of one row in chunks; SLOTS: x-parts from the per-lane slots, else formed inline */
template <int NG, int K0, bool SLOTS>
__device__ __forceinline__ void row_chunks(unsigned gvA, unsigned cvA, double x, double y, double *c) {
    if constexpr (K0 < NG) {
        constexpr int N = (NG - K0) < PMVS_CHUNK ? (NG - K0) : PMVS_CHUNK;
        ColumnTaps<N> t;
        if constexpr (SLOTS) column_coords<N>(gvA, cvA, K0, y, t);
        else direct_coords<N>(gvA + 96u * K0, x, y, t);
        column_blend<N>(t, c + K0);
        row_chunks<NG, K0 + N, SLOTS>(gvA, cvA, x, y, c);
    }
}

/* FAST: no gradient weight and the difference weight's exponent provably in [-700, 0] — the distance weight is
 * always read (the table holds ones when it is disabled) and the difference weight always evaluated (negK = 0 when it
 * is disabled), so the row loop carries no configuration flags. */
template <int V, bool FAST, bool SLOTS>
__device__ __noinline__ void fitness_columns(const DevScene &S, const EvalCtx &E, const double *__restrict__ sDistW,
                                             const double *__restrict__ sExpT, const WarpWork &W, int nx, int ny, double &fitOut,
                                             double &swOut) {
    constexpr int NG = V - 1;                       /* non-reference views */
    static_assert(!SLOTS || NG <= PMVS_SLOT_VIEWS, "slot mode holds at most PMVS_SLOT_VIEWS non-reference views");
    const int lane = threadIdx.x & 31;
    const int G = nx <= 16 ? 32 / nx : 1;          /* row groups sharing the warp (narrow windows) */
    const unsigned hA = smem_addr(W.H), ysA = smem_addr(W.ys), xsA = smem_addr(W.xs), distA = smem_addr(sDistW), viewA = smem_addr(E.view);
    const unsigned gvA = smem_addr(W.gv), rfA = smem_addr(W.rowf), riA = smem_addr(W.rowi), tabA = smem_addr(sExpT);
    const unsigned cvA = smem_addr(W.colv) + 8u * lane;
    const double invV = 1.0 / (double)V;
    const bool useDist = FAST || S.cfg.adaptiveDistanceEnable, useDiff = FAST || S.cfg.adaptiveDifferenceEnable;
    const bool useGrad = !FAST && S.cfg.adaptiveGradientEnable;
    const double gradW = S.cfg.gradientWeighting;
    /* exp argument = negK * (sum of deviations)^2 */
    const double negK = S.cfg.adaptiveDifferenceEnable ? -(invV * invV) / S.cfg.diffWeighting : 0.0;
    const bool expSafe = FAST;
    const double *__restrict__ refEdge = E.refEdge;
    const int refCols = E.refCols, refV = E.refView;

    /* per-evaluation tables: compact non-reference view list, per-row constants of the reference view */
    if constexpr (SLOTS) {
        if (lane < NG) {
            const int v = lane + (lane >= refV ? 1 : 0);
            const ViewS &vw = E.view[v];
            double *g = W.gv + 6 * lane;
            g[0] = W.H[9 * v + 1];
            g[1] = W.H[9 * v + 4];
            g[2] = W.H[9 * v + 7];
            *(const uint32_t **)(g + 3) = vw.quad;
            *(int *)(g + 4) = vw.cols;
        }
    } else {
        fill_view_table(E, W, NG, refV);
    }
    for (int j = lane; j < ny; j += 32) {
        const double y = W.ys[j];
        const double ty = __dadd_rd(y, PMVS_MAGIC_FLOOR);
        const int py = __double2loint(ty);
        W.rowf[j] = y - (ty - PMVS_MAGIC_FLOOR);
        W.rowi[j] = make_int2(py * refCols, 2 * (__double2int_rn(y) - py));
    }
    __syncwarp();

    double fit = 0, sw = 0;
    for (int i0 = 0; i0 < nx; i0 += 32) {          /* column passes (windows wider than the warp) */
        const int span = nx - i0 < 32 ? nx - i0 : 32;
        const int i = i0 + (G > 1 ? lane % nx : lane), g = G > 1 ? lane / nx : 0;
        const bool active = G > 1 ? g < G : lane < span;
        const double x = lds_f64(xsA + 8u * (active ? i : i0));
#pragma unroll
        for (int k = 0; k < (SLOTS ? NG : 0); ++k) {
            const unsigned h = hA + 72u * (unsigned)(k + (k >= refV ? 1 : 0));
            sts_f64_v(cvA + 256u * (3 * k), fma(lds_f64(h), x, lds_f64(h + 16u)));
            sts_f64_v(cvA + 256u * (3 * k + 1), fma(lds_f64(h + 24u), x, lds_f64(h + 40u)));
            sts_f64_v(cvA + 256u * (3 * k + 2), fma(lds_f64(h + 48u), x, lds_f64(h + 64u)));
        }
        RefColumn rc;
        const double tx = __dadd_rd(x, PMVS_MAGIC_FLOOR);
        const int pxr = __double2loint(tx), rx = __double2int_rn(x);
        rc.quad = E.refQuad + pxr;
        rc.fx = x - (tx - PMVS_MAGIC_FLOOR);
        rc.selx = rx - pxr;
        const int jEnd = active ? ny : 0;
        for (int j = g; j < jEnd; j += 2 * G) {
            const bool two = j + G < jEnd;
            const int j2 = two ? j + G : j;
            const double y0 = lds_f64(ysA + 8u * j), y1 = lds_f64(ysA + 8u * j2);
            const int2 ri0 = lds_s32x2(riA + 8u * j), ri1 = lds_s32x2(riA + 8u * j2);
            bool keep0, keep1;
            double s0, s1;
            if constexpr (NG <= PMVS_TWO_ROW_VIEWS) {
                /* both rows' tap loads are in flight before the first one is consumed */
                ColumnTaps<NG> ta, tb;
                double ca[V], cb[V];
                if constexpr (SLOTS) {
                    column_coords<NG>(gvA, cvA, 0, y0, ta);
                    column_coords<NG>(gvA, cvA, 0, y1, tb);
                } else {
                    direct_coords<NG>(gvA, x, y0, ta);
                    direct_coords<NG>(gvA, x, y1, tb);
                }
                ca[NG] = ref_sample(rc, ri0, lds_f64(rfA + 8u * j), keep0);
                cb[NG] = ref_sample(rc, ri1, lds_f64(rfA + 8u * j2), keep1);
                column_blend<NG>(ta, ca);
                column_blend<NG>(tb, cb);
                s0 = sum_abs_dev<V>(ca, invV);
                s1 = sum_abs_dev<V>(cb, invV);
            } else {
                /* many views: one row at a time (a real loop, so the two rows' colours are never live together) */
                s0 = s1 = 0;
                keep0 = keep1 = false;
#pragma unroll 1
                for (int rrow = 0; rrow < 2; ++rrow) {
                    double c[V];
                    bool kp;
                    c[NG] = ref_sample(rc, rrow ? ri1 : ri0, lds_f64(rfA + 8u * (rrow ? j2 : j)), kp);
                    row_chunks<NG, 0, SLOTS>(gvA, cvA, x, rrow ? y1 : y0, c);
                    const double sv = sum_abs_dev<V>(c, invV);
                    if (rrow) { s1 = sv; keep1 = kp; }
                    else { s0 = sv; keep0 = kp; }
                }
            }
            keep1 = keep1 && two;
            double w0 = 1.0, w1 = 1.0;
            if (useDist) {                                                                    /* patch.cpp:1030-1032 */
                w0 = PMVS_DIST_GLOBAL ? __ldg(S.distW + (i * ny + j)) : lds_f64(distA + 8u * (i * ny + j));
                w1 = PMVS_DIST_GLOBAL ? __ldg(S.distW + (i * ny + j2)) : lds_f64(distA + 8u * (i * ny + j2));
            }
            if (useDiff) {                                                                    /* patch.cpp:1033-1035 */
                const double x0 = s0 * s0 * negK, x1 = s1 * s1 * negK;
                if (expSafe) { w0 *= exp_table(x0, tabA); w1 *= exp_table(x1, tabA); }
                else { w0 *= exp_nonpos(x0); w1 *= exp_nonpos(x1); }
            }
            if (useGrad) {                                                                    /* patch.cpp:1036-1038 */
                const int rofs0 = ri0.x + (ri0.y ? refCols : 0) + pxr + rc.selx, rofs1 = ri1.x + (ri1.y ? refCols : 0) + pxr + rc.selx;
                w0 *= exp_nonpos(-1.0 / (__ldg(refEdge + rofs0) * gradW));
                w1 *= exp_nonpos(-1.0 / (__ldg(refEdge + rofs1) * gradW));
            }
            w0 = keep0 ? w0 : 0.0;
            w1 = keep1 ? w1 : 0.0;
            sw += w0;
            fit = fma(w0, s0, fit);
            sw += w1;
            fit = fma(w1, s1, fit);
        }
    }
    fitOut = warp_sum(fit) * invV;
    swOut = warp_sum(sw);
}

/*
 * The same loop for any number of views (V > 16: BASELINE.json's 32- and 64-view configurations). The colours of one
 * pixel no longer fit in registers and per-lane slots for every view would not fit in shared memory, so a pixel takes
 * two passes over the views in chunks of four — first the cross-view sum, then, with the mean known, the absolute
 * deviations (the samples are recomputed: same arithmetic, same values) — and the x-part of each homography row is
 * formed inline, X = fma(h1, y, fma(h0, x, h2)): the expression the slots hold, so a sample is bit-identical to the
 * V <= 16 loop's. Two rows per trip share the homography loads. Configuration flags are read per trip (the trip is
 * tens of chunks long).
 */
template <int N>
__device__ __forceinline__ void direct_coords(unsigned gvK, double x, double y, ColumnTaps<N> &t) {   /* gvK: table entry of the first view */
    double w[N], r[N], e[N];
#pragma unroll
    for (int k = 0; k < N; ++k) w[k] = fma(lds_f64(gvK + 96u * k + 56u), y, fma(lds_f64(gvK + 96u * k + 48u), x, lds_f64(gvK + 96u * k + 64u)));
#pragma unroll
    for (int k = 0; k < N; ++k) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[k]) : "d"(w[k]));
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = fma(-w[k], r[k], 1.0);
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = fma(e[k], e[k], e[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = fma(r[k], e[k], r[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        t.ix[k] = fma(lds_f64(gvK + 96u * k + 8u), y, fma(lds_f64(gvK + 96u * k), x, lds_f64(gvK + 96u * k + 16u))) * r[k];
        t.fy[k] = fma(lds_f64(gvK + 96u * k + 32u), y, fma(lds_f64(gvK + 96u * k + 24u), x, lds_f64(gvK + 96u * k + 40u))) * r[k];     /* iy */
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        w[k] = __dadd_rd(t.ix[k], PMVS_MAGIC_FLOOR);
        r[k] = __dadd_rd(t.fy[k], PMVS_MAGIC_FLOOR);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const uint32_t *__restrict__ quad = (const uint32_t *)lds_u64(gvK + 96u * k + 72u);
        const int cols = lds_s32(gvK + 96u * k + 80u);
        t.px[k] = __double2loint(w[k]);
        t.q[k] = __ldg(quad + (__double2loint(r[k]) * cols + t.px[k]));
    }
#pragma unroll
    for (int k = 0; k < N; ++k) t.fy[k] = t.fy[k] - (r[k] - PMVS_MAGIC_FLOOR);
}

/* PASS 0: sum of the colours of views [k0, k0+N) of both rows; PASS 1: sum of |colour - mean| */
template <int N, int PASS>
__device__ __forceinline__ void many_chunk(unsigned gvK, double x, double y0, double y1, double m0, double m1, double &a0, double &a1) {
    ColumnTaps<N> ta, tb;
    double ca[N], cb[N];
    direct_coords<N>(gvK, x, y0, ta);
    direct_coords<N>(gvK, x, y1, tb);
    column_blend<N>(ta, ca);
    column_blend<N>(tb, cb);
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < N; ++k) { ca[k] = fabs(ca[k] - m0); cb[k] = fabs(cb[k] - m1); }
    }
    a0 += tree_sum<N>(ca);
    a1 += tree_sum<N>(cb);
}

__device__ __noinline__ void fitness_columns_many(const DevScene &S, const EvalCtx &E, const double *__restrict__ sDistW,
                                                  const double *__restrict__ sExpT, const WarpWork &W, int nx, int ny, double &fitOut,
                                                  double &swOut) {
    const int V = E.V, NG = V - 1, refV = E.refView;
    const int lane = threadIdx.x & 31;
    const int G = nx <= 16 ? 32 / nx : 1;
    const unsigned ysA = smem_addr(W.ys), xsA = smem_addr(W.xs), distA = smem_addr(sDistW), gvA = smem_addr(W.gv);
    const unsigned rfA = smem_addr(W.rowf), riA = smem_addr(W.rowi), tabA = smem_addr(sExpT);
    const double invV = 1.0 / (double)V;
    const double negK = S.cfg.adaptiveDifferenceEnable ? -(invV * invV) / S.cfg.diffWeighting : 0.0;
    const bool expSafe = -(255.0 * 255.0) / S.cfg.diffWeighting >= -700.0;
    const int refCols = E.refCols;
    fill_view_table(E, W, NG, refV);
    for (int j = lane; j < ny; j += 32) {
        const double y = W.ys[j];
        const double ty = __dadd_rd(y, PMVS_MAGIC_FLOOR);
        const int py = __double2loint(ty);
        W.rowf[j] = y - (ty - PMVS_MAGIC_FLOOR);
        W.rowi[j] = make_int2(py * refCols, 2 * (__double2int_rn(y) - py));
    }
    __syncwarp();
    double fit = 0, sw = 0;
    for (int i0 = 0; i0 < nx; i0 += 32) {
        const int span = nx - i0 < 32 ? nx - i0 : 32;
        const int i = i0 + (G > 1 ? lane % nx : lane), g = G > 1 ? lane / nx : 0;
        const bool active = G > 1 ? g < G : lane < span;
        const double x = lds_f64(xsA + 8u * (active ? i : i0));
        RefColumn rc;
        const double tx = __dadd_rd(x, PMVS_MAGIC_FLOOR);
        const int pxr = __double2loint(tx), rx = __double2int_rn(x);
        rc.quad = E.refQuad + pxr;
        rc.fx = x - (tx - PMVS_MAGIC_FLOOR);
        rc.selx = rx - pxr;
        const int jEnd = active ? ny : 0;
        for (int j = g; j < jEnd; j += 2 * G) {
            const bool two = j + G < jEnd;
            const int j2 = two ? j + G : j;
            const double y0 = lds_f64(ysA + 8u * j), y1 = lds_f64(ysA + 8u * j2);
            const int2 ri0 = lds_s32x2(riA + 8u * j), ri1 = lds_s32x2(riA + 8u * j2);
            bool keep0, keep1;
            const double cr0 = ref_sample(rc, ri0, lds_f64(rfA + 8u * j), keep0);
            const double cr1 = ref_sample(rc, ri1, lds_f64(rfA + 8u * j2), keep1);
            double s0 = cr0, s1 = cr1;
            int k0 = 0;
#pragma unroll 1
            for (; k0 + 4 <= NG; k0 += 4) many_chunk<4, 0>(gvA + 96u * k0, x, y0, y1, 0.0, 0.0, s0, s1);
#pragma unroll 1
            for (; k0 < NG; ++k0) many_chunk<1, 0>(gvA + 96u * k0, x, y0, y1, 0.0, 0.0, s0, s1);
            const double m0 = s0 * invV, m1 = s1 * invV;
            s0 = fabs(cr0 - m0);
            s1 = fabs(cr1 - m1);
#pragma unroll 1
            for (k0 = 0; k0 + 4 <= NG; k0 += 4) many_chunk<4, 1>(gvA + 96u * k0, x, y0, y1, m0, m1, s0, s1);
#pragma unroll 1
            for (; k0 < NG; ++k0) many_chunk<1, 1>(gvA + 96u * k0, x, y0, y1, m0, m1, s0, s1);
            keep1 = keep1 && two;
            double w0 = 1.0, w1 = 1.0;
            if (S.cfg.adaptiveDistanceEnable) {                                               /* patch.cpp:1030-1032 */
                w0 = PMVS_DIST_GLOBAL ? __ldg(S.distW + (i * ny + j)) : lds_f64(distA + 8u * (i * ny + j));
                w1 = PMVS_DIST_GLOBAL ? __ldg(S.distW + (i * ny + j2)) : lds_f64(distA + 8u * (i * ny + j2));
            }
            if (S.cfg.adaptiveDifferenceEnable) {                                             /* patch.cpp:1033-1035 */
                const double x0 = s0 * s0 * negK, x1 = s1 * s1 * negK;
                if (expSafe) { w0 *= exp_table(x0, tabA); w1 *= exp_table(x1, tabA); }
                else { w0 *= exp_nonpos(x0); w1 *= exp_nonpos(x1); }
            }
            if (S.cfg.adaptiveGradientEnable) {                                               /* patch.cpp:1036-1038 */
                const int rofs0 = ri0.x + (ri0.y ? refCols : 0) + pxr + rc.selx, rofs1 = ri1.x + (ri1.y ? refCols : 0) + pxr + rc.selx;
                w0 *= exp_nonpos(-1.0 / (__ldg(E.refEdge + rofs0) * S.cfg.gradientWeighting));
                w1 *= exp_nonpos(-1.0 / (__ldg(E.refEdge + rofs1) * S.cfg.gradientWeighting));
            }
            w0 = keep0 ? w0 : 0.0;
            w1 = keep1 ? w1 : 0.0;
            sw += w0;
            fit = fma(w0, s0, fit);
            sw += w1;
            fit = fma(w1, s1, fit);
        }
    }
    fitOut = warp_sum(fit) * invV;
    swOut = warp_sum(sw);
}

__device__ __forceinline__ bool ref_window_ok(const DevScene &S, const EvalCtx &E, const double *center, double *pt);

/* =====================================================================================================
 * View-lane window loop (fitness_vl): the unchecked sample loop of getFitness (patch.cpp:979-1042) for 2..9 views
 * over the patch's reference window (RefWin).
 *
 * A group of GL lanes owns one window column at a time (32/GL columns per pass) and GL consecutive rows of it.
 * In the SAMPLE phase lane s of the group is a VIEW: it keeps the homography constants of non-reference view
 * c*GL + s (c < NCH chunks) in registers — h1, h4, h7 and the x-parts A = h0 x + h2, B = h3 x + h5, C = h6 x + h8 of
 * its current column — and samples that view at the group's GL rows, in the row order s^0, s^1, .. s^(GL-1). No
 * shared-memory traffic for homographies, no per-lane slots. Then t = 1..GL-1 butterfly shuffles (lane s sends its
 * sample of row s^t to lane s^t, a FIXED register: no selects) transpose the GL x GL block, and in the PIXEL phase
 * lane s is a ROW: it holds the colours of all views at (column, row s), adds the reference view's colour from the
 * RefWin table and does the cross-view mean / absolute deviations (patch.cpp:1019-1027), the difference weight
 * (:1033-1035) and the accumulation (:1039-1040) for that one pixel. The views reach the lanes in the order
 * s, s^1, s^2, s^3, i.e. as the unordered pairs {0,1}, {2,3}: the pairwise sums (c0 + c1) + (c2 + c3) are bit-identical
 * on every lane because IEEE addition is commutative — results do not depend on the lane a pixel lands on.
 * Per sample the arithmetic is column_coords / column_blend's (same expressions, same values). The distance weight
 * is applied in its separable form gx[column] * gy[row] (RefWin), the background mask is one bit test.
 * FULL: NG == GL * NCH (no padding views). Padding lanes sample view 0 again and their colours are discarded.
 * =================================================================================================== */
__device__ __forceinline__ double2 lds_f64x2(unsigned a) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
/* butterfly shuffle of a double IN PLACE (both halves keep their registers: no pair-repacking moves around the SHFLs) */
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    asm volatile("{ .reg .b32 lo, hi;\n\t"
                 "mov.b64 {lo, hi}, %0;\n\t"
                 "shfl.sync.bfly.b32 lo, lo, %1, 0x1f, 0xffffffff;\n\t"
                 "shfl.sync.bfly.b32 hi, hi, %1, 0x1f, 0xffffffff;\n\t"
                 "mov.b64 %0, {lo, hi}; }"
                 : "+d"(v)
                 : "r"(m));
    return v;
}

/* exp(x) of the difference weight for -700 <= x <= 0 as exp_table, with its constants in the constant bank (operands
 * of the fma, no register or uniform-register traffic) */
__constant__ double kExpT[8] = {92.33248261689366, -0.01083042469326756, -2.9815858269852933e-12, 8.33333333333333321769e-03,
                                4.16666666666666643537e-02, 1.66666666666666657415e-01, 0.5, 1.0};
__device__ __forceinline__ double exp_table_c(double x, unsigned tabA) {
    const double t = fma(x, kExpT[0], 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, kExpT[1], x);
    r = fma(kd, kExpT[2], r);
    const double T = lds_f64(tabA + 8u * (unsigned)(k & 63));
    double q = fma(r, kExpT[3], kExpT[4]);
    q = fma(r, q, kExpT[5]);
    q = fma(r, q, kExpT[6]);
    q = fma(r, q, kExpT[7]);
    const double p = fma(T * r, q, T);
    return __hiloint2double(__double2hiint(p) + ((k >> 6) << 20), __double2loint(p));
}

/* N tap loads issued back to back: inside a (possibly divergent) device function every LDG needs the global-memory
 * descriptor re-validated in a uniform register (LDC + 2 R2UR); adjacent loads share one */
template <int N>
__device__ __forceinline__ void ldg_group(const uint32_t *const *ad, uint32_t *q) {
    if constexpr (N == 8) {
        asm volatile("ld.global.nc.u32 %0, [%8];\n\tld.global.nc.u32 %1, [%9];\n\tld.global.nc.u32 %2, [%10];\n\tld.global.nc.u32 %3, [%11];\n\t"
                     "ld.global.nc.u32 %4, [%12];\n\tld.global.nc.u32 %5, [%13];\n\tld.global.nc.u32 %6, [%14];\n\tld.global.nc.u32 %7, [%15];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                     : "l"(ad[0]), "l"(ad[1]), "l"(ad[2]), "l"(ad[3]), "l"(ad[4]), "l"(ad[5]), "l"(ad[6]), "l"(ad[7]));
    } else if constexpr (N == 4) {
        asm volatile("ld.global.nc.u32 %0, [%4];\n\tld.global.nc.u32 %1, [%5];\n\tld.global.nc.u32 %2, [%6];\n\tld.global.nc.u32 %3, [%7];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3])
                     : "l"(ad[0]), "l"(ad[1]), "l"(ad[2]), "l"(ad[3]));
    } else {
#pragma unroll
        for (int n = 0; n < N; ++n) q[n] = __ldg(ad[n]);
    }
}

/* N samples staged together (the stages of column_coords / column_blend): all projective coordinates, all reciprocals,
 * all tap loads in flight before the first blend */
template <int N>
struct VlTaps {
    double ix[N], fy[N];
    uint32_t q[N];
    int px[N];
};

#ifndef PMVS_VL_DP4A
#define PMVS_VL_DP4A 1
#endif
#ifndef PMVS_VL_F2I
#define PMVS_VL_F2I 1
#endif
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {      /* sum of (unsigned byte of a) * (signed byte of b) + c */
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#ifndef PMVS_VL_RB
#define PMVS_VL_RB 2            /* row blocks (of GL rows) per trip of the view-lane loop */
#endif
#ifndef PMVS_VL_PIPE
#define PMVS_VL_PIPE 0          /* 1: software-pipelined trips (tap loads of trip t+1 issued before the pixel phase of trip t) */
#endif

template <int GL, int NCH, int RB, bool FULL, bool TILE = false>
__device__ __noinline__ void fitness_vl(const DevScene &S, const EvalCtx &E, const RefWin &R, const double *__restrict__ Hw,
                                        const double *__restrict__ sExpT, double &fitOut, double &swOut) {
    constexpr int CP = 32 / GL;                     /* columns per pass */
    constexpr int LG = GL == 4 ? 2 : (GL == 2 ? 1 : 0);
    constexpr int NS = NCH * GL * RB;               /* samples of one lane per trip */
    constexpr int FW = 64 / GL;                     /* width of one lane's field of the row mask */
    constexpr unsigned STEP = 16u * GL * RB;        /* bytes of ysg one trip covers */
    const int lane = threadIdx.x & 31, s = lane & (GL - 1), ci = lane >> LG;
    const int V = E.V, NG = V - 1, refV = E.refView, nx = R.nx, nyp = R.nyp;
    const int ny = (R.ny + GL * RB - 1) / (GL * RB) * (GL * RB);       /* rows incl. padding (padding rows carry mask bit 0) */
    const unsigned END = 16u * (unsigned)ny;
    const unsigned hA = smem_addr(Hw), xsA = smem_addr(R.xs), gxA = smem_addr(R.gx), ysgA = smem_addr(R.ysg), mkA = smem_addr(R.mask);
    const unsigned rcA = smem_addr(R.refc), tabA = smem_addr(sExpT), viewA = smem_addr(E.view);
    const double invV = 1.0 / (double)V;
    const double negK = S.cfg.adaptiveDifferenceEnable ? -(invV * invV) / S.cfg.diffWeighting : 0.0;

    /* this lane's views */
    double h1[NCH], h4[NCH], h7[NCH];
    const uint32_t *quad[NCH];
    int cols[NCH];
    unsigned hv[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        int k = c * GL + s;
        if (!FULL && k >= NG) k = 0;
        const int v = k + (k >= refV ? 1 : 0);
        hv[c] = hA + 72u * (unsigned)v;
        h1[c] = lds_f64(hv[c] + 8u);
        h4[c] = lds_f64(hv[c] + 32u);
        h7[c] = lds_f64(hv[c] + 56u);
        const unsigned va = viewA + (unsigned)(sizeof(ViewS) * v);
        quad[c] = (const uint32_t *)lds_u64(va + (unsigned)offsetof(ViewS, quad));
        cols[c] = lds_s32(va + (unsigned)offsetof(ViewS, cols));
    }
#if PMVS_TILE
    unsigned tb[NCH];                /* TILE: shared address of tile word (0, 0) of this lane's view, origin folded in */
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        int k = c * GL + s;
        if (k >= NG || k >= 4) k = 0;
        tb[c] = smem_addr(R.tile) + 4u * (unsigned)(k * PMVS_TILE * PMVS_TILE) - 4u * (unsigned)(R.toy[k] * PMVS_TILE + R.tox[k]);
    }
#endif
    /* which of the views that reach this lane in the pixel phase are real (padding views excluded) */
    bool real[NCH][GL];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int t = 0; t < GL; ++t) real[c][t] = FULL || (c * GL + (s ^ t)) < NG;
    /* row addresses of the sample phase: rows (s ^ t) of the current block; of the pixel phase: row s */
    unsigned ya[GL];
#pragma unroll
    for (int t = 0; t < GL; ++t) ya[t] = ysgA + 16u * (unsigned)(s ^ t);

    /* ---- the phases of one trip ---- */
    double A[NCH], B[NCH], C[NCH];
    /* x-parts of this lane's views for column pass i0 (sample side) */
    auto sample_column = [&](int i0) {
        const int i = i0 + ci;
        const int ic = i < nx ? i : nx - 1;
        const double x = lds_f64(xsA + 8u * ic);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            A[c] = fma(lds_f64(hv[c]), x, lds_f64(hv[c] + 16u));
            B[c] = fma(lds_f64(hv[c] + 24u), x, lds_f64(hv[c] + 40u));
            C[c] = fma(lds_f64(hv[c] + 48u), x, lds_f64(hv[c] + 64u));
        }
    };
    /* sample phase: my views at rows (s ^ t) of each block of the trip at byte offset jo; tap loads issued */
    auto sample = [&](unsigned jo, VlTaps<NS> &tp) {
        double w[NS], r[NS], e[NS], y[RB * GL];
#pragma unroll
        for (int b = 0; b < RB; ++b)
#pragma unroll
            for (int t = 0; t < GL; ++t) y[b * GL + t] = lds_f64(ya[t] + jo + 16u * GL * b);
#pragma unroll
        for (int n = 0; n < NS; ++n) w[n] = fma(h7[n / (GL * RB)], y[n % (GL * RB)], C[n / (GL * RB)]);
#pragma unroll
        for (int n = 0; n < NS; ++n) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[n]) : "d"(w[n]));
#pragma unroll
        for (int n = 0; n < NS; ++n) e[n] = fma(-w[n], r[n], 1.0);
#pragma unroll
        for (int n = 0; n < NS; ++n) e[n] = fma(e[n], e[n], e[n]);
#pragma unroll
        for (int n = 0; n < NS; ++n) r[n] = fma(r[n], e[n], r[n]);
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            tp.ix[n] = fma(h1[n / (GL * RB)], y[n % (GL * RB)], A[n / (GL * RB)]) * r[n];
            tp.fy[n] = fma(h4[n / (GL * RB)], y[n % (GL * RB)], B[n / (GL * RB)]) * r[n];       /* iy */
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) {
#if !PMVS_VL_F2I
            w[n] = __dadd_rd(tp.ix[n], PMVS_MAGIC_FLOOR);
#endif
            r[n] = __dadd_rd(tp.fy[n], PMVS_MAGIC_FLOOR);
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) {
#if PMVS_VL_F2I
            tp.px[n] = __double2int_rd(tp.ix[n]);          /* floor on the conversion unit (F2I.F64.FLOOR): the FP64 / integer dispatch is the loop's bound */
#else
            tp.px[n] = __double2loint(w[n]);
#endif
#if PMVS_TILE
            if (TILE) tp.q[n] = (uint32_t)lds_s32(tb[n / (GL * RB)] + 4u * (unsigned)(__double2loint(r[n]) * PMVS_TILE + tp.px[n]));
            else
#endif
            tp.q[n] = __ldg(quad[n / (GL * RB)] + (__double2loint(r[n]) * cols[n / (GL * RB)] + tp.px[n]));
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) tp.fy[n] = tp.fy[n] - (r[n] - PMVS_MAGIC_FLOOR);
    };
    /* blend (column_blend) + transpose: afterwards col[(c, b, t)] = view c*GL + (s ^ t) at row GL*b + s of the trip */
    auto blend = [&](const VlTaps<NS> &tp, double *col) {
#pragma unroll
        for (int n = 0; n < NS; ++n) {
#if PMVS_VL_DP4A
            /* tap differences as byte dot products (IDP.4A.U8.S8): taps (g00, g01, g10, g11) . signed weights */
            const int ndx = dp4a_us(tp.q[n], 0x0000ff01, 0);                   /* g00 - g01 */
            const int ndxy = dp4a_us(tp.q[n], (int)0xff0101ffu, 0);            /* -g00 + g01 + g10 - g11 */
            const int idy = dp4a_us(tp.q[n], 0x000100ff, 0);                   /* g10 - g00 */
            const int g00 = (int)(tp.q[n] & 0xffu);
#else
            const int g00 = (int)__byte_perm(tp.q[n], 0, 0x4440), g01 = (int)__byte_perm(tp.q[n], 0, 0x4441);
            const int g10 = (int)__byte_perm(tp.q[n], 0, 0x4442), g11 = (int)__byte_perm(tp.q[n], 0, 0x4443);
            const int ndx = g00 - g01, idy = g10 - g00, ndxy = g10 - g11 - ndx;
#endif
            const int k0 = tp.px[n] * ndx + g00, k1 = tp.px[n] * ndxy + idy;
            col[n] = fma(tp.fy[n], fma(-tp.ix[n], cvt_a(ndxy), cvt_b(k1)), fma(-tp.ix[n], cvt_a(ndx), cvt_b(k0)));
        }
#pragma unroll
        for (int n = 0; n < NS; ++n)
            if (n % GL) col[n] = shfl_xor_f64(col[n], n % GL);
    };
    /* pixel side: column constants, accumulators */
    double fit = 0, sw = 0, cfit = 0, csw = 0, gxv = 0;
    unsigned mlo = 0, mhi = 0, rcol = 0;
    auto pixel_column = [&](int i0) {
        const int i = i0 + ci;
        const bool colOk = i < nx;
        const int ic = colOk ? i : nx - 1;
        gxv = colOk ? lds_f64(gxA + 8u * ic) : 0.0;
        const unsigned long long mk = lds_u64(mkA + 8u * ic) >> (FW * s);     /* bit b = row GL * b + s of this column */
        mlo = (unsigned)mk;
        mhi = (unsigned)(mk >> 32);
        rcol = rcA + 8u * (unsigned)(ic * nyp + s);
    };
    /* pixel phase of the trip at byte offset jo: mean / deviations / weights / accumulation for row GL*b + s */
    auto pixel = [&](const double *col, unsigned jo) {
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const double cref = lds_f64(rcol + 8u * GL * b);
            /* row factor of the distance weight, zero where the reference pixel is background (patch.cpp:986) */
            double gy = 0.0;
            if (mlo & (1u << b)) gy = lds_f64(ya[0] + jo + 16u * GL * b + 8u);
            double cv[NCH][GL];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int t = 0; t < GL; ++t) cv[c][t] = (FULL || real[c][t]) ? col[(c * RB + b) * GL + t] : 0.0;
            double sum = tree_sum<GL>(cv[0]);
#pragma unroll
            for (int c = 1; c < NCH; ++c) sum += tree_sum<GL>(cv[c]);
            sum += cref;
            const double mean = sum * invV;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int t = 0; t < GL; ++t) cv[c][t] = (FULL || real[c][t]) ? fabs(cv[c][t] - mean) : 0.0;
            double dev = tree_sum<GL>(cv[0]);
#pragma unroll
            for (int c = 1; c < NCH; ++c) dev += tree_sum<GL>(cv[c]);
            dev += fabs(cref - mean);
            const double wgt = gy * exp_table_c(dev * dev * negK, tabA);                  /* patch.cpp:1030-1035 */
            csw += wgt;
            cfit = fma(wgt, dev, cfit);
        }
        rcol += 8u * GL * RB;
        if (GL == 1) mlo = __funnelshift_r(mlo, mhi, RB), mhi >>= RB;
        else mlo >>= RB;
    };
    auto close_column = [&]() {
        sw = fma(gxv, csw, sw);
        fit = fma(gxv, cfit, fit);
        cfit = csw = 0;
    };

    /* PMVS_VL_FENCE: a never-taken branch after the sample phase. ptxas schedules inside basic blocks: without the block
     * boundary it sinks every tap load next to its first use (a rolling schedule, ~20 instructions between LDG and PRMT) and
     * the warp stalls on each of them; with it, all tap loads of a trip are in flight before the first one is consumed */
#ifndef PMVS_VL_FENCE
#define PMVS_VL_FENCE 0
#endif
    const bool never = S.cfg.patchRadius < 0;
#define PMVS_FENCE_BLOCK() do { if (PMVS_VL_FENCE && never) { fit = -fit; sw = -sw; } } while (0)
#if PMVS_VL_PIPE
    /* trips in (column pass, row trip) order; the sample phase runs one trip ahead of the pixel phase */
    const int tripsPerPass = (int)(END / STEP), T = ((nx + CP - 1) / CP) * tripsPerPass;
    VlTaps<NS> tp;
    int cs = 0, cp = 0;
    unsigned js = 0, jp = 0;
    sample_column(0);
    pixel_column(0);
    sample(0u, tp);
    for (int t = 0; t < T - 1; ++t) {
        double col[NS];
        blend(tp, col);
        js += STEP;
        if (js >= END) {
            js = 0;
            cs += CP;
            sample_column(cs);
        }
        sample(js, tp);
        PMVS_FENCE_BLOCK();
        pixel(col, jp);
        jp += STEP;
        if (jp >= END) {
            close_column();
            jp = 0;
            cp += CP;
            pixel_column(cp);
        }
    }
    {
        double col[NS];
        blend(tp, col);
        pixel(col, jp);
        close_column();
    }
#else
    for (int i0 = 0; i0 < nx; i0 += CP) {
        sample_column(i0);
        pixel_column(i0);
        for (unsigned jo = 0; jo < END; jo += STEP) {
            VlTaps<NS> tp;
            double col[NS];
            sample(jo, tp);
            PMVS_FENCE_BLOCK();
            blend(tp, col);
            pixel(col, jo);
        }
        close_column();
    }
#endif
    fitOut = warp_sum(fit) * invV;
    swOut = warp_sum(sw);
}

/*
 * The view-lane loop for any number of views (V = 2..64) and for gradient weighting: GL = 4 lanes per window column, four
 * rows per trip, the non-reference views in chunks of four — lane s of a quad samples view 4c + s of chunk c (homography
 * row read from shared memory per chunk and trip: 9 loads and 3 fma for four samples, where the column-lane loop paid them
 * per sample), the quad transposes, and lane s then holds the four colours of chunk c at row s. They go into the
 * cross-view sum and into the warp's colour stash (shared memory, 32 B per lane and chunk); once every chunk is in, the mean
 * is known and a second walk over the STASH — not over the images — forms the absolute deviations (patch.cpp:1022-1027):
 * one sampling pass instead of the two of fitness_columns_many. GRAD: the per-pixel weight table of the reference window
 * (distance x gradient x mask) replaces the separable distance factors and the mask bits.
 */
__device__ __forceinline__ void sts_f64x2(unsigned a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory"); }
__device__ __forceinline__ double2 lds_f64x2_v(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}

template <bool GRAD>
__device__ __noinline__ void fitness_vl_many(const DevScene &S, const EvalCtx &E, const RefWin &R, const double *__restrict__ Hw,
                                             const double *__restrict__ sExpT, double *stash, double &fitOut, double &swOut) {
    const int lane = threadIdx.x & 31, s = lane & 3, ci = lane >> 2;
    const int V = E.V, NG = V - 1, NCH = (NG + 3) >> 2, refV = E.refView, nx = R.nx, nyp = R.nyp;
    const int ny = (R.ny + 3) & ~3;
    const unsigned END = 16u * (unsigned)ny;
    const unsigned hA = smem_addr(Hw), xsA = smem_addr(R.xs), gxA = smem_addr(R.gx), ysgA = smem_addr(R.ysg), mkA = smem_addr(R.mask);
    const unsigned rcA = smem_addr(R.refc), pwA = GRAD ? smem_addr(R.pw) : 0u, tabA = smem_addr(sExpT), viewA = smem_addr(E.view);
    const unsigned stA = smem_addr(stash) + 16u * (unsigned)lane;        /* chunk c: pair 0 at stA + 1024 c, pair 1 at stA + 1024 c + 512 */
    const double invV = 1.0 / (double)V;
    const double negK = S.cfg.adaptiveDifferenceEnable ? -(invV * invV) / S.cfg.diffWeighting : 0.0;
    unsigned ya[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) ya[t] = ysgA + 16u * (unsigned)(s ^ t);

    double fit = 0, sw = 0;
    for (int i0 = 0; i0 < nx; i0 += 8) {
        const int i = i0 + ci;
        const bool colOk = i < nx;
        const int ic = colOk ? i : nx - 1;
        const double x = lds_f64(xsA + 8u * ic);
        const double gxv = colOk ? (GRAD ? 1.0 : lds_f64(gxA + 8u * ic)) : 0.0;
        unsigned mlo = 0;
        if (!GRAD) mlo = (unsigned)(lds_u64(mkA + 8u * ic) >> (16 * s));
        unsigned pix = 8u * (unsigned)(ic * nyp + s);                      /* byte offset of (column, row s) in refc / pw */
        double cfit = 0, csw = 0;
        for (unsigned jo = 0; jo < END; jo += 64u) {
            double y[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) y[t] = lds_f64(ya[t] + jo);
            double sum = 0;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                int k = 4 * c + s;
                if (k >= NG) k = 0;                                         /* padding lane: samples view 0 again, discarded below */
                const int v = k + (k >= refV ? 1 : 0);
                const unsigned h = hA + 72u * (unsigned)v, va = viewA + (unsigned)(sizeof(ViewS) * v);
                const double h1 = lds_f64(h + 8u), h4 = lds_f64(h + 32u), h7 = lds_f64(h + 56u);
                const double A = fma(lds_f64(h), x, lds_f64(h + 16u)), B = fma(lds_f64(h + 24u), x, lds_f64(h + 40u)),
                             Cc = fma(lds_f64(h + 48u), x, lds_f64(h + 64u));
                const uint32_t *__restrict__ quad = (const uint32_t *)lds_u64(va + (unsigned)offsetof(ViewS, quad));
                const int cols = lds_s32(va + (unsigned)offsetof(ViewS, cols));
                double w[4], r[4], e[4], ix[4], fy[4];
                int px[4];
                uint32_t q[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) w[t] = fma(h7, y[t], Cc);
#pragma unroll
                for (int t = 0; t < 4; ++t) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[t]) : "d"(w[t]));
#pragma unroll
                for (int t = 0; t < 4; ++t) e[t] = fma(-w[t], r[t], 1.0);
#pragma unroll
                for (int t = 0; t < 4; ++t) e[t] = fma(e[t], e[t], e[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) r[t] = fma(r[t], e[t], r[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    ix[t] = fma(h1, y[t], A) * r[t];
                    fy[t] = fma(h4, y[t], B) * r[t];
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double tr = __dadd_rd(fy[t], PMVS_MAGIC_FLOOR);
                    px[t] = __double2int_rd(ix[t]);
                    q[t] = __ldg(quad + (__double2loint(tr) * cols + px[t]));
                    fy[t] = fy[t] - (tr - PMVS_MAGIC_FLOOR);
                }
                double col[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int ndx = dp4a_us(q[t], 0x0000ff01, 0), ndxy = dp4a_us(q[t], (int)0xff0101ffu, 0), idy = dp4a_us(q[t], 0x000100ff, 0);
                    const int g00 = (int)(q[t] & 0xffu);
                    const int k0 = px[t] * ndx + g00, k1 = px[t] * ndxy + idy;
                    col[t] = fma(fy[t], fma(-ix[t], (double)ndxy, (double)k1), fma(-ix[t], (double)ndx, (double)k0));
                }
#pragma unroll
                for (int t = 1; t < 4; ++t) col[t] = shfl_xor_f64(col[t], t);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (4 * c + (s ^ t) >= NG) col[t] = 0.0;
                sum += (col[0] + col[1]) + (col[2] + col[3]);
                sts_f64x2(stA + 1024u * (unsigned)c, col[0], col[1]);
                sts_f64x2(stA + 1024u * (unsigned)c + 512u, col[2], col[3]);
            }
            const double cref = lds_f64(rcA + pix);
            sum += cref;
            const double mean = sum * invV;
            double dev = 0;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                const double2 a = lds_f64x2_v(stA + 1024u * (unsigned)c), b = lds_f64x2_v(stA + 1024u * (unsigned)c + 512u);
                double d0 = fabs(a.x - mean), d1 = fabs(a.y - mean), d2 = fabs(b.x - mean), d3 = fabs(b.y - mean);
                if (4 * c + 4 > NG) {                                        /* last chunk: padding views do not count */
                    if (4 * c + (s ^ 0) >= NG) d0 = 0.0;
                    if (4 * c + (s ^ 1) >= NG) d1 = 0.0;
                    if (4 * c + (s ^ 2) >= NG) d2 = 0.0;
                    if (4 * c + (s ^ 3) >= NG) d3 = 0.0;
                }
                dev += (d0 + d1) + (d2 + d3);
            }
            dev += fabs(cref - mean);
            double wgt;
            if (GRAD) wgt = lds_f64(pwA + pix);                                              /* patch.cpp:1030-1032, :1036-1038, :986 */
            else {
                wgt = 0.0;
                if (mlo & 1u) wgt = lds_f64(ya[0] + jo + 8u);
                mlo >>= 1;
            }
            wgt *= exp_table_c(dev * dev * negK, tabA);                                      /* patch.cpp:1033-1035 */
            csw += wgt;
            cfit = fma(wgt, dev, cfit);
            pix += 32u;
        }
        sw = fma(gxv, csw, sw);
        fit = fma(gxv, cfit, fit);
    }
    fitOut = warp_sum(fit) * invV;
    swOut = warp_sum(sw);
}

/*
 * Many views without a colour stash (fitness_vl_rows): a group of GLN lanes (a quad below) owns one window pixel at a time (32 / GLN columns per pass,
 * rows one after the other); lane s samples the views {4c + s} — at most 4 G4 of them — and keeps THEIR colours in
 * registers. The cross-view sum and the sum of absolute deviations are quad all-reduces (two butterfly shuffles each; the
 * pairwise order (l0 + l1) + (l2 + l3) is the same on every lane), so nothing has to be transposed or stored, one sampling
 * pass, and the shared memory stays what the scene's tables need (two CTAs per SM at 64 views). The per-pixel tail
 * (weights, exp, accumulation) runs on every lane of the group (lane 0 accumulates) or, at 8 lanes per pixel, once per block of
 * 8 rows with lane s taking row s of the block.
 */
template <int GLN, int G4, bool GRAD>
__device__ __noinline__ void fitness_vl_rows(const DevScene &S, const EvalCtx &E, const RefWin &R, const double *__restrict__ Hw,
                                             const double *__restrict__ sExpT, double &fitOut, double &swOut) {
    constexpr int NV = 4 * G4;                       /* views of one lane */
    constexpr bool SPLIT = GLN >= 8;                 /* per-pixel tail once per block of GLN rows instead of on every lane */
    constexpr int LGN = GLN == 8 ? 3 : (GLN == 4 ? 2 : (GLN == 2 ? 1 : 0));
    const int lane = threadIdx.x & 31, s = lane & (GLN - 1), ci = lane >> LGN;
    const int V = E.V, NG = V - 1, refV = E.refView, nx = R.nx, ny = R.ny, nyp = R.nyp;
    const unsigned hA = smem_addr(Hw), xsA = smem_addr(R.xs), gxA = smem_addr(R.gx), ysgA = smem_addr(R.ysg), mkA = smem_addr(R.mask);
    const unsigned rcA = smem_addr(R.refc), pwA = GRAD ? smem_addr(R.pw) : 0u, tabA = smem_addr(sExpT), viewA = smem_addr(E.view);
    const double invV = 1.0 / (double)V;
    const double negK = S.cfg.adaptiveDifferenceEnable ? -(invV * invV) / S.cfg.diffWeighting : 0.0;
    /* this lane's views are {4n + s}: table addresses (homography row, view record) are formed per sample — arrays of them
     * would cost 2 registers per view; padding entries (4n + s >= NG) sample view 0 again and are masked out */
    const int nReal = NG > s ? (NG - s + GLN - 1) >> LGN : 0;   /* n < nReal  <=>  GLN n + s < NG */
    const bool never = S.cfg.patchRadius < 0;
    double fit = 0, sw = 0;
    for (int i0 = 0; i0 < nx; i0 += 32 / GLN) {
        const int i = i0 + ci;
        const bool colOk = i < nx;
        const int ic = colOk ? i : nx - 1;
        const double x = lds_f64(xsA + 8u * ic);
        /* SPLIT: every lane accumulates the rows it ran the tail for; else lane 0 of the group accumulates */
        const double gxv = (colOk && (SPLIT || s == 0)) ? (GRAD ? 1.0 : lds_f64(gxA + 8u * ic)) : 0.0;
        const unsigned long long mk = GRAD ? 0ull : lds_u64(mkA + 8u * ic);
        double cfit = 0, csw = 0, mydev = 0;
        for (int j = 0; j < ny; ++j) {
            const double y = lds_f64(ysgA + 16u * j);
            double col[NV];
#pragma unroll
            for (int g = 0; g < G4; ++g) {
                double w[4], r[4], e[4], ix[4], fy[4];
                int px[4];
                uint32_t q[4];
                unsigned hv[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int n = 4 * g + t;
                    const int k = n < nReal ? GLN * n + s : 0;
                    hv[t] = (unsigned)(k + (k >= refV ? 1 : 0));     /* view index (the reference view is skipped) */
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const unsigned h = hA + 72u * hv[t];
                    w[t] = fma(lds_f64_v(h + 56u), y, fma(lds_f64_v(h + 48u), x, lds_f64_v(h + 64u)));      /* volatile: a plain load is row-invariant and would be hoisted out of the row loop into 9 registers per view */
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[t]) : "d"(w[t]));
#pragma unroll
                for (int t = 0; t < 4; ++t) e[t] = fma(-w[t], r[t], 1.0);
#pragma unroll
                for (int t = 0; t < 4; ++t) e[t] = fma(e[t], e[t], e[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) r[t] = fma(r[t], e[t], r[t]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const unsigned h = hA + 72u * hv[t];
                    ix[t] = fma(lds_f64_v(h + 8u), y, fma(lds_f64_v(h), x, lds_f64_v(h + 16u))) * r[t];
                    fy[t] = fma(lds_f64_v(h + 32u), y, fma(lds_f64_v(h + 24u), x, lds_f64_v(h + 40u))) * r[t];
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const unsigned va = viewA + (unsigned)sizeof(ViewS) * hv[t];
                    const uint32_t *__restrict__ quad = (const uint32_t *)lds_u64(va + (unsigned)offsetof(ViewS, quad));
                    const int cols = lds_s32(va + (unsigned)offsetof(ViewS, cols));
                    const double tr = __dadd_rd(fy[t], PMVS_MAGIC_FLOOR);
                    px[t] = __double2int_rd(ix[t]);
                    q[t] = __ldg(quad + (__double2loint(tr) * cols + px[t]));
                    fy[t] = fy[t] - (tr - PMVS_MAGIC_FLOOR);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int ndx = dp4a_us(q[t], 0x0000ff01, 0), ndxy = dp4a_us(q[t], (int)0xff0101ffu, 0), idy = dp4a_us(q[t], 0x000100ff, 0);
                    const int g00 = (int)(q[t] & 0xffu);
                    const int k0 = px[t] * ndx + g00, k1 = px[t] * ndxy + idy;
                    const double cc = fma(fy[t], fma(-ix[t], (double)ndxy, (double)k1), fma(-ix[t], (double)ndx, (double)k0));
                    col[4 * g + t] = (4 * g + t) < nReal ? cc : 0.0;
                }
                /* a never-taken branch = a basic-block boundary: ptxas would otherwise interleave all 4 G4 samples and spill */
                if (G4 > 2 && never) col[4 * g] = -col[4 * g];
            }
            double sum = tree_sum<NV>(col);
            if (GLN > 1) sum += shfl_xor_f64(sum, 1);
            if (GLN > 2) sum += shfl_xor_f64(sum, 2);
            if (GLN > 4) sum += shfl_xor_f64(sum, 4);
            const unsigned pix = 8u * (unsigned)(ic * nyp + j);
            const double cref = lds_f64(rcA + pix);
            sum += cref;
            const double mean = sum * invV;
#pragma unroll
            for (int n = 0; n < NV; ++n) col[n] = n < nReal ? fabs(col[n] - mean) : 0.0;
            double dev = tree_sum<NV>(col);
            if (GLN > 1) dev += shfl_xor_f64(dev, 1);
            if (GLN > 2) dev += shfl_xor_f64(dev, 2);
            if (GLN > 4) dev += shfl_xor_f64(dev, 4);
            dev += fabs(cref - mean);
            /* The per-pixel tail (weights, exp, accumulation) is the same on all GLN lanes of the group. SPLIT (8 lanes per pixel):
             * instead of computing it GLN times, lane s keeps the deviation of row (block of GLN rows) + s and runs the tail once
             * per block for ITS row. Measured (profiles/r2_ab_runs.txt, run 10): +1.6 % at 8 lanes (config 5), -2.1 % at 2 lanes
             * (config 4: the extra live state spills), so the narrower groups keep the tail on every lane, lane 0 accumulating. */
            if (SPLIT) {
                if ((j & (GLN - 1)) == s) mydev = dev;
                if ((j & (GLN - 1)) != GLN - 1 && j != ny - 1) continue;
            } else mydev = dev;
            const int jr = SPLIT ? (j & ~(GLN - 1)) + s : j;
            double wgt = 0.0;
            if (!SPLIT || jr < ny) {
                if (GRAD) wgt = lds_f64(pwA + 8u * (unsigned)(ic * nyp + jr));                   /* patch.cpp:1030-1032, :1036-1038, :986 */
                else if ((mk >> (16 * (jr & 3) + (jr >> 2))) & 1ull) wgt = lds_f64(ysgA + 16u * jr + 8u);
            }
            wgt *= exp_table_c(mydev * mydev * negK, tabA);                                      /* patch.cpp:1033-1035 */
            csw += wgt;
            cfit = fma(wgt, mydev, cfit);
        }
        sw = fma(gxv, csw, sw);
        fit = fma(gxv, cfit, fit);
    }
    fitOut = warp_sum(fit) * invV;
    swOut = warp_sum(sw);
}

#if PMVS_TILE
/* Stage the tiles of the current swarm (CTA-collective; warp 0's homography area Hc is free before the swarm starts):
 * origins from the canonical hypothesis' window centre in each view, rows copied by cp.async.bulk (one bulk copy of
 * 4 T bytes per tile row, the TMA unit), completion on one mbarrier. */
__device__ __forceinline__ void stage_tiles(const DevScene &S, const EvalCtx &E, RefWin &R, const double *center, const double *normal, double *Hc) {
    const int tid = threadIdx.x, T = PMVS_TILE;
    if (tid == 0) R.tileOk = 0;
    if (!R.ok || E.V != 5 || E.refView < 0) { __syncthreads(); return; }
    if (tid < 32) {
        warp_homographies(E, center, normal, Hc);
        if (tid < 4) {
            const int v = tid + (tid >= E.refView ? 1 : 0);
            const double *H = Hc + 9 * v;
            const double x = R.pt[0], y = R.pt[1];
            const double w = H[6] * x + H[7] * y + H[8];
            int ox = ((int)floor((H[0] * x + H[1] * y + H[2]) / w) - T / 2) & ~3, oy = (int)floor((H[3] * x + H[4] * y + H[5]) / w) - T / 2;
            const ViewS &vw = E.view[v];
            ox = ox < 0 ? 0 : (ox > ((vw.cols - T) & ~3) ? ((vw.cols - T) & ~3) : ox);
            oy = oy < 0 ? 0 : (oy > vw.rows - T ? vw.rows - T : oy);
            R.tox[tid] = ox;
            R.toy[tid] = oy;
            if (vw.cols < T + 4 || vw.rows < T || (vw.cols & 3)) atomicExch(&R.tileOk, -1);      /* bulk copies need 16-byte aligned rows */
        }
    }
    __syncthreads();
    if (R.tileOk < 0) { if (tid == 0) R.tileOk = 0; __syncthreads(); return; }
    const unsigned mbar = smem_addr(&R.mbar);
    if (tid == 0) {
        if (R.tilePhase < 0) {       /* first use by this CTA */
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
            R.tilePhase = 0;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       /* earlier generic-proxy reads of the tiles / the init before the async writes */
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(4 * T * T * 4) : "memory");
    }
    __syncthreads();
    for (int r = tid; r < 4 * T; r += blockDim.x) {
        const int k = r / T, row = r - k * T, v = k + (k >= E.refView ? 1 : 0);
        const ViewS &vw = E.view[v];
        const uint32_t *src = vw.quad + ((size_t)(R.toy[k] + row) * vw.cols + R.tox[k]);
        const unsigned dst = smem_addr(R.tile) + 4u * (unsigned)(k * T * T + row * T);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(4 * T), "r"(mbar)
                     : "memory");
    }
    const int parity = R.tilePhase & 1;
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    __syncthreads();
    if (tid == 0) { R.tilePhase ^= 1; R.tileOk = 1; }
    __syncthreads();
}
#endif

/* dispatch on the number of non-reference views NG = 1..8; false: not covered */
__device__ __forceinline__ bool fitness_vl_dispatch(const DevScene &S, const EvalCtx &E, const RefWin &R, const double *Hw, const double *sExpT,
                                                    double *stash, double &fit, double &sw, bool tileHit = false) {
    const int NGv = E.V - 1;
    if (R.pw || NGv > PMVS_VL_VIEWS - 1) {    /* gradient weighting (per-pixel weights) or more views than the register form holds */
#if PMVS_VL_MANY_STASH
        if (stash) {
            if (R.pw) fitness_vl_many<true>(S, E, R, Hw, sExpT, stash, fit, sw);
            else fitness_vl_many<false>(S, E, R, Hw, sExpT, stash, fit, sw);
            return true;
        }
#endif
        /* lanes per pixel: few views -> fewer lanes share a pixel (the per-pixel tail is computed by all of them) */
        const int forced = (S.useVL >> 4) & 15;
        /* measured (profiles/r2_ab_runs.txt, run 8): the FEWEST lanes per pixel that keep a lane's colours in registers win —
         * lanes of a group sample different views, i.e. different images, so every extra lane per pixel splits the warp's tap
         * load into more 32-byte sectors (config 3: 14.5 / 12.8 / 8.9 k patches/s at 1 / 2 / 4 lanes, config 4: 18.9 / 16.2 /
         * 11.9 k at 2 / 4 / 8); above 48 views 8 lanes x 8 views (no spills) edge out 4 x 16 (config 5: 3.39 vs 3.33 k) */
        const int gln = forced ? forced : (NGv <= 16 ? 1 : (NGv <= 32 ? 2 : (NGv <= 48 ? 4 : 8)));
#define PMVS_ROWS(GLN_, G4_) do { if (R.pw) fitness_vl_rows<GLN_, G4_, true>(S, E, R, Hw, sExpT, fit, sw); else fitness_vl_rows<GLN_, G4_, false>(S, E, R, Hw, sExpT, fit, sw); } while (0)
        if (gln == 1) {
            if (NGv <= 4) PMVS_ROWS(1, 1);
            else if (NGv <= 8) PMVS_ROWS(1, 2);
            else if (NGv <= 12) PMVS_ROWS(1, 3);
            else if (NGv <= 16) PMVS_ROWS(1, 4);
            else PMVS_ROWS(4, 4);
        } else if (gln == 2) {
            if (NGv <= 8) PMVS_ROWS(2, 1);
            else if (NGv <= 16) PMVS_ROWS(2, 2);
            else if (NGv <= 24) PMVS_ROWS(2, 3);
            else if (NGv <= 32) PMVS_ROWS(2, 4);
            else PMVS_ROWS(4, 4);
        } else if (gln == 4) {
            if (NGv <= 16) PMVS_ROWS(4, 1);
            else if (NGv <= 32) PMVS_ROWS(4, 2);
            else if (NGv <= 48) PMVS_ROWS(4, 3);
            else PMVS_ROWS(4, 4);
        } else {
            if (NGv <= 32) PMVS_ROWS(8, 1);
            else PMVS_ROWS(8, 2);
        }
#undef PMVS_ROWS
        return true;
    }
    switch (E.V - 1) {
    case 1: fitness_vl<1, 1, 8, true>(S, E, R, Hw, sExpT, fit, sw); return true;
    case 2: fitness_vl<2, 1, 4, true>(S, E, R, Hw, sExpT, fit, sw); return true;
    case 3: fitness_vl<4, 1, PMVS_VL_RB, false>(S, E, R, Hw, sExpT, fit, sw); return true;
    case 4:
#if PMVS_TILE
        if (tileHit) { fitness_vl<4, 1, PMVS_VL_RB, true, true>(S, E, R, Hw, sExpT, fit, sw); return true; }
#endif
        fitness_vl<4, 1, PMVS_VL_RB, true>(S, E, R, Hw, sExpT, fit, sw); return true;
    case 5: case 6: case 7: fitness_vl<4, 2, 1, false>(S, E, R, Hw, sExpT, fit, sw); return true;
    case 8: fitness_vl<4, 2, 1, true>(S, E, R, Hw, sExpT, fit, sw); return true;
    default: return false;
    }
}
/* configurations the view-lane loop covers: difference-weight exponent provably in [-700, 0] (|deviation| <= 255 per view) */
__device__ __forceinline__ bool vl_config_ok(const DevScene &S) {
    return S.useVL && (!S.cfg.adaptiveDifferenceEnable || -(255.0 * 255.0) / S.cfg.diffWeighting >= -700.0);
}

/*
 * Build the patch's reference window for the hypothesis centre `center` (any point of the viewing ray). Collective over
 * `nthreads` threads with index tid (a whole CTA: WARP = false, barriers are __syncthreads; one warp: WARP = true).
 * R.ok = 0 when the fast path cannot be used for this swarm run (window test of patch.cpp:951-962 fails at the canonical
 * point, window sample count differs from patchSize^2, reference camera not among the views).
 */
template <bool WARP>
__device__ __forceinline__ void ref_win_sync() {
    if (WARP) __syncwarp();
    else __syncthreads();
}
template <bool WARP>
__device__ __forceinline__ void build_ref_win(const DevScene &S, const EvalCtx &E, RefWin &R, const double *center, int tid, int nthreads) {
    const int radius = S.cfg.patchRadius, ps = S.cfg.patchSize;
    if (tid == 0) {
        R.ok = 0;
        R.nx = R.ny = 0;
        if (E.valid && E.refView >= 0 && E.V >= 2 && E.V <= PMVS_MAX_VIEWS && vl_config_ok(S) &&
            (!S.cfg.adaptiveGradientEnable || (R.pw != nullptr && E.refEdge != nullptr))) {
            double pt[2];
            const double c[3] = {center[0], center[1], center[2]};
            if (ref_window_ok(S, E, c, pt)) {
                R.pt[0] = pt[0];
                R.pt[1] = pt[1];
                R.ok = 1;
            }
        }
    }
    ref_win_sync<WARP>();
    if (!R.ok) return;                                   /* uniform */
    if (tid < 32) {
        /* axes: xs directly, ys through the gx area (filled with the weights afterwards) */
        const double pt[2] = {R.pt[0], R.pt[1]};
        const int nxy = warp_window_axes(pt, radius, ps, R.xs, R.gx);
        const int nx = nxy & 0xffff, ny = nxy >> 16;
        if (nx == ps && ny == ps) {
            for (int j = tid; j < PMVS_PS_PAD8(ps); j += 32)
                R.ysg[j] = j < ny ? make_double2(R.gx[j], __ldg(S.distG + j)) : make_double2(R.gx[ny - 1], 0.0);
            __syncwarp();
            for (int i = tid; i < nx; i += 32) {
                R.gx[i] = __ldg(S.distG + i);
                R.mask[i] = 0ull;
            }
        }
        if (tid == 0) {
            R.nx = nx;
            R.ny = ny;
            R.gl = E.V == 2 ? 1 : (E.V == 3 ? 2 : 4);
            if (nx != ps || ny != ps) R.ok = 0;          /* a rounding step of the ++x recurrence dropped a sample: rare, slow path */
        }
    }
    ref_win_sync<WARP>();
    if (!R.ok) return;
    const int nx = R.nx, ny = R.ny, nyPad = PMVS_PS_PAD8(ps), refCols = E.refCols, gl = R.gl, fw = 64 / gl;
    const uint32_t *__restrict__ refQuad = E.refQuad;
    for (int idx = tid; idx < nx * nyPad; idx += nthreads) {
        const int i = idx / nyPad, j = idx - i * nyPad;
        if (j >= ny) {
            R.refc[i * R.nyp + j] = 0.0;
            continue;
        }
        const double x = R.xs[i], y = R.ysg[j].x;
        const double tx = __dadd_rd(x, PMVS_MAGIC_FLOOR), ty = __dadd_rd(y, PMVS_MAGIC_FLOOR);
        const int px = __double2loint(tx), py = __double2loint(ty);
        RefColumn rc;
        rc.quad = refQuad + px;
        rc.fx = x - (tx - PMVS_MAGIC_FLOOR);
        rc.selx = __double2int_rn(x) - px;
        bool keep;
        const double c = ref_sample(rc, make_int2(py * refCols, 2 * (__double2int_rn(y) - py)), y - (ty - PMVS_MAGIC_FLOOR), keep);
        R.refc[i * R.nyp + j] = c;
        if (keep) atomicOr(R.mask + i, 1ull << (fw * (j & (gl - 1)) + j / gl));
        if (R.pw) {      /* distance x gradient weight of this pixel, zero on background (patch.cpp:1030-1032, :1036-1038, :986) */
            const size_t rofs = (size_t)__double2int_rn(y) * refCols + __double2int_rn(x);
            double wgt = R.gx[i] * R.ysg[j].y;
            wgt *= exp(-1.0 / (__ldg(E.refEdge + rofs) * S.cfg.gradientWeighting));
            R.pw[i * R.nyp + j] = keep ? wgt : 0.0;
        }
    }
    if (R.pw)
        for (int idx = tid; idx < nx * (nyPad - ny); idx += nthreads) {
            const int i = idx / (nyPad - ny), j = ny + idx % (nyPad - ny);
            R.pw[i * R.nyp + j] = 0.0;
        }
    ref_win_sync<WARP>();
}

template <int V>
__device__ __forceinline__ bool fitness_columns_dispatch(int nV, const DevScene &S, const EvalCtx &E, const double *sDistW,
                                                         const double *sExpT, const WarpWork &W, int nx, int ny, double &fit,
                                                         double &sw) {
    if constexpr (V >= 2) {
        if (nV == V) {
            /* |deviation| <= 255 per view bounds the exponent of the difference weight */
            const double xmin = -(255.0 * 255.0) / S.cfg.diffWeighting;
            const bool fast = !S.cfg.adaptiveGradientEnable && S.cfg.adaptiveDistanceEnable && (!S.cfg.adaptiveDifferenceEnable || xmin >= -700.0);
            /* per-lane slots when the scene keeps a slot area that holds this patch's views, else x-parts inline */
            if constexpr (V - 1 <= PMVS_SLOT_VIEWS) {
                if (V - 1 <= W.slotViews) {
                    if (fast) fitness_columns<V, true, true>(S, E, sDistW, sExpT, W, nx, ny, fit, sw);
                    else fitness_columns<V, false, true>(S, E, sDistW, sExpT, W, nx, ny, fit, sw);
                    return true;
                }
            }
            if (fast) fitness_columns<V, true, false>(S, E, sDistW, sExpT, W, nx, ny, fit, sw);
            else fitness_columns<V, false, false>(S, E, sDistW, sExpT, W, nx, ny, fit, sw);
            return true;
        }
        return fitness_columns_dispatch<V - 1>(nV, S, E, sDistW, sExpT, W, nx, ny, fit, sw);
    } else {
        return false;
    }
}

/*
 * PAIS::getFitness (patch.cpp:914-1047) for the hypothesis (theta, phi, depth); one warp, result in every lane.
 * Hypotheses whose window corners all project at least PMVS_EDGE_EPS inside every view take the unchecked loop
 * (a projective map with w > 0 keeps the window inside the convex hull of its corners); anything else takes the
 * loop with the reference's per-sample test, so the DBL_MAX sentinel is reproduced exactly.
 */
#define PMVS_EDGE_EPS 1e-6
/* everything of getFitness after the homographies and the reference projection: window axes, corner test, sample loop */
template <int VCAP>
__device__ __noinline__ double warp_window(const DevScene &S, const EvalCtx &E, const double *sDistW, const WarpWork &W, const double *pt) {
    const int lane = threadIdx.x & 31;
    const int radius = S.cfg.patchRadius, ps = S.cfg.patchSize;
    const int nxy = warp_window_axes(pt, radius, ps, W.xs, W.ys);
    const int nx = nxy & 0xffff, ny = nxy >> 16;

    /* corner test, lane = (view, corner) */
    bool inside = true;
    for (int t = lane; t < 4 * E.V; t += 32) {
        const int v = t >> 2, cidx = t & 3;
        const double *H = W.H + 9 * v;
        const ViewS &vw = E.view[v];
        const double loX = 2.0 + PMVS_EDGE_EPS, hiX = (double)(vw.cols - 3) - PMVS_EDGE_EPS;
        const double loY = 2.0 + PMVS_EDGE_EPS, hiY = (double)(vw.rows - 3) - PMVS_EDGE_EPS;
        const double x = W.xs[(cidx & 1) ? nx - 1 : 0], y = W.ys[(cidx & 2) ? ny - 1 : 0];
        /* w > 0: lo <= n/w < hi  <=>  lo*w <= n < hi*w; the 1e-6 margin dwarfs the rounding of the products */
        const double w = H[6] * x + H[7] * y + H[8];
        const double nxw = H[0] * x + H[1] * y + H[2], nyw = H[3] * x + H[4] * y + H[5];
        if (!(w > 0.0 && nxw >= loX * w && nxw < hiX * w && nyw >= loY * w && nyw < hiY * w)) inside = false;
    }
    inside = __all_sync(PMVS_FULL, inside);
    double fit, sw;
    bool ok;
    if (inside) {
        ok = true;
        /* lane-per-column loop: every V in 2..16 has its own instantiation, reached from the VCAP = 8 / 16 entry */
        if (VCAP == 8 && W.gv && nx > 0 && E.refView >= 0 && fitness_columns_dispatch<8>(E.V, S, E, sDistW, sDistW + PMVS_DIST_PAD(ps), W, nx, ny, fit, sw)) {}
        else if (VCAP == 16 && W.gv && nx > 0 && E.refView >= 0 && E.V > 8 && fitness_columns_dispatch<16>(E.V, S, E, sDistW, sDistW + PMVS_DIST_PAD(ps), W, nx, ny, fit, sw)) {}
        else if (VCAP == 0 && W.gv && nx > 0 && E.refView >= 0) fitness_columns_many(S, E, sDistW, sDistW + PMVS_DIST_PAD(ps), W, nx, ny, fit, sw);
        else ok = fitness_samples<VCAP, false>(S, E, sDistW, W.H, W.xs, W.ys, nx, ny, fit, sw);
    } else {
        /* Exact early-out. The four window corners are window samples: if the reference pixel of one is not masked
         * (patch.cpp:986) and some view's sample there fails the bounds test (:999), the reference returns DBL_MAX
         * whatever the other pixels hold — and so would the checked loop below, after walking the whole window at a
         * third of the unchecked loop's speed. Near the image borders most failed corner tests end here (the host
         * driver's late expansion rounds spent a fifth of their time, and most of their barrier waits, on such walks).
         * Same sample_view<true> and same (x, y) as the loop: the decision is the loop's own. */
        bool hit = false;
        for (int t = lane; t < 4 * E.V; t += 32) {
            const int v = t >> 2, cidx = t & 3;
            const double x = W.xs[(cidx & 1) ? nx - 1 : 0], y = W.ys[(cidx & 2) ? ny - 1 : 0];
            const size_t rofs = (size_t)__double2int_rn(y) * E.refCols + __double2int_rn(x);
            double c;
            if ((__ldg(E.refQuad + rofs) & 0xff) != 0 && !sample_view<true>(W.H + 9 * v, E.view[v], x, y, c)) hit = true;
        }
        if (__any_sync(PMVS_FULL, hit)) ok = false;
        else ok = fitness_samples<VCAP, true>(S, E, sDistW, W.H, W.xs, W.ys, nx, ny, fit, sw);
    }
    __syncwarp();
    if (!ok) return DBL_MAX;                                                          /* :999-1002 */
    return fit / sw;                                                                  /* :1046 */
}
/* the view-lane route: the hypothesis sits on the patch's reference window and all four window corners project inside
 * every view (same test as warp_window). Returns false when the hypothesis has to take the per-hypothesis path. */
__device__ __noinline__ bool warp_window_vl(const DevScene &S, const EvalCtx &E, const double *sExpT, const WarpWork &W, const double *pt,
                                            double &result) {
    const RefWin &R = *W.rw;
    if (!(fabs(pt[0] - R.pt[0]) <= PMVS_PT_TOL && fabs(pt[1] - R.pt[1]) <= PMVS_PT_TOL)) return false;
    const int lane = threadIdx.x & 31;
    const int nx = R.nx, ny = R.ny;
    bool inside = true;
    for (int t = lane; t < 4 * E.V; t += 32) {
        const int v = t >> 2, cidx = t & 3;
        const double *H = W.H + 9 * v;
        const ViewS &vw = E.view[v];
        const double loX = 2.0 + PMVS_EDGE_EPS, hiX = (double)(vw.cols - 3) - PMVS_EDGE_EPS;
        const double loY = 2.0 + PMVS_EDGE_EPS, hiY = (double)(vw.rows - 3) - PMVS_EDGE_EPS;
        const double x = R.xs[(cidx & 1) ? nx - 1 : 0], y = R.ysg[(cidx & 2) ? ny - 1 : 0].x;
        const double w = H[6] * x + H[7] * y + H[8];
        const double nxw = H[0] * x + H[1] * y + H[2], nyw = H[3] * x + H[4] * y + H[5];
        if (!(w > 0.0 && nxw >= loX * w && nxw < hiX * w && nyw >= loY * w && nyw < hiY * w)) inside = false;
    }
    if (!__all_sync(PMVS_FULL, inside)) return false;
#if PMVS_FOOTPRINT
    if (E.V <= 8) {
        RefWin &Rw = const_cast<RefWin &>(R);
        for (int t = lane; t < 4 * E.V; t += 32) {
            const int v = t >> 2, cidx = t & 3;
            const double *H = W.H + 9 * v;
            const double x = R.xs[(cidx & 1) ? nx - 1 : 0], y = R.ysg[(cidx & 2) ? ny - 1 : 0].x;
            const double w = H[6] * x + H[7] * y + H[8];
            const int px = (int)floor((H[0] * x + H[1] * y + H[2]) / w), py = (int)floor((H[3] * x + H[4] * y + H[5]) / w);
            atomicMin(&Rw.foot[v][0], px);
            atomicMin(&Rw.foot[v][1], py);
            atomicMax(&Rw.foot[v][2], px + 1);
            atomicMax(&Rw.foot[v][3], py + 1);
        }
        __syncwarp();
        /* first evaluation of the swarm defines the tile centres */
        int first = 0;
        if (lane == 0) first = atomicCAS(&Rw.footn[5], 0, 1) == 0;
        first = __shfl_sync(PMVS_FULL, first, 0);
        int cxs = 0, cys = 0;
        {
            const int v = lane < E.V ? lane : 0;
            const double *H = W.H + 9 * v;
            const double x = R.xs[nx / 2], y = R.ysg[ny / 2].x;
            const double w = H[6] * x + H[7] * y + H[8];
            cxs = (int)floor((H[0] * x + H[1] * y + H[2]) / w);
            cys = (int)floor((H[3] * x + H[4] * y + H[5]) / w);
            if (first && lane < E.V) { Rw.footc[lane][0] = cxs; Rw.footc[lane][1] = cys; atomicExch(&Rw.footn[6], 1); }
        }
        if (first) __threadfence_block();
        __syncwarp();
        if (*(volatile int *)&Rw.footn[6]) {
            bool in48 = true, in64 = true, in96 = true, in128 = true;
            for (int t = lane; t < 4 * E.V; t += 32) {
                const int v = t >> 2, cidx = t & 3;
                const double *H = W.H + 9 * v;
                const double x = R.xs[(cidx & 1) ? nx - 1 : 0], y = R.ysg[(cidx & 2) ? ny - 1 : 0].x;
                const double w = H[6] * x + H[7] * y + H[8];
                const int px = (int)floor((H[0] * x + H[1] * y + H[2]) / w), py = (int)floor((H[3] * x + H[4] * y + H[5]) / w);
                const int dx = px - *(volatile int *)&Rw.footc[v][0], dy = py - *(volatile int *)&Rw.footc[v][1];
                const int m = max(max(dx, -dx - 1), max(dy, -dy - 1)) + 1;      /* half-size needed incl. the +1 tap */
                in48 = in48 && m <= 24; in64 = in64 && m <= 32; in96 = in96 && m <= 48; in128 = in128 && m <= 64;
            }
            in48 = __all_sync(PMVS_FULL, in48); in64 = __all_sync(PMVS_FULL, in64); in96 = __all_sync(PMVS_FULL, in96); in128 = __all_sync(PMVS_FULL, in128);
            if (lane == 0) {
                atomicAdd(&Rw.footn[0], 1);
                if (in48) atomicAdd(&Rw.footn[1], 1);
                if (in64) atomicAdd(&Rw.footn[2], 1);
                if (in96) atomicAdd(&Rw.footn[3], 1);
                if (in128) atomicAdd(&Rw.footn[4], 1);
            }
        }
    }
#endif
    bool tileHit = false;
#if PMVS_TILE
    if (R.tileOk && E.V == 5) {          /* every window corner of every non-reference view inside that view's tile */
        bool in = true;
        for (int t = lane; t < 4 * E.V; t += 32) {
            const int v = t >> 2, cidx = t & 3;
            if (v == E.refView) continue;
            const int k = v - (v > E.refView ? 1 : 0);
            const double *H = W.H + 9 * v;
            const double x = R.xs[(cidx & 1) ? nx - 1 : 0], y = R.ysg[(cidx & 2) ? ny - 1 : 0].x;
            const double w = H[6] * x + H[7] * y + H[8];
            const double ix = (H[0] * x + H[1] * y + H[2]) / w - R.tox[k], iy = (H[3] * x + H[4] * y + H[5]) / w - R.toy[k];
            if (!(ix >= 1.0 && ix < PMVS_TILE - 1.0 && iy >= 1.0 && iy < PMVS_TILE - 1.0)) in = false;
        }
        tileHit = __all_sync(PMVS_FULL, in);
    }
#endif
    double fit, sw;
    if (!fitness_vl_dispatch(S, E, R, W.H, sExpT, W.stash, fit, sw, tileHit)) return false;
    __syncwarp();
    result = fit / sw;                                                                /* patch.cpp:1046 */
    return true;
}
__device__ __forceinline__ double warp_window_any(const DevScene &S, const EvalCtx &E, const double *sDistW, const WarpWork &W, const double *pt) {
    if (W.rw && W.rw->ok) {
        double f;
        if (warp_window_vl(S, E, sDistW + PMVS_DIST_PAD(S.cfg.patchSize), W, pt, f)) return f;
    }
    if (E.V <= 8) return warp_window<8>(S, E, sDistW, W, pt);
    if (E.V <= 16) return warp_window<16>(S, E, sDistW, W, pt);
    return warp_window<0>(S, E, sDistW, W, pt);
}

/* the reference-image position of the hypothesis and the tests of patch.cpp:951-962; false = DBL_MAX */
__device__ __forceinline__ bool ref_window_ok(const DevScene &S, const EvalCtx &E, const double *center, double *pt) {
    const int radius = S.cfg.patchRadius;
    project_pt(E.refR, E.refT, E.refFocal, E.refPP, E.sc, center, pt);                /* :951-954 */
    if (!in_image(pt[0], pt[1], E.refCols, E.refRows)) return false;
    if (pt[0] - radius < 2 || pt[0] + radius >= E.refCols - 3 || pt[1] - radius < 2 || pt[1] + radius >= E.refRows - 3)
        return false;                                                                 /* :957-962 */
    return true;
}

/*
 * PAIS::getFitness (patch.cpp:914-1047) for the hypothesis (theta, phi, depth); one warp, result in every lane.
 * Hypotheses whose window corners all project at least PMVS_EDGE_EPS inside every view take the unchecked loop
 * (a projective map with w > 0 keeps the window inside the convex hull of its corners); anything else takes the
 * loop with the reference's per-sample test, so the DBL_MAX sentinel is reproduced exactly.
 */
__device__ __noinline__ double warp_fitness_any(const DevScene &S, const EvalCtx &E, const double *sDistW, const WarpWork &W,
                                                double theta, double phi, double depth) {
    double n[3];
    spherical2Normal(theta, phi, n);                                                  /* :935-936 */
    if (dot3(n, E.refOptN) > 0) return DBL_MAX;                                       /* :939-941 */
    double center[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) center[k] = E.ray[k] * depth + E.refC[k];             /* :944 */
    if (!E.valid) return DBL_MAX;
    __syncwarp();
    warp_homographies(E, center, n, W.H);                                             /* :947-948 */
    double pt[2];
    if (!ref_window_ok(S, E, center, pt)) return DBL_MAX;
    return warp_window_any(S, E, sDistW, W, pt);
}

/*
 * The same for m <= PMVS_EVAL_BATCH hypotheses of one patch (the particles a warp evaluates in one generation), with
 * the scalar part — normal, plane, homographies, reference projection — done ONCE for all of them: lane = (hypothesis,
 * view), m * V <= 32, instead of m passes that each keep V lanes busy. Per lane the arithmetic is the single-hypothesis
 * path's, so results are bit-identical. Homographies land in W.H + 9 V k; the window part then runs per hypothesis.
 */
struct ParticleS {   /* pso/particle.h:5-28 */
    double pos[3], vec[3], pBest[3], nBest[3], fitness, pbf, _r0, _r1;
};
struct HypoS {
    double pt[2];
    int run, _pad;
};
/* hypotheses = positions of the particles part[p0 + k * stride], k < m; each result lands in that particle's `fitness`
 * (lane 0 writes it; the caller reads it back after a __syncwarp). Returns how many ran a full window. */
__device__ __noinline__ unsigned warp_fitness_batch(const DevScene &S, const EvalCtx &E, const double *sDistW, const WarpWork &W, int m,
                                                    ParticleS *part, int p0, int stride) {
    const int lane = threadIdx.x & 31;
    const int V = E.V;
    unsigned ran = 0;
    if (!E.valid || V < 1 || m * V > 32) {          /* not batchable: one at a time */
        for (int k = 0; k < m; ++k) {
            ParticleS &q = part[p0 + k * stride];
            const double f = warp_fitness_any(S, E, sDistW, W, q.pos[0], q.pos[1], q.pos[2]);
            ran += f != DBL_MAX ? 1u : 0u;
            if (lane == 0) q.fitness = f;
        }
        __syncwarp();
        return ran;
    }
    HypoS *hyp = (HypoS *)W.hyp;
    __syncwarp();
    {
        const int k = lane / V, v = lane - k * V;
        if (k < m) {
            const double *pos = part[p0 + k * stride].pos;
            double n[3], center[3], pt[2];
            spherical2Normal(pos[0], pos[1], n);
            bool run = !(dot3(n, E.refOptN) > 0);
#pragma unroll
            for (int q = 0; q < 3; ++q) center[q] = E.ray[q] * pos[2] + E.refC[q];
            if (run) {
                const double d = -dot3(center, n);
                double Mref[9], inv[9];
                plane_matrix(E.refKR, E.refKT, n, d, E.sc, Mref);
                inv3(Mref, inv);
                const ViewS &vw = E.view[v];
                double *Hi = W.H + 9 * (k * V + v);
                if (vw.isRef) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) Hi[q] = (q % 4 == 0) ? 1.0 : 0.0;
                } else {
                    double M[9];
                    plane_matrix(vw.KR, vw.KT, n, d, E.sc, M);
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            double acc = M[r * 3] * inv[c];
                            acc += M[r * 3 + 1] * inv[3 + c];
                            acc += M[r * 3 + 2] * inv[6 + c];
                            Hi[r * 3 + c] = acc;
                        }
                }
                run = ref_window_ok(S, E, center, pt);
            }
            if (v == 0) {
                hyp[k].pt[0] = pt[0];
                hyp[k].pt[1] = pt[1];
                hyp[k].run = run ? 1 : 0;
            }
        }
    }
    __syncwarp();
    for (int k = 0; k < m; ++k) {
        double f = DBL_MAX;
        if (hyp[k].run) {
            WarpWork Wk = W;
            Wk.H = W.H + 9 * (k * V);
            const double pt[2] = {hyp[k].pt[0], hyp[k].pt[1]};
            f = warp_window_any(S, E, sDistW, Wk, pt);
        }
        ran += f != DBL_MAX ? 1u : 0u;
        if (lane == 0) part[p0 + k * stride].fitness = f;
    }
    __syncwarp();
    return ran;
}

/* Fill the evaluation context for (refCam, LOD, camIdx[0..V)). Collective over `nthreads` threads with index tid. */
__device__ __forceinline__ void build_eval_ctx(const DevScene &S, EvalCtx &E, const double *ray, int refCam, int LOD, int V,
                                               const uint16_t *camIdx, int tid, int nthreads) {
    if (tid == 0) {
        bool valid = refCam >= 0 && refCam < S.nCams && LOD >= 0 && LOD < PMVS_MAX_LEVELS && V > 0 && V <= PMVS_MAX_VIEWS;
        if (valid) {
            const DevCamera &rc = S.cams[refCam];
            valid = LOD <= rc.maxLOD;
            for (int k = 0; k < 3; ++k) {
                E.ray[k] = ray[k];
                E.refC[k] = rc.center[k];
                E.refOptN[k] = rc.optN[k];
                E.refT[k] = rc.t[k];
                E.refKT[k] = rc.KT[k];
            }
            for (int k = 0; k < 9; ++k) {
                E.refR[k] = rc.R[k];
                E.refKR[k] = rc.KR[k];
            }
            E.refFocal[0] = rc.focal[0];
            E.refFocal[1] = rc.focal[1];
            E.refPP[0] = rc.pp[0];
            E.refPP[1] = rc.pp[1];
            if (valid) {
                E.sc = S.lodScale[LOD];
                E.refQuad = rc.level[LOD].quad;
                E.refEdge = rc.level[LOD].edge;
                E.refCols = rc.level[LOD].cols;
                E.refRows = rc.level[LOD].rows;
            }
        }
        E.refCam = refCam;
        E.LOD = LOD;
        E.V = valid ? V : 0;
        E.valid = valid ? 1 : 0;
    }
    for (int v = tid; v < V && v < PMVS_MAX_VIEWS; v += nthreads) {
        ViewS &vw = E.view[v];
        const int ci = camIdx[v];
        if (ci < S.nCams && LOD >= 0 && LOD < PMVS_MAX_LEVELS) {
            const DevCamera &cam = S.cams[ci];
            for (int k = 0; k < 9; ++k) vw.KR[k] = cam.KR[k];
            for (int k = 0; k < 3; ++k) vw.KT[k] = cam.KT[k];
            const bool has = LOD <= cam.maxLOD;
            vw.quad = has ? cam.level[LOD].quad : nullptr;
            vw.cols = has ? cam.level[LOD].cols : 0;
            vw.rows = has ? cam.level[LOD].rows : 0;
            vw.isRef = (ci == refCam);
        } else {
            vw.quad = nullptr;
            vw.cols = vw.rows = 0;
            vw.isRef = 0;
        }
    }
}
/* after build_eval_ctx + barrier: a view without the level (cameras of different sizes) invalidates the context —
 * the reference would index a missing pyramid level there. */
__device__ __forceinline__ void finish_eval_ctx(EvalCtx &E, int tid) {
    if (tid == 0) {
        E.refView = -1;
        for (int v = 0; v < E.V; ++v) {
            if (E.view[v].quad == nullptr) E.valid = 0;
            if (E.view[v].isRef && E.refView < 0) E.refView = v;
        }
    }
}

/* =====================================================================================================
 * GLN-PSO (pso/psosolver.cpp). State in shared memory; one warp evaluates one particle at a time, one thread moves one.
 * Arithmetic is the reference's expression order without contraction, so given identical fitness values the
 * swarm is bit-identical to the unmodified reference solver (tests/test_pso_kat.py).
 * =================================================================================================== */
struct PsoS {
    double L[3], U[3], inter[3];
    double iw, gBestFitness;
    uint64_t key;
    int P, maxIter, iteration, gBestIdx, localK, converged, drawBase, _pad;
};

/*
 * moveParticles (psosolver.cpp:220-265) incl. getLocalBest (:151-191) and setNearNeighborBest (:193-218), with
 * thread = particle: every thread scans the other particles j = 0..P-1 in the reference's order, so the selections
 * (stable k-nearest by pBest distance, first maximum of the fitness-distance ratio) need no reductions and are the
 * reference's sequential semantics verbatim. The four scans of a particle (local best, FDR in each dimension) are
 * independent: each kind of scan runs on its own warp, results meet in shared memory.
 */
struct MoveS {
    int lBest[PMVS_MAX_PARTICLES];
    int nIdx[3][PMVS_MAX_PARTICLES];
};

__device__ __forceinline__ int pso_local_best(const PsoS &ps, const ParticleS *part, int i) {
    const int P = ps.P, K = ps.localK;
    const double b0 = part[i].pBest[0], b1 = part[i].pBest[1], b2 = part[i].pBest[2];
    /* the localK nearest in stable ascending order of (dist, index): insertion into a sorted register list */
    double bd[5];
    int bj[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { bd[k] = DBL_MAX; bj[k] = 0x7fffffff; }
    for (int j = 0; j < P; ++j) {
        double d;
        if (j == i) d = DBL_MAX;
        else {
            const double a0 = b0 - part[j].pBest[0], a1 = b1 - part[j].pBest[1], a2 = b2 - part[j].pBest[2];
            d = a0 * a0;
            d += a1 * a1;
            d += a2 * a2;
        }
        int id = j;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const bool before = (d < bd[k]) || (d == bd[k] && id < bj[k]);
            const double td = before ? bd[k] : d;
            const int tj = before ? bj[k] : id;
            bd[k] = before ? d : bd[k];
            bj[k] = before ? id : bj[k];
            d = td;
            id = tj;
        }
    }
    double minFitness = DBL_MAX;
    int best = i;                      /* default: the particle's own pBest (:183) */
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k < K && bj[k] < P) {
            const double f = part[bj[k]].pbf;
            if (f < minFitness) { minFitness = f; best = bj[k]; }
        }
    return best;
}

__device__ __forceinline__ int pso_near_neighbor(const PsoS &ps, const ParticleS *part, int i, int d) {
    const double fitness = part[i].fitness, x = part[i].pos[d];
    double maxFDR = -DBL_MAX;
    int best = -1;                     /* -1: nBest[d] keeps its previous value (:204-217) */
    for (int j = 0; j < ps.P; ++j) {
        if (j == i) continue;
        const double FDR = (fitness - part[j].pbf) / fabs(x - part[j].pBest[d]);
        if (FDR > maxFDR) { maxFDR = FDR; best = j; }
    }
    return best;
}

/* the four neighbourhood scans of every particle, one scan kind per warp (no divergence); no barrier inside */
__device__ __forceinline__ void pso_scans(const PsoS &ps, const ParticleS *part, MoveS &mv) {
    const int P = ps.P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    for (int kind = warp; kind < 4; kind += NW)
        for (int i = lane; i < P; i += 32) {
            if (kind == 0) mv.lBest[i] = pso_local_best(ps, part, i);
            else mv.nIdx[kind - 1][i] = pso_near_neighbor(ps, part, i, kind - 1);
        }
}

/* velocity / position update of generation `it`, thread = particle (psosolver.cpp:225-262); ends with a barrier */
__device__ __forceinline__ void pso_apply_moves(PsoS &ps, ParticleS *part, const MoveS &mv, int it) {
    const int P = ps.P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        ParticleS &me = part[i];
        const uint64_t c0 = (uint64_t)ps.drawBase + 4ull * ((uint64_t)it * P + i);
        const double pVecW = 1.2 * pmvs_random(ps.key, c0);
        const double gVecW = 1.5 * pmvs_random(ps.key, c0 + 1);
        const double lVecW = 1.0 * pmvs_random(ps.key, c0 + 2);
        const double nVecW = 1.0 * pmvs_random(ps.key, c0 + 3);
        const ParticleS &gb = part[ps.gBestIdx];
        const ParticleS &lb = part[mv.lBest[i]];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int nj = mv.nIdx[d][i];
            if (nj >= 0) me.nBest[d] = part[nj].pBest[d];
            const double x = me.pos[d];
            double v = ps.iw * me.vec[d] + pVecW * (me.pBest[d] - x) + gVecW * (gb.pBest[d] - x) + lVecW * (lb.pBest[d] - x) +
                       nVecW * (me.nBest[d] - x);
            double nx = x + v;
            if (nx > ps.U[d]) nx = ps.U[d];
            if (nx < ps.L[d]) nx = ps.L[d];
            me.vec[d] = v;
            me.pos[d] = nx;
        }
    }
    __syncthreads();
}

/* updateGbest (psosolver.cpp:137-149), sequential like the reference (<=: the highest index wins ties; NaN never wins) */
__device__ __forceinline__ void pso_update_gbest(PsoS &ps, const ParticleS *part) {
    for (int j = 0; j < ps.P; ++j)
        if (part[j].pbf <= ps.gBestFitness) {
            ps.gBestFitness = part[j].pbf;
            ps.gBestIdx = j;
        }
}

/*
 * PsoSolver ctor + setParticle + run(true) (psosolver.cpp:7-43, :94-110, :267-306). CTA-collective.
 * eval(pos) is warp-collective and returns the fitness in every lane. Returns evaluations spent.
 */
template <class Eval>
__device__ unsigned pso_run(PsoS &ps, ParticleS *part, MoveS &mv, Eval &eval, const double *init, bool hasInit) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int P = ps.P;
    unsigned evals = 0;
    for (int i = tid; i < P; i += blockDim.x) {                       /* initParticles :94-110 */
        ParticleS &q = part[i];
        for (int d = 0; d < 3; ++d) {
            const uint64_t c = 2ull * ((uint64_t)d * P + i);
            q.pos[d] = (ps.inter[d] * pmvs_random(ps.key, c)) + ps.L[d];
            q.vec[d] = (2.0 * ps.inter[d] * pmvs_random(ps.key, c + 1)) - ps.inter[d];
            q.pBest[d] = q.pos[d];
            q.nBest[d] = 0;
        }
        q.fitness = 1.7976931348623158e+308;
        q.pbf = 1.7976931348623158e+308;
    }
    if (tid == 0) ps.drawBase = 6 * P + (hasInit ? 3 : 0);           /* draws consumed before the first iteration */
    __syncthreads();
    if (tid == 0 && hasInit) {                                        /* setParticle :267-284 */
        ParticleS &q = part[0];
        for (int d = 0; d < 3; ++d) {
            q.pos[d] = init[d];
            q.pBest[d] = q.pos[d];
            q.vec[d] = (2.0 * ps.inter[d] * pmvs_random(ps.key, 6ull * P + d)) - ps.inter[d];
        }
    }
    __syncthreads();
    for (int p0 = warp; p0 < P; p0 += NW * PMVS_EVAL_BATCH) {         /* initFitness :112-119 */
        int m = (P - p0 + NW - 1) / NW;
        if (m > PMVS_EVAL_BATCH) m = PMVS_EVAL_BATCH;
        eval.batch(m, part, p0, NW);                                  /* fitness of part[p0 + k NW], k < m */
        if (lane == 0)
            for (int k = 0; k < m; ++k) part[p0 + k * NW].pbf = part[p0 + k * NW].fitness;
    }
    evals += P;
    __syncthreads();
    /*
     * run (:286-306), three barriers per generation. After the evaluations of generation it-1:
     *   phase A  one warp does the sequential bookkeeping — updateGbest (:137-149) and the inertia step (:304) of
     *            generation it-1, then the convergence test of generation it (:295, sequential sums :70-92) — while
     *            four other warps run the neighbourhood scans of generation it (they read pBest / pBestFitness only);
     *   phase B  velocity / position update;   phase C  evaluations.
     */
    const int bookWarp = NW > 4 ? 4 : 0;
    int it = 0;
    for (;;) {
        if (it < ps.maxIter) pso_scans(ps, part, mv);
        if (warp == bookWarp) {
            if (lane == 0) {
                if (it == 0) {                                        /* :288-291 */
                    ps.gBestIdx = 0;
                    ps.gBestFitness = part[0].pbf;
                } else {
                    const double niw = ps.iw - 1.0 / ps.maxIter;      /* :304 (after updateGbest, which does not read iw) */
                    ps.iw = niw > 0.4 ? niw : 0.4;
                }
                pso_update_gbest(ps, part);
            }
            __syncwarp();
            double index = 0;
            if (lane == 0) {
                const double *g = part[ps.gBestIdx].pBest;
                for (int i = 0; i < P; ++i)
                    for (int d = 0; d < 3; ++d) index += fabs(part[i].pos[d] - g[d]);
                index /= (3 * P);
            } else if (lane == 1) {
                for (int i = 0; i < P; ++i)
                    for (int d = 0; d < 3; ++d) index += fabs(part[i].vec[d]);
                index /= (3 * P);
            }
            const double disp = __shfl_sync(PMVS_FULL, index, 0), velo = __shfl_sync(PMVS_FULL, index, 1);
            if (lane == 0) ps.converged = (disp < 0.01 && velo < 0.01) ? 1 : 0;
        }
        __syncthreads();
        if (it >= ps.maxIter || ps.converged) break;
        pso_apply_moves(ps, part, mv, it);                                               /* moveParticles */
        for (int p0 = warp; p0 < P; p0 += NW * PMVS_EVAL_BATCH) {                        /* updateFitness :121-135 */
            int m = (P - p0 + NW - 1) / NW;
            if (m > PMVS_EVAL_BATCH) m = PMVS_EVAL_BATCH;
            eval.batch(m, part, p0, NW);
            if (lane == 0)
                for (int k = 0; k < m; ++k) {
                    ParticleS &q = part[p0 + k * NW];
                    if (q.fitness < q.pbf) {
                        q.pbf = q.fitness;
                        for (int d = 0; d < 3; ++d) q.pBest[d] = q.pos[d];
                    }
                }
        }
        evals += P;
        __syncthreads();
        ++it;
    }
    __syncthreads();            /* everyone has read the loop-exit flags */
    if (tid == 0) ps.iteration = it;
    __syncthreads();
    return evals;
}

__device__ __forceinline__ void pso_setup(PsoS &ps, const double *L, const double *U, int maxIter, int P, uint64_t key) {
    for (int d = 0; d < 3; ++d) {
        ps.L[d] = L[d];
        ps.U[d] = U[d];
        ps.inter[d] = U[d] - L[d];
    }
    ps.iw = 0.8;
    ps.gBestFitness = DBL_MAX;
    ps.key = key;
    ps.P = P;
    ps.maxIter = maxIter;
    ps.iteration = 0;
    ps.gBestIdx = 0;
    ps.localK = P < 5 ? P : 5;
    ps.converged = 0;
}
