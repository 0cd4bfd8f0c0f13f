/*
 * pmvs_oracle.cpp — CPU oracle for the pais-mvs patch-refinement hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This file is a plain f64 C++ restatement of the reference algorithm
 * (adahbingee/pais-mvs). Nothing under pais-mvs_b200/ may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - PSO half: PINNED. orc_pso_* is checked bit-for-bit against the UNMODIFIED reference
 *     TMVS/pso/psosolver.cpp + particle.cpp compiled in place (oracle/_ref/libpso_ref.so) and against
 *     golden vectors generated from it (tests/golden/pso_kat.json).
 *   - Cost / visibility half (patch.cpp): PARITY UNPINNED. The reference ships no tests, golden vectors
 *     or sample data, and patch.cpp cannot be compiled here (needs OpenCV 2.4.2 C++). The restatement
 *     follows the cited lines and is cross-checked by an independent NumPy implementation
 *     (tests/np_reference.py) and closed-form cases only.
 *   - OpenCV 2.4.2 arithmetic on the path (Mat_::inv 3x3, fitEllipse, cvRound) is restated from the
 *     published OpenCV 2.4 algorithms; OpenCV sources are not under /root/reference.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/pmvs_b200.h"

namespace {

/* ------------------------------------------------------------------------------------------------
 * Counter-based RNG replacing srand(time)+rand() (TMVS/pso/psosolver.cpp:60-68). rand() returns a
 * 31-bit value as with glibc (RAND_MAX = 2^31-1); random() = rand()/RAND_MAX in [0,1].
 * Independent restatement of the definition in pais-mvs_b200/csrc/pmvs_rng.h.
 * ---------------------------------------------------------------------------------------------- */
const uint64_t GOLD = 0x9E3779B97F4A7C15ULL;
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline uint64_t stream_key(uint64_t seed, int patchId, int run) {
    uint64_t k = mix64(seed + GOLD * (uint64_t)(uint32_t)(patchId + 1));
    return mix64(k + GOLD * (uint64_t)(uint32_t)(run + 1));
}
inline uint32_t rand31(uint64_t key, uint64_t ctr) { return (uint32_t)(mix64(key + GOLD * (ctr + 1)) >> 33); }

struct Rng {
    uint64_t key, ctr;
    double random() { return ((double)rand31(key, ctr++)) / 2147483647.0; }
};

/* ------------------------------------------------------------------------------------------------ */
struct Level { int cols, rows; std::vector<uint8_t> grey; std::vector<double> edge; };
struct Cam {
    PmvsCamera c;
    std::vector<Level> lv;
};
struct Scene {
    PmvsConfig cfg;
    std::vector<Cam> cams;
    std::vector<double> distW;      /* patchSize*patchSize, index x*patchSize+y (mvs.cpp:104-109) */
    double lodScale[PMVS_MAX_LEVELS];
    uint64_t seed;
};

/* cvRound: round-half-to-even (SSE2 cvtsd2si under the default rounding mode, OpenCV 2.4 core/types_c.h). */
inline int cvRound(double v) { return (int)std::nearbyint(v); }

/* MVS::initPatchDistanceWeighting, TMVS/mvs/mvs.cpp:97-114 */
void initDistWeight(Scene &s) {
    const int ps = s.cfg.patchSize, r = s.cfg.patchRadius;
    s.distW.assign((size_t)ps * ps, 0.0);
    double sigma = s.cfg.distWeighting;
    double s2 = 1.0 / (2.0 * sigma * sigma);
    double sc = 1.0 / (2.0 * M_PI * sigma * sigma);
    for (int x = 0; x < ps; ++x)
        for (int y = 0; y < ps; ++y) {
            double e = -(pow((double)(x - r), 2) + pow((double)(y - r), 2)) * s2;
            s.distW[(size_t)x * ps + y] = sc * exp(e);
        }
    /* cv::sum on CV_64F (OpenCV 2.4 stat.cpp sum_): s0 += src[i] + src[i+1] + src[i+2] + src[i+3] four at a time over the
     * continuous matrix, remainder one by one; `Mat / n` is a multiplication by 1./n (matop.cpp operator/(Mat,double)) */
    double n = 0;
    size_t i = 0;
    for (; i + 4 <= s.distW.size(); i += 4) n += s.distW[i] + s.distW[i + 1] + s.distW[i + 2] + s.distW[i + 3];
    for (; i < s.distW.size(); ++i) n += s.distW[i];
    const double rn = 1. / n;
    for (i = 0; i < s.distW.size(); ++i) s.distW[i] = s.distW[i] * rn;
}

void applyConfig(Scene &s, const PmvsConfig &cfg) {
    s.cfg = cfg;
    s.cfg.patchSize = (cfg.patchRadius << 1) + 1;          /* mvs.cpp:67 */
    for (int l = 0; l < PMVS_MAX_LEVELS; ++l) s.lodScale[l] = pow(s.cfg.lodRatio, l);
    initDistWeight(s);
}

/* Utility::spherical2Normal / normal2Spherical, TMVS/mvs/utility.h:17-29 */
inline void spherical2Normal(const double in[2], double out[3]) {
    out[0] = sin(in[0]) * cos(in[1]);
    out[1] = sin(in[0]) * sin(in[1]);
    out[2] = cos(in[0]);
}
inline void normal2Spherical(const double in[3], double out[2]) {
    out[0] = acos(in[2]);
    out[1] = atan2(in[1], in[0]);
}
inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* Camera::inImage(Vec2d), TMVS/mvs/camera.h:116-131 */
inline bool inImage(const Cam &cam, double x, double y, int LOD) {
    if (LOD > cam.c.maxLOD) return false;
    if (std::isnan(x) || std::isnan(y)) return false;
    const Level &L = cam.lv[LOD];
    return !(x < 0 || x >= L.cols || y < 0 || y >= L.rows);
}
inline bool inImageI(const Cam &cam, int x, int y, int LOD) {   /* camera.h:133-148 */
    if (LOD > cam.c.maxLOD) return false;
    const Level &L = cam.lv[LOD];
    return !(x < 0 || x >= L.cols || y < 0 || y >= L.rows);
}

/* Camera::project (no distortion), TMVS/mvs/camera.cpp:138-160 */
inline bool project(const Scene &s, const Cam &cam, const double X[3], double out[2], int LOD) {
    const double *R = cam.c.R, *t = cam.c.t;
    double x2 = (R[0] * X[0] + R[1] * X[1] + R[2] * X[2]) + t[0];
    double y2 = (R[3] * X[0] + R[4] * X[1] + R[5] * X[2]) + t[1];
    double z2 = (R[6] * X[0] + R[7] * X[1] + R[8] * X[2]) + t[2];
    out[0] = cam.c.focal[0] * (x2 / z2);
    out[1] = cam.c.focal[1] * (y2 / z2);
    out[0] += cam.c.principal[0];
    out[1] += cam.c.principal[1];
    double sc = s.lodScale[LOD];
    out[0] *= sc;
    out[1] *= sc;
    return inImage(cam, out[0], out[1], LOD);
}

/* OpenCV 2.4 cv::invert for 3x3 CV_64F (modules/core/src/lapack.cpp, closed-form adjugate with det3;
 * singular -> zero matrix). Used by Mat_<double>::inv() at TMVS/mvs/patch.cpp:314. */
inline void inv3(const double S[9], double D[9]) {
#define Sd(r, c) S[(r)*3 + (c)]
    double d = Sd(0, 0) * (Sd(1, 1) * Sd(2, 2) - Sd(1, 2) * Sd(2, 1)) -
               Sd(0, 1) * (Sd(1, 0) * Sd(2, 2) - Sd(1, 2) * Sd(2, 0)) +
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

/* M = d*L*KR - L*KT*n^T with L = diag(s,s,1): the bracket of TMVS/mvs/patch.cpp:314 and :328 */
inline void planeMatrix(const double KR[9], const double KT[3], const double n[3], double d, double sc, double M[9]) {
    const double L[3] = {sc, sc, 1.0};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[r * 3 + c] = d * (L[r] * KR[r * 3 + c]) - (L[r] * KT[r]) * n[c];   /* gemm applies alpha after the dot product */
}

/* Patch::getHomographies, TMVS/mvs/patch.cpp:290-330 */
void getHomographies(const Scene &s, int refCamIdx, const uint16_t *camIdx, int camNum, int LOD, const double center[3],
                     const double normal[3], double *H /* camNum*9 */) {
    const Cam &ref = s.cams[refCamIdx];
    const double d = -dot3(center, normal);
    const double sc = s.lodScale[LOD];
    double Mref[9], inv[9];
    planeMatrix(ref.c.KR, ref.c.KT, normal, d, sc, Mref);
    inv3(Mref, inv);
    for (int i = 0; i < camNum; ++i) {
        double *Hi = H + 9 * i;
        if (camIdx[i] == refCamIdx) {
            for (int k = 0; k < 9; ++k) Hi[k] = (k % 4 == 0) ? 1.0 : 0.0;
            continue;
        }
        const Cam &cam = s.cams[camIdx[i]];
        double M[9];
        planeMatrix(cam.c.KR, cam.c.KT, normal, d, sc, M);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double acc = 0;
                for (int k = 0; k < 3; ++k) acc += M[r * 3 + k] * inv[k * 3 + c];
                Hi[r * 3 + c] = acc;
            }
    }
}

/* What getFitness reads from the Patch */
struct PatchView {
    double ray[3];
    int refCamIdx, LOD, camNum;
    const uint16_t *camIdx;
};

/* PAIS::getFitness, TMVS/mvs/patch.cpp:914-1047. pos = (theta, phi, depth). */
double getFitness(const Scene &s, const PatchView &patch, const double pos[3]) {
    const int patchRadius = s.cfg.patchRadius;
    const int patchSize = s.cfg.patchSize;
    const int LOD = patch.LOD;
    const Cam &refCam = s.cams[patch.refCamIdx];
    const int camNum = patch.camNum;
    const Level &refL = refCam.lv[LOD];

    double normal[3];
    const double sph[2] = {pos[0], pos[1]};
    spherical2Normal(sph, normal);                                   /* :935-936 */
    if (dot3(normal, refCam.c.opticalNormal) > 0) return DBL_MAX;    /* :939-941 */

    double center[3];
    for (int k = 0; k < 3; ++k) center[k] = patch.ray[k] * pos[2] + refCam.c.center[k];   /* :944 */

    std::vector<double> H((size_t)camNum * 9);
    getHomographies(s, patch.refCamIdx, patch.camIdx, camNum, LOD, center, normal, H.data());   /* :947-948 */

    double pt[2];
    if (!project(s, refCam, center, pt, LOD)) return DBL_MAX;        /* :951-954 */
    if (pt[0] - patchRadius < 2 || pt[0] + patchRadius >= refL.cols - 3 || pt[1] - patchRadius < 2 ||
        pt[1] + patchRadius >= refL.rows - 3)
        return DBL_MAX;                                              /* :957-962 */

    std::vector<double> c(camNum);
    double fitness = 0, sumWeight = 0;
    const double diffWeighting = s.cfg.diffWeighting, gradientWeighting = s.cfg.gradientWeighting;
    size_t it = 0;                                                   /* distance-table iterator :975 */
    (void)patchSize;

    for (double x = pt[0] - patchRadius; x <= pt[0] + patchRadius; ++x) {          /* :979 */
        for (double y = pt[1] - patchRadius; y <= pt[1] + patchRadius; ++y, ++it) { /* :980 */
            double mean = 0, avgSad = 0;
            if (refL.grey[(size_t)cvRound(y) * refL.cols + cvRound(x)] == 0) continue;   /* :986 */
            for (int i = 0; i < camNum; ++i) {
                const Level &L = s.cams[patch.camIdx[i]].lv[LOD];
                const double *Hi = &H[9 * i];
                double w = (Hi[6] * x + Hi[7] * y + Hi[8]);
                double ix = (Hi[0] * x + Hi[1] * y + Hi[2]) / w;
                double iy = (Hi[3] * x + Hi[4] * y + Hi[5]) / w;
                /* :999. NaN coordinates are undefined behaviour in the reference ((int)NaN indexes memory);
                 * the oracle and the CUDA path both treat them as out-of-bounds. */
                if (ix < 2 || ix >= L.cols - 3 || iy < 2 || iy >= L.rows - 3 || w == 0 || std::isnan(ix) || std::isnan(iy))
                    return DBL_MAX;
                int px0 = (int)ix, py0 = (int)iy;                    /* :1005-1012 */
                int px1 = px0 + 1, py2 = py0 + 1;
                const uint8_t *g = L.grey.data();
                const size_t W = L.cols;
                c[i] = (double)g[(size_t)py0 * W + px0] * (px1 - ix) * (py2 - iy) +
                       (double)g[(size_t)py0 * W + px1] * (ix - px0) * (py2 - iy) +
                       (double)g[(size_t)py2 * W + px0] * (px1 - ix) * (iy - py0) +
                       (double)g[(size_t)py2 * W + px1] * (ix - px0) * (iy - py0);   /* :1014-1017 */
                mean += c[i];
            }
            mean /= camNum;                                          /* :1022 */
            for (int i = 0; i < camNum; ++i) avgSad += fabs(c[i] - mean);
            avgSad /= camNum;                                        /* :1027 */

            double weight = 1;
            if (s.cfg.adaptiveDistanceEnable) weight *= s.distW[it];                          /* :1030-1032 */
            if (s.cfg.adaptiveDifferenceEnable) weight *= exp(-avgSad * avgSad / diffWeighting);  /* :1033-1035 */
            if (s.cfg.adaptiveGradientEnable)
                weight *= exp(-1.0 / (refL.edge[(size_t)cvRound(y) * refL.cols + cvRound(x)] * gradientWeighting)); /* :1036-1038 */
            sumWeight += weight;
            fitness += weight * avgSad;
        }
    }
    return fitness / sumWeight;                                      /* :1046 */
}

/* ------------------------------------------------------------------------------------------------
 * GLN-PSO restatement of TMVS/pso/psosolver.cpp + particle.cpp (whole files).
 * ---------------------------------------------------------------------------------------------- */
typedef double (*FitFn)(const double *pos, void *obj);

struct Particle {   /* particle.h:5-28 */
    double pBest[3], nBest[3], pos[3], vec[3];
    const double *lBest;
    double fitness, pBestFitness;
};

struct PsoSolver {
    int dim, iteration, maxIteration, particleNum, localK;
    double convergenceThreshold, iw, pw, gw, lw, nw;
    double rangeL[3], rangeU[3], rangeInter[3];
    std::vector<Particle> particles;
    const double *gBest;
    double gBestFitness;
    int gBestIteration;
    bool enableGLNPSO;
    FitFn fn;
    void *obj;
    Rng rng;
    uint32_t evals;

    /* ctor psosolver.cpp:7-43, defaults psosolver.h:104-111 */
    PsoSolver(const double *L, const double *U, FitFn fn_, void *obj_, int maxIter, int P, Rng rng_)
        : dim(3), iteration(0), maxIteration(maxIter), particleNum(P), localK(std::min(P, 5)), convergenceThreshold(0.01),
          iw(0.8), pw(1.2), gw(1.5), lw(1.0), nw(1.0), gBest(NULL), gBestFitness(DBL_MAX), gBestIteration(-1),
          enableGLNPSO(false), fn(fn_), obj(obj_), rng(rng_), evals(0) {
        for (int i = 0; i < dim; ++i) {
            rangeL[i] = L[i];
            rangeU[i] = U[i];
            rangeInter[i] = U[i] - L[i];
        }
        initParticles();
    }
    double random() { return rng.random(); }

    void initParticles() {   /* psosolver.cpp:94-110 */
        Particle z;
        memset(&z, 0, sizeof(z));
        z.fitness = z.pBestFitness = 1.7976931348623158e+308;   /* particle.cpp:11-12 */
        z.lBest = NULL;
        particles.assign(particleNum, z);
        for (int d = 0; d < dim; d++)
            for (int i = 0; i < particleNum; i++) {
                particles[i].pos[d] = (rangeInter[d] * random()) + rangeL[d];
                particles[i].vec[d] = (2.0 * rangeInter[d] * random()) - rangeInter[d];
                particles[i].pBest[d] = particles[i].pos[d];
            }
    }
    void setParticle(const double *pos) {   /* psosolver.cpp:267-284, vec == NULL, idx == 0 */
        for (int d = 0; d < dim; d++) {
            particles[0].pos[d] = pos[d];
            particles[0].pBest[d] = particles[0].pos[d];
            particles[0].vec[d] = (2.0 * rangeInter[d] * random()) - rangeInter[d];
        }
    }
    double eval(const Particle &p) { ++evals; return fn(p.pos, obj); }
    void initFitness() {   /* :112-119 */
        for (int i = 0; i < particleNum; i++) {
            Particle &p = particles[i];
            p.fitness = eval(p);
            p.pBestFitness = p.fitness;
        }
    }
    void updateFitness() {   /* :121-135 */
        for (int i = 0; i < particleNum; i++) {
            Particle &p = particles[i];
            p.fitness = eval(p);
            if (p.fitness < p.pBestFitness) {
                p.pBestFitness = p.fitness;
                for (int d = 0; d < dim; d++) p.pBest[d] = p.pos[d];
            }
        }
    }
    void updateGbest() {   /* :137-149 */
        for (int j = 0; j < particleNum; j++) {
            const Particle &p = particles[j];
            if (p.pBestFitness <= gBestFitness) {
                gBestFitness = p.pBestFitness;
                gBest = p.pBest;
                gBestIteration = iteration;
            }
        }
    }
    double getDispersionIDX() const {   /* :70-80 */
        double index = 0;
        for (int i = 0; i < particleNum; i++)
            for (int j = 0; j < dim; j++) index += fabs(particles[i].pos[j] - gBest[j]);
        index /= (dim * particleNum);
        return index;
    }
    double getVelocityIDX() const {   /* :82-92 */
        double index = 0;
        for (int i = 0; i < particleNum; i++)
            for (int j = 0; j < dim; j++) index += fabs(particles[i].vec[j]);
        index /= (dim * particleNum);
        return index;
    }
    const double *getLocalBest(int idx) const {   /* :151-191 */
        struct LP { double dist; int i; };
        std::vector<LP> cont(particleNum);
        const double *pos = particles[idx].pBest;
        for (int i = 0; i < particleNum; i++) {
            cont[i].dist = 0;
            cont[i].i = i;
            if (i == idx) { cont[i].dist = DBL_MAX; continue; }
            for (int d = 0; d < dim; d++) cont[i].dist += (pos[d] - particles[i].pBest[d]) * (pos[d] - particles[i].pBest[d]);
        }
        /* std::sort at :176. MSVC 2010 (the reference's compiler) uses insertion sort for <= 32 elements and
         * libstdc++ for <= 16, both stable; equal distances therefore keep index order. Restated as a stable sort. */
        std::stable_sort(cont.begin(), cont.end(), [](const LP &a, const LP &b) { return a.dist < b.dist; });
        double minFitness = DBL_MAX;
        const double *lBest = pos;
        for (int k = 0; k < localK; k++) {
            const Particle &p = particles[cont[k].i];
            if (p.pBestFitness < minFitness) {
                minFitness = p.pBestFitness;
                lBest = p.pBest;
            }
        }
        return lBest;
    }
    void setNearNeighborBest(int idx) {   /* :193-218 */
        const double fitness = particles[idx].fitness;
        const double *pos = particles[idx].pos;
        double *nBest = particles[idx].nBest;
        for (int d = 0; d < dim; d++) {
            double maxFDR = -DBL_MAX;
            for (int i = 0; i < particleNum; i++) {
                if (i == idx) continue;
                const Particle &p = particles[i];
                double FDR = (fitness - p.pBestFitness) / fabs(pos[d] - p.pBest[d]);
                if (FDR > maxFDR) {
                    maxFDR = FDR;
                    nBest[d] = p.pBest[d];
                }
            }
        }
    }
    void moveParticles() {   /* :220-265, particle order = single-thread order */
        for (int i = 0; i < particleNum; i++) {
            double pVecW, gVecW, lVecW = 0, nVecW = 0;
            Particle &p = particles[i];
            pVecW = pw * random();
            gVecW = gw * random();
            if (enableGLNPSO) {
                lVecW = lw * random();
                nVecW = nw * random();
                p.lBest = getLocalBest(i);
                setNearNeighborBest(i);
            }
            for (int d = 0; d < dim; d++) {
                if (enableGLNPSO)
                    p.vec[d] = iw * p.vec[d] + pVecW * (p.pBest[d] - p.pos[d]) + gVecW * (gBest[d] - p.pos[d]) +
                               lVecW * (p.lBest[d] - p.pos[d]) + nVecW * (p.nBest[d] - p.pos[d]);
                else
                    p.vec[d] = iw * p.vec[d] + pVecW * (p.pBest[d] - p.pos[d]) + gVecW * (gBest[d] - p.pos[d]);
                p.pos[d] += p.vec[d];
                if (p.pos[d] > rangeU[d]) p.pos[d] = rangeU[d];
                if (p.pos[d] < rangeL[d]) p.pos[d] = rangeL[d];
            }
        }
    }
    void run(bool glnpso, double minIw = 0.4) {   /* :286-306 */
        enableGLNPSO = glnpso;
        initFitness();
        gBest = particles[0].pBest;
        gBestFitness = particles[0].pBestFitness;
        updateGbest();
        for (iteration = 0; iteration < maxIteration; iteration++) {
            if (getDispersionIDX() < convergenceThreshold && getVelocityIDX() < convergenceThreshold) break;
            moveParticles();
            updateFitness();
            updateGbest();
            iw = std::max(iw - 1.0 / maxIteration, minIw);
        }
    }
};

/* ------------------------------------------------------------------------------------------------
 * OpenCV 2.4 fitEllipse (imgproc/src/shapedescr.cpp cvFitEllipse2, "New fitellipse algorithm, contributed by
 * Dr. Daniel Weiss") with cvSolve(CV_SVD) = one-sided Jacobi SVD (core/src/lapack.cpp JacobiSVDImpl_) followed by
 * truncated back-substitution (SVBkSb, threshold = 2*DBL_EPSILON*sum(w)). Restated from the published algorithm.
 * ---------------------------------------------------------------------------------------------- */
/* Solve min ||A x - b|| for A (m x n, row-major, n <= 5, m <= 8) via Jacobi SVD; min-norm on rank deficiency. */
void svdSolve(const double *A, const double *b, int m, int n, double *x) {
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
    const int max_iter = std::max(m, 30);
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
    /* x = sum_i (u_i . b / w_i) v_i over w_i > threshold; u_i = At[i]/w_i (order of the singular values is
     * irrelevant to the sum, so the descending sort of the original is omitted). */
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

/* returns min(width,height)/max(width,height) exactly as TMVS/mvs/patch.cpp:285-287 (float sizes, float division) */
double fitEllipseRatio(const float *px, const float *py, int n, float *outW, float *outH) {
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
    svdSolve(Ad, bd, n, 5, gfp);
    double A2[4] = {2 * gfp[0], gfp[2], gfp[2], 2 * gfp[1]}, b2[2] = {gfp[3], gfp[4]};
    svdSolve(A2, b2, 2, 2, rp);
    for (int i = 0; i < n; ++i) {
        float x = px[i] - cx, y = py[i] - cy;
        bd[i] = 1.0;
        Ad[i * 3] = (x - rp[0]) * (x - rp[0]);
        Ad[i * 3 + 1] = (y - rp[1]) * (y - rp[1]);
        Ad[i * 3 + 2] = (x - rp[0]) * (y - rp[1]);
    }
    svdSolve(Ad, bd, n, 3, gfp);
    rp[4] = -0.5 * atan2(gfp[2], gfp[1] - gfp[0]);
    t = sin(-2.0 * rp[4]);
    if (fabs(t) > fabs(gfp[2]) * min_eps) t = gfp[2] / t;
    else t = gfp[1] - gfp[0];
    rp[2] = fabs(gfp[0] + gfp[1] - t);
    if (rp[2] > min_eps) rp[2] = sqrt(2.0 / rp[2]);
    rp[3] = fabs(gfp[0] + gfp[1] + t);
    if (rp[3] > min_eps) rp[3] = sqrt(2.0 / rp[3]);
    float w = (float)(rp[2] * 2), h = (float)(rp[3] * 2);
    if (outW) *outW = w;
    if (outH) *outH = h;
    return (double)(std::min(w, h) / std::max(w, h));
}

/* Patch::getHomographyRegionRatio, TMVS/mvs/patch.cpp:269-288 */
double regionRatio(const Scene &s, const double pt[2], const double *H) {
    const int r = s.cfg.patchRadius;
    double x[] = {pt[0] - r, pt[0] - r, pt[0] + r, pt[0] + r, pt[0] - r, pt[0], pt[0] + r, pt[0]};
    double y[] = {pt[1] - r, pt[1] + r, pt[1] + r, pt[1] - r, pt[1], pt[1] + r, pt[1], pt[1] - r};
    float fx[8], fy[8];
    for (int i = 0; i < 8; ++i) {
        double w = H[6] * x[i] + H[7] * y[i] + H[8];
        fx[i] = (float)((H[0] * x[i] + H[1] * y[i] + H[2]) / w);
        fy[i] = (float)((H[3] * x[i] + H[4] * y[i] + H[5]) / w);
    }
    return fitEllipseRatio(fx, fy, 8, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * Patch (abstractpatch.h:21-53 + patch.h:17-20) and its lifecycle.
 * ---------------------------------------------------------------------------------------------- */
struct Patch {
    double center[3], normal[3], normalS[2], ray[3], depth, depthRange[2];
    double fitness, priority, correlation;
    int LOD, refCamIdx, type, id;
    bool drop;
    std::vector<uint16_t> camIdx;
    std::vector<double> corrTable;
    std::vector<double> imgPoint;   /* 2 per entry */
    int psoRuns, psoIterations;
    uint32_t evals, status, windowEvals;
    int camNum() const { return (int)camIdx.size(); }
};

void setNormalS(Patch &p, const double nS[2]) {   /* abstractpatch.cpp:47-50 */
    p.normalS[0] = nS[0];
    p.normalS[1] = nS[1];
    spherical2Normal(p.normalS, p.normal);
}

/* Patch::setReferenceCameraIndex, patch.cpp:415-445 */
void setReferenceCameraIndex(const Scene &s, Patch &p) {
    if (p.drop) return;
    const int camNum = p.camNum();
    if (camNum < s.cfg.minCamNum) { p.drop = true; return; }
    p.refCamIdx = -1;
    double maxCorr = -DBL_MAX;
    for (int i = 0; i < camNum; i++) {
        const double *on = s.cams[p.camIdx[i]].c.opticalNormal;
        const double neg[3] = {-on[0], -on[1], -on[2]};
        double corr = dot3(p.normal, neg);
        if (corr > maxCorr) { maxCorr = corr; p.refCamIdx = p.camIdx[i]; }
    }
    if (p.refCamIdx < 0) { p.refCamIdx = p.camIdx[0]; p.drop = true; }
}

/* Patch::setDepthAndRay, patch.cpp:447-461 */
void setDepthAndRay(const Scene &s, Patch &p) {
    if (p.drop) return;
    if (p.refCamIdx < 0) { p.drop = true; return; }
    const double *C = s.cams[p.refCamIdx].c.center;
    for (int k = 0; k < 3; ++k) p.ray[k] = p.center[k] - C[k];
    p.depth = sqrt(p.ray[0] * p.ray[0] + p.ray[1] * p.ray[1] + p.ray[2] * p.ray[2]);   /* cv::norm(Vec3d) */
    double inv = 1.0 / p.depth;
    for (int k = 0; k < 3; ++k) p.ray[k] = p.ray[k] * inv;
}

/* Patch::setDepthRange, patch.cpp:463-509 */
void setDepthRange(const Scene &s, Patch &p) {
    if (p.drop) return;
    const int camNum = p.camNum();
    if (camNum < s.cfg.minCamNum) { p.drop = true; return; }
    const Cam &refCam = s.cams[p.refCamIdx];
    double c2[3];
    for (int k = 0; k < 3; ++k) c2[k] = p.ray[k] * (p.depth + 1.0) + refCam.c.center[k];
    double maxWorldDist = -DBL_MAX;
    for (int i = 0; i < camNum; i++) {
        if (p.camIdx[i] == p.refCamIdx) continue;
        const Cam &cam = s.cams[p.camIdx[i]];
        double p1[2], p2[2];
        project(s, cam, p.center, p1, 0);
        project(s, cam, c2, p2, 0);
        double dx = p1[0] - p2[0], dy = p1[1] - p2[1];
        double imgDist = sqrt(dx * dx + dy * dy);
        double worldDist = 1.0 / imgDist;
        if (worldDist > maxWorldDist && imgDist >= 0.01) maxWorldDist = worldDist;
    }
    if (maxWorldDist == -DBL_MAX) { p.drop = true; return; }
    p.depthRange[0] = std::max(p.depth - maxWorldDist * s.cfg.depthRangeScalar, 0.0);
    p.depthRange[1] = p.depth + std::min(maxWorldDist * s.cfg.depthRangeScalar, s.cfg.neighborRadius * 100);
}

/* Patch::setLOD, patch.cpp:511-610 */
void setLOD(const Scene &s, Patch &p) {
    if (p.drop) return;
    if (p.refCamIdx < 0) { p.drop = true; return; }
    const int r = s.cfg.patchRadius;
    const Cam &refCam = s.cams[p.refCamIdx];
    double mean = 0, variance = 0;
    int count;
    std::vector<uint8_t> tex((size_t)s.cfg.patchSize * s.cfg.patchSize);
    double pt[2];
    p.LOD = s.cfg.minLOD - 1;
    while (variance < s.cfg.textureVariation) {
        p.LOD++;
        if (p.LOD >= refCam.c.maxLOD) { p.LOD = refCam.c.maxLOD; return; }
        if (!project(s, refCam, p.center, pt, p.LOD)) { p.LOD = std::max(p.LOD - 1, 0); return; }
        mean = 0; variance = 0; count = 0;
        const Level &L = refCam.lv[p.LOD];
        for (int x = cvRound(pt[0]) - r; x <= cvRound(pt[0]) + r; x++)
            for (int y = cvRound(pt[1]) - r; y <= cvRound(pt[1]) + r; y++) {
                if (!inImageI(refCam, x, y, p.LOD)) { p.LOD = std::max(p.LOD - 1, 0); return; }
                tex[count] = L.grey[(size_t)y * L.cols + x];
                mean += tex[count];
                count++;
            }
        mean /= count;
        for (int i = 0; i < count; i++) variance += (tex[i] - mean) * (tex[i] - mean);
        variance /= count;
    }
}

/* Patch::setPriority, patch.cpp:612-625 */
void setPriority(const Scene &s, Patch &p) {
    if (p.drop) return;
    double camRatio = ((double)p.camNum()) / ((double)s.cams.size());
    p.priority = p.fitness * exp(-p.correlation / 1.0 - camRatio / 1.0) * (p.LOD + 1.0);
}

/* Patch::setImagePoint, patch.cpp:627-653 (colour lookup omitted: RGB images are outside the hot path) */
void setImagePoint(const Scene &s, Patch &p) {
    if (p.drop) return;
    const int camNum = p.camNum();
    if (camNum == 0) return;
    p.imgPoint.resize((size_t)camNum * 2);
    for (int i = 0; i < camNum; ++i) project(s, s.cams[p.camIdx[i]], p.center, &p.imgPoint[2 * i], 0);
}

/* Patch::getHomographyPatch, patch.cpp:332-386. Returns false (=> drop) on overflow. */
bool getHomographyPatch(const Scene &s, const double pt[2], const Level &img, const double *H, std::vector<double> &hp) {
    const int r = s.cfg.patchRadius, ps = s.cfg.patchSize;
    hp.assign((size_t)ps * ps, 0.0);
    int count = 0;
    double sum = 0;
    for (double x = pt[0] - r; x <= pt[0] + r; ++x)
        for (double y = pt[1] - r; y <= pt[1] + r; ++y) {
            double w = (H[6] * x + H[7] * y + H[8]);
            double ix = (H[0] * x + H[1] * y + H[2]) / w;
            double iy = (H[3] * x + H[4] * y + H[5]) / w;
            if (ix < 0 || ix >= img.cols - 1 || iy < 0 || iy >= img.rows - 1 || w == 0 || std::isnan(ix) || std::isnan(iy)) return false;
            int px0 = (int)ix, py0 = (int)iy, px1 = px0 + 1, py2 = py0 + 1;
            const uint8_t *g = img.grey.data();
            const size_t W = img.cols;
            double v = (double)g[(size_t)py0 * W + px0] * (px1 - ix) * (py2 - iy) + (double)g[(size_t)py0 * W + px1] * (ix - px0) * (py2 - iy) +
                       (double)g[(size_t)py2 * W + px0] * (px1 - ix) * (iy - py0) + (double)g[(size_t)py2 * W + px1] * (ix - px0) * (iy - py0);
            if (count < ps * ps) hp[count] = v;
            sum += v * v;
            ++count;
        }
    const double rn = 1. / sqrt(sum);                    /* hp /= sqrt(sum): Mat /= s is convertTo(.., 1./s) (OpenCV 2.4 mat.hpp) */
    for (size_t i = 0; i < hp.size(); ++i) hp[i] = hp[i] * rn;
    return true;
}

/* Patch::setCorrelationTable, patch.cpp:221-267 */
void setCorrelationTable(const Scene &s, Patch &p, const std::vector<double> &H) {
    const int camNum = p.camNum();
    const Cam &refCam = s.cams[p.refCamIdx];
    p.corrTable.assign((size_t)camNum * camNum, 0.0);
    double pt[2];
    project(s, refCam, p.center, pt, p.LOD);
    std::vector<std::vector<double> > HP(camNum);
    for (int i = 0; i < camNum; i++) {
        if (p.drop) break;                                    /* patch.cpp:334 */
        const Level &img = s.cams[p.camIdx[i]].lv[p.LOD];
        if (!getHomographyPatch(s, pt, img, &H[9 * i], HP[i])) p.drop = true;
    }
    if (p.drop) { p.correlation = 0; return; }
    for (int i = 0; i < camNum; ++i) {
        p.corrTable[(size_t)i * camNum + i] = 0;
        for (int j = i + 1; j < camNum; ++j) {
            double corr = 0;
            for (size_t k = 0; k < HP[i].size(); ++k) corr += HP[i][k] * HP[j][k];
            p.corrTable[(size_t)i * camNum + j] = corr;
            p.corrTable[(size_t)j * camNum + i] = corr;
        }
    }
    p.correlation = 0;
    for (int i = 0; i < camNum; ++i)
        for (int j = 0; j < camNum; ++j) p.correlation += p.corrTable[(size_t)i * camNum + j];
    p.correlation /= (camNum * camNum - camNum);
}

/* Patch::removeInvisibleCamera, patch.cpp:655-721 */
void removeInvisibleCamera(const Scene &s, Patch &p) {
    if (p.drop) return;
    const int camNum = p.camNum();
    const Cam &refCam = s.cams[p.refCamIdx];
    std::vector<double> H((size_t)camNum * 9);
    getHomographies(s, p.refCamIdx, p.camIdx.data(), camNum, p.LOD, p.center, p.normal, H.data());
    setCorrelationTable(s, p, H);
    /* NOTE: setCorrelationTable may set drop, but the reference carries on (corrTable is all zero then). */
    double maxCorr = -DBL_MAX;
    int maxIdx = 0;
    for (int i = 0; i < camNum; ++i) {
        double corrSum = 0;
        for (int j = 0; j < camNum; ++j) corrSum += p.corrTable[(size_t)i * camNum + j];
        if (corrSum >= maxCorr) { maxIdx = i; maxCorr = corrSum; }
    }
    double pt[2];
    project(s, refCam, p.center, pt, p.LOD);
    std::vector<int> removeIdx;
    for (int i = 0; i < camNum; ++i) {
        if (regionRatio(s, pt, &H[9 * i]) < s.cfg.minRegionRatio) { removeIdx.push_back(p.camIdx[i]); continue; }
        const double *on = s.cams[p.camIdx[i]].c.opticalNormal;
        const double neg[3] = {-on[0], -on[1], -on[2]};
        if (dot3(p.normal, neg) < 0) { removeIdx.push_back(p.camIdx[i]); continue; }
        if (i == maxIdx) continue;
        if (p.corrTable[(size_t)maxIdx * camNum + i] < s.cfg.minCorrelation) { removeIdx.push_back(p.camIdx[i]); continue; }
    }
    for (size_t i = 0; i < removeIdx.size(); i++) {
        std::vector<uint16_t>::iterator it = std::find(p.camIdx.begin(), p.camIdx.end(), (uint16_t)removeIdx[i]);
        if (it != p.camIdx.end()) p.camIdx.erase(it);
    }
    if (p.camNum() < s.cfg.minCamNum) p.drop = true;
}

/* Patch::expandVisibleCamera, patch.cpp:723-761 */
void expandVisibleCamera(const Scene &s, Patch &p) {
    if (p.drop) return;
    std::vector<int> exp;
    for (size_t i = 0; i < s.cams.size(); ++i) {
        const double *on = s.cams[i].c.opticalNormal;
        const double neg[3] = {-on[0], -on[1], -on[2]};
        if (dot3(p.normal, neg) >= s.cfg.visibleCorrelation) exp.push_back((int)i);
    }
    if ((int)exp.size() < s.cfg.minCamNum) {
        for (size_t i = 0; i < p.camIdx.size(); ++i) {
            const double *on = s.cams[p.camIdx[i]].c.opticalNormal;
            const double neg[3] = {-on[0], -on[1], -on[2]};
            if (dot3(p.normal, neg) >= s.cfg.visibleCorrelation / 2.0) exp.push_back(p.camIdx[i]);
        }
        std::sort(exp.begin(), exp.end());
        exp.resize(std::unique(exp.begin(), exp.end()) - exp.begin());
    }
    if ((int)exp.size() > PMVS_MAX_VIEWS) {   /* capacity limit of the C-ABI records, not in the reference */
        p.status |= PMVS_S_TOO_MANY_VIEWS;
        p.camIdx.clear();
        p.drop = true;
        return;
    }
    p.camIdx.assign(exp.begin(), exp.end());
    if (p.camNum() < s.cfg.minCamNum) p.drop = true;
}

struct FitCtx { const Scene *s; PatchView pv; uint32_t windowEvals; };
double fitCallback(const double *pos, void *obj) {
    FitCtx *c = (FitCtx *)obj;
    const double f = getFitness(*c->s, c->pv, pos);
    if (f != DBL_MAX) {
#pragma omp atomic
        c->windowEvals++;
    }
    return f;
}

/* optional: the UNMODIFIED reference solver (oracle/_ref/libpso_ref.so), registered at run time */
typedef int (*RefPsoFn)(const double *L, const double *U, FitFn fn, void *obj, int maxIter, int P, const double *init,
                        uint64_t key, double *gbest, double *gbestFitness, int *iterations, int nThreads);
RefPsoFn g_refPso = NULL;

/* Patch::psoOptimization, patch.cpp:180-219 */
void psoOptimization(const Scene &s, Patch &p, bool useRefPso, int nThreads) {
    double rangeL[] = {0.0, p.normalS[1] - M_PI / 2.0, p.depthRange[0]};
    double rangeU[] = {M_PI, p.normalS[1] + M_PI / 2.0, p.depthRange[1]};
    double init[] = {p.normalS[0], p.normalS[1], p.depth};
    int maxIter, P;
    if (p.type == PMVS_TYPE_SEED) {
        maxIter = s.cfg.maxIteration * 2;
        P = s.cfg.particleNum * 2;
    } else {
        rangeL[0] = std::max(0.0, p.normalS[0] - M_PI / s.cfg.reduceNormalRange);
        rangeU[0] = std::min(M_PI, p.normalS[0] + M_PI / s.cfg.reduceNormalRange);
        rangeL[1] = p.normalS[1] - M_PI / s.cfg.reduceNormalRange;
        rangeU[1] = p.normalS[1] + M_PI / s.cfg.reduceNormalRange;
        maxIter = s.cfg.maxIteration;
        P = s.cfg.particleNum;
    }
    FitCtx ctx;
    ctx.s = &s;
    ctx.windowEvals = 0;
    memcpy(ctx.pv.ray, p.ray, sizeof(p.ray));
    ctx.pv.refCamIdx = p.refCamIdx;
    ctx.pv.LOD = p.LOD;
    ctx.pv.camNum = p.camNum();
    ctx.pv.camIdx = p.camIdx.data();
    const uint64_t key = stream_key(s.seed, p.id, p.psoRuns);
    double gb[3], gbf;
    int iters;
    if (useRefPso && g_refPso) {
        g_refPso(rangeL, rangeU, fitCallback, &ctx, maxIter, P, init, key, gb, &gbf, &iters, nThreads);
        p.evals += (uint32_t)(P * (iters + 1));
    } else {
        Rng rng = {key, 0};
        PsoSolver solver(rangeL, rangeU, fitCallback, &ctx, maxIter, P, rng);
        solver.setParticle(init);
        solver.run(true);
        gbf = solver.gBestFitness;
        memcpy(gb, solver.gBest, sizeof(gb));
        iters = solver.iteration;
        p.evals += solver.evals;
    }
    p.fitness = gbf;
    setNormalS(p, gb);
    p.depth = gb[2];
    const double *C = s.cams[p.refCamIdx].c.center;
    for (int k = 0; k < 3; ++k) p.center[k] = p.ray[k] * p.depth + C[k];
    p.psoIterations = iters;
    p.psoRuns++;
    p.windowEvals += ctx.windowEvals;
}

/* Patch::refine, patch.cpp:114-176 */
void refine(const Scene &s, Patch &p, bool useRefPso, int nThreads) {
    if (p.camNum() < s.cfg.minCamNum) { p.fitness = DBL_MAX; p.priority = DBL_MAX; p.drop = true; return; }
    setReferenceCameraIndex(s, p);
    setDepthAndRay(s, p);
    setDepthRange(s, p);
    setLOD(s, p);
    if (p.drop) return;
    int beforeRefCamIdx = p.refCamIdx, afterRefCamIdx = -1;
    int beforeCamNum = p.camNum(), afterCamNum = -1;
    int count = 0;
    int totalCamNum = beforeCamNum;
    while ((beforeRefCamIdx != afterRefCamIdx || beforeCamNum != afterCamNum) && count++ <= totalCamNum) {
        if (p.camNum() < s.cfg.minCamNum) { p.fitness = DBL_MAX; p.priority = DBL_MAX; p.drop = true; return; }
        beforeRefCamIdx = p.refCamIdx;
        beforeCamNum = p.camNum();
        psoOptimization(s, p, useRefPso, nThreads);
        if (p.fitness > s.cfg.maxFitness) { p.drop = true; return; }
        removeInvisibleCamera(s, p);
        setReferenceCameraIndex(s, p);
        setDepthAndRay(s, p);
        setDepthRange(s, p);
        setLOD(s, p);
        if (p.type == PMVS_TYPE_EXPAND) break;
        afterRefCamIdx = p.refCamIdx;
        afterCamNum = p.camNum();
    }
    setPriority(s, p);
    setImagePoint(s, p);
}

void patchFromIn(const PmvsPatchIn &in, Patch &p) {   /* AbstractPatch::init, abstractpatch.cpp:25-40 */
    memset(p.ray, 0, sizeof(p.ray));
    memcpy(p.center, in.center, sizeof(p.center));
    memcpy(p.normal, in.normal, sizeof(p.normal));
    memcpy(p.normalS, in.normalS, sizeof(p.normalS));
    p.depth = 0;
    p.depthRange[0] = p.depthRange[1] = 0;
    p.fitness = DBL_MAX;
    p.priority = DBL_MAX;
    p.correlation = 0;
    p.LOD = -1;
    p.refCamIdx = -1;
    p.type = in.type;
    p.id = in.id;
    p.drop = false;
    p.camIdx.assign(in.camIdx, in.camIdx + std::max(0, std::min(in.nCam, PMVS_MAX_VIEWS)));
    p.psoRuns = 0;
    p.psoIterations = 0;
    p.evals = 0;
    p.status = 0;
    p.windowEvals = 0;
}

void patchToOut(const Patch &p, PmvsPatchOut &o) {
    memset(&o, 0, sizeof(o));
    memcpy(o.center, p.center, sizeof(o.center));
    memcpy(o.normal, p.normal, sizeof(o.normal));
    memcpy(o.normalS, p.normalS, sizeof(o.normalS));
    memcpy(o.ray, p.ray, sizeof(o.ray));
    o.depth = p.depth;
    o.depthRange[0] = p.depthRange[0];
    o.depthRange[1] = p.depthRange[1];
    o.fitness = p.fitness;
    o.priority = p.priority;
    o.correlation = p.correlation;
    o.LOD = p.LOD;
    o.refCamIdx = p.refCamIdx;
    o.nCam = p.camNum();
    o.drop = p.drop ? 1 : 0;
    o.psoRuns = p.psoRuns;
    o.psoIterations = p.psoIterations;
    o.evaluations = p.evals;
    o.status = p.status;
    o.windowEvaluations = p.windowEvals;
    for (int i = 0; i < o.nCam; ++i) o.camIdx[i] = p.camIdx[i];
    o.nImgPoint = (int)(p.imgPoint.size() / 2);
    for (int i = 0; i < o.nImgPoint; ++i) { o.imgPoint[i][0] = p.imgPoint[2 * i]; o.imgPoint[i][1] = p.imgPoint[2 * i + 1]; }
}

/* analytic test functions for the PSO known-answer tests */
double testFn(const double *x, void *obj) {
    int id = *(int *)obj;
    switch (id) {
    default:
    case 0: return (x[0] - 0.3) * (x[0] - 0.3) + (x[1] + 0.2) * (x[1] + 0.2) + (x[2] - 1.5) * (x[2] - 1.5);   /* sphere */
    case 1: { double a = x[1] - x[0] * x[0], b = 1 - x[0], c = x[2] - x[1] * x[1], d = 1 - x[1]; return 100 * a * a + b * b + 100 * c * c + d * d; }   /* rosenbrock */
    case 2: { double s = 30; for (int i = 0; i < 3; ++i) s += x[i] * x[i] - 10 * cos(2 * M_PI * x[i]); return s; }   /* rastrigin */
    case 3: return (x[0] > 0.5) ? DBL_MAX : fabs(x[0]) + fabs(x[1]) + fabs(x[2]);   /* with invalid (DBL_MAX) region */
    case 4: return floor(4 * fabs(x[0])) + floor(4 * fabs(x[1])) + floor(4 * fabs(x[2]));   /* plateaus: exercises ties */
    }
}

}   // namespace

/* ================================================================================================
 * C API (ctypes) — test infrastructure
 * ============================================================================================== */
extern "C" {

struct orc_scene { Scene s; };

orc_scene *orc_create(const PmvsConfig *cfg, int nCams, const PmvsCamera *cams, uint64_t seed) {
    orc_scene *o = new orc_scene();
    applyConfig(o->s, *cfg);
    o->s.seed = seed;
    o->s.cams.resize(nCams);
    for (int i = 0; i < nCams; ++i) {
        Cam &c = o->s.cams[i];
        c.c = cams[i];
        c.lv.resize(cams[i].maxLOD + 1);
        for (int l = 0; l <= cams[i].maxLOD; ++l) {
            const PmvsLevel &L = cams[i].level[l];
            c.lv[l].cols = L.cols;
            c.lv[l].rows = L.rows;
            c.lv[l].grey.resize((size_t)L.cols * L.rows);
            for (int y = 0; y < L.rows; ++y) memcpy(&c.lv[l].grey[(size_t)y * L.cols], L.grey + (size_t)y * L.pitch, L.cols);
            if (L.edge) c.lv[l].edge.assign(L.edge, L.edge + (size_t)L.cols * L.rows);
            c.c.level[l].grey = NULL;
            c.c.level[l].edge = NULL;
        }
    }
    return o;
}
void orc_destroy(orc_scene *o) { delete o; }
void orc_set_config(orc_scene *o, const PmvsConfig *cfg) { applyConfig(o->s, *cfg); }
void orc_set_neighbor_radius(orc_scene *o, double r) { o->s.cfg.neighborRadius = r; }
void orc_set_ref_pso(void *fn) { g_refPso = (RefPsoFn)fn; }
void orc_dist_weight(orc_scene *o, double *out) { memcpy(out, o->s.distW.data(), o->s.distW.size() * sizeof(double)); }

/* seam 1 */
void orc_fitness_batch(orc_scene *o, int n, const PmvsHypothesis *in, double *out, int nThreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nThreads > 0 ? nThreads : 1)
    for (int i = 0; i < n; ++i) {
        PatchView pv;
        memcpy(pv.ray, in[i].ray, sizeof(pv.ray));
        pv.refCamIdx = in[i].refCamIdx;
        pv.LOD = in[i].LOD;
        pv.camNum = in[i].nCam;
        pv.camIdx = in[i].camIdx;
        const double pos[3] = {in[i].theta, in[i].phi, in[i].depth};
        out[i] = getFitness(o->s, pv, pos);
    }
}

void orc_homographies(orc_scene *o, const PmvsHypothesis *in, double *H) {
    double normal[3], center[3];
    const double sph[2] = {in->theta, in->phi};
    spherical2Normal(sph, normal);
    for (int k = 0; k < 3; ++k) center[k] = in->ray[k] * in->depth + o->s.cams[in->refCamIdx].c.center[k];
    getHomographies(o->s, in->refCamIdx, in->camIdx, in->nCam, in->LOD, center, normal, H);
}

/* seam 2. patchThreads > 1 parallelises over patches (a "best-case CPU" arrangement, not the reference's);
 * psoMode: 0 = restated PSO, 1 = unmodified reference PSO (serial), 2 = unmodified reference PSO with the
 * reference's own OpenMP-over-particles threading (psoThreads threads; RNG draw order is then racy, timing only). */
void orc_refine_batch(orc_scene *o, int n, const PmvsPatchIn *in, PmvsPatchOut *out, uint32_t flags, int psoMode,
                      int patchThreads, int psoThreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(patchThreads > 0 ? patchThreads : 1)
    for (int i = 0; i < n; ++i) {
        Patch p;
        patchFromIn(in[i], p);
        if ((flags & PMVS_F_EXPAND_VISIBLE) && p.type == PMVS_TYPE_EXPAND) expandVisibleCamera(o->s, p);
        refine(o->s, p, psoMode != 0, psoMode == 2 ? psoThreads : 1);
        if (flags & PMVS_F_POST_REMOVE_INVISIBLE) removeInvisibleCamera(o->s, p);
        patchToOut(p, out[i]);
    }
}

/* standalone pieces for unit tests */
double orc_fit_ellipse_ratio(const float *px, const float *py, int n, float *w, float *h) { return fitEllipseRatio(px, py, n, w, h); }
double orc_region_ratio(orc_scene *o, const double *pt, const double *H) { return regionRatio(o->s, pt, H); }
int orc_project(orc_scene *o, int cam, const double *X, int LOD, double *out) { return project(o->s, o->s.cams[cam], X, out, LOD) ? 1 : 0; }
void orc_inv3(const double *S, double *D) { inv3(S, D); }
uint32_t orc_rand31(uint64_t seed, int patchId, int run, uint64_t ctr) { return rand31(stream_key(seed, patchId, run), ctr); }
uint64_t orc_stream_key(uint64_t seed, int patchId, int run) { return stream_key(seed, patchId, run); }

/* PSO on analytic functions (known-answer tests): restated solver */
void orc_pso_test(int fnId, const double *L, const double *U, int maxIter, int P, const double *init, uint64_t key, int glnpso,
                  double *gbest, double *gbestFitness, int *iterations, double *particlesOut /* P*8: pos3 vec3 fit pbf */) {
    Rng rng = {key, 0};
    PsoSolver solver(L, U, testFn, &fnId, maxIter, P, rng);
    if (init) solver.setParticle(init);
    solver.run(glnpso != 0);
    memcpy(gbest, solver.gBest, 3 * sizeof(double));
    *gbestFitness = solver.gBestFitness;
    *iterations = solver.iteration;
    if (particlesOut)
        for (int i = 0; i < P; ++i) {
            const Particle &p = solver.particles[i];
            double *q = particlesOut + 8 * i;
            q[0] = p.pos[0]; q[1] = p.pos[1]; q[2] = p.pos[2];
            q[3] = p.vec[0]; q[4] = p.vec[1]; q[5] = p.vec[2];
            q[6] = p.fitness; q[7] = p.pBestFitness;
        }
}
/* MVS::neighborPatchFiltering's per-patch neighbour list length (TMVS/mvs/mvs.cpp:470-499): the reference builds the
 * list of all other patches with dist = cv::norm(center - centerN) (:483), sorts it ascending and keeps entries until
 * dist > neighborRadius (:496); the length equals the number of j != i with dist <= radius. cv::norm(Vec3d) is
 * sqrt(x*x + y*y + z*z) accumulated left to right (compiled here with -ffp-contract=off). */
void orc_neighbor_counts(int n, const double *centers, double radius, int *counts) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        int c = 0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double dx = centers[3 * i] - centers[3 * j], dy = centers[3 * i + 1] - centers[3 * j + 1], dz = centers[3 * i + 2] - centers[3 * j + 2];
            double s = dx * dx;
            s += dy * dy;
            s += dz * dz;
            const double dist = sqrt(s);
            if (!(dist > radius)) ++c;       /* :496 breaks on dist > radius */
        }
        counts[i] = c;
    }
}
double orc_test_fn(int fnId, const double *x) { return testFn(x, &fnId); }
void *orc_test_fn_ptr(void) { return (void *)testFn; }

}   // extern "C"
