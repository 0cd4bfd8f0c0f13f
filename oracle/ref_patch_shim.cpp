/*
 * ref_patch_shim.cpp — harness around the UNMODIFIED reference patch model, compiled in place from /root/reference by
 * oracle/Makefile into oracle/_ref/libtmvs_ref.so together with TMVS/mvs/{patch,abstractpatch,camera,cellmap,mvs}.cpp and
 * TMVS/pso/{psosolver,particle}.cpp against the OpenCV stand-in oracle/cvshim/.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/pmvs_oracle.cpp header). No reference source is copied into the repo: this file
 * only (1) injects a scene (cameras + pyramids in the C-ABI's PmvsCamera layout) into the reference's MVS singleton and
 * Camera objects, (2) puts a Patch into the state a PmvsHypothesis / PmvsPatchIn describes and calls the reference's own
 * PAIS::getFitness / Patch::getHomographies / Patch::refine / Patch::removeInvisibleCamera on it, (3) interposes
 * rand()/srand() with the repo's counter-based stream (pmvs_rng.h) exactly like ref_pso_shim.cpp, keyed per patch and
 * per solver run, and (4) supplies the link-time pieces that are outside the path (log file, file I/O, viewer hook).
 * Access to the reference's private members is by `#define private public` in THIS translation unit only (member
 * layout does not depend on access specifiers in the Itanium ABI).
 */
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <omp.h>
#include <unistd.h>

#include "../include/pmvs_b200.h"
#include "../pais-mvs_b200/csrc/pmvs_rng.h"
#include <opencv2/opencv.hpp>

#define private public
#define protected public
#include "mvs/patch.h"
#undef private
#undef protected

using namespace PAIS;

/* ---- pieces outside the path that the reference objects link against ----------------------------------------- */
void addPatchView(const Patch &) {}                                   /* viewer hook (TMVS.cpp:17-24): no viewer */
namespace PAIS {
ofstream *LogManager::instance = NULL;                                /* io/logmanager.cpp is not compiled: no log.txt */
bool LogManager::create() { return false; }
void LogManager::log(const char *, ...) {}
void LogManager::warning(const char *, ...) {}
void LogManager::error(const char *, ...) {}
void LogManager::close() {}
void FileLoader::loadNVM(const char *, MVS &) {}
void FileLoader::loadNVM2(const char *, MVS &) {}
void FileLoader::loadMVS(const char *, MVS &) {}
void FileWriter::writeMVS(const char *, const MVS &) {}
void FileWriter::writePLY(const char *, const MVS &) {}
void FileWriter::wirtePSR(const char *, const MVS &) {}
void FileWriter::writeDeletedPatchMVS(const char *, const MVS &) {}
void FileWriter::writeDeletedPatchPLY(const char *, const MVS &) {}
}

/* ---- OpenCV stand-ins with bodies ---------------------------------------------------------------------------- */
namespace {
/* cvSolve(CV_SVD) = one-sided Jacobi SVD + truncated back-substitution (lapack.cpp JacobiSVDImpl_, SVBkSb), as
 * restated in pmvs_oracle.cpp (the reference itself does not contain this code) */
void svd_solve(const double *A, const double *b, int m, int n, double *x) {
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
                if (std::fabs(p) <= eps * std::sqrt(a * bb)) continue;
                p *= 2;
                double beta = a - bb, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = std::sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = std::sqrt((gamma + beta) / (gamma * 2));
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
        W[i] = std::sqrt(sd);
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
}   // namespace

namespace cv {
/* cvFitEllipse2 (shapedescr.cpp) for n <= 8 float points: returns the box the reference reads (size only matters) */
RotatedRect fitEllipse(const std::vector<Point2f> &pts) {
    const int n = (int)pts.size();
    const double min_eps = 1e-8;
    double gfp[5], rp[5], t;
    std::vector<double> Ad((size_t)n * 5), bd((size_t)n);
    float cx = 0, cy = 0;
    for (int i = 0; i < n; ++i) { cx += pts[i].x; cy += pts[i].y; }
    cx /= n;
    cy /= n;
    for (int i = 0; i < n; ++i) {
        float x = pts[i].x - cx, y = pts[i].y - cy;
        bd[i] = 10000.0;
        Ad[i * 5] = -(double)x * x;
        Ad[i * 5 + 1] = -(double)y * y;
        Ad[i * 5 + 2] = -(double)x * y;
        Ad[i * 5 + 3] = x;
        Ad[i * 5 + 4] = y;
    }
    svd_solve(&Ad[0], &bd[0], n, 5, gfp);
    double A2[4] = {2 * gfp[0], gfp[2], gfp[2], 2 * gfp[1]}, b2[2] = {gfp[3], gfp[4]};
    svd_solve(A2, b2, 2, 2, rp);
    for (int i = 0; i < n; ++i) {
        float x = pts[i].x - cx, y = pts[i].y - cy;
        bd[i] = 1.0;
        Ad[i * 3] = (x - rp[0]) * (x - rp[0]);
        Ad[i * 3 + 1] = (y - rp[1]) * (y - rp[1]);
        Ad[i * 3 + 2] = (x - rp[0]) * (y - rp[1]);
    }
    svd_solve(&Ad[0], &bd[0], n, 3, gfp);
    rp[4] = -0.5 * atan2(gfp[2], gfp[1] - gfp[0]);
    t = sin(-2.0 * rp[4]);
    if (std::fabs(t) > std::fabs(gfp[2]) * min_eps) t = gfp[2] / t;
    else t = gfp[1] - gfp[0];
    rp[2] = std::fabs(gfp[0] + gfp[1] - t);
    if (rp[2] > min_eps) rp[2] = std::sqrt(2.0 / rp[2]);
    rp[3] = std::fabs(gfp[0] + gfp[1] + t);
    if (rp[3] > min_eps) rp[3] = std::sqrt(2.0 / rp[3]);
    RotatedRect box;
    box.center = Point2f((float)rp[0] + cx, (float)rp[1] + cy);
    box.size.width = (float)(rp[2] * 2);
    box.size.height = (float)(rp[3] * 2);
    if (box.size.width > box.size.height) {
        std::swap(box.size.width, box.size.height);
        box.angle = (float)(90 + rp[4] * 180 / M_PI);
    } else {
        box.angle = (float)(rp[4] * 180 / M_PI);
    }
    return box;
}
/* only reached through Camera's file constructor / the show* helpers, which this harness never calls */
void resize(const Mat &src, Mat &dst, Size, double, double, int) { dst = src.clone(); }
void Sobel(const Mat &src, Mat &dst, int, int, int, int) { dst = Mat(src.rows, src.cols, CV_64FC1); }
}   // namespace cv

/* ---- rand()/srand(): the repo's counter-based stream, one key per (patch id, solver run) ------------------------ */
namespace {
uint64_t gSeed = 0;
int gPatchId = 0, gRun = 0;
uint64_t gKey = 0, gCtr = 0;
bool gExpansion = false;          /* inside MVS::expansionPatches(): the patch under refinement is the one constructed last */
}
namespace PAIS { struct IdPeek { static int last() { return AbstractPatch::globalId - 1; } }; }
extern "C" int rand(void) { return (int)pmvs_rand31(gKey, gCtr++); }
extern "C" void srand(unsigned int) {          /* PsoSolver::setRandomSeed (psosolver.cpp:60-64): a new solver starts */
    if (gExpansion) {                          /* expandCell (mvs.cpp:566-577): Patch expPatch(center, parent) took id = globalId++ */
        const int id = PAIS::IdPeek::last();
        if (id != gPatchId) { gPatchId = id; gRun = 0; }
    }
    gKey = pmvs_stream_key(gSeed, gPatchId, gRun++);
    gCtr = 0;
}

namespace {
struct Quiet {                                   /* the reference printf()s progress and its whole config */
    int saved;
    Quiet() {
        fflush(stdout);
        saved = dup(1);
        int nul = open_null();
        dup2(nul, 1);
        close(nul);
    }
    ~Quiet() {
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
    static int open_null() {
        FILE *f = fopen("/dev/null", "w");
        int fd = dup(fileno(f));
        fclose(f);
        return fd;
    }
};

Mat_<double> mat3x3(const double *v) {
    Mat_<double> m(3, 3);
    for (int i = 0; i < 9; ++i) m.at<double>(i / 3, i % 3) = v[i];
    return m;
}
Mat_<double> mat3x1(const double *v) {
    Mat_<double> m(3, 1);
    for (int i = 0; i < 3; ++i) m.at<double>(i, 0) = v[i];
    return m;
}

/* a Patch in the state AbstractPatch::init leaves it (abstractpatch.cpp:25-40) plus the given inputs */
Patch make_patch(const double *center, const double *normal, const double *normalS, int nCam, const uint16_t *camIdx, int type, int id) {
    std::vector<int> none;
    Patch p(Vec3d(0, 0, 0), Vec2d(0, 0), none, DBL_MAX, 0.0, id);       /* loader ctor with no cameras: every setter returns early */
    p.init();
    p.center = Vec3d(center[0], center[1], center[2]);
    p.normal = Vec3d(normal[0], normal[1], normal[2]);
    p.normalS = Vec2d(normalS[0], normalS[1]);
    p.camIdx.assign(camIdx, camIdx + nCam);
    p.type = type;
    p.drop = false;
    return p;
}
}   // namespace

extern "C" {

/* Build the reference's MVS singleton for `cfg` and inject the cameras (values bit for bit those of the C-ABI records). */
int ref_scene_create(const PmvsConfig *cfg, int nCams, const PmvsCamera *cams, uint64_t seed) {
    Quiet q;
    gSeed = seed;
    omp_set_num_threads(1);                       /* the reference's OpenMP loops race on rand() (psosolver.cpp:222-237) */
    MvsConfig mc;
    static_assert(sizeof(MvsConfig) == sizeof(PmvsConfig), "MvsConfig layout");
    memcpy(&mc, cfg, sizeof(mc));
    MVS &mvs = MVS::getInstance(mc);
    mvs.neighborRadius = cfg->neighborRadius;     /* derived at run time by setNeighborRadius (mvs.cpp:147-152) */
    mvs.cameras.clear();
    for (int i = 0; i < nCams; ++i) {
        const PmvsCamera &c = cams[i];
        Camera cam;
        cam._isAvaliable = true;
        cam.maxLOD = c.maxLOD;
        snprintf(cam.fileName, MAX_FILE_NAME_LENGTH, "cam%d", i);
        cam.focal = Vec2d(c.focal[0], c.focal[1]);
        cam.radialDistortion = 0;
        cam.principlePoint = Vec2d(c.principal[0], c.principal[1]);
        cam.center = Vec3d(c.center[0], c.center[1], c.center[2]);
        cam.rotation = mat3x3(c.R);
        cam.translation = mat3x1(c.t);
        cam.KR = mat3x3(c.KR);
        cam.KT = mat3x1(c.KT);
        cam.opticalNormal = Vec3d(c.opticalNormal[0], c.opticalNormal[1], c.opticalNormal[2]);
        cam.imgPyramid.resize(c.maxLOD + 1);
        cam.edgePyramid.resize(c.maxLOD + 1);
        for (int l = 0; l <= c.maxLOD; ++l) {
            const PmvsLevel &L = c.level[l];
            Mat_<uchar> g(L.rows, L.cols);
            for (int r = 0; r < L.rows; ++r) memcpy(g.ptr<uchar>(r), L.grey + (size_t)r * L.pitch, (size_t)L.cols);
            cam.imgPyramid[l] = g;
            Mat_<double> e(L.rows, L.cols);
            if (L.edge) memcpy(e.data, L.edge, sizeof(double) * (size_t)L.rows * L.cols);
            cam.edgePyramid[l] = e;
        }
        const Mat_<uchar> &g0 = cam.imgPyramid[0];
        cam.imgRGB = Mat_<Vec3b>(g0.rows, g0.cols);
        for (int r = 0; r < g0.rows; ++r)
            for (int x = 0; x < g0.cols; ++x) {
                const uchar v = g0.at<uchar>(r, x);
                cam.imgRGB.at<Vec3b>(r, x) = Vec3b(v, v, v);
            }
        mvs.cameras.push_back(cam);
    }
    return 0;
}

/* threads of the reference's own OpenMP loops (over particles: psosolver.cpp:113,122,222). More than one makes its rand()
 * calls race, as in the reference: timing only. */
int ref_set_threads(int n) {
    omp_set_num_threads(n > 1 ? n : 1);
    return 0;
}

int ref_set_neighbor_radius(double r) {
    MVS::getInstance().neighborRadius = r;
    return 0;
}

/* MVS::initPatchDistanceWeighting (mvs.cpp:97-114) as the reference left it in patchDistWeight */
int ref_dist_weight(double *out) {
    const Mat_<double> &w = MVS::getInstance().getPatchDistanceWeighting();
    for (int x = 0; x < w.rows; ++x)
        for (int y = 0; y < w.cols; ++y) out[x * w.cols + y] = w.at<double>(x, y);
    return w.rows;
}

/* Utility::normal2Spherical / spherical2Normal (utility.h:17-29) */
void ref_normal2spherical(const double *n, double *s) {
    Vec2d o;
    Utility::normal2Spherical(Vec3d(n[0], n[1], n[2]), o);
    s[0] = o[0];
    s[1] = o[1];
}
void ref_spherical2normal(const double *s, double *n) {
    Vec3d o;
    Utility::spherical2Normal(Vec2d(s[0], s[1]), o);
    n[0] = o[0]; n[1] = o[1]; n[2] = o[2];
}

/* Camera::project (camera.cpp:138-160) of camera `cam` at level LOD; returns inImage */
int ref_project(int cam, const double *X, int LOD, double *out) {
    Vec2d p;
    const bool in = MVS::getInstance().getCamera(cam).project(Vec3d(X[0], X[1], X[2]), p, LOD);
    out[0] = p[0];
    out[1] = p[1];
    return in ? 1 : 0;
}

static Patch hypothesis_patch(const PmvsHypothesis &h) {
    const double zero3[3] = {0, 0, 0}, zero2[2] = {0, 0};
    Patch p = make_patch(zero3, zero3, zero2, h.nCam, h.camIdx, Patch::TYPE_EXPAND, 0);
    p.ray = Vec3d(h.ray[0], h.ray[1], h.ray[2]);
    p.refCamIdx = h.refCamIdx;
    p.LOD = h.LOD;
    return p;
}

/* PAIS::getFitness (patch.cpp:914-1047) for n hypotheses */
int ref_fitness_batch(int n, const PmvsHypothesis *in, double *out) {
    Quiet q;
    for (int i = 0; i < n; ++i) {
        Patch p = hypothesis_patch(in[i]);
        Particle pt(3);
        pt.pos[0] = in[i].theta;
        pt.pos[1] = in[i].phi;
        pt.pos[2] = in[i].depth;
        out[i] = PAIS::getFitness(pt, &p);
    }
    return 0;
}

/* Patch::getHomographies (patch.cpp:290-330) for the hypothesis' centre and normal: V x 9 doubles */
int ref_homographies(const PmvsHypothesis *h, double *H) {
    Quiet q;
    Patch p = hypothesis_patch(*h);
    Vec3d normal;
    Utility::spherical2Normal(Vec2d(h->theta, h->phi), normal);
    const Vec3d center = p.getRay() * h->depth + MVS::getInstance().getCamera(h->refCamIdx).getCenter();
    vector<Mat_<double> > Hs;
    p.getHomographies(center, normal, Hs);
    for (int v = 0; v < (int)Hs.size(); ++v)
        for (int k = 0; k < 9; ++k) H[9 * v + k] = Hs[v].at<double>(k / 3, k % 3);
    return (int)Hs.size();
}

/* Patch::getHomographyRegionRatio (patch.cpp:269-288) */
double ref_region_ratio(const double *pt, const double *H) {
    const double zero3[3] = {0, 0, 0}, zero2[2] = {0, 0};
    Patch p = make_patch(zero3, zero3, zero2, 0, NULL, Patch::TYPE_EXPAND, 0);
    return p.getHomographyRegionRatio(Vec2d(pt[0], pt[1]), mat3x3(H));
}

/* seam 2 on the reference itself: [expandVisibleCamera ->] refine() [-> removeInvisibleCamera()], mvs.cpp:214-215 / :572-574 */
int ref_refine_batch(int n, const PmvsPatchIn *in, PmvsPatchOut *out, uint32_t flags) {
    Quiet q;
    for (int i = 0; i < n; ++i) {
        const PmvsPatchIn &a = in[i];
        PmvsPatchOut &o = out[i];
        memset(&o, 0, sizeof(o));
        gPatchId = a.id;
        gRun = 0;
        Patch p = make_patch(a.center, a.normal, a.normalS, a.nCam, a.camIdx, a.type, a.id);
        if ((flags & PMVS_F_EXPAND_VISIBLE) && a.type == PMVS_TYPE_EXPAND) p.expandVisibleCamera();       /* patch.cpp:36-43 */
        p.refine();
        if (flags & PMVS_F_POST_REMOVE_INVISIBLE) p.removeInvisibleCamera();
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
        o.drop = p.drop ? 1 : 0;
        o.psoRuns = gRun;
        o.nCam = (int)p.camIdx.size() > PMVS_MAX_VIEWS ? PMVS_MAX_VIEWS : (int)p.camIdx.size();
        for (int k = 0; k < o.nCam; ++k) o.camIdx[k] = (uint16_t)p.camIdx[k];
        o.nImgPoint = (int)p.imgPoint.size() > PMVS_MAX_VIEWS ? PMVS_MAX_VIEWS : (int)p.imgPoint.size();
        for (int k = 0; k < o.nImgPoint; ++k) {
            o.imgPoint[k][0] = p.imgPoint[k][0];
            o.imgPoint[k][1] = p.imgPoint[k][1];
        }
    }
    return 0;
}


/* ---- the caller side, unmodified: MVS::expansionPatches (mvs.cpp:233-275) with everything under it — expandNeighborCell,
 * expandCell, getExpansionPatchCenter, skipNeighborCell, runtimeFiltering, insertPatch, deletePatch, the queue pops, cell
 * maps, setNeighborRadius — after the seed loop of MVS::refineSeedPatches (:196-231), which is replayed here line by line
 * only because the RNG stream has to be keyed by the seed's id (the reference seeds by wall clock). */
int ref_run_reconstruction(int nSeeds, const PmvsPatchIn *seeds) {
    Quiet q;
    MVS &mvs = MVS::getInstance();
    mvs.patches.clear();
    mvs.deletedPatches.clear();
    mvs.cellMaps.clear();
    mvs.queue.clear();
    for (int i = 0; i < nSeeds; ++i) {
        const PmvsPatchIn &a = seeds[i];
        mvs.patches.insert(std::pair<int, Patch>(a.id, make_patch(a.center, a.normal, a.normalS, a.nCam, a.camIdx, Patch::TYPE_SEED, a.id)));
    }
    AbstractPatch::globalId = nSeeds;                       /* seeds took ids 0..nSeeds-1 (abstractpatch.cpp:7-18) */
    mvs.setNeighborRadius();                                /* refineSeedPatches, mvs.cpp:202 */
    gExpansion = false;
    for (std::map<int, Patch>::iterator it = mvs.patches.begin(); it != mvs.patches.end();) {
        Patch &pth = it->second;
        if (pth.getCameraNumber() < mvs.minCamNum) { it = mvs.deletePatch(pth); continue; }
        gPatchId = pth.getId();
        gRun = 0;
        pth.refine();
        pth.removeInvisibleCamera();
        if (!mvs.runtimeFiltering(pth)) { it = mvs.deletePatch(pth); continue; }
        ++it;
    }
    mvs.setNeighborRadius();
    gExpansion = true;
    gPatchId = -1;
    mvs.expansionPatches();
    gExpansion = false;
    return (int)mvs.patches.size();
}
double ref_neighbor_radius(void) { return MVS::getInstance().neighborRadius; }
/* patch k in id order: id, geometry, scores, cameras, image points; returns 0 past the end */
int ref_get_patch(int k, int *id, double *center, double *normal, double *scores /*fitness, priority, correlation*/, int *nCam, int *camIdx,
                  double *imgPoints, int *expanded) {
    const MVS &mvs = MVS::getInstance();
    if (k < 0 || k >= (int)mvs.patches.size()) return 0;
    std::map<int, Patch>::const_iterator it = mvs.patches.begin();
    std::advance(it, k);
    const Patch &p = it->second;
    *id = p.getId();
    for (int d = 0; d < 3; ++d) { center[d] = p.getCenter()[d]; normal[d] = p.getNormal()[d]; }
    scores[0] = p.getFitness();
    scores[1] = p.getPriority();
    scores[2] = p.getCorrelation();
    *nCam = p.getCameraNumber();
    for (int i = 0; i < *nCam && i < PMVS_MAX_VIEWS; ++i) camIdx[i] = p.getCameraIndices()[i];
    const int np = (int)p.getImagePoints().size();
    for (int i = 0; i < np && i < PMVS_MAX_VIEWS; ++i) { imgPoints[2 * i] = p.getImagePoints()[i][0]; imgPoints[2 * i + 1] = p.getImagePoints()[i][1]; }
    *expanded = p.isExpanded() ? 1 : 0;
    return 1 + np;
}

}   // extern "C"
