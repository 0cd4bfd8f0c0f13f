/*
 * tmvs_lib.cpp — host-side reconstruction driver (see tmvs.h). Every function cites the reference lines it stands for
 * (paths relative to the reference tree). All patch refinement goes through pmvs_refine_batch(); there is no host
 * implementation of refine() here.
 */
#include "tmvs.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <thread>

namespace tmvs {

static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* ---------------------------------------------------------------------------------------------------------
 * config
 * ------------------------------------------------------------------------------------------------------- */
void setInitConfig(MvsConfig &c) {   /* TMVS/TMVS.cpp:26-52 */
    memset(&c, 0, sizeof(c));
    c.cellSize = 4;
    c.patchRadius = 15;
    c.patchSize = 31;
    c.reduceNormalRange = 2;
    c.adaptiveDistanceEnable = 1;
    c.adaptiveDifferenceEnable = 1;
    c.adaptiveGradientEnable = 0;
    c.distWeighting = c.patchRadius / 3.0;
    c.diffWeighting = 128 * 128;
    c.gradientWeighting = 10.0;
    c.minCamNum = 3;
    c.textureVariation = 36;
    c.visibleCorrelation = 0.7;
    c.minCorrelation = 0.7;
    c.maxFitness = 10.0;
    c.minLOD = 0;
    c.maxLOD = 15;
    c.lodRatio = 0.8;
    c.maxCellPatchNum = 3;
    c.neighborRadius = 0.005;
    c.neighborRadiusScalar = 0.0025;
    c.minRegionRatio = 0.55;
    c.depthRangeScalar = 1;
    c.particleNum = 5;
    c.maxIteration = 10;
    c.expansionStrategy = EXPANSION_BEST_FIRST;
}

bool loadConfig(const char *fileName, MvsConfig &c) {   /* TMVS/io/fileloader.cpp:474-565 */
    std::ifstream file(fileName);
    if (!file.is_open()) return false;
    std::string line;
    while (std::getline(file, line)) {
        if (!line.empty() && line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key, val;
        if (!(ss >> key)) continue;
        if (!(ss >> val)) continue;
        const double f = atof(val.c_str());
        const int i = atoi(val.c_str());
        if (key == "patchRadius") { c.patchRadius = i; c.patchSize = (i << 1) + 1; }
        else if (key == "reduceNormalRange") c.reduceNormalRange = f;
        else if (key == "adaptiveDistanceEnable") c.adaptiveDistanceEnable = (uint8_t)(i != 0);
        else if (key == "adaptiveDifferenceEnable") c.adaptiveDifferenceEnable = (uint8_t)(i != 0);
        else if (key == "adaptiveGradientEnable") c.adaptiveGradientEnable = (uint8_t)(i != 0);
        else if (key == "distWeighting") c.distWeighting = f;
        else if (key == "diffWeighting") c.diffWeighting = f;
        else if (key == "gradientWeighting") c.gradientWeighting = f;   /* documented in README.md:140-142; the reference parser forgot it */
        else if (key == "visibleCorrelation") c.visibleCorrelation = f;
        else if (key == "depthRangeScalar") c.depthRangeScalar = f;
        else if (key == "particleNum") c.particleNum = i;
        else if (key == "maxIteration") c.maxIteration = i;
        else if (key == "cellSize") c.cellSize = i;
        else if (key == "maxCellPatchNum") c.maxCellPatchNum = i;
        else if (key == "expansionStrategy") c.expansionStrategy = i;
        else if (key == "textureVariation") c.textureVariation = f;
        else if (key == "minLOD") c.minLOD = i;
        else if (key == "maxLOD") c.maxLOD = i;
        else if (key == "lodRatio") c.lodRatio = f;
        else if (key == "minCamNum") c.minCamNum = i;
        else if (key == "minCorrelation") c.minCorrelation = f;
        else if (key == "minRegionRatio") c.minRegionRatio = f;
        else if (key == "maxFitness") c.maxFitness = f;
        else if (key == "neighborRadiusScalar") c.neighborRadiusScalar = f;
    }
    return true;
}

/* ---------------------------------------------------------------------------------------------------------
 * images (the reference uses cv::imread, camera.cpp:51-69; resize / Sobel run on the GPU, csrc/pmvs_pyramid.cuh)
 * ------------------------------------------------------------------------------------------------------- */
static bool pnmToken(std::istream &in, std::string &tok) {
    tok.clear();
    int ch;
    while ((ch = in.get()) != EOF) {
        if (ch == '#') { while ((ch = in.get()) != EOF && ch != '\n') {} continue; }
        if (isspace(ch)) { if (!tok.empty()) return true; continue; }
        tok.push_back((char)ch);
    }
    return !tok.empty();
}

bool readPnm(const std::string &path, int &cols, int &rows, std::vector<uint8_t> &grey, std::vector<uint8_t> &rgb) {
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in.is_open()) return false;
    std::string magic, t;
    if (!pnmToken(in, magic) || (magic != "P5" && magic != "P6")) return false;
    if (!pnmToken(in, t)) return false;
    cols = atoi(t.c_str());
    if (!pnmToken(in, t)) return false;
    rows = atoi(t.c_str());
    if (!pnmToken(in, t)) return false;   /* the single whitespace after maxval was consumed by pnmToken */
    if (atoi(t.c_str()) != 255 || cols <= 0 || rows <= 0) return false;
    const size_t n = (size_t)cols * rows;
    grey.resize(n);
    rgb.resize(n * 3);
    if (magic == "P5") {
        in.read((char *)grey.data(), (std::streamsize)n);
        if ((size_t)in.gcount() != n) return false;
        for (size_t i = 0; i < n; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = grey[i];
    } else {
        in.read((char *)rgb.data(), (std::streamsize)(n * 3));
        if ((size_t)in.gcount() != n * 3) return false;
        /* cv::imread(file, 0): fixed-point BT.601 as OpenCV's RGB2GRAY (R*4899 + G*9617 + B*1868 + 8192) >> 14 */
        for (size_t i = 0; i < n; ++i) grey[i] = (uint8_t)((rgb[3 * i] * 4899 + rgb[3 * i + 1] * 9617 + rgb[3 * i + 2] * 1868 + 8192) >> 14);
    }
    return true;
}

/* ---------------------------------------------------------------------------------------------------------
 * Camera
 * ------------------------------------------------------------------------------------------------------- */
static void quaternionToRotation(const double q[4], double R[9]) {   /* camera.cpp:6-35 */
    const double qq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double qw = 1, qx = 0, qy = 0, qz = 0;
    if (qq > 0) { qw = q[0] / qq; qx = q[1] / qq; qy = q[2] / qq; qz = q[3] / qq; }
    R[0] = qw * qw + qx * qx - qz * qz - qy * qy;
    R[1] = 2 * qx * qy - 2 * qz * qw;
    R[2] = 2 * qy * qw + 2 * qz * qx;
    R[3] = 2 * qx * qy + 2 * qw * qz;
    R[4] = qy * qy + qw * qw - qz * qz - qx * qx;
    R[5] = 2 * qz * qy - 2 * qx * qw;
    R[6] = 2 * qx * qz - 2 * qy * qw;
    R[7] = 2 * qy * qz + 2 * qw * qx;
    R[8] = qz * qz + qw * qw - qy * qy - qx * qx;
}

bool Camera::project(const double X[3], double out[2], int LOD, double lodRatio) const {
    const double x2 = (R[0] * X[0] + R[1] * X[1] + R[2] * X[2]) + t[0];
    const double y2 = (R[3] * X[0] + R[4] * X[1] + R[5] * X[2]) + t[1];
    const double z2 = (R[6] * X[0] + R[7] * X[1] + R[8] * X[2]) + t[2];
    out[0] = focal[0] * (x2 / z2) + principal[0];
    out[1] = focal[1] * (y2 / z2) + principal[1];
    const double sc = LOD == 0 ? 1.0 : pow(lodRatio, LOD);      /* pow(x, 0) == 1 exactly */
    out[0] *= sc;
    out[1] *= sc;
    if (LOD > maxLOD || std::isnan(out[0]) || std::isnan(out[1])) return false;   /* camera.h:116-131 */
    const int lc = pyramid.empty() ? cols : pyramid[LOD].cols, lr = pyramid.empty() ? rows : pyramid[LOD].rows;
    return !(out[0] < 0 || out[0] >= lc || out[1] < 0 || out[1] >= lr);
}

/* the Camera ctor, camera.cpp:45-136 */
bool MVS::addCamera(Camera &cam, bool loadImage) {
    cam.available = false;
    if (loadImage) {
        std::vector<uint8_t> grey;
        std::string base = imageDir + cam.fileName, stem = base;
        const size_t dot = base.find_last_of('.'), slash = base.find_last_of("/\\");
        if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) stem = base.substr(0, dot);
        const std::string tries[] = {base, stem + ".ppm", stem + ".pgm"};
        bool ok = false;
        for (const std::string &p : tries)
            if (readPnm(p, cam.cols, cam.rows, grey, cam.rgb)) { ok = true; break; }
        if (!ok) {
            err = "can't read image " + base + " (PGM/PPM expected; see tools/convert_images.py)";
            return false;
        }
        /* camera.cpp:63-92: level sizes here, the levels themselves (INTER_AREA resize, edge pyramid) are built on the
         * GPU by pmvs_create from level 0 — the host only ever reads level 0 (runtimeFiltering, mvs.cpp:851-863) */
        int32_t lc[PMVS_MAX_LEVELS], lr[PMVS_MAX_LEVELS];
        if (pmvs_pyramid_levels(cam.cols, cam.rows, cfg.lodRatio, cfg.maxLOD, &cam.maxLOD, lc, lr) != PMVS_OK) {
            err = "bad image size / lodRatio for " + base;
            return false;
        }
        cam.pyramid.resize(cam.maxLOD + 1);
        for (int i = 0; i <= cam.maxLOD; ++i) { cam.pyramid[i].cols = lc[i]; cam.pyramid[i].rows = lr[i]; }
        cam.pyramid[0].grey = grey;
    }
    if (cam.principal[0] < 0 && cam.principal[1] < 0) {   /* camera.cpp:101-106 */
        cam.principal[0] = cam.cols >> 1;
        cam.principal[1] = cam.rows >> 1;
    }
    quaternionToRotation(cam.quaternion, cam.R);
    for (int r = 0; r < 3; ++r) cam.t[r] = -(cam.R[3 * r] * cam.center[0] + cam.R[3 * r + 1] * cam.center[1] + cam.R[3 * r + 2] * cam.center[2]);
    const double K[9] = {cam.focal[0], 0, cam.principal[0], 0, cam.focal[1], cam.principal[1], 0, 0, 1};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) cam.KR[3 * r + c] = K[3 * r] * cam.R[c] + K[3 * r + 1] * cam.R[3 + c] + K[3 * r + 2] * cam.R[6 + c];
        cam.KT[r] = K[3 * r] * cam.t[0] + K[3 * r + 1] * cam.t[1] + K[3 * r + 2] * cam.t[2];
    }
    for (int c = 0; c < 3; ++c) cam.opticalNormal[c] = cam.R[6 + c];   /* R^T * (0,0,1), camera.cpp:130-133 */
    cam.available = true;
    cameras.push_back(cam);
    return true;
}

/* ---------------------------------------------------------------------------------------------------------
 * CellMap
 * ------------------------------------------------------------------------------------------------------- */
void CellMap::init(int imgW, int imgH, int cellSize) {
    width = (int)std::ceil((double)imgW / (double)cellSize);
    height = (int)std::ceil((double)imgH / (double)cellSize);
    cells.assign((size_t)width * height, CellIds());
}
bool CellMap::insert(int x, int y, int id) {
    if (!inMap(x, y)) return false;
    cells[(size_t)y * width + x].push_back(id);
    return true;
}
bool CellMap::drop(int x, int y, int id) {
    if (!inMap(x, y)) return false;
    return cells[(size_t)y * width + x].erase(id);
}

/* ---------------------------------------------------------------------------------------------------------
 * MVS
 * ------------------------------------------------------------------------------------------------------- */
MVS::MVS(const MvsConfig &c) {
    memset(&cfg, 0, sizeof(cfg));
    setConfig(c);
}
MVS::~MVS() {
    for (size_t g = 0; g < ctxs.size(); ++g)
        if (ctxs[g]) pmvs_destroy(ctxs[g]);
}

void MVS::setConfig(const MvsConfig &c) {   /* mvs.cpp:42-72: neighborRadius is NOT copied (derived at run time) */
    const double keep = cfg.neighborRadius;
    cfg = c;
    cfg.patchSize = (cfg.patchRadius << 1) + 1;
    cfg.neighborRadius = keep;
    for (size_t g = 0; g < ctxs.size(); ++g) {
        if (pmvs_set_config(ctxs[g], &cfg) != PMVS_OK) err = pmvs_last_error(ctxs[g]);
        pmvs_set_neighbor_radius(ctxs[g], cfg.neighborRadius);
    }
}

static void normal2Spherical(const double n[3], double s[2]) {   /* utility.h:17-22 */
    s[0] = acos(n[2]);
    s[1] = atan2(n[1], n[0]);
}
static void spherical2Normal(const double s[2], double n[3]) {   /* utility.h:25-29 */
    n[0] = sin(s[0]) * cos(s[1]);
    n[1] = sin(s[0]) * sin(s[1]);
    n[2] = cos(s[0]);
}

void MVS::setEstimatedNormal(Patch &p) const {   /* patch.cpp:390-413 */
    if (p.drop) return;
    if ((int)p.camIdx.size() < cfg.minCamNum) { p.drop = true; return; }
    double n[3] = {0, 0, 0};
    for (size_t i = 0; i < p.camIdx.size(); ++i) {
        const Camera &cam = cameras[p.camIdx[i]];
        double d[3] = {cam.center[0] - p.center[0], cam.center[1] - p.center[1], cam.center[2] - p.center[2]};
        const double inv = 1.0 / sqrt(dot3(d, d));
        for (int k = 0; k < 3; ++k) n[k] += d[k] * inv;
    }
    const double inv = 1.0 / sqrt(dot3(n, n));
    for (int k = 0; k < 3; ++k) p.normal[k] = n[k] * inv;
    normal2Spherical(p.normal, p.normalS);
}

/* symmetric 3x3 pseudo-inverse times b (stands for Mat::inv(DECOMP_SVD)*b at patch.cpp:106) by cyclic Jacobi */
static void solveSym3(const double A[9], const double b[3], double x[3]) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = A[3 * i + j];
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(a[p][q]) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    const double lmax = std::max(fabs(a[0][0]), std::max(fabs(a[1][1]), fabs(a[2][2])));
    x[0] = x[1] = x[2] = 0;
    for (int k = 0; k < 3; ++k) {
        if (fabs(a[k][k]) <= lmax * 3 * DBL_EPSILON) continue;
        const double coef = (v[0][k] * b[0] + v[1][k] * b[1] + v[2][k] * b[2]) / a[k][k];
        for (int i = 0; i < 3; ++i) x[i] += coef * v[i][k];
    }
}

void MVS::reCentering() {   /* mvs.cpp:135-145 -> Patch::reCentering patch.cpp:67-112 */
    for (std::map<int, Patch>::iterator it = patches.begin(); it != patches.end(); ++it) {
        Patch &p = it->second;
        double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
        const int camNum = (int)p.camIdx.size();
        for (int i = 0; i < camNum && 2 * i + 1 < (int)p.imgPoint.size(); ++i) {
            const Camera &cam = cameras[p.camIdx[i]];
            const double q[3] = {(p.imgPoint[2 * i] - cam.principal[0]) / cam.focal[0] - cam.t[0],
                                 (p.imgPoint[2 * i + 1] - cam.principal[1]) / cam.focal[1] - cam.t[1], 1.0 - cam.t[2]};
            double n[3];
            for (int c = 0; c < 3; ++c) n[c] = (cam.R[c] * q[0] + cam.R[3 + c] * q[1] + cam.R[6 + c] * q[2]) - cam.center[c];   /* R^T * (p - t) - C */
            const double inv = 1.0 / sqrt(dot3(n, n));
            for (int c = 0; c < 3; ++c) n[c] *= inv;
            const double *C = cam.center;
            A[0] += 1 - n[0] * n[0]; A[1] += -n[0] * n[1]; A[2] += -n[0] * n[2];
            A[3] += -n[0] * n[1]; A[4] += 1 - n[1] * n[1]; A[5] += -n[1] * n[2];
            A[6] += -n[0] * n[2]; A[7] += -n[1] * n[2]; A[8] += 1 - n[2] * n[2];
            b[0] += (1 - n[0] * n[0]) * C[0] - n[0] * n[1] * C[1] - n[0] * n[2] * C[2];
            b[1] += -n[0] * n[1] * C[0] + (1 - n[1] * n[1]) * C[1] - n[1] * n[2] * C[2];
            b[2] += -n[0] * n[2] * C[0] - n[1] * n[2] * C[1] + (1 - n[2] * n[2]) * C[2];
        }
        solveSym3(A, b, p.center);
        setEstimatedNormal(p);
    }
}

void MVS::setNeighborRadius() {   /* mvs.cpp:147-152 + getBoundingVolume :967-990 */
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (std::map<int, Patch>::const_iterator it = patches.begin(); it != patches.end(); ++it)
        for (int i = 0; i < 3; ++i) {
            mn[i] = std::min(mn[i], it->second.center[i]);
            mx[i] = std::max(mx[i], it->second.center[i]);
        }
    const double volume = fabs((mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]));
    cfg.neighborRadius = pow(volume, 1.0 / 3.0) * cfg.neighborRadiusScalar;
    for (size_t g = 0; g < ctxs.size(); ++g) pmvs_set_neighbor_radius(ctxs[g], cfg.neighborRadius);
    if (verbose) printf("neighborRadius %f\n", cfg.neighborRadius);
}

bool MVS::isNeighbor(const Patch &a, const Patch &b, double neighborRadius) {   /* patch.cpp:6-23 */
    const double d[3] = {a.center[0] - b.center[0], a.center[1] - b.center[1], a.center[2] - b.center[2]};
    double dist = 0;
    dist += fabs(dot3(d, a.normal));
    dist += fabs(dot3(d, b.normal));
    return dist <= neighborRadius;
}

bool MVS::runtimeFiltering(const Patch &p) const {   /* mvs.cpp:838-898 */
    if (p.drop) return false;
    const int camNum = (int)p.camIdx.size();
    if (camNum < cfg.minCamNum) return false;
    if (p.fitness > cfg.maxFitness) return false;
    if (p.fitness == 0.0) return false;
    if (p.priority > 10000) return false;
    if (std::isnan(p.fitness) || std::isnan(p.priority) || std::isnan(p.correlation)) return false;
    if (p.correlation < cfg.minCorrelation) return false;
    for (size_t i = 0; i < cameras.size(); ++i) {   /* inside every image and not on background */
        const Camera &cam = cameras[i];
        double pt[2];
        if (!cam.project(p.center, pt, 0, cfg.lodRatio)) return false;
        if (!cam.pyramid.empty()) {
            /* cvRound may land on cols/rows (the reference then reads one past the row/image): clamp */
            const int x = std::min(cvRound(pt[0]), cam.cols - 1), y = std::min(cvRound(pt[1]), cam.rows - 1);
            if (cam.pyramid[0].grey[(size_t)y * cam.cols + x] == 0) return false;
        }
    }
    int count = 0;
    for (int i = 0; i < camNum; ++i) {
        const double *on = cameras[p.camIdx[i]].opticalNormal;
        const double neg[3] = {-on[0], -on[1], -on[2]};
        if (dot3(p.normal, neg) > 0) count++;
    }
    if (count < cfg.minCamNum) return false;
    if (cellMaps.empty()) return true;
    int fullCellCounter = 0;
    for (int i = 0; i < camNum; ++i) {
        if (2 * i + 1 >= (int)p.imgPoint.size()) continue;
        const int cx = (int)(p.imgPoint[2 * i] / cfg.cellSize), cy = (int)(p.imgPoint[2 * i + 1] / cfg.cellSize);
        const CellMap &m = cellMaps[p.camIdx[i]];
        if (!m.inMap(cx, cy)) continue;   /* the reference indexes the cell unchecked */
        const CellIds &cell = m.cell(cx, cy);
        if (std::find(cell.begin(), cell.end(), p.id) != cell.end()) return true;
        if ((int)cell.size() >= cfg.maxCellPatchNum) ++fullCellCounter;
    }
    if (fullCellCounter >= camNum) return false;
    return true;
}

void MVS::getExpansionPatchCenter(const Camera &cam, const Patch &parent, int cx, int cy, double center[3]) const {   /* mvs.cpp:809-836 */
    const double px = (cx + 0.5) * cfg.cellSize, py = (cy + 0.5) * cfg.cellSize;
    const double q[3] = {(px - cam.principal[0]) / cam.focal[0] - cam.t[0], (py - cam.principal[1]) / cam.focal[1] - cam.t[1], 1.0 - cam.t[2]};
    double v12[3], v13[3];
    for (int c = 0; c < 3; ++c) {
        const double p3d = cam.R[c] * q[0] + cam.R[3 + c] * q[1] + cam.R[6 + c] * q[2];   /* R^T * (p - t) */
        v12[c] = p3d - cam.center[c];
        v13[c] = parent.center[c] - cam.center[c];
    }
    const double u = dot3(parent.normal, v13) / dot3(parent.normal, v12);
    for (int c = 0; c < 3; ++c) center[c] = cam.center[c] + u * v12[c];
}

bool MVS::skipNeighborCell(const CellIds &cell, const Patch &ref) const {   /* mvs.cpp:792-807 */
    const int n = (int)cell.size();
    if (n >= cfg.maxCellPatchNum) return true;
    for (int k = 0; k < n; ++k) {
        const Patch *q = lookup(cell[k]);
        if (!q) continue;
        if (q->correlation > cfg.minCorrelation) return true;
        if (isNeighbor(ref, *q, cfg.neighborRadius)) return true;
    }
    return false;
}

void MVS::setCellMaps() {   /* mvs.cpp:116-133 (+ initCellMaps :74-88) */
    cellMaps.assign(cameras.size(), CellMap());
    for (size_t i = 0; i < cameras.size(); ++i) cellMaps[i].init(cameras[i].cols, cameras[i].rows, cfg.cellSize);
    for (std::map<int, Patch>::iterator it = patches.begin(); it != patches.end(); ++it) {
        const Patch &p = it->second;
        for (size_t i = 0; i < p.camIdx.size() && 2 * i + 1 < p.imgPoint.size(); ++i)
            cellMaps[p.camIdx[i]].insert((int)(p.imgPoint[2 * i] / cfg.cellSize), (int)(p.imgPoint[2 * i + 1] / cfg.cellSize), p.id);
    }
}

void MVS::insertPatch(const Patch &p) {   /* mvs.cpp:579-601 */
    if (!runtimeFiltering(p)) return;
    /* ids grow monotonically during an expansion: the end() hint makes the insertion O(1) */
    const size_t before = patches.size();
    const std::map<int, Patch>::iterator ins = patches.insert(patches.end(), std::pair<int, Patch>(p.id, p));
    if (!idIndex.empty() && patches.size() != before && p.id >= 0) {        /* keep the id index of a running expansion current */
        if ((size_t)p.id >= idIndex.size()) idIndex.resize(std::max((size_t)p.id + 1, idIndex.size() * 2), nullptr);
        idIndex[p.id] = &ins->second;
    }
    queuePush(p.id);
    for (size_t i = 0; i < p.camIdx.size() && 2 * i + 1 < p.imgPoint.size(); ++i)
        cellMaps[p.camIdx[i]].insert((int)(p.imgPoint[2 * i] / cfg.cellSize), (int)(p.imgPoint[2 * i + 1] / cfg.cellSize), p.id);
}

void MVS::deletePatch(int id) {   /* mvs.cpp:607-634 */
    std::map<int, Patch>::iterator it = patches.find(id);
    if (it == patches.end()) return;
    if (!cellMaps.empty()) {
        const Patch &p = it->second;
        for (size_t i = 0; i < p.camIdx.size() && 2 * i + 1 < p.imgPoint.size(); ++i)
            cellMaps[p.camIdx[i]].drop((int)(p.imgPoint[2 * i] / cfg.cellSize), (int)(p.imgPoint[2 * i + 1] / cfg.cellSize), p.id);
    }
    deletedPatches.push_back(it->second);
    if (id >= 0 && (size_t)id < idIndex.size()) idIndex[id] = nullptr;
    patches.erase(it);
}

/* The reference scans the whole queue for every pop (mvs.cpp:656-788), which is quadratic in the number of patches and
 * would dominate once refine() runs on the GPU. Same selection rule, indexed: best/worst-first keep the live entries
 * ordered by (priority, insertion sequence) — the reference's strict comparison picks the earliest-inserted among equal
 * priorities — breadth/depth-first pop from the two ends. Entries whose patch is gone or already expanded are dropped
 * lazily when they surface, like the reference erases them while scanning. */
void MVS::queuePush(int id) {
    queue.push_back(id);
    const Patch *qp = lookup(id);
    const double pr = qp ? qp->priority : 0.0;
    const double key = cfg.expansionStrategy == EXPANSION_WORST_FIRST ? -pr : pr;
    /* never selected by the strict comparisons against the initial +-DBL_MAX (:682, :717): NaN; DBL_MAX and above under
     * best-first; -DBL_MAX and below under worst-first */
    const bool selectable = !std::isnan(pr) && (cfg.expansionStrategy == EXPANSION_WORST_FIRST ? pr > -DBL_MAX : pr < DBL_MAX);
    if (selectable) prioQueue.push(std::make_pair(std::make_pair(key, queueSeq), id));
    fifo.push_back(id);
    ++queueSeq;
}

void MVS::queueClear() {
    queue.clear();
    prioQueue = decltype(prioQueue)();
    fifo.clear();
    queueSeq = 0;
}

int MVS::getPatchIdFromQueue() {
    const bool byPriority = cfg.expansionStrategy != EXPANSION_BREATH_FIRST && cfg.expansionStrategy != EXPANSION_DEPTH_FIRST;
    for (;;) {
        int id;
        if (byPriority) {
            if (prioQueue.empty()) return -1;
            id = prioQueue.top().second;
            prioQueue.pop();
        } else {
            if (fifo.empty()) return -1;
            if (cfg.expansionStrategy == EXPANSION_BREATH_FIRST) { id = fifo.front(); fifo.pop_front(); }
            else { id = fifo.back(); fifo.pop_back(); }
        }
        const Patch *qp = lookup(id);
        if (!qp || qp->expanded) continue;
        return id;
    }
}

/* --- GPU ------------------------------------------------------------------------------------------------ */
/* one context per GPU: the scene (cameras, pyramids, tables) is replicated, patches are sharded */
bool MVS::ensureContext() {
    if (!ctxs.empty() || refineOverride) return true;
    std::vector<PmvsCamera> recs(cameras.size());
    for (size_t i = 0; i < cameras.size(); ++i) {
        const Camera &c = cameras[i];
        PmvsCamera &r = recs[i];
        memset(&r, 0, sizeof(r));
        memcpy(r.focal, c.focal, sizeof(r.focal));
        memcpy(r.principal, c.principal, sizeof(r.principal));
        memcpy(r.center, c.center, sizeof(r.center));
        memcpy(r.R, c.R, sizeof(r.R));
        memcpy(r.t, c.t, sizeof(r.t));
        memcpy(r.KR, c.KR, sizeof(r.KR));
        memcpy(r.KT, c.KT, sizeof(r.KT));
        memcpy(r.opticalNormal, c.opticalNormal, sizeof(r.opticalNormal));
        r.maxLOD = c.maxLOD;
        for (int l = 0; l <= c.maxLOD; ++l) {
            r.level[l].cols = c.pyramid[l].cols;
            r.level[l].rows = c.pyramid[l].rows;
            r.level[l].pitch = c.pyramid[l].cols;
            r.level[l].grey = c.pyramid[l].grey.empty() ? nullptr : c.pyramid[l].grey.data();   /* NULL: built on the device */
            r.level[l].edge = nullptr;
        }
    }
    const std::chrono::steady_clock::time_point tc0 = std::chrono::steady_clock::now();
    /* one context per GPU, created side by side: each uploads the level-0 images and builds its own pyramids (seconds for many
     * large images), so N GPUs cost one creation time, not N */
    const int G = numGpus > 0 ? numGpus : 1;
    std::vector<pmvs_ctx *> made(G, nullptr);
    std::vector<int> rcs(G, PMVS_OK);
    {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g)
            th.emplace_back([&, g]() { rcs[g] = pmvs_create(&made[g], &cfg, (int)recs.size(), recs.data(), device + g, rngSeed); });
        for (size_t k = 0; k < th.size(); ++k) th[k].join();
    }
    for (int g = 0; g < G; ++g)
        if (rcs[g] != PMVS_OK) {
            err = std::string("pmvs_create (device ") + std::to_string(device + g) + "): " + (made[g] ? pmvs_last_error(made[g]) : "out of memory");
            for (int k = 0; k < G; ++k)
                if (made[k]) pmvs_destroy(made[k]);
            return false;
        }
    ctxs.assign(made.begin(), made.end());
    contextSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count();
    return true;
}

void MVS::patchColor(Patch &p) const {   /* patch.cpp:648-652 */
    if (p.refCamIdx < 0 || p.refCamIdx >= (int)cameras.size()) return;
    const Camera &rc = cameras[p.refCamIdx];
    double pt[2];
    if (rc.rgb.empty() || !rc.project(p.center, pt, 0, cfg.lodRatio)) return;
    const int x = std::min(cvRound(pt[0]), rc.cols - 1), y = std::min(cvRound(pt[1]), rc.rows - 1);
    const uint8_t *px = &rc.rgb[((size_t)y * rc.cols + x) * 3];
    p.color[0] = px[2];
    p.color[1] = px[1];
    p.color[2] = px[0];
}

/* Patch::refine() (+ trailing removeInvisibleCamera) for a batch, on the GPU. parentCams != NULL: expansion patches,
 * batch[i]->camIdx is replaced by the parent's list and expandVisibleCamera runs on the device. */
bool MVS::refineBatch(std::vector<Patch *> &batch, unsigned flags, const std::vector<std::vector<int> > *parentCams) {
    if (batch.empty()) return true;
    if (!ensureContext()) return false;
    const int n = (int)batch.size();
    std::vector<PmvsPatchIn> in(n);
    std::vector<PmvsPatchOut> out(n);
    /* marshalling on all host cores: at 10^6 candidates per second the 1.3 KB result records are GB/s of memcpy */
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (int i = 0; i < n; ++i) {
        const Patch &p = *batch[i];
        PmvsPatchIn &r = in[i];
        memset(&r, 0, sizeof(r));
        memcpy(r.center, p.center, sizeof(r.center));
        memcpy(r.normal, p.normal, sizeof(r.normal));
        memcpy(r.normalS, p.normalS, sizeof(r.normalS));
        r.type = p.type;
        r.id = p.id;
        const std::vector<int> &cams = parentCams ? (*parentCams)[i] : p.camIdx;
        if (cams.size() > PMVS_MAX_VIEWS) {
#pragma omp critical(tmvs_warn_views)
            {
                if (!warnedViews) {
                    warnedViews = true;
                    fprintf(stderr, "tmvs: patch %d lists %zu cameras; only the first %d are used (PMVS_MAX_VIEWS)\n", p.id, cams.size(), PMVS_MAX_VIEWS);
                }
            }
        }
        r.nCam = (int)std::min<size_t>(cams.size(), PMVS_MAX_VIEWS);
        for (int k = 0; k < r.nCam; ++k) r.camIdx[k] = (uint16_t)cams[k];
    }
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    /* contiguous shards, one host thread per GPU; results do not depend on the split (the PSO stream is keyed by patch id) */
    const int G = (int)ctxs.size();
    std::vector<int> rcs(G, PMVS_OK);
    if (refineOverride) {             /* CPU tests of the driver's control flow: a stand-in for the library call */
        rcs.assign(1, refineOverride(refineUser, n, in.data(), out.data(), flags));
    } else if (G == 1 || n < 2 * G) rcs[0] = pmvs_refine_batch(ctxs[0], n, in.data(), out.data(), flags);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) {
            const int lo = (int)((long)n * g / G), hi = (int)((long)n * (g + 1) / G);
            th.push_back(std::thread([&, g, lo, hi]() { rcs[g] = pmvs_refine_batch(ctxs[g], hi - lo, in.data() + lo, out.data() + lo, flags); }));
        }
        for (int g = 0; g < G; ++g) th[g].join();
    }
    gpuSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (size_t g = 0; g < rcs.size(); ++g)
        if (rcs[g] != PMVS_OK) { err = std::string("pmvs_refine_batch: ") + (refineOverride ? "stand-in failed" : pmvs_last_error(ctxs[g])); return false; }
    refinedCount += n;
    long sEv = 0, sWev = 0, sIt = 0, sRuns = 0, sDrop = 0, sViews = 0, sLOD[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel for schedule(static) reduction(+ : sEv, sWev, sIt, sRuns, sDrop, sViews, sLOD[:8]) if (n >= 4096)
    for (int i = 0; i < n; ++i) {
        Patch &p = *batch[i];
        const PmvsPatchOut &o = out[i];
        memcpy(p.center, o.center, sizeof(p.center));
        memcpy(p.normal, o.normal, sizeof(p.normal));
        memcpy(p.normalS, o.normalS, sizeof(p.normalS));
        p.fitness = o.fitness;
        p.priority = o.priority;
        p.correlation = o.correlation;
        p.LOD = o.LOD;
        p.refCamIdx = o.refCamIdx;
        p.drop = o.drop != 0;
        sEv += o.evaluations;
        sWev += o.windowEvaluations;
        sIt += o.psoIterations;
        sRuns += o.psoRuns;
        sDrop += o.drop != 0;
        sViews += o.nCam;
        if (o.LOD >= 0 && o.LOD < 8) sLOD[o.LOD]++;
        p.camIdx.assign(o.camIdx, o.camIdx + o.nCam);
        p.imgPoint.resize((size_t)o.nImgPoint * 2);
        for (int k = 0; k < o.nImgPoint; ++k) { p.imgPoint[2 * k] = o.imgPoint[k][0]; p.imgPoint[2 * k + 1] = o.imgPoint[k][1]; }
        if (o.nImgPoint > 0) patchColor(p);
    }
    statEvaluations += sEv;
    statWindowEvaluations += sWev;
    statIterations += sIt;
    statRuns += sRuns;
    statDropped += sDrop;
    statViews += sViews;
    for (int l = 0; l < 8; ++l) statLOD[l] += sLOD[l];
    return true;
}

bool MVS::refineSeedPatches() {   /* mvs.cpp:196-231 */
    if (patches.empty()) { printf("No seed patches\n"); return true; }
    setNeighborRadius();
    std::vector<int> few;
    std::vector<Patch *> batch;
    for (std::map<int, Patch>::iterator it = patches.begin(); it != patches.end(); ++it) {
        if ((int)it->second.camIdx.size() < cfg.minCamNum) few.push_back(it->first);
        else batch.push_back(&it->second);
    }
    for (size_t k = 0; k < few.size(); ++k) deletePatch(few[k]);
    if (!ensureContext()) return false;
    for (size_t g = 0; g < ctxs.size(); ++g) pmvs_set_neighbor_radius(ctxs[g], cfg.neighborRadius);
    if (!refineBatch(batch, PMVS_F_POST_REMOVE_INVISIBLE, nullptr)) return false;
    std::vector<int> bad;
    for (std::map<int, Patch>::iterator it = patches.begin(); it != patches.end(); ++it)
        if (!runtimeFiltering(it->second)) bad.push_back(it->first);
    for (size_t k = 0; k < bad.size(); ++k) deletePatch(bad[k]);
    setNeighborRadius();
    return true;
}

bool MVS::expansionPatches() {   /* mvs.cpp:233-275, 529-577 — in rounds */
    setCellMaps();
    IdIndexGuard index(*this);                     /* skipNeighborCell looks every member of every visited cell up */
    queueClear();                                  /* initPriorityQueue, mvs.cpp:90-95 */
    for (std::map<int, Patch>::const_iterator it = patches.begin(); it != patches.end(); ++it) queuePush(it->first);
    setNeighborRadius();
    if (!ensureContext()) return false;
    size_t saveTime = 0;
    struct Cand { int parent, cam, cx, cy; };
    typedef std::chrono::steady_clock Clock;
    double tPop = 0, tGen = 0, tGenOpen = 0, tCommit = 0, tSave = 0;
    Clock::time_point lastSave = Clock::now();
    long gpuCalls = 0, refinedMax = 0, generatedCands = 0;
    /* per-pass scratch over the cell maps, reset lazily by a pass stamp: candidates already generated for a cell, and
     * (merged mode) the chain of candidates expected to land in it. Two generations are kept: the pass being generated and
     * the pass whose candidates are in flight on the GPUs (pipelined rounds) — what that round is expected to fill is
     * left alone too, so a round does not refine the 3-D neighbours its predecessor is already refining */
    struct CellRec { int stamp[2], pend[2], head[2]; };     /* one record per cell: one cache line per visit */
    typedef std::vector<CellRec> CellScratch;
    std::vector<CellScratch> scratch(cameras.size());
    for (size_t i = 0; i < cameras.size(); ++i) {
        const size_t nc = (size_t)cellMaps[i].width * cellMaps[i].height;
        const CellRec fresh = {{-1, -1}, {0, 0}, {-1, -1}};
        scratch[i].assign(nc, fresh);
    }
    struct Snap { double c[3], n[3]; };             /* a candidate as generated: unrefined centre, its parent's normal */
    std::vector<std::pair<int, int> > chain[2];    /* (candidate index, next entry) */
    std::vector<Snap> snap[2];
    int passStamp = 0, inflightPass = -1;
    /* merged mode: parents whose expectation failed (a candidate of theirs was rejected) try their remaining
     * (slot, neighbour) combinations in the next round's pass, like the serial reference moves on to the next camera */
    struct Carry { int parent; std::vector<unsigned char> tried; };
    std::vector<Carry> carry;
    /* One round's working set. In the merged mode two rounds are in flight (below): the GPUs refine round k+1 while the host
     * commits round k and pops and generates round k+2 (mvs.cpp:233-275 restructured). */
    struct RoundWork {
        std::vector<int> parents;
        std::vector<std::vector<unsigned char> > tried;     /* per parent: (slot, neighbour) combinations already refined */
        std::vector<int> nGenerated, nRejected, candParentIdx;
        std::vector<Cand> cands;
        std::vector<Patch> cpatch;
        std::vector<std::vector<int> > parentCams;
        size_t maxSlots = 0, nCands = 0, accepted = 0;
        int pass = -1;                                      /* stamp of the generate() pass that produced the candidates */
        std::vector<Patch *> batch;
        bool ok = true;
    };
    auto popParents = [&](RoundWork &W) {
        /* 1. pop up to roundSize parents in strategy order */
        Clock::time_point tp0 = Clock::now();
        for (size_t k = 0; k < carry.size(); ++k)
            if (lookup(carry[k].parent)) { W.parents.push_back(carry[k].parent); W.tried.push_back(carry[k].tried); }
        carry.clear();
        const size_t nCarried = W.parents.size();
        while ((int)(W.parents.size() - nCarried) < roundSize) {
            const int id = getPatchIdFromQueue();
            if (id < 0) break;
            Patch *pp = const_cast<Patch *>(lookup(id));
            if (!pp) continue;
            pp->expanded = true;
            if (!runtimeFiltering(*pp)) { deletePatch(id); continue; }   /* mvs.cpp:255-260 */
            W.parents.push_back(id);
        }
        tPop += std::chrono::duration<double>(Clock::now() - tp0).count();
        if (W.parents.empty()) return;
        W.maxSlots = 0;
        for (size_t k = 0; k < W.parents.size(); ++k) {
            const Patch *pp = lookup(W.parents[k]);
            if (pp) W.maxSlots = std::max(W.maxSlots, pp->camIdx.size());
        }
        W.tried.resize(W.parents.size());
        for (size_t k = 0; k < W.parents.size(); ++k) W.tried[k].resize(W.maxSlots * 4, 0);
        W.nGenerated.assign(W.parents.size(), 0);
        W.nRejected.assign(W.parents.size(), 0);
    };
        /* 2.-4. The reference visits a parent's visible cameras one after the other (expandNeighborCell,
         * mvs.cpp:535-563) and inserts each refined candidate before looking at the next camera, so the (up to) five
         * views of the same 3-D neighbour are refined once: the first one fills the cells the others would target.
         * Refining every candidate of every slot blindly would refine that neighbour five times (2.7x the work).
         * --slot-passes: one pass per camera slot i — candidates of slot i for all parents -> GPU -> serial commit in
         * parent order -> slot i+1 sees the updated cell maps (the reference's visiting order; the passes of slots 1..
         * are sparse, each paying a full kernel latency). Default (mergeSlots): one pass, see below. */
        /* mergeSlots: ONE pass per round over all camera slots (slot-major, the order the per-slot passes commit in).
         * What the per-slot passes learn from the commits in between — "this cell now holds a neighbour of the parent"
         * (skipNeighborCell, mvs.cpp:803-804) — is predicted instead: a candidate is expected to land in the cells its
         * unrefined centre projects to in its parent's cameras, with its parent's normal. A later candidate whose
         * target cell holds such an expected neighbour of ITS parent is not generated. Mispredictions (refinement moved
         * the centre across a cell border, or the expected patch was rejected) cost a wasted refinement or leave a cell
         * to a later round; the commit below still re-checks every target cell against the real state. The sparse
         * passes of slots 1.. (tens of candidates, a full kernel latency each) disappear: 5x fewer, 5x larger calls. */
    /* candidates of the camera slots [slot0, slot1) of W's parents; false: no parent has such a slot */
    auto generate = [&](RoundWork &W, size_t slot0, size_t slot1) -> bool {
        W.cands.clear();
        W.cpatch.clear();
        W.parentCams.clear();
        W.candParentIdx.clear();
        {
            Clock::time_point tg0 = Clock::now();
            /* a cell can take at most maxCellPatchNum patches (skipNeighborCell, mvs.cpp:794-795): do not refine more
             * candidates for a cell than it still has room for — the serial reference would have skipped them */
            ++passStamp;
            if (inflightPass >= 0 && ((inflightPass ^ passStamp) & 1) == 0) ++passStamp;      /* the two generations use different slots */
            const int cs = passStamp & 1, ps = cs ^ 1;
            chain[cs].clear();
            snap[cs].clear();
            auto cellAt = [&](int cam, int x, int y) -> CellRec & {
                CellRec &sc = scratch[cam][(size_t)y * cellMaps[cam].width + x];
                if (sc.stamp[cs] != passStamp) { sc.stamp[cs] = passStamp; sc.pend[cs] = 0; sc.head[cs] = -1; }
                return sc;
            };
            auto neighborOfSnap = [&](const Patch &a, const Snap &b) -> bool {      /* isNeighbor (patch.cpp:6-23) against a snapshot */
                const double d[3] = {a.center[0] - b.c[0], a.center[1] - b.c[1], a.center[2] - b.c[2]};
                double dist = 0;
                dist += fabs(dot3(d, a.normal));
                dist += fabs(dot3(d, b.n));
                return dist <= cfg.neighborRadius;
            };
            /* the read-only part of the visit — is the neighbour cell inside the map and not to be skipped
             * (skipNeighborCell, mvs.cpp:792-807: the patch look-ups and isNeighbor tests) — for every (parent, slot,
             * neighbour) on all host cores; the order-dependent part below stays serial */
            const size_t nSl = slot1 - slot0;
            /* per (parent, slot): camera and cell of the parent's image point + which of the four neighbour cells are open —
             * flat arrays, so the serial pass below walks memory in order and touches a parent only for its open cells */
            struct Visit { int cam, cx, cy; unsigned open; };
            std::vector<Visit> visit(W.parents.size() * nSl);
            std::vector<const Patch *> parentOf(W.parents.size());
#pragma omp parallel for schedule(dynamic, 32) if (W.parents.size() >= 128)
            for (long k = 0; k < (long)W.parents.size(); ++k) {
                const Patch *pp = lookup(W.parents[k]);          /* O(1) id index: a std::map walk per (slot, parent) was the largest host cost */
                parentOf[k] = pp;
                for (size_t slot = slot0; slot < slot1; ++slot) {
                    Visit &vs = visit[(size_t)k * nSl + (slot - slot0)];
                    vs.cam = -1;
                    vs.open = 0;
                    if (!pp) continue;
                    const Patch &pth = *pp;
                    if (slot >= pth.camIdx.size() || 2 * slot + 1 >= pth.imgPoint.size()) continue;
                    const CellMap &m = cellMaps[pth.camIdx[slot]];
                    const int cx = (int)(pth.imgPoint[2 * slot] / cfg.cellSize), cy = (int)(pth.imgPoint[2 * slot + 1] / cfg.cellSize);
                    const int nx[4] = {cx - 1, cx, cx + 1, cx}, ny[4] = {cy, cy - 1, cy, cy + 1};
                    vs.cam = pth.camIdx[slot];
                    vs.cx = cx;
                    vs.cy = cy;
                    for (int j = 0; j < 4; ++j)
                        if (m.inMap(nx[j], ny[j]) && !skipNeighborCell(m.cell(nx[j], ny[j]), pth)) vs.open |= 1u << j;
                }
            }
            tGenOpen += std::chrono::duration<double>(Clock::now() - tg0).count();
            bool anySlot = false;
            for (size_t slot = slot0; slot < slot1; ++slot)
            for (size_t k = 0; k < W.parents.size(); ++k) {
                const Visit &vs = visit[k * nSl + (slot - slot0)];
                if (vs.cam < 0) continue;
                anySlot = true;
                if (!vs.open) continue;
                const Patch &pth = *parentOf[k];
                const int ci = vs.cam;
                const CellMap &m = cellMaps[ci];
                const int cx = vs.cx, cy = vs.cy;
                const int nx[4] = {cx - 1, cx, cx + 1, cx}, ny[4] = {cy, cy - 1, cy, cy + 1};
                for (int j = 0; j < 4; ++j) {
                    if (!((vs.open >> j) & 1u)) continue;
                    if (W.tried[k][slot * 4 + j]) continue;
                    CellRec &cr = cellAt(ci, nx[j], ny[j]);
                    const bool prevLive = inflightPass >= 0 && cr.stamp[ps] == inflightPass;
                    if ((int)m.cell(nx[j], ny[j]).size() + cr.pend[cs] + (prevLive ? cr.pend[ps] : 0) >= cfg.maxCellPatchNum) continue;
                    if (mergeSlots) {
                        bool taken = false;
                        for (int q = cr.head[cs]; q >= 0 && !taken; q = chain[cs][q].second)
                            taken = neighborOfSnap(pth, snap[cs][chain[cs][q].first]);
                        if (prevLive)
                            for (int q = cr.head[ps]; q >= 0 && !taken; q = chain[ps][q].second)
                                taken = neighborOfSnap(pth, snap[ps][chain[ps][q].first]);
                        if (taken) continue;
                    }
                    ++cr.pend[cs];
                    Patch e;                                   /* Patch(center, parent), patch.cpp:36-43 */
                    e.type = PMVS_TYPE_EXPAND;
                    e.id = nextId++;
                    getExpansionPatchCenter(cameras[ci], pth, nx[j], ny[j], e.center);
                    memcpy(e.normal, pth.normal, sizeof(e.normal));
                    normal2Spherical(e.normal, e.normalS);
                    const Cand c = {W.parents[k], ci, nx[j], ny[j]};
                    W.tried[k][slot * 4 + j] = 1;
                    W.nGenerated[k]++;
                    W.candParentIdx.push_back((int)k);
                    W.cands.push_back(c);
                    W.cpatch.push_back(e);
                    W.parentCams.push_back(pth.camIdx);
                    if (mergeSlots) {
                        Snap sn;
                        memcpy(sn.c, e.center, sizeof(sn.c));
                        memcpy(sn.n, pth.normal, sizeof(sn.n));
                        snap[cs].push_back(sn);
                        const int me = (int)snap[cs].size() - 1;
                        for (size_t v = 0; v < pth.camIdx.size(); ++v) {
                            double pt[2];
                            const int cv = pth.camIdx[v];
                            if (!cameras[cv].project(e.center, pt, 0, cfg.lodRatio)) continue;
                            const int ex = (int)(pt[0] / cfg.cellSize), ey = (int)(pt[1] / cfg.cellSize);
                            if (!cellMaps[cv].inMap(ex, ey)) continue;
                            CellRec &er = cellAt(cv, ex, ey);
                            chain[cs].push_back(std::make_pair(me, er.head[cs]));
                            er.head[cs] = (int)chain[cs].size() - 1;
                        }
                    }
                }
            }
            W.pass = passStamp;
            tGen += std::chrono::duration<double>(Clock::now() - tg0).count();
            return anySlot;
        }
    };
    /* serial commit in parent order; the target cell is re-checked as the reference would have seen it */
    auto commit = [&](RoundWork &W) {
        {
            Clock::time_point tc0 = Clock::now();
            for (size_t k = 0; k < W.cands.size(); ++k) {
                const Patch *pp = lookup(W.cands[k].parent);
                if (!pp) continue;
                if (skipNeighborCell(cellMaps[W.cands[k].cam].cell(W.cands[k].cx, W.cands[k].cy), *pp)) continue;
                const size_t before = patches.size();
                insertPatch(W.cpatch[k]);
                W.accepted += patches.size() - before;
                if (patches.size() == before) W.nRejected[W.candParentIdx[k]]++;      /* runtimeFiltering turned it down */
            }
            W.nCands += W.cands.size();
            tCommit += std::chrono::duration<double>(Clock::now() - tc0).count();
        }
    };
    auto finishRound = [&](RoundWork &W, int round) {
        if (mergeSlots)
            for (size_t k = 0; k < W.parents.size(); ++k)
                if (W.nRejected[k] > 0 && W.nGenerated[k] > 0) { Carry c; c.parent = W.parents[k]; c.tried.swap(W.tried[k]); carry.push_back(c); }
        if (verbose)
            printf("round %d: parents %zu candidates %zu accepted %zu patches %zu queue %zu\n", round, W.parents.size(), W.nCands, W.accepted,
                   patches.size(), byPriorityQueueSize());
        /* mvs.cpp:265-268 checkpoints every 500 accepted patches; at GPU speed that is every few milliseconds and the
         * rewrite of the whole file becomes quadratic, so checkpoints are additionally spaced autosaveSeconds apart */
        if (patches.size() / 500 > saveTime && std::chrono::duration<double>(Clock::now() - lastSave).count() >= autosaveSeconds) {
            Clock::time_point ts0 = Clock::now();
            lastSave = ts0;
            saveTime = patches.size() / 500;
            writeMVS((outDir + "auto_save.mvs").c_str());
            tSave += std::chrono::duration<double>(Clock::now() - ts0).count();
        }
    };
    const unsigned refineFlags = PMVS_F_EXPAND_VISIBLE | PMVS_F_POST_REMOVE_INVISIBLE;     /* mvs.cpp:572-574 on the GPU */
    if (mergeSlots && pipelineRounds) {
        /* Two rounds in flight. While the GPUs refine round k+1 the host commits round k and then pops and generates round
         * k+2 from that state; round k+1 itself was generated while round k was on the GPUs, knowing what round k was
         * expected to fill (the two-generation scratch above). Nothing waits for the commit: the commit's own re-check
         * (skipNeighborCell, mvs.cpp:792-807) turns the rare duplicate down. */
        std::thread gpu;
        auto launch = [&](RoundWork *W) {
            ++gpuCalls;
            refinedMax = std::max(refinedMax, (long)W->cands.size());
            generatedCands += (long)W->cands.size();
            W->batch.resize(W->cpatch.size());
            for (size_t k = 0; k < W->cpatch.size(); ++k) W->batch[k] = &W->cpatch[k];
            inflightPass = W->pass;
            gpu = std::thread([this, W, refineFlags]() { W->ok = refineBatch(W->batch, refineFlags, &W->parentCams); });
        };
        auto popAndGenerate = [&]() -> std::unique_ptr<RoundWork> {
            std::unique_ptr<RoundWork> W(new RoundWork());
            popParents(*W);
            if (!W->parents.empty()) generate(*W, 0, W->maxSlots);
            return W;
        };
        std::unique_ptr<RoundWork> cur = popAndGenerate();
        bool curLaunched = false;
        if (!cur->cands.empty()) { launch(cur.get()); curLaunched = true; }
        std::unique_ptr<RoundWork> next = popAndGenerate();
        for (int round = 0; !cur->parents.empty(); ++round) {
            if (curLaunched) {
                gpu.join();
                inflightPass = -1;
                if (!cur->ok) return false;
            }
            bool nextLaunched = false;
            if (!next->cands.empty()) { launch(next.get()); nextLaunched = true; }
            if (curLaunched) commit(*cur);
            finishRound(*cur, round);
            std::unique_ptr<RoundWork> after = popAndGenerate();
            cur = std::move(next);
            curLaunched = nextLaunched;
            next = std::move(after);
            if (cur->parents.empty() && !next->parents.empty()) {      /* the queue was empty until the last commit refilled it */
                cur = std::move(next);
                curLaunched = false;
                if (!cur->cands.empty()) { launch(cur.get()); curLaunched = true; }
                next = popAndGenerate();
            }
        }
        if (gpu.joinable()) gpu.join();
    } else
    for (int round = 0;; ++round) {
        RoundWork W;
        popParents(W);
        if (W.parents.empty()) break;
        for (size_t slot0 = 0; slot0 < W.maxSlots; slot0 = mergeSlots ? W.maxSlots : slot0 + 1) {
            const size_t slot1 = mergeSlots ? W.maxSlots : slot0 + 1;
            if (!generate(W, slot0, slot1)) break;
            if (W.cands.empty()) continue;
            ++gpuCalls;
            refinedMax = std::max(refinedMax, (long)W.cands.size());
            std::vector<Patch *> batch(W.cpatch.size());
            for (size_t k = 0; k < W.cpatch.size(); ++k) batch[k] = &W.cpatch[k];
            if (!refineBatch(batch, refineFlags, &W.parentCams)) return false;
            commit(W);
        }
        finishRound(W, round);
    }

    printf("expansion host seconds: pop %.3f generate %.3f commit %.3f auto_save %.3f; gpu calls %ld (largest %ld candidates)\n", tPop, tGen, tCommit,
           tSave, gpuCalls, refinedMax);
    if (verbose || getenv("TMVS_HOST_TIMERS"))
        printf("generate: %.3f s of it in the parallel open-cell scan; pipelined rounds refined %ld candidates\n", tGenOpen, generatedCands);
    setNeighborRadius();
    return true;
}

/* ---------------------------------------------------------------------------------------------------------
 * `-f` filters (mvs.cpp:279-525)
 * ------------------------------------------------------------------------------------------------------- */
MVS::IdIndexGuard::IdIndexGuard(const MVS &mvs) : m(mvs) {
    m.idIndex.clear();
    if (m.patches.empty()) return;
    m.idIndex.assign((size_t)m.patches.rbegin()->first + 1, nullptr);
    for (std::map<int, Patch>::const_iterator it = m.patches.begin(); it != m.patches.end(); ++it)
        if (it->first >= 0) m.idIndex[it->first] = &it->second;
}

void MVS::ensureFilterMaps() {   /* mvs.cpp:280-283, :328-331, :400-403, :449-452 */
    if (cellMaps.empty()) {
        setNeighborRadius();
        setCellMaps();
    }
}

void MVS::cellFiltering() {   /* mvs.cpp:279-325 */
    ensureFilterMaps();
    IdIndexGuard index(*this);
    for (size_t i = 0; i < cameras.size(); ++i) {
        CellMap &map = cellMaps[i];
        for (int x = 0; x < map.width; ++x)
            for (int y = 0; y < map.height; ++y) {
                const CellIds &cell = map.cell(x, y);
                const int pthNum = (int)cell.size();
                std::vector<int> removeIdx;
                for (int j = 0; j < pthNum; ++j) {
                    double corrSum = 0;
                    for (int k = 0; k < pthNum; ++k) {
                        if (j == k) continue;
                        const Patch *pk = lookup(cell[k]);
                        if (!pk) continue;
                        corrSum += pk->correlation;
                    }
                    const Patch *pj = lookup(cell[j]);
                    if (!pj) continue;
                    if (pj->correlation * (double)pj->camIdx.size() < corrSum) removeIdx.push_back(cell[j]);
                }
                for (size_t j = 0; j < removeIdx.size(); ++j) deletePatch(removeIdx[j]);
            }
    }
}

void MVS::neighborCellFiltering(double neighborRatio) {   /* mvs.cpp:327-397 */
    ensureFilterMaps();
    IdIndexGuard index(*this);
    for (size_t i = 0; i < cameras.size(); ++i) {
        CellMap &map = cellMaps[i];
        for (int x = 0; x < map.width; ++x)
            for (int y = 0; y < map.height; ++y) {
                const CellIds &cell = map.cell(x, y);
                std::vector<int> removeIdx;
                const int nx[9] = {x, x - 1, x + 1, x - 1, x + 1, x + 1, x, x - 1, x};
                const int ny[9] = {y, y - 1, y - 1, y + 1, y + 1, y, y + 1, y, y - 1};
                const int pthNum = (int)cell.size();
                for (int j = 0; j < pthNum; ++j) {
                    const Patch *pc = lookup(cell[j]);
                    if (!pc) continue;
                    const Patch &centerPth = *pc;
                    int neighborPthSum = 0, neighborPthNum = 0;
                    for (int q = 0; q < 9; ++q) {
                        if (!map.inMap(nx[q], ny[q])) continue;
                        const CellIds &neighborCell = map.cell(nx[q], ny[q]);
                        neighborPthSum += (int)neighborCell.size();
                        for (size_t k = 0; k < neighborCell.size(); ++k) {
                            const Patch *pn = lookup(neighborCell[k]);
                            if (!pn) continue;
                            if (isNeighbor(centerPth, *pn, cfg.neighborRadius)) ++neighborPthNum;
                        }
                    }
                    if ((double)neighborPthNum / (double)neighborPthSum < neighborRatio) removeIdx.push_back(centerPth.id);
                }
                for (size_t k = 0; k < removeIdx.size(); ++k) deletePatch(removeIdx[k]);
            }
    }
}

void MVS::visibilityFiltering() {   /* mvs.cpp:399-446 */
    ensureFilterMaps();
    IdIndexGuard index(*this);
    for (std::map<int, Patch>::iterator it = patches.begin(); it != patches.end();) {
        const Patch &pth = it->second;
        const int camNum = (int)pth.camIdx.size();
        int visibleCount = camNum;
        for (int i = 0; i < camNum && 2 * (size_t)i + 1 < pth.imgPoint.size(); ++i) {
            const Camera &cam = cameras[pth.camIdx[i]];
            const double d[3] = {pth.center[0] - cam.center[0], pth.center[1] - cam.center[1], pth.center[2] - cam.center[2]};
            const double depth = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                  /* cv::norm */
            const int cx = (int)(pth.imgPoint[2 * i] / cfg.cellSize), cy = (int)(pth.imgPoint[2 * i + 1] / cfg.cellSize);
            const CellMap &map = cellMaps[pth.camIdx[i]];
            if (!map.inMap(cx, cy)) continue;          /* the reference indexes the cell unchecked (out of range there is undefined) */
            const CellIds &cell = map.cell(cx, cy);
            for (size_t p = 0; p < cell.size(); ++p) {
                if (cell[p] == pth.id) continue;
                const Patch *pn = lookup(cell[p]);
                if (!pn) continue;
                const double e[3] = {pn->center[0] - cam.center[0], pn->center[1] - cam.center[1], pn->center[2] - cam.center[2]};
                const double neighborDepth = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
                if (depth > neighborDepth) { --visibleCount; break; }
            }
        }
        if (visibleCount < cfg.minCamNum) {
            const int id = pth.id;
            ++it;                                   /* std::map::erase only invalidates the erased element */
            deletePatch(id);
            continue;
        }
        ++it;
    }
}

bool MVS::neighborPatchFiltering(double neighborRatio) {   /* mvs.cpp:448-525 */
    ensureFilterMaps();
    const int n = (int)patches.size();
    if (n == 0) return true;
    std::vector<int> ids;
    std::vector<double> centers;
    ids.reserve(n);
    centers.reserve(3 * (size_t)n);
    for (std::map<int, Patch>::const_iterator it = patches.begin(); it != patches.end(); ++it) {
        ids.push_back(it->first);
        centers.insert(centers.end(), it->second.center, it->second.center + 3);
    }
    /* the reference's per-patch distance list + sort + radius cut (:470-499) only feeds a count: one GPU pair scan,
     * rows sharded over the devices (independent units, no exchange) */
    std::vector<int> counts((size_t)n, 0);
    const int G = numGpus > 0 ? numGpus : 1;
    std::vector<int> rc((size_t)G, PMVS_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        const int base = n / G, extra = n % G;
        const int first = g * base + std::min(g, extra), count = base + (g < extra ? 1 : 0);
        th.emplace_back([&, g, first, count]() {
            rc[g] = pmvs_neighbor_counts(device + g, n, centers.data(), cfg.neighborRadius, first, count, counts.data() + first);
        });
    }
    for (size_t g = 0; g < th.size(); ++g) th[g].join();
    for (int g = 0; g < G; ++g)
        if (rc[g] != PMVS_OK) { err = "pmvs_neighbor_counts failed (no CUDA device? there is no CPU path)"; return false; }
    double avg = 0;                                                  /* :502-506 */
    for (int i = 0; i < n; ++i) avg += (double)counts[i];
    avg /= (double)n;
    lastAvgNeighborNum = avg;
    printf("\naverage neighbor number: %f\n", avg);
    /* the reference pushes PatchNeighbor records from OpenMP threads (:497-501), so its deletion order — and with it
     * the order of deletedPatches — is nondeterministic; the SET removed is not. Here: ascending id. */
    for (int i = 0; i < n; ++i)
        if ((double)counts[i] < avg * neighborRatio) deletePatch(ids[i]);
    return true;
}

static void writePatchRecord(std::ofstream &file, const Patch &p) {   /* filewriter.cpp:49-69 */
    const int camNum = (int)p.camIdx.size();
    file.write((const char *)p.center, 3 * sizeof(double));
    file.write((const char *)p.normalS, 2 * sizeof(double));
    file.write((const char *)&camNum, sizeof(int));
    for (int k = 0; k < camNum; ++k) file.write((const char *)&p.camIdx[k], sizeof(int));
    file.write((const char *)&p.fitness, sizeof(double));
    file.write((const char *)&p.correlation, sizeof(double));
}

bool MVS::writeDeletedPatchMVS(const char *fileName) const {   /* filewriter.cpp:173-204 */
    std::ofstream file(fileName, std::ios::binary);
    if (!file.is_open()) return false;
    file << "MVS_V3" << "\n";
    file.write((const char *)&cfg, sizeof(MvsConfig));
    file << "CAMERAS " << (int)cameras.size() << "\n";
    for (size_t i = 0; i < cameras.size(); ++i) {
        const Camera &c = cameras[i];
        const int len = (int)c.fileName.size();
        file.write((const char *)&len, sizeof(int));
        file.write(c.fileName.data(), len);
        file.write((const char *)c.center, 3 * sizeof(double));
        file.write((const char *)c.focal, 2 * sizeof(double));
        file.write((const char *)c.principal, 2 * sizeof(double));
        file.write((const char *)c.quaternion, 4 * sizeof(double));
        file.write((const char *)&c.radialDistortion, sizeof(double));
    }
    file << "PATCHES " << (int)deletedPatches.size() << "\n";
    for (size_t i = 0; i < deletedPatches.size(); ++i) writePatchRecord(file, deletedPatches[i]);
    return (bool)file;
}

/* one vertex line of filewriter.cpp:128-135: `file << double` with the default ostream state is printf's %g (precision
 * 6), the colour goes out r g b from the b,g,r pixel. Formatted into one buffer per file instead of nine stream
 * insertions per patch (tests/test_host_cpu.py compares the bytes). */
static void plyLine(std::string &buf, const Patch &p) {
    char line[256];
    const int n = snprintf(line, sizeof(line), "%g %g %g %g %g %g %d %d %d\n", p.center[0], p.center[1], p.center[2], p.normal[0], p.normal[1],
                           p.normal[2], int(p.color[2]), int(p.color[1]), int(p.color[0]));
    buf.append(line, (size_t)n);
}

bool MVS::writeDeletedPatchPLY(const char *fileName) const {   /* filewriter.cpp:206-241 */
    std::ofstream file(fileName);
    if (!file.is_open()) return false;
    file << "ply\nformat ascii 1.0\nelement vertex " << deletedPatches.size() << "\n";
    file << "property float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n";
    file << "property uchar diffuse_red\nproperty uchar diffuse_green\nproperty uchar diffuse_blue\nend_header\n";
    std::string buf;
    buf.reserve(deletedPatches.size() * 80);
    for (size_t i = 0; i < deletedPatches.size(); ++i) plyLine(buf, deletedPatches[i]);
    file.write(buf.data(), (std::streamsize)buf.size());
    return (bool)file;
}

/* ---------------------------------------------------------------------------------------------------------
 * loaders / writers
 * ------------------------------------------------------------------------------------------------------- */
static std::string dirOf(const char *fileName) {   /* FileLoader::getDir, fileloader.cpp:9-13 */
    const std::string s(fileName);
    const size_t found = s.find_last_of("/\\");
    return found == std::string::npos ? std::string() : s.substr(0, found + 1);
}

bool MVS::loadNVM(const char *fileName, bool nvm2) {   /* fileloader.cpp:251-325 (NVM), :327-401 (NVM2) */
    cameras.clear();
    patches.clear();
    std::ifstream file(fileName);
    if (!file.is_open()) { err = std::string("can't open NVM file ") + fileName; return false; }
    if (imageDir.empty()) imageDir = dirOf(fileName);
    std::string line;
    int stage = 0;   /* 0: header, 1: camera count, 2: point count */
    while (std::getline(file, line)) {
        std::istringstream ss(line);
        std::string tok;
        if (!(ss >> tok)) continue;
        if (stage == 0) {
            if (tok == "NVM_V3") stage = 1;
            continue;
        }
        if (stage == 1) {
            const int num = atoi(tok.c_str());
            for (int i = 0; i < num; ++i) {
                if (!std::getline(file, line)) { err = "NVM: truncated camera list"; return false; }
                std::istringstream cs(line);
                Camera cam;
                cam.principal[0] = cam.principal[1] = -1;
                cam.radialDistortion = 0;
                if (!(cs >> cam.fileName >> cam.focal[0])) { err = "NVM: bad camera line"; return false; }
                if (nvm2) cs >> cam.focal[1] >> cam.principal[0] >> cam.principal[1];
                else cam.focal[1] = cam.focal[0];
                cs >> cam.quaternion[0] >> cam.quaternion[1] >> cam.quaternion[2] >> cam.quaternion[3];
                cs >> cam.center[0] >> cam.center[1] >> cam.center[2];
                if (!nvm2) cs >> cam.radialDistortion;
                if (cs.fail()) { err = "NVM: bad camera line"; return false; }
                if (!addCamera(cam, true)) return false;
            }
            stage = 2;
            continue;
        }
        if (stage == 2) {
            const int num = atoi(tok.c_str());
            for (int i = 0; i < num; ++i) {   /* loadNvmPatch, fileloader.cpp:112-165 */
                if (!std::getline(file, line)) { err = "NVM: truncated point list"; return false; }
                std::istringstream ps(line);
                Patch p;
                int r, g, b, camNum;
                if (!(ps >> p.center[0] >> p.center[1] >> p.center[2] >> r >> g >> b >> camNum)) { err = "NVM: bad point line"; return false; }
                p.color[2] = (uint8_t)r;
                p.color[1] = (uint8_t)g;
                p.color[0] = (uint8_t)b;
                for (int k = 0; k < camNum; ++k) {
                    int idx, feat;
                    double x, y;
                    if (!(ps >> idx >> feat >> x >> y) || idx < 0 || idx >= (int)cameras.size()) { err = "NVM: bad measurement"; return false; }
                    if (std::find(p.camIdx.begin(), p.camIdx.end(), idx) != p.camIdx.end()) continue;   /* one measurement per camera: a view enters the cost once */
                    p.camIdx.push_back(idx);
                    p.imgPoint.push_back(x + cameras[idx].cols / 2);
                    p.imgPoint.push_back(y + cameras[idx].rows / 2);
                }
                p.type = PMVS_TYPE_SEED;
                p.id = nextId++;
                setEstimatedNormal(p);               /* seed ctor, patch.cpp:26-34 */
                patches.insert(std::pair<int, Patch>(p.id, p));
            }
            break;
        }
    }
    reCentering();   /* mvs.cpp:161-164 */
    return true;
}

bool MVS::loadMVS(const char *fileName) {   /* fileloader.cpp:403-472 */
    cameras.clear();
    patches.clear();
    std::ifstream file(fileName, std::ios::binary);
    if (!file.is_open()) { err = std::string("can't open MVS file ") + fileName; return false; }
    std::string line;
    int stage = 0;
    while (std::getline(file, line)) {
        std::istringstream ss(line);
        std::string tok;
        if (!(ss >> tok)) continue;
        if (stage == 0) {
            if (tok == "MVS_V2") stage = 1;
            else if (tok == "MVS_V3") {
                MvsConfig c;
                file.read((char *)&c, sizeof(c));
                if (!file) { err = "MVS: truncated config"; return false; }
                setConfig(c);
                stage = 1;
            }
            continue;
        }
        int num = 0;
        ss >> num;
        if (stage == 1) {   /* "CAMERAS n", loadMvsCamera :173-206 */
            for (int i = 0; i < num; ++i) {
                Camera cam;
                int len = 0;
                file.read((char *)&len, sizeof(int));
                if (!file || len < 0 || len > 65536) { err = "MVS: bad camera record"; return false; }
                cam.fileName.resize((size_t)len);
                file.read(&cam.fileName[0], len);
                file.read((char *)cam.center, 3 * sizeof(double));
                file.read((char *)cam.focal, 2 * sizeof(double));
                file.read((char *)cam.principal, 2 * sizeof(double));
                file.read((char *)cam.quaternion, 4 * sizeof(double));
                file.read((char *)&cam.radialDistortion, sizeof(double));
                if (!file) { err = "MVS: truncated camera record"; return false; }
                if (!addCamera(cam, true)) return false;
            }
            stage = 2;
            continue;
        }
        if (stage == 2) {   /* "PATCHES n", loadMvsPatch :208-232 */
            for (int i = 0; i < num; ++i) {
                Patch p;
                int camNum = 0;
                file.read((char *)p.center, 3 * sizeof(double));
                file.read((char *)p.normalS, 2 * sizeof(double));
                file.read((char *)&camNum, sizeof(int));
                if (!file || camNum < 0 || camNum > 65536) { err = "MVS: bad patch record"; return false; }
                for (int k = 0; k < camNum; ++k) {
                    int idx = 0;
                    file.read((char *)&idx, sizeof(int));
                    if (!file || idx < 0 || idx >= (int)cameras.size()) { err = "MVS: patch record names a camera the file does not hold"; return false; }
                    p.camIdx.push_back(idx);
                }
                file.read((char *)&p.fitness, sizeof(double));
                file.read((char *)&p.correlation, sizeof(double));
                if (!file) { err = "MVS: truncated patch record"; return false; }
                spherical2Normal(p.normalS, p.normal);
                p.type = PMVS_TYPE_SEED;      /* loader ctor marks TYPE_SEED, patch.cpp:45-59 */
                p.id = nextId++;
                /* the loader ctor (patch.cpp:45-59) goes on to setReferenceCameraIndex (:415-445: arg-max of
                 * normal . -opticalNormal, strict >, first wins) and setImagePoint (:627-653: projections and the
                 * colour under the reference camera's projection) — what `-f` writes into its PLY files. Depth range,
                 * LOD and priority are recomputed by refine() before anything reads them. */
                double maxCorr = -DBL_MAX;
                for (size_t k = 0; k < p.camIdx.size(); ++k) {
                    if (p.camIdx[k] < 0 || p.camIdx[k] >= (int)cameras.size()) continue;
                    const double *on = cameras[p.camIdx[k]].opticalNormal;
                    const double neg[3] = {-on[0], -on[1], -on[2]};
                    const double corr = dot3(p.normal, neg);
                    if (corr > maxCorr) { maxCorr = corr; p.refCamIdx = p.camIdx[k]; }
                }
                for (size_t k = 0; k < p.camIdx.size(); ++k) {
                    double pt[2] = {0, 0};
                    if (p.camIdx[k] >= 0 && p.camIdx[k] < (int)cameras.size()) cameras[p.camIdx[k]].project(p.center, pt, 0, cfg.lodRatio);
                    p.imgPoint.push_back(pt[0]);
                    p.imgPoint.push_back(pt[1]);
                }
                patchColor(p);
                patches.insert(std::pair<int, Patch>(p.id, p));
            }
            stage = 3;
        }
    }
    return true;
}

bool MVS::writeMVS(const char *fileName) const {   /* filewriter.cpp:71-102, :26-69 */
    std::ofstream file(fileName, std::ios::binary);
    if (!file.is_open()) return false;
    file << "MVS_V3" << "\n";
    file.write((const char *)&cfg, sizeof(MvsConfig));
    file << "CAMERAS " << (int)cameras.size() << "\n";
    for (size_t i = 0; i < cameras.size(); ++i) {
        const Camera &c = cameras[i];
        const int len = (int)c.fileName.size();
        file.write((const char *)&len, sizeof(int));
        file.write(c.fileName.data(), len);
        file.write((const char *)c.center, 3 * sizeof(double));
        file.write((const char *)c.focal, 2 * sizeof(double));
        file.write((const char *)c.principal, 2 * sizeof(double));
        file.write((const char *)c.quaternion, 4 * sizeof(double));
        file.write((const char *)&c.radialDistortion, sizeof(double));
    }
    file << "PATCHES " << (int)patches.size() << "\n";
    for (std::map<int, Patch>::const_iterator it = patches.begin(); it != patches.end(); ++it) {
        const Patch &p = it->second;
        const int camNum = (int)p.camIdx.size();
        file.write((const char *)p.center, 3 * sizeof(double));
        file.write((const char *)p.normalS, 2 * sizeof(double));
        file.write((const char *)&camNum, sizeof(int));
        for (int k = 0; k < camNum; ++k) file.write((const char *)&p.camIdx[k], sizeof(int));
        file.write((const char *)&p.fitness, sizeof(double));
        file.write((const char *)&p.correlation, sizeof(double));
    }
    return (bool)file;
}

bool MVS::writePLY(const char *fileName) const {   /* filewriter.cpp:104-139 */
    std::ofstream file(fileName);
    if (!file.is_open()) return false;
    file << "ply\nformat ascii 1.0\nelement vertex " << patches.size() << "\n";
    file << "property float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n";
    file << "property uchar diffuse_red\nproperty uchar diffuse_green\nproperty uchar diffuse_blue\nend_header\n";
    std::string buf;
    buf.reserve(patches.size() * 80);
    for (std::map<int, Patch>::const_iterator it = patches.begin(); it != patches.end(); ++it) plyLine(buf, it->second);
    file.write(buf.data(), (std::streamsize)buf.size());
    return (bool)file;
}

bool MVS::writePSR(const char *fileName) const {   /* filewriter.cpp:141-171 */
    std::ofstream file(fileName, std::ios::binary);
    if (!file.is_open()) return false;
    for (std::map<int, Patch>::const_iterator it = patches.begin(); it != patches.end(); ++it) {
        const Patch &p = it->second;
        const float v[6] = {(float)p.center[0], (float)p.center[1], (float)p.center[2], (float)p.normal[0], (float)p.normal[1], (float)p.normal[2]};
        file.write((const char *)v, sizeof(v));
    }
    return (bool)file;
}

}   // namespace tmvs
