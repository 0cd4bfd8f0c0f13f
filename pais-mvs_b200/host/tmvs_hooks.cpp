/*
 * tmvs_hooks.cpp — plain-C test hooks over the host driver's caller-side functions (last row:
 * getExpansionPatchCenter, skipNeighborCell, runtimeFiltering, insertPatch / deletePatch, the queue pops, isNeighbor,
 * reCentering, setNeighborRadius, Camera::project, cell maps). Built as pais-mvs_b200/lib/libtmvs_host.so so that
 * tests/test_host_parity_cpu.py can drive them from ctypes against the test suite's restatement of the reference
 * lines. Nothing here is on the product path: `tmvs` does not link this file and no GPU is touched (no context is created).
 */
#include <cmath>
#include <cstring>

#include "tmvs.h"

using namespace tmvs;

namespace {
Patch makePatch(int id, const double *center, const double *normal, double fitness, double priority, double correlation, int nCam,
                const int *camIdx, const double *imgPoint, int expanded, int drop) {
    Patch p;
    p.id = id;
    for (int k = 0; k < 3; ++k) { p.center[k] = center[k]; p.normal[k] = normal[k]; }
    p.fitness = fitness;
    p.priority = priority;
    p.correlation = correlation;
    p.camIdx.assign(camIdx, camIdx + nCam);
    if (imgPoint) p.imgPoint.assign(imgPoint, imgPoint + 2 * nCam);
    p.expanded = expanded != 0;
    p.drop = drop != 0;
    return p;
}
}   // namespace

/* Stand-in for Patch::refine() in driver tests: a deterministic function of the candidate alone. The centre slides
 * along the ray from the first parent camera onto the plane z = planeZ, the normal is (0,0,-1), every camera sees the
 * patch, and the priority varies with the position so that the queue strategies have something to order. A centre
 * that leaves a camera's image is dropped. Restated in tests/test_expansion_cpu.py. */
struct PlaneRefiner {
    MVS *mvs;
    double planeZ;
    long calls = 0, refined = 0;
};
int planeRefine(void *user, int n, const PmvsPatchIn *in, PmvsPatchOut *out, unsigned) {
    PlaneRefiner &r = *(PlaneRefiner *)user;
    const MVS &m = *r.mvs;
    r.calls++;
    r.refined += n;
    for (int i = 0; i < n; ++i) {
        PmvsPatchOut &o = out[i];
        memset(&o, 0, sizeof(o));
        const double *C = m.cameras[in[i].camIdx[0]].center;
        const double t = (r.planeZ - C[2]) / (in[i].center[2] - C[2]);
        for (int k = 0; k < 3; ++k) o.center[k] = C[k] + t * (in[i].center[k] - C[k]);
        o.normal[0] = 0; o.normal[1] = 0; o.normal[2] = -1;
        o.normalS[0] = acos(-1.0); o.normalS[1] = 0;
        o.fitness = 1.0;
        o.correlation = 0.95;
        o.priority = 1.0 + (o.center[0] * 0.37 + o.center[1] * 0.11);
        o.LOD = 0;
        o.refCamIdx = in[i].camIdx[0];
        o.psoRuns = 1;
        o.nCam = (int)m.cameras.size();
        o.nImgPoint = o.nCam;
        for (int c = 0; c < o.nCam; ++c) {
            o.camIdx[c] = (uint16_t)c;
            if (!m.cameras[c].project(o.center, o.imgPoint[c], 0, m.cfg.lodRatio)) o.drop = 1;
        }
        if (o.drop) { o.fitness = o.priority = 1.7976931348623157e308; o.nImgPoint = 0; }
    }
    return PMVS_OK;
}

extern "C" {

/* runs MVS::expansionPatches over the patches put so far with the plane stand-in; returns the number of refine calls.
 * mergeSlots: 0 one pass per camera slot, 1 merged pass with pipelined rounds (the default of tmvs), 2 merged pass, --no-pipeline */
long tmvs_hook_expand_plane(void *h, double planeZ, int roundSize, int mergeSlots, long *refined) {
    MVS &m = *(MVS *)h;
    PlaneRefiner r;
    r.mvs = &m;
    r.planeZ = planeZ;
    m.refineOverride = planeRefine;
    m.refineUser = &r;
    m.roundSize = roundSize;
    m.mergeSlots = mergeSlots != 0;
    m.pipelineRounds = mergeSlots != 2;
    m.autosaveSeconds = 1e18;
    const bool ok = m.expansionPatches();
    m.refineOverride = nullptr;
    m.refineUser = nullptr;
    if (refined) *refined = r.refined;
    return ok ? r.calls : -1;
}
int tmvs_hook_patch_ids(void *h, int *out, int cap) {
    const MVS &m = *(MVS *)h;
    int k = 0;
    for (std::map<int, Patch>::const_iterator it = m.patches.begin(); it != m.patches.end() && k < cap; ++it) out[k++] = it->first;
    return (int)m.patches.size();
}

void *tmvs_hook_create(const PmvsConfig *cfg) { return new MVS(*cfg); }
void tmvs_hook_destroy(void *h) { delete (MVS *)h; }
void tmvs_hook_set_neighbor_radius_value(void *h, double r) { ((MVS *)h)->cfg.neighborRadius = r; }

/* Camera ctor (camera.cpp:45-136) from NVM-style parameters; grey = level 0 (may be NULL: no background test) */
int tmvs_hook_add_camera(void *h, double focal, const double *quaternion, const double *center, int cols, int rows, const uint8_t *grey) {
    MVS &m = *(MVS *)h;
    Camera cam;
    cam.fileName = "hook";
    cam.focal[0] = cam.focal[1] = focal;
    cam.principal[0] = cam.principal[1] = -1;   /* NVM: derived from the image size, camera.cpp:101-106 */
    cam.radialDistortion = 0;
    for (int k = 0; k < 4; ++k) cam.quaternion[k] = quaternion[k];
    for (int k = 0; k < 3; ++k) cam.center[k] = center[k];
    cam.cols = cols;
    cam.rows = rows;
    if (grey) {
        cam.pyramid.resize(1);
        cam.pyramid[0].cols = cols;
        cam.pyramid[0].rows = rows;
        cam.pyramid[0].grey.assign(grey, grey + (size_t)cols * rows);
    }
    return m.addCamera(cam, false) ? (int)m.cameras.size() - 1 : -1;
}

int tmvs_hook_project(void *h, int cam, const double *X, int LOD, double *out) {
    const MVS &m = *(MVS *)h;
    return m.cameras[cam].project(X, out, LOD, m.cfg.lodRatio) ? 1 : 0;
}

/* puts a patch into the container without filtering (what loadMVS / the seed pass leave behind) */
void tmvs_hook_put_patch(void *h, int id, const double *center, const double *normal, double fitness, double priority, double correlation,
                         int nCam, const int *camIdx, const double *imgPoint, int expanded) {
    MVS &m = *(MVS *)h;
    m.patches[id] = makePatch(id, center, normal, fitness, priority, correlation, nCam, camIdx, imgPoint, expanded, 0);
    if (m.nextId <= id) m.nextId = id + 1;      /* the loaders number patches with nextId++ */
}
void tmvs_hook_set_cell_maps(void *h) { ((MVS *)h)->setCellMaps(); }
void tmvs_hook_init_queue(void *h) {   /* initPriorityQueue, mvs.cpp:90-95: every patch in id order */
    MVS &m = *(MVS *)h;
    m.queueClear();
    for (std::map<int, Patch>::const_iterator it = m.patches.begin(); it != m.patches.end(); ++it) m.queuePush(it->first);
}

int tmvs_hook_runtime_filtering(void *h, int id, const double *center, const double *normal, double fitness, double priority,
                                double correlation, int nCam, const int *camIdx, const double *imgPoint, int drop) {
    const MVS &m = *(MVS *)h;
    return m.runtimeFiltering(makePatch(id, center, normal, fitness, priority, correlation, nCam, camIdx, imgPoint, 0, drop)) ? 1 : 0;
}
int tmvs_hook_insert_patch(void *h, int id, const double *center, const double *normal, double fitness, double priority, double correlation,
                           int nCam, const int *camIdx, const double *imgPoint, int drop) {
    MVS &m = *(MVS *)h;
    const size_t before = m.patches.size();
    m.insertPatch(makePatch(id, center, normal, fitness, priority, correlation, nCam, camIdx, imgPoint, 0, drop));
    return m.patches.size() > before ? 1 : 0;
}
void tmvs_hook_delete_patch(void *h, int id) { ((MVS *)h)->deletePatch(id); }
void tmvs_hook_set_expanded(void *h, int id) {
    MVS &m = *(MVS *)h;
    std::map<int, Patch>::iterator it = m.patches.find(id);
    if (it != m.patches.end()) it->second.expanded = true;
}
int tmvs_hook_pop(void *h) { return ((MVS *)h)->getPatchIdFromQueue(); }
int tmvs_hook_patch_count(void *h) { return (int)((MVS *)h)->patches.size(); }
int tmvs_hook_deleted_count(void *h) { return (int)((MVS *)h)->deletedPatches.size(); }

int tmvs_hook_cell(void *h, int cam, int cx, int cy, int *out, int cap) {   /* -1: outside the map */
    const MVS &m = *(MVS *)h;
    const CellMap &cm = m.cellMaps[cam];
    if (!cm.inMap(cx, cy)) return -1;
    const CellIds &c = cm.cell(cx, cy);
    for (int k = 0; k < (int)c.size() && k < cap; ++k) out[k] = c[k];
    return (int)c.size();
}
void tmvs_hook_map_size(void *h, int cam, int *wh) {
    const MVS &m = *(MVS *)h;
    wh[0] = m.cellMaps[cam].width;
    wh[1] = m.cellMaps[cam].height;
}

void tmvs_hook_expansion_center(void *h, int cam, int parentId, int cx, int cy, double *center) {
    const MVS &m = *(MVS *)h;
    m.getExpansionPatchCenter(m.cameras[cam], m.patches.at(parentId), cx, cy, center);
}
int tmvs_hook_skip_neighbor_cell(void *h, int cam, int cx, int cy, int refId) {
    const MVS &m = *(MVS *)h;
    return m.skipNeighborCell(m.cellMaps[cam].cell(cx, cy), m.patches.at(refId)) ? 1 : 0;
}
int tmvs_hook_is_neighbor(const double *c1, const double *n1, const double *c2, const double *n2, double radius) {
    Patch a, b;
    for (int k = 0; k < 3; ++k) { a.center[k] = c1[k]; a.normal[k] = n1[k]; b.center[k] = c2[k]; b.normal[k] = n2[k]; }
    return MVS::isNeighbor(a, b, radius) ? 1 : 0;
}

void tmvs_hook_recentering(void *h) { ((MVS *)h)->reCentering(); }
double tmvs_hook_set_neighbor_radius(void *h) {
    MVS &m = *(MVS *)h;
    m.setNeighborRadius();
    return m.cfg.neighborRadius;
}
int tmvs_hook_get_patch(void *h, int id, double *center, double *normal, double *normalS) {   /* -1 missing, else drop flag */
    const MVS &m = *(MVS *)h;
    std::map<int, Patch>::const_iterator it = m.patches.find(id);
    if (it == m.patches.end()) return -1;
    for (int k = 0; k < 3; ++k) { center[k] = it->second.center[k]; normal[k] = it->second.normal[k]; }
    normalS[0] = it->second.normalS[0];
    normalS[1] = it->second.normalS[1];
    return it->second.drop ? 1 : 0;
}

}   /* extern "C" */

/* CellIds (tmvs.h) against std::vector<int> under random push / erase / copy traffic crossing the inline <-> heap boundary;
 * returns 0 when every observable (size, order, contents, copies) agreed */
extern "C" int tmvs_hook_cellids_selftest(unsigned seed, int ops) {
    unsigned long long st = seed * 2654435761ull + 12345;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(st >> 33); };
    CellIds c;
    std::vector<int> model;
    for (int k = 0; k < ops; ++k) {
        const unsigned r = rnd() % 100;
        if (r < 55 || model.empty()) {
            const int id = (int)(rnd() % 12);                 /* few distinct ids: duplicates occur, erase removes the first */
            c.push_back(id);
            model.push_back(id);
        } else if (r < 90) {
            const int id = (int)(rnd() % 12);
            const bool a = c.erase(id);
            std::vector<int>::iterator it = std::find(model.begin(), model.end(), id);
            const bool b = it != model.end();
            if (b) model.erase(it);
            if (a != b) return 1;
        } else {
            CellIds d(c), e;                                 /* copy construction, assignment over inline and heap states */
            e = d;
            e = e;
            c = e;
            std::vector<CellIds> v(3, c);
            v.resize(40, CellIds());
            c = v[2];
        }
        if (c.size() != model.size() || c.empty() != model.empty()) return 2;
        for (size_t i = 0; i < model.size(); ++i)
            if (c[i] != model[i]) return 3;
        if ((size_t)(c.end() - c.begin()) != model.size()) return 4;
    }
    return 0;
}
