/*
 * tmvs_hooks.cpp — plain-C test hooks over the host driver's caller-side functions (last row:
 * getExpansionPatchCenter, skipNeighborCell, runtimeFiltering, insertPatch / deletePatch, the queue pops, isNeighbor,
 * reCentering, setNeighborRadius, Camera::project, cell maps). Built as pais-mvs_b200/lib/libtmvs_host.so so that
 * tests/test_host_parity_cpu.py can drive them from ctypes against the test suite's restatement of the reference
 * lines. Nothing here is on the product path: `tmvs` does not link this file and no GPU is touched (no context is created).
 */
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

extern "C" {

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
    const std::vector<int> &c = cm.cell(cx, cy);
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
