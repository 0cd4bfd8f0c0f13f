// Sanitizer harness for the host expansion driver (tools/host_sanitize.sh): six cameras, 32 seeds on a plane, the plane stand-in for
// refine() (test hook), MVS::expansionPatches in the mode given as argv[1] (1: two rounds in flight, 2: --no-pipeline, 0: --slot-passes).
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "../pais-mvs_b200/host/tmvs.h"
using namespace tmvs;
extern "C" {
void *tmvs_hook_create(const PmvsConfig *cfg);
void tmvs_hook_destroy(void *h);
int tmvs_hook_add_camera(void *h, double focal, const double *quaternion, const double *center, int cols, int rows, const uint8_t *grey);
void tmvs_hook_put_patch(void *h, int id, const double *center, const double *normal, double fitness, double priority, double correlation, int nCam, const int *camIdx, const double *imgPoint, int expanded);
long tmvs_hook_expand_plane(void *h, double planeZ, int roundSize, int mergeSlots, long *refined);
int tmvs_hook_patch_count(void *h);
int tmvs_hook_project(void *h, int cam, const double *X, int LOD, double *out);
int tmvs_hook_cellids_selftest(unsigned seed, int ops);
}
int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 1;
    MvsConfig cfg; setInitConfig(cfg);
    cfg.cellSize = 4; cfg.maxCellPatchNum = 3; cfg.minCamNum = 3; cfg.maxFitness = 10.0; cfg.minCorrelation = 0.9; cfg.neighborRadiusScalar = 0.01;
    void *h = tmvs_hook_create((const PmvsConfig *)&cfg);
    const int NC = 6, cols = 640, rows = 480;
    for (int i = 0; i < NC; ++i) {
        const double ang = (-20 + 40.0 * i / (NC - 1)) * M_PI / 180;
        double center[3] = {-2.0 + 4.0 * i / (NC - 1), 0.05 * i, -10.0};
        // look-at quaternion: rotation about y by -ang (camera looks at the origin)
        double q[4] = {1, 0, 0, 0}; (void)ang;
        if (tmvs_hook_add_camera(h, 1.2 * cols, q, center, cols, rows, nullptr) != i) { printf("camera failed\n"); return 1; }
    }
    int ci[NC]; for (int i = 0; i < NC; ++i) ci[i] = i;
    int put = 0;
    for (int pid = 0; pid < 32; ++pid) {
        double c[3] = {1.5 * ((pid * 37 % 100) / 50.0 - 1), 1.0 * ((pid * 61 % 100) / 50.0 - 1), 0.0}, n[3] = {0, 0, -1}, pts[2 * NC];
        bool ok = true;
        for (int i = 0; i < NC; ++i) ok = ok && tmvs_hook_project(h, i, c, 0, pts + 2 * i);
        if (!ok) continue;
        tmvs_hook_put_patch(h, pid, c, n, 1.0, 1.0 + 0.1 * pid, 0.95, NC, ci, pts, 0);
        ++put;
    }
    long refined = 0;
    long calls = tmvs_hook_expand_plane(h, 0.0, 256, mode, &refined);
    printf("mode %d: seeds %d calls %ld refined %ld patches %d selftest %d\n", mode, put, calls, refined, tmvs_hook_patch_count(h), tmvs_hook_cellids_selftest(7, 5000));
    tmvs_hook_destroy(h);
    return 0;
}
