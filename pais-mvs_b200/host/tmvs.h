/*
 * tmvs.h — host side of `tmvs -r`: the reference's MVS / Camera / Patch / CellMap / FileLoader / FileWriter roles for the
 * reconstruction command (TMVS/TMVS.cpp:76-122), in C++ like the reference, with Patch::refine() served by the
 * B200 library (include/pmvs_b200.h) in batches.
 *
 * Kept from the reference: MvsConfig (byte-identical, mvs.h:19-72), the NVM / NVM2 / MVS_V2 / MVS_V3 input formats, the
 * MVS_V3 / PLY / PSR output formats, config.txt, compiled defaults, seed refinement, cell maps, runtime filtering,
 * expansion strategies. Changed on purpose: expansion runs in ROUNDS (pop K parents, generate their candidates — the
 * cells a candidate is expected to fill are not targeted again in the round —, refine the batch on the GPU, commit
 * serially in slot / parent order with every target cell re-checked) instead of one patch at a time;
 * images are read as PGM/PPM (no OpenCV here; tools/convert_images.py makes them).
 */
#ifndef TMVS_HOST_H
#define TMVS_HOST_H

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <vector>

#include "../../include/pmvs_b200.h"

namespace tmvs {

typedef PmvsConfig MvsConfig;

enum { EXPANSION_BEST_FIRST = 0, EXPANSION_WORST_FIRST = 1, EXPANSION_BREATH_FIRST = 2, EXPANSION_DEPTH_FIRST = 3 };   /* mvs.h */

struct Level {
    int cols = 0, rows = 0;
    std::vector<uint8_t> grey;
    std::vector<double> edge;
};

/* Camera, TMVS/mvs/camera.h:15-72 */
struct Camera {
    std::string fileName;
    double focal[2], principal[2], quaternion[4], center[3], radialDistortion;
    double R[9], t[3], KR[9], KT[3], opticalNormal[3];
    int cols = 0, rows = 0, maxLOD = 0;
    std::vector<Level> pyramid;
    std::vector<uint8_t> rgb;          /* rows*cols*3, R,G,B */
    bool available = false;
    bool project(const double X[3], double out[2], int LOD, double lodRatio) const;   /* camera.cpp:138-160 */
};

/* the AbstractPatch/Patch state the host keeps (abstractpatch.h:21-53) */
struct Patch {
    int id = -1, type = PMVS_TYPE_SEED;
    double center[3] = {0, 0, 0}, normal[3] = {0, 0, 0}, normalS[2] = {0, 0};
    double fitness = 1.7976931348623157e308, priority = 1.7976931348623157e308, correlation = 0;
    int LOD = -1, refCamIdx = -1;
    bool drop = false, expanded = false;
    uint8_t color[3] = {0, 0, 0};      /* b, g, r like cv::Vec3b */
    std::vector<int> camIdx;
    std::vector<double> imgPoint;      /* 2 per entry */
};

/* The patch ids of one cell (the reference keeps a std::vector<int> per cell, cellmap.h). Cells hold a handful of ids
 * (maxCellPatchNum = 3 by default), and a reconstruction touches millions of them: up to five ids live inline in the 24
 * bytes a vector header would take — no allocation, one cache line per visit — more move to the heap. */
class CellIds {
    int32_t w_[6];                       /* w_[0]: count; inline: ids in w_[1..5]; heap: std::vector<int>* in w_[2..3] */
    enum { INLINE = 5 };
    std::vector<int> *heap() const {
        std::vector<int> *h;
        memcpy(&h, &w_[2], sizeof(h));
        return h;
    }
    void setHeap(std::vector<int> *h) { memcpy(&w_[2], &h, sizeof(h)); }
    bool onHeap() const { return w_[0] > INLINE; }

public:
    CellIds() { w_[0] = 0; }
    CellIds(const CellIds &o) {
        memcpy(w_, o.w_, sizeof(w_));
        if (o.onHeap()) setHeap(new std::vector<int>(*o.heap()));
    }
    CellIds &operator=(const CellIds &o) {
        if (this == &o) return *this;
        if (onHeap()) delete heap();
        memcpy(w_, o.w_, sizeof(w_));
        if (o.onHeap()) setHeap(new std::vector<int>(*o.heap()));
        return *this;
    }
    ~CellIds() {
        if (onHeap()) delete heap();
    }
    size_t size() const { return (size_t)w_[0]; }
    bool empty() const { return w_[0] == 0; }
    const int *begin() const { return onHeap() ? heap()->data() : &w_[1]; }
    const int *end() const { return begin() + w_[0]; }
    int operator[](size_t i) const { return begin()[i]; }
    void push_back(int id) {
        if (w_[0] < INLINE) { w_[1 + w_[0]++] = id; return; }
        if (w_[0] == INLINE) {
            std::vector<int> *h = new std::vector<int>(&w_[1], &w_[1] + INLINE);
            setHeap(h);
        }
        heap()->push_back(id);
        ++w_[0];
    }
    bool erase(int id) {                 /* first occurrence, order kept (std::find + vector::erase in the reference) */
        const int *b = begin(), *e = end(), *it = std::find(b, e, id);
        if (it == e) return false;
        const size_t at = (size_t)(it - b);
        if (onHeap()) {
            std::vector<int> *h = heap();
            h->erase(h->begin() + (long)at);
            if ((int)h->size() == INLINE) {                 /* back inline */
                int tmp[INLINE];
                for (int k = 0; k < INLINE; ++k) tmp[k] = (*h)[(size_t)k];
                delete h;
                for (int k = 0; k < INLINE; ++k) w_[1 + k] = tmp[k];
            }
        } else {
            for (size_t k = at; k + 1 < (size_t)w_[0]; ++k) w_[1 + k] = w_[2 + k];
        }
        --w_[0];
        return true;
    }
};

struct CellMap {   /* TMVS/mvs/cellmap.{h,cpp} */
    int width = 0, height = 0;
    std::vector<CellIds> cells;
    void init(int imgW, int imgH, int cellSize);
    bool inMap(int x, int y) const { return !(x < 0 || y < 0 || x >= width || y >= height); }
    bool insert(int x, int y, int id);
    bool drop(int x, int y, int id);
    const CellIds &cell(int x, int y) const { return cells[(size_t)y * width + x]; }
};

void setInitConfig(MvsConfig &c);                                   /* TMVS.cpp:26-52 */
bool loadConfig(const char *fileName, MvsConfig &c);                /* fileloader.cpp:474-565 */

/* image reader (camera.cpp:51-69; the pyramid itself is built on the GPU) */
bool readPnm(const std::string &path, int &cols, int &rows, std::vector<uint8_t> &grey, std::vector<uint8_t> &rgb);

class MVS {
public:
    MvsConfig cfg;
    std::vector<Camera> cameras;
    std::map<int, Patch> patches;
    std::vector<Patch> deletedPatches;
    std::vector<CellMap> cellMaps;
    std::vector<int> queue;            /* insertion order (the reference's container) */
    int nextId = 0;
    int roundSize = 1024;              /* parents popped per expansion round */
    bool pipelineRounds = true;        /* merged mode: two rounds in flight (commit k and generate k+2 on the host while the GPUs refine round k+1) */
    bool mergeSlots = true;            /* one GPU pass per round over all camera slots (expected-neighbour prediction); false: one per slot */
    int device = 0;                    /* first device */
    int numGpus = 1;                   /* devices device .. device+numGpus-1, candidates sharded by index */
    uint64_t rngSeed = 42;
    double autosaveSeconds = 5.0;      /* minimum spacing of auto_save.mvs checkpoints */
    bool verbose = false;
    bool warnedViews = false;            /* one warning when a seed lists more than PMVS_MAX_VIEWS cameras */
    std::string outDir;                  /* prefix of the files the driver writes on its own (auto_save.mvs) */
    std::string imageDir;              /* prefix for camera image files */
    long refinedCount = 0;             /* patches sent through refine() */
    double gpuSeconds = 0;             /* inside pmvs_refine_batch calls */
    /* what the refined patches cost (sums over PmvsPatchOut): swarm evaluations, evaluations that ran the window loop,
     * iterations, swarm runs, dropped records, visible cameras kept, level of detail */
    long statEvaluations = 0, statWindowEvaluations = 0, statIterations = 0, statRuns = 0, statDropped = 0, statViews = 0, statLOD[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double contextSeconds = 0;         /* pmvs_create: CUDA context, module load, pyramid upload + build */

    explicit MVS(const MvsConfig &c);
    ~MVS();
    void setConfig(const MvsConfig &c);                             /* mvs.cpp:42-72 */

    bool loadNVM(const char *fileName, bool nvm2);                  /* fileloader.cpp:251-401 + mvs.cpp:161-169 */
    bool loadMVS(const char *fileName);                             /* fileloader.cpp:403-472 */
    bool writeMVS(const char *fileName) const;                      /* filewriter.cpp:71-102 */
    bool writePLY(const char *fileName) const;                      /* filewriter.cpp:104-139 */
    bool writePSR(const char *fileName) const;                      /* filewriter.cpp:141-171 */

    bool refineSeedPatches();                                       /* mvs.cpp:196-231 */
    bool expansionPatches();                                        /* mvs.cpp:233-275, in rounds */

    /* `-f` post-process (TMVS.cpp:124-170). The three PMVS filters are the reference's serial cell-map walks; the
     * PCMVS filter's O(N^2) pair scan runs on the GPU(s) (pmvs_neighbor_counts), rows sharded over numGpus devices. */
    void cellFiltering();                                           /* mvs.cpp:279-325 */
    void visibilityFiltering();                                     /* mvs.cpp:399-446 */
    void neighborCellFiltering(double neighborRatio);               /* mvs.cpp:327-397 */
    bool neighborPatchFiltering(double neighborRatio);              /* mvs.cpp:448-525 */
    bool writeDeletedPatchMVS(const char *fileName) const;          /* filewriter.cpp:173-204 */
    bool writeDeletedPatchPLY(const char *fileName) const;          /* filewriter.cpp:206-241 */
    void clearDeletedPatches() { deletedPatches.clear(); }          /* mvs.cpp:154-156 */
    double lastAvgNeighborNum = 0;                                  /* "average neighbor number", mvs.cpp:506-511 */

    /* tests only: a host stand-in for pmvs_refine_batch, so that the driver's control flow (rounds, candidate generation,
     * commit order) can be checked on a GPU-less box; never set by tmvs */
    typedef int (*RefineOverride)(void *user, int n, const PmvsPatchIn *in, PmvsPatchOut *out, unsigned flags);
    RefineOverride refineOverride = nullptr;
    void *refineUser = nullptr;

    /* pieces exposed for tests */
    bool addCamera(Camera &cam, bool loadImage);
    void reCentering();                                             /* mvs.cpp:135-145, patch.cpp:67-112 */
    void setNeighborRadius();                                       /* mvs.cpp:147-152 */
    bool runtimeFiltering(const Patch &p) const;                    /* mvs.cpp:838-898 */
    void getExpansionPatchCenter(const Camera &cam, const Patch &parent, int cx, int cy, double center[3]) const;   /* mvs.cpp:809-836 */
    bool skipNeighborCell(const CellIds &cell, const Patch &ref) const;                                    /* mvs.cpp:792-807 */
    static bool isNeighbor(const Patch &a, const Patch &b, double neighborRadius);                                  /* patch.cpp:6-23 */
    int getPatchIdFromQueue();                                      /* mvs.cpp:636-788, indexed */
    size_t byPriorityQueueSize() const { return prioQueue.size() > fifo.size() ? prioQueue.size() : fifo.size(); }
    const std::string &lastError() const { return err; }
    void setCellMaps();                                             /* mvs.cpp:116-133 */
    void insertPatch(const Patch &p);                               /* mvs.cpp:579-601 */
    void deletePatch(int id);                                       /* mvs.cpp:607-634 */
    void setEstimatedNormal(Patch &p) const;                        /* patch.cpp:390-413 */
    void queuePush(int id);                                         /* initPriorityQueue / insertPatch's queue.push_back */
    void queueClear();

private:
    /* id -> patch while a filter runs (std::map nodes do not move; deletePatch clears the entry, insertPatch drops the
     * index): the cell walks of mvs.cpp:279-446 look every cell member up, millions of times */
    mutable std::vector<const Patch *> idIndex;
    struct IdIndexGuard {
        const MVS &m;
        explicit IdIndexGuard(const MVS &mvs);
        ~IdIndexGuard() { m.idIndex.clear(); }
    };
    const Patch *lookup(int id) const {
        if (idIndex.empty()) {
            std::map<int, Patch>::const_iterator it = patches.find(id);
            return it == patches.end() ? nullptr : &it->second;
        }
        return id >= 0 && (size_t)id < idIndex.size() ? idIndex[id] : nullptr;
    }
    /* min-heap on (key, insertion sequence): the same total order an ordered set would give (sequence numbers are unique),
     * with O(1) average insertion into a contiguous array */
    typedef std::pair<std::pair<double, long>, int> PrioEntry;
    std::priority_queue<PrioEntry, std::vector<PrioEntry>, std::greater<PrioEntry> > prioQueue;
    std::deque<int> fifo;
    long queueSeq = 0;
    std::vector<pmvs_ctx *> ctxs;       /* one per GPU */
    std::string err;
    bool ensureContext();
    void ensureFilterMaps();                                        /* the `if (cellMaps.empty())` prologue of every filter */
    bool refineBatch(std::vector<Patch *> &batch, unsigned flags, const std::vector<std::vector<int> > *parentCams);
    void patchColor(Patch &p) const;                                /* patch.cpp:648-652 */
};

}   // namespace tmvs
#endif
