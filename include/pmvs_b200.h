/*
 * pmvs_b200.h — C-ABI of the B200-native patch-refinement path of pais-mvs.
 *
 * The reference (adahbingee/pais-mvs, C++/OpenCV, CPU only) has no FFI layer; the two seams
 * this library replaces are (citations relative to the reference tree):
 *
 *   seam 1  double PAIS::getFitness(const Particle&, void *obj)   TMVS/mvs/patch.h:66,
 *           TMVS/mvs/patch.cpp:914-1047 — the PSO cost callback (TMVS/pso/psosolver.h:72).
 *           -> pmvs_fitness_batch()
 *   seam 2  void Patch::refine() TMVS/mvs/patch.h:50, patch.cpp:114-176, followed at both call
 *           sites (TMVS/mvs/mvs.cpp:214-215 and :573-574) by Patch::removeInvisibleCamera()
 *           patch.cpp:655-721.  -> pmvs_refine_batch()
 *
 * Plain C: pointers and sizes only, no C++/torch types. All calls are blocking (internally they
 * run on the context's own CUDA stream). Return value: 0 on success, negative PMVS_E_* on
 * argument/CUDA errors. Per-patch algorithmic failure stays in-band exactly like the reference:
 * out.drop = 1 and fitness/priority = DBL_MAX (patch.cpp:118-123).
 *
 * There is no CPU fallback: every compute entry point fails with PMVS_E_CUDA when no sm_100
 * device is usable.
 */
#ifndef PMVS_B200_H
#define PMVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMVS_MAX_VIEWS   64   /* max visible cameras of one patch (BASELINE.json config 5) */
#define PMVS_MAX_LEVELS  16   /* MvsConfig.maxLOD default 15 -> 16 pyramid levels (TMVS.cpp:41) */

#define PMVS_OK            0
#define PMVS_E_ARG        -1
#define PMVS_E_CUDA       -2
#define PMVS_E_NOMEM      -3
#define PMVS_E_UNSUPPORTED -4

/* Patch types (TMVS/mvs/patch.h:17-18). */
#define PMVS_TYPE_SEED    0
#define PMVS_TYPE_EXPAND  1

/* pmvs_refine_batch flags */
#define PMVS_F_POST_REMOVE_INVISIBLE  1u  /* also run the caller's trailing removeInvisibleCamera()
                                             (mvs.cpp:215, :574) inside the same launch */
#define PMVS_F_EXPAND_VISIBLE         2u  /* for TYPE_EXPAND inputs, treat in.camIdx as the PARENT's
                                             cameras and run Patch::expandVisibleCamera()
                                             (patch.cpp:723-761, part of the expansion ctor :36-43) */

/* per-patch status bits (out.status); 0 = nothing unusual */
#define PMVS_S_TOO_MANY_VIEWS      1u  /* expandVisibleCamera found > PMVS_MAX_VIEWS cameras: patch dropped */
#define PMVS_S_TOO_MANY_PARTICLES  2u  /* reserved (rounds before 2: TYPE_SEED with 2*particleNum > 64); seeds now run the
                                          reference's 2*particleNum particles for every accepted particleNum (<= 64) */
#define PMVS_S_BAD_CAMERA          4u  /* a camIdx entry is not a camera of the scene: patch dropped (device-resident
                                          inputs only; pmvs_refine_batch rejects such host records with PMVS_E_ARG) */

/*
 * Byte-for-byte the reference's MvsConfig (TMVS/mvs/mvs.h:19-72) as laid out by MSVC/gcc x64:
 * 160 bytes; this is also what MVS_V3 files carry (TMVS/io/filewriter.cpp:71-102).
 */
typedef struct PmvsConfig {
    int32_t cellSize;              /*   0 */
    int32_t patchRadius;           /*   4 */
    int32_t patchSize;             /*   8  2*radius+1, recomputed by MVS::setConfig (mvs.cpp:67) */
    int32_t minCamNum;             /*  12 */
    double  textureVariation;      /*  16 */
    double  visibleCorrelation;    /*  24 */
    double  minCorrelation;        /*  32 */
    double  maxFitness;            /*  40 */
    double  lodRatio;              /*  48 */
    int32_t minLOD;                /*  56 */
    int32_t maxLOD;                /*  60 */
    int32_t maxCellPatchNum;       /*  64 */
    int32_t _pad0;                 /*  68 */
    double  reduceNormalRange;     /*  72 */
    uint8_t adaptiveDistanceEnable;   /* 80 */
    uint8_t adaptiveDifferenceEnable; /* 81 */
    uint8_t adaptiveGradientEnable;   /* 82 */
    uint8_t _pad1[5];
    double  distWeighting;         /*  88 */
    double  diffWeighting;         /*  96 */
    double  gradientWeighting;     /* 104 */
    double  neighborRadius;        /* 112 */
    double  neighborRadiusScalar;  /* 120 */
    double  minRegionRatio;        /* 128 */
    double  depthRangeScalar;      /* 136 */
    int32_t particleNum;           /* 144 */
    int32_t maxIteration;          /* 148 */
    int32_t expansionStrategy;     /* 152 */
    int32_t _pad2;                 /* 156 */
} PmvsConfig;

/* One pyramid level of one camera (TMVS/mvs/camera.cpp:63-92). Host pointers, borrowed for the
 * duration of pmvs_create. grey: u8 row-major, `pitch` bytes per row. edge: f64 row-major,
 * cols doubles per row, may be NULL when adaptiveGradientEnable is 0 (only read at patch.cpp:1037). */
typedef struct PmvsLevel {
    int32_t cols, rows;
    int64_t pitch;
    const uint8_t *grey;
    const double  *edge;
} PmvsLevel;

/* Writable host level for pmvs_build_pyramid. */
typedef struct PmvsLevelOut {
    int32_t cols, rows;
    int64_t pitch;
    uint8_t *grey;
    double  *edge;
} PmvsLevelOut;

/* Camera (TMVS/mvs/camera.h:15-72) after construction (camera.cpp:45-136): matrices row-major. */
typedef struct PmvsCamera {
    double focal[2];
    double principal[2];
    double center[3];
    double R[9];              /* rotation           camera.cpp:6-35   */
    double t[3];              /* translation -R*C   camera.cpp:120    */
    double KR[9];             /* K*R                camera.cpp:123    */
    double KT[3];             /* K*t                camera.cpp:124    */
    double opticalNormal[3];  /* R^T*(0,0,1)        camera.cpp:130-133 */
    int32_t maxLOD;           /* camera.cpp:63-64; levels 0..maxLOD present */
    int32_t _pad;
    PmvsLevel level[PMVS_MAX_LEVELS];
} PmvsCamera;

/* One hypothesis for seam 1: what getFitness reads from the Particle (pos[0..2] = theta, phi,
 * depth; patch.cpp:936,944) and from the Patch (ray, refCamIdx, camIdx, LOD; patch.cpp:922-932). */
typedef struct PmvsHypothesis {
    double  ray[3];
    double  theta, phi, depth;
    int32_t refCamIdx;
    int32_t LOD;
    int32_t nCam;
    uint16_t camIdx[PMVS_MAX_VIEWS];
    int32_t _pad;
} PmvsHypothesis;

/* Input of seam 2: the state a Patch has when refine() is entered (after one of the ctors
 * patch.cpp:26-59): center, normal + its spherical form (AbstractPatch::setNormal,
 * abstractpatch.cpp:42-50), ordered visible-camera list, type, id (keys the RNG stream). */
typedef struct PmvsPatchIn {
    double  center[3];
    double  normal[3];
    double  normalS[2];
    int32_t type;             /* PMVS_TYPE_SEED / PMVS_TYPE_EXPAND */
    int32_t id;
    int32_t nCam;
    int32_t _pad;
    uint16_t camIdx[PMVS_MAX_VIEWS];
} PmvsPatchIn;

/* Output of seam 2: every AbstractPatch field refine()/removeInvisibleCamera() write
 * (abstractpatch.h:21-53) plus Patch::drop (patch.h:19). */
typedef struct PmvsPatchOut {
    double  center[3];
    double  normal[3];
    double  normalS[2];
    double  ray[3];
    double  depth;
    double  depthRange[2];
    double  fitness;
    double  priority;
    double  correlation;
    int32_t LOD;
    int32_t refCamIdx;
    int32_t nCam;
    int32_t drop;
    int32_t psoRuns;          /* number of psoOptimization() calls made (1 for expansion patches) */
    int32_t psoIterations;    /* PsoSolver::getIteration() of the last run (patch.cpp:216) */
    uint32_t evaluations;     /* getFitness calls spent on this patch (all runs) */
    uint32_t status;
    uint16_t camIdx[PMVS_MAX_VIEWS];
    int32_t nImgPoint;        /* imgPoint.size(): set by setImagePoint (patch.cpp:627-653) BEFORE the
                                 trailing removeInvisibleCamera, so it can exceed nCam */
    uint32_t windowEvaluations; /* getFitness calls that walked the whole (2r+1)^2 window, i.e. did not return the
                                 DBL_MAX sentinel early (patch.cpp:939-962, :999-1002): the gather-traffic unit */
    double  imgPoint[PMVS_MAX_VIEWS][2];
} PmvsPatchOut;

typedef struct pmvs_ctx pmvs_ctx;   /* opaque; owns all device memory; one per GPU; not thread-safe */

/* Number of pyramid levels of a cols x rows image and their sizes (camera.cpp:63-64; cv::resize rounds the size).
 * Host-only helper. levelCols / levelRows: PMVS_MAX_LEVELS entries or NULL. */
int pmvs_pyramid_levels(int cols, int rows, double lodRatio, int cfgMaxLOD, int *maxLOD, int32_t *levelCols,
                        int32_t *levelRows);

/* The Camera ctor's pyramid (camera.cpp:63-92) built on `device`: level l = INTER_AREA resize of level 0 by
 * lodRatio^l, edge = min-max normalised [-1,0,1] gradient magnitude. levels[0..maxLOD]: caller-allocated host arrays
 * with cols/rows from pmvs_pyramid_levels (levels[0].grey may be NULL); edge arrays required iff withEdge. */
int pmvs_build_pyramid(int device, const uint8_t *grey0, int cols, int rows, int64_t pitch, double lodRatio, int maxLOD,
                       int withEdge, PmvsLevelOut *levels);

/* The pair scan of the PCMVS neighbour filter, MVS::neighborPatchFiltering (TMVS/mvs/mvs.cpp:448-525), on `device`:
 * counts[k] = #{ j != first+k : cv::norm(center[first+k] - center[j]) <= radius }, k in [0, count) — the length of the
 * reference's PatchNeighbor::nid list (:489-499), bit-exact. centers: n x 3 f64 (host); [first, first+count) is the
 * shard of rows this device scans against all n points (rows are independent). */
int pmvs_neighbor_counts(int device, int n, const double *centers, double radius, int first, int count, int *counts);

/* Upload config + cameras + pyramids to `device` and build the derived tables
 * (distance weighting MVS::initPatchDistanceWeighting mvs.cpp:97-114; lodRatio^l).
 * rngSeed keys the counter-based replacement of the reference's srand(time)+rand()
 * (psosolver.cpp:60-68). Level 0 of every camera must be given; a level l >= 1 whose `grey` is NULL is built on the
 * device from level 0 (its cols/rows may be 0 or must equal pmvs_pyramid_levels'), and a NULL `edge` is built on the
 * device when cfg->adaptiveGradientEnable is set. */
int pmvs_create(pmvs_ctx **out, const PmvsConfig *cfg, int nCams, const PmvsCamera *cams,
                int device, uint64_t rngSeed);

/* MVS::setNeighborRadius (mvs.cpp:147-152): derived at run time, read by setDepthRange (patch.cpp:508). */
int pmvs_set_neighbor_radius(pmvs_ctx *ctx, double neighborRadius);

/* Replace the config (MVS::setConfig mvs.cpp:42-72); patchRadius/weights may change, cameras stay. */
int pmvs_set_config(pmvs_ctx *ctx, const PmvsConfig *cfg);

/* seam 1: n independent getFitness evaluations. Host pointers. */
int pmvs_fitness_batch(pmvs_ctx *ctx, int n, const PmvsHypothesis *in, double *outFitness);

/* seam 2: n independent Patch::refine() calls (+ trailing removeInvisibleCamera with
 * PMVS_F_POST_REMOVE_INVISIBLE). Host pointers; out is caller-allocated (n records). */
int pmvs_refine_batch(pmvs_ctx *ctx, int n, const PmvsPatchIn *in, PmvsPatchOut *out, uint32_t flags);

/* Device-resident variant used for kernel-only timing and by multi-GPU drivers: d_in/d_out are
 * device pointers on the context's device; the launch is enqueued on `cudaStream` (a cudaStream_t
 * passed as void*, NULL = the context's stream) and NOT synchronised. */
int pmvs_refine_batch_device(pmvs_ctx *ctx, int n, const PmvsPatchIn *d_in, PmvsPatchOut *d_out,
                             uint32_t flags, void *cudaStream);

/* The exchange record of a multi-GPU driver (SURVEY.md 8b `pmvs_allgather_patches`, 8e): packs the n device-resident
 * results into 8 doubles per patch {centre 3, normal 3, fitness, drop} — the payload ranks all-gather between expansion
 * rounds (the collective itself is the host's: ncclAllGather / torch.distributed on d_records). d_counters: NULL or three
 * uint64 {evaluations, windowEvaluations, kept patches} the kernel ADDS this batch's totals to. Enqueued on `cudaStream`
 * (NULL = the context's stream), not synchronised. */
int pmvs_pack_records_device(pmvs_ctx *ctx, int n, const PmvsPatchOut *d_out, double *d_records, uint64_t *d_counters,
                             void *cudaStream);

/* Number of kernel launches issued by this context so far (bench.py "gpu_launches"). */
int64_t pmvs_launch_count(const pmvs_ctx *ctx);

/* Test support (not a reference seam): the GLN-PSO solver alone (TMVS/pso/psosolver.cpp) on the analytic functions
 * 0 sphere, 1 rosenbrock, 3 |x| with a DBL_MAX region, 4 plateaus, so it can be checked bit-for-bit against the
 * unmodified reference solver. L,U,init: n*3; hasInit,maxIter,P,fn,keys: n. Outputs: gbest n*3, gbestFitness n,
 * iterations n, particles n*64*8 {pos3,vec3,fitness,pBestFitness} (may be NULL). */
int pmvs_pso_test(pmvs_ctx *ctx, int n, const double *L, const double *U, const double *init, const int *hasInit,
                  const int *maxIter, const int *P, const int *fn, const uint64_t *keys, double *gbest,
                  double *gbestFitness, int *iterations, double *particles);

void pmvs_destroy(pmvs_ctx *ctx);
const char *pmvs_last_error(const pmvs_ctx *ctx);   /* never NULL; "" when no error */
const char *pmvs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PMVS_B200_H */
