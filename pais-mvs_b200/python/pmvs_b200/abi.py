"""ctypes mirrors of include/pmvs_b200.h (the C-ABI records). Layouts are asserted against the header's
documented offsets; the reference types they stand for are cited in the header."""
import ctypes as C

MAX_VIEWS = 64
MAX_LEVELS = 16

TYPE_SEED, TYPE_EXPAND = 0, 1
F_POST_REMOVE_INVISIBLE = 1
F_EXPAND_VISIBLE = 2
S_TOO_MANY_VIEWS = 1
S_BAD_CAMERA = 4
E_ARG, E_CUDA, E_NOMEM, E_UNSUPPORTED = -1, -2, -3, -4

DBL_MAX = 1.7976931348623157e308


class PmvsConfig(C.Structure):
    """MvsConfig, TMVS/mvs/mvs.h:19-72 (160 bytes)."""
    _fields_ = [
        ("cellSize", C.c_int32), ("patchRadius", C.c_int32), ("patchSize", C.c_int32), ("minCamNum", C.c_int32),
        ("textureVariation", C.c_double), ("visibleCorrelation", C.c_double), ("minCorrelation", C.c_double),
        ("maxFitness", C.c_double), ("lodRatio", C.c_double),
        ("minLOD", C.c_int32), ("maxLOD", C.c_int32), ("maxCellPatchNum", C.c_int32), ("_pad0", C.c_int32),
        ("reduceNormalRange", C.c_double),
        ("adaptiveDistanceEnable", C.c_uint8), ("adaptiveDifferenceEnable", C.c_uint8),
        ("adaptiveGradientEnable", C.c_uint8), ("_pad1", C.c_uint8 * 5),
        ("distWeighting", C.c_double), ("diffWeighting", C.c_double), ("gradientWeighting", C.c_double),
        ("neighborRadius", C.c_double), ("neighborRadiusScalar", C.c_double), ("minRegionRatio", C.c_double),
        ("depthRangeScalar", C.c_double),
        ("particleNum", C.c_int32), ("maxIteration", C.c_int32), ("expansionStrategy", C.c_int32), ("_pad2", C.c_int32),
    ]


class PmvsLevel(C.Structure):
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("pitch", C.c_int64),
                ("grey", C.c_void_p), ("edge", C.c_void_p)]


class PmvsLevelOut(C.Structure):
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("pitch", C.c_int64),
                ("grey", C.c_void_p), ("edge", C.c_void_p)]


class PmvsCamera(C.Structure):
    _fields_ = [
        ("focal", C.c_double * 2), ("principal", C.c_double * 2), ("center", C.c_double * 3),
        ("R", C.c_double * 9), ("t", C.c_double * 3), ("KR", C.c_double * 9), ("KT", C.c_double * 3),
        ("opticalNormal", C.c_double * 3), ("maxLOD", C.c_int32), ("_pad", C.c_int32),
        ("level", PmvsLevel * MAX_LEVELS),
    ]


class PmvsHypothesis(C.Structure):
    _fields_ = [
        ("ray", C.c_double * 3), ("theta", C.c_double), ("phi", C.c_double), ("depth", C.c_double),
        ("refCamIdx", C.c_int32), ("LOD", C.c_int32), ("nCam", C.c_int32),
        ("camIdx", C.c_uint16 * MAX_VIEWS), ("_pad", C.c_int32),
    ]


class PmvsPatchIn(C.Structure):
    _fields_ = [
        ("center", C.c_double * 3), ("normal", C.c_double * 3), ("normalS", C.c_double * 2),
        ("type", C.c_int32), ("id", C.c_int32), ("nCam", C.c_int32), ("_pad", C.c_int32),
        ("camIdx", C.c_uint16 * MAX_VIEWS),
    ]


class PmvsPatchOut(C.Structure):
    _fields_ = [
        ("center", C.c_double * 3), ("normal", C.c_double * 3), ("normalS", C.c_double * 2), ("ray", C.c_double * 3),
        ("depth", C.c_double), ("depthRange", C.c_double * 2),
        ("fitness", C.c_double), ("priority", C.c_double), ("correlation", C.c_double),
        ("LOD", C.c_int32), ("refCamIdx", C.c_int32), ("nCam", C.c_int32), ("drop", C.c_int32),
        ("psoRuns", C.c_int32), ("psoIterations", C.c_int32), ("evaluations", C.c_uint32), ("status", C.c_uint32),
        ("camIdx", C.c_uint16 * MAX_VIEWS),
        ("nImgPoint", C.c_int32), ("windowEvaluations", C.c_uint32),
        ("imgPoint", (C.c_double * 2) * MAX_VIEWS),
    ]


assert C.sizeof(PmvsConfig) == 160
assert PmvsConfig.textureVariation.offset == 16 and PmvsConfig.minLOD.offset == 56
assert PmvsConfig.reduceNormalRange.offset == 72 and PmvsConfig.adaptiveDistanceEnable.offset == 80
assert PmvsConfig.distWeighting.offset == 88 and PmvsConfig.neighborRadius.offset == 112
assert PmvsConfig.particleNum.offset == 144 and PmvsConfig.expansionStrategy.offset == 152


def default_config():
    """setInitConfig, TMVS/TMVS.cpp:26-52 (compiled defaults)."""
    c = PmvsConfig()
    c.cellSize = 4
    c.patchRadius = 15
    c.patchSize = 31
    c.reduceNormalRange = 2
    c.adaptiveDistanceEnable = 1
    c.adaptiveDifferenceEnable = 1
    c.adaptiveGradientEnable = 0
    c.distWeighting = c.patchRadius / 3.0
    c.diffWeighting = 128.0 * 128.0
    c.gradientWeighting = 10.0
    c.minCamNum = 3
    c.textureVariation = 36
    c.visibleCorrelation = 0.7
    c.minCorrelation = 0.7
    c.maxFitness = 10.0
    c.minLOD = 0
    c.maxLOD = 15
    c.lodRatio = 0.8
    c.maxCellPatchNum = 3
    c.neighborRadius = 0.005
    c.neighborRadiusScalar = 0.0025
    c.minRegionRatio = 0.55
    c.depthRangeScalar = 1
    c.particleNum = 5
    c.maxIteration = 10
    c.expansionStrategy = 0
    return c


def readme_config():
    """The README sample config.txt (README.md:110-207) applied over the compiled defaults."""
    c = default_config()
    c.depthRangeScalar = 8
    c.particleNum = 15
    c.maxIteration = 30
    c.cellSize = 2
    c.minCorrelation = 0.9
    c.minRegionRatio = 0.15
    c.neighborRadiusScalar = 0.01
    c.distWeighting = 5
    return c
