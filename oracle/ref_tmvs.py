"""ctypes front-end of oracle/_ref/libtmvs_ref.so: the UNMODIFIED reference patch model (TMVS/mvs/patch.cpp, abstractpatch.cpp,
camera.cpp, cellmap.cpp, mvs.cpp + TMVS/pso/*.cpp) compiled in place against oracle/cvshim (oracle/Makefile, ref_patch_shim.cpp).

TEST INFRASTRUCTURE ONLY: the pin of the f64 restatement (oracle/pmvs_oracle.cpp). The reference keeps its scene in a process-wide
singleton (MVS::getInstance), so there is ONE reference scene per process: RefScene() replaces the previous one.
"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "pais-mvs_b200", "python"))
from pmvs_b200 import abi  # noqa: E402

_LIB = None
PATH = os.path.join(HERE, "_ref", "libtmvs_ref.so")


def lib():
    """None when the library was never built (no /root/reference on this box and no prebuilt oracle/_ref)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(PATH):
            import orc
            orc.build()
        if not os.path.exists(PATH):
            return None
        L = C.CDLL(PATH)
        L.ref_scene_create.argtypes = [C.POINTER(abi.PmvsConfig), C.c_int, C.POINTER(abi.PmvsCamera), C.c_uint64]
        L.ref_set_neighbor_radius.argtypes = [C.c_double]
        L.ref_dist_weight.argtypes = [C.POINTER(C.c_double)]
        L.ref_normal2spherical.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_spherical2normal.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_project.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.ref_fitness_batch.argtypes = [C.c_int, C.POINTER(abi.PmvsHypothesis), C.POINTER(C.c_double)]
        L.ref_homographies.argtypes = [C.POINTER(abi.PmvsHypothesis), C.POINTER(C.c_double)]
        L.ref_region_ratio.restype = C.c_double
        L.ref_region_ratio.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_refine_batch.argtypes = [C.c_int, C.POINTER(abi.PmvsPatchIn), C.POINTER(abi.PmvsPatchOut), C.c_uint32]
        L.ref_run_reconstruction.argtypes = [C.c_int, C.POINTER(abi.PmvsPatchIn)]
        L.ref_neighbor_radius.restype = C.c_double
        L.ref_get_patch.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


class RefScene:
    def __init__(self, cfg, records, seed=42):
        self.L = lib()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libtmvs_ref.so is not available")
        self.n_cams = len(records)
        self.ps = 2 * cfg.patchRadius + 1
        self.L.ref_scene_create(C.byref(cfg), self.n_cams, records, seed)

    def set_threads(self, n):
        """OpenMP threads of the reference's own loops (particles); > 1 is for timing only (racy rand(), as in the reference)."""
        self.L.ref_set_threads(int(n))

    def set_neighbor_radius(self, r):
        self.L.ref_set_neighbor_radius(r)

    def dist_weight(self):
        out = (C.c_double * (self.ps * self.ps))()
        self.L.ref_dist_weight(out)
        return list(out)

    def normal2spherical(self, n):
        s = (C.c_double * 2)()
        self.L.ref_normal2spherical((C.c_double * 3)(*n), s)
        return list(s)

    def spherical2normal(self, s):
        n = (C.c_double * 3)()
        self.L.ref_spherical2normal((C.c_double * 2)(*s), n)
        return list(n)

    def project(self, cam, X, lod=0):
        out = (C.c_double * 2)()
        ok = self.L.ref_project(cam, (C.c_double * 3)(*X), lod, out)
        return ok, list(out)

    def fitness_batch(self, hyps):
        n = len(hyps)
        out = (C.c_double * n)()
        self.L.ref_fitness_batch(n, hyps, out)
        return list(out)

    def homographies(self, hyp):
        out = (C.c_double * (9 * hyp.nCam))()
        self.L.ref_homographies(C.byref(hyp), out)
        return list(out)

    def region_ratio(self, pt, H):
        return self.L.ref_region_ratio((C.c_double * 2)(*pt), (C.c_double * 9)(*H))

    def refine_batch(self, patches, flags=0):
        n = len(patches)
        out = (abi.PmvsPatchOut * n)()
        self.L.ref_refine_batch(n, patches, out, flags)
        return out

    def run_reconstruction(self, seeds):
        """The reference's own seed loop + MVS::expansionPatches() (mvs.cpp:196-275, unmodified) from these seed records
        (ids 0..n-1). Returns the final patch container in id order: dicts of exact values."""
        n = self.L.ref_run_reconstruction(len(seeds), seeds)
        out = []
        for k in range(n):
            pid, nc, ex = C.c_int(), C.c_int(), C.c_int()
            c, nr, sc = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
            ci, ip = (C.c_int * abi.MAX_VIEWS)(), (C.c_double * (2 * abi.MAX_VIEWS))()
            r = self.L.ref_get_patch(k, C.byref(pid), c, nr, sc, C.byref(nc), ci, ip, C.byref(ex))
            assert r > 0
            out.append(dict(id=pid.value, center=list(c), normal=list(nr), fitness=sc[0], priority=sc[1], correlation=sc[2],
                            cam_idx=list(ci[:nc.value]), img_point=[(ip[2 * i], ip[2 * i + 1]) for i in range(r - 1)], expanded=bool(ex.value)))
        return out

    def neighbor_radius(self):
        return self.L.ref_neighbor_radius()
