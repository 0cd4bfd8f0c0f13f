"""ctypes front-end of the CPU oracle (oracle/liborc.so, oracle/_ref/libpso_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under pais-mvs_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "pais-mvs_b200", "python"))
from pmvs_b200 import abi  # noqa: E402

_LIB = None
_REF = None


def build(force=False):
    """Compile liborc.so (and, when /root/reference exists, the unmodified reference PSO) via oracle/Makefile."""
    so = os.path.join(HERE, "liborc.so")
    src = os.path.join(HERE, "pmvs_oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "pmvs_b200.h")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or stale or (os.path.isdir("/root/reference/TMVS/pso") and
                          not os.path.exists(os.path.join(HERE, "_ref", "libpso_ref.so"))):
        subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liborc.so"))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.PmvsConfig), C.c_int, C.POINTER(abi.PmvsCamera), C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_config.argtypes = [C.c_void_p, C.POINTER(abi.PmvsConfig)]
        L.orc_set_neighbor_radius.argtypes = [C.c_void_p, C.c_double]
        L.orc_set_ref_pso.argtypes = [C.c_void_p]
        L.orc_dist_weight.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_fitness_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.PmvsHypothesis), C.POINTER(C.c_double), C.c_int]
        L.orc_homographies.argtypes = [C.c_void_p, C.POINTER(abi.PmvsHypothesis), C.POINTER(C.c_double)]
        L.orc_refine_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.PmvsPatchIn), C.POINTER(abi.PmvsPatchOut),
                                       C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.orc_fit_ellipse_ratio.restype = C.c_double
        L.orc_fit_ellipse_ratio.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float),
                                            C.POINTER(C.c_float)]
        L.orc_region_ratio.restype = C.c_double
        L.orc_region_ratio.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_project.restype = C.c_int
        L.orc_project.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.orc_inv3.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_rand31.restype = C.c_uint32
        L.orc_rand31.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint64]
        L.orc_stream_key.restype = C.c_uint64
        L.orc_stream_key.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_neighbor_counts.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_int)]
        L.orc_pso_test.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int,
                                   C.POINTER(C.c_double), C.c_uint64, C.c_int, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_test_fn.restype = C.c_double
        L.orc_test_fn.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.orc_test_fn_ptr.restype = C.c_void_p
        _LIB = L
    return _LIB


def ref_lib():
    """The UNMODIFIED reference PSO (TMVS/pso/*.cpp compiled in place); None when it was never built."""
    global _REF
    if _REF is None:
        p = os.path.join(HERE, "_ref", "libpso_ref.so")
        if not os.path.exists(p):
            build()
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_pso_solve_ptr.restype = C.c_void_p
        R.ref_pso_solve_basic.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.POINTER(C.c_double), C.c_uint64, C.c_int, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _REF = R
    return _REF


class Oracle:
    """One oracle scene (cfg + cameras + pyramids copied in)."""

    def __init__(self, cfg, records, seed=42, use_ref_pso=False):
        self.L = lib()
        self.n_cams = len(records)
        self.h = self.L.orc_create(C.byref(cfg), self.n_cams, records, seed)
        self.pso_mode = 0
        if use_ref_pso:
            R = ref_lib()
            if R is None:
                raise RuntimeError("oracle/_ref/libpso_ref.so is not available")
            self.L.orc_set_ref_pso(R.ref_pso_solve_ptr())
            self.pso_mode = 1

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_config(self, cfg):
        self.L.orc_set_config(self.h, C.byref(cfg))

    def set_neighbor_radius(self, r):
        self.L.orc_set_neighbor_radius(self.h, r)

    def dist_weight(self, patch_size):
        out = (C.c_double * (patch_size * patch_size))()
        self.L.orc_dist_weight(self.h, out)
        return list(out)

    def fitness_batch(self, hyps, threads=1):
        n = len(hyps)
        out = (C.c_double * n)()
        self.L.orc_fitness_batch(self.h, n, hyps, out, threads)
        return list(out)

    def homographies(self, hyp):
        out = (C.c_double * (9 * hyp.nCam))()
        self.L.orc_homographies(self.h, C.byref(hyp), out)
        return list(out)

    def refine_batch(self, patches, flags=0, pso_mode=None, patch_threads=1, pso_threads=1):
        n = len(patches)
        out = (abi.PmvsPatchOut * n)()
        mode = self.pso_mode if pso_mode is None else pso_mode
        self.L.orc_refine_batch(self.h, n, patches, out, flags, mode, patch_threads, pso_threads)
        return out


def neighbor_counts(centers, radius):
    """orc_neighbor_counts: centers [n,3] float64 -> int32 [n] (mvs.cpp:470-499)."""
    import numpy as np
    c = np.ascontiguousarray(centers, dtype=np.float64)
    out = np.zeros(len(c), dtype=np.int32)
    lib().orc_neighbor_counts(len(c), c.ctypes.data_as(C.POINTER(C.c_double)), float(radius), out.ctypes.data_as(C.POINTER(C.c_int)))
    return out
