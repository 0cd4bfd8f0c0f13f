"""Loader of the C-ABI library (pais-mvs_b200/lib/libpmvs_b200.so, declared in include/pmvs_b200.h).

There is no fallback: if the library is missing, or no sm_100 device is usable, the calls raise."""
import ctypes as C
import os
import subprocess

from . import abi

PKG_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
LIB_PATH = os.environ.get("PMVS_LIB") or os.path.join(PKG_ROOT, "lib", "libpmvs_b200.so")      # PMVS_LIB: build variants (tuning)
CSRC = os.path.join(PKG_ROOT, "csrc")

# every symbol include/pmvs_b200.h declares
SYMBOLS = ["pmvs_create", "pmvs_set_neighbor_radius", "pmvs_set_config", "pmvs_fitness_batch", "pmvs_refine_batch",
           "pmvs_refine_batch_device", "pmvs_pack_records_device", "pmvs_launch_count", "pmvs_pso_test", "pmvs_destroy", "pmvs_last_error",
           "pmvs_version", "pmvs_pyramid_levels", "pmvs_build_pyramid", "pmvs_neighbor_counts"]

_LIB = None


class PmvsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pmvs error %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    """nvcc build of the library for sm_100a (cross-compiles without a GPU)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-s", "-C", CSRC])
    return LIB_PATH


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise PmvsError(abi_E_CUDA, "%s is missing: run `make -C %s` (or __graft_entry__.build()); there is no CPU path"
                        % (LIB_PATH, CSRC))
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.pmvs_create.restype = C.c_int
    L.pmvs_create.argtypes = [C.POINTER(vp), C.POINTER(abi.PmvsConfig), C.c_int, C.POINTER(abi.PmvsCamera), C.c_int, C.c_uint64]
    L.pmvs_set_neighbor_radius.restype = C.c_int
    L.pmvs_set_neighbor_radius.argtypes = [vp, C.c_double]
    L.pmvs_set_config.restype = C.c_int
    L.pmvs_set_config.argtypes = [vp, C.POINTER(abi.PmvsConfig)]
    L.pmvs_fitness_batch.restype = C.c_int
    L.pmvs_fitness_batch.argtypes = [vp, C.c_int, C.POINTER(abi.PmvsHypothesis), C.POINTER(C.c_double)]
    L.pmvs_refine_batch.restype = C.c_int
    L.pmvs_refine_batch.argtypes = [vp, C.c_int, C.POINTER(abi.PmvsPatchIn), C.POINTER(abi.PmvsPatchOut), C.c_uint32]
    L.pmvs_refine_batch_device.restype = C.c_int
    L.pmvs_refine_batch_device.argtypes = [vp, C.c_int, vp, vp, C.c_uint32, vp]
    L.pmvs_pack_records_device.restype = C.c_int
    L.pmvs_pack_records_device.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.pmvs_launch_count.restype = C.c_int64
    L.pmvs_launch_count.argtypes = [vp]
    L.pmvs_pso_test.restype = C.c_int
    L.pmvs_pso_test.argtypes = [vp, C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int)] * 4 + \
        [C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.pmvs_destroy.restype = None
    L.pmvs_destroy.argtypes = [vp]
    L.pmvs_last_error.restype = C.c_char_p
    L.pmvs_last_error.argtypes = [vp]
    L.pmvs_version.restype = C.c_char_p
    L.pmvs_pyramid_levels.restype = C.c_int
    L.pmvs_pyramid_levels.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.pmvs_neighbor_counts.restype = C.c_int
    L.pmvs_neighbor_counts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.pmvs_build_pyramid.restype = C.c_int
    L.pmvs_build_pyramid.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_int, C.c_int,
                                     C.POINTER(abi.PmvsLevelOut)]
    _LIB = L
    return L


abi_E_CUDA = -2
