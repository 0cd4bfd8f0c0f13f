"""Host-side mirror of the reference interface for the patch-refinement path.

`PatchRefiner` plays the part of the reference's MVS singleton for this path (TMVS/mvs/mvs.h:161-239: it owns the
config and the cameras) and exposes the two seams the C-ABI replaces: `fitness()` = PAIS::getFitness
(TMVS/mvs/patch.cpp:914) and `refine()` = Patch::refine() + removeInvisibleCamera() as called from
MVS::refineSeedPatches / MVS::expandCell (TMVS/mvs/mvs.cpp:214-215, :573-574). All compute happens in
libpmvs_b200.so on the GPU; errors surface as PmvsError, per-patch failure stays in-band (out.drop) like the reference.
"""
import ctypes as C

from . import abi
from .lib import PmvsError, load


class PatchRefiner:
    def __init__(self, cfg, camera_records, device=0, seed=42):
        self.L = load()
        self.h = C.c_void_p()
        self.n_cams = len(camera_records)
        rc = self.L.pmvs_create(C.byref(self.h), C.byref(cfg), self.n_cams, camera_records, device, seed)
        if rc != 0:
            msg = self.L.pmvs_last_error(self.h).decode() if self.h else "allocation failed"
            self.close()
            raise PmvsError(rc, msg)

    def _check(self, rc):
        if rc != 0:
            raise PmvsError(rc, self.L.pmvs_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.pmvs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_config(self, cfg):                       # MVS::setConfig, mvs.cpp:42-72
        self._check(self.L.pmvs_set_config(self.h, C.byref(cfg)))

    def set_neighbor_radius(self, r):                # MVS::setNeighborRadius, mvs.cpp:147-152
        self._check(self.L.pmvs_set_neighbor_radius(self.h, r))

    def fitness(self, hyps):
        n = len(hyps)
        out = (C.c_double * max(n, 1))()
        self._check(self.L.pmvs_fitness_batch(self.h, n, hyps, out))
        return list(out)[:n]

    def refine(self, patches, flags=abi.F_POST_REMOVE_INVISIBLE, out=None):
        n = len(patches)
        if out is None:
            out = (abi.PmvsPatchOut * max(n, 1))()
        self._check(self.L.pmvs_refine_batch(self.h, n, patches, out, flags))
        return out

    def refine_device(self, n, d_in, d_out, flags=abi.F_POST_REMOVE_INVISIBLE, stream=None):
        """Device-resident variant: d_in / d_out are device addresses (ints), launch is asynchronous."""
        self._check(self.L.pmvs_refine_batch_device(self.h, n, C.c_void_p(d_in), C.c_void_p(d_out), flags,
                                                    C.c_void_p(stream) if stream else None))

    def pack_records_device(self, n, d_out, d_records, d_counters=None, stream=None):
        """{centre, normal, fitness, drop} as 8 doubles per patch from device-resident results (the all-gather payload)."""
        self._check(self.L.pmvs_pack_records_device(self.h, n, C.c_void_p(d_out), C.c_void_p(d_records),
                                                    C.c_void_p(d_counters) if d_counters else None, C.c_void_p(stream) if stream else None))

    def launch_count(self):
        return int(self.L.pmvs_launch_count(self.h))

    def pso_test(self, problems):
        """problems: list of dict(L, U, init|None, maxIter, P, fn, key). Returns list of dict."""
        n = len(problems)
        D = C.c_double
        L = (D * (3 * n))(*[v for p in problems for v in p["L"]])
        U = (D * (3 * n))(*[v for p in problems for v in p["U"]])
        init = (D * (3 * n))(*[v for p in problems for v in (p["init"] if p.get("init") is not None else (0, 0, 0))])
        has = (C.c_int * n)(*[1 if p.get("init") is not None else 0 for p in problems])
        mi = (C.c_int * n)(*[p["maxIter"] for p in problems])
        P = (C.c_int * n)(*[p["P"] for p in problems])
        fn = (C.c_int * n)(*[p["fn"] for p in problems])
        keys = (C.c_uint64 * n)(*[p["key"] for p in problems])
        gb = (D * (3 * n))()
        gf = (D * n)()
        it = (C.c_int * n)()
        parts = (D * (n * 64 * 8))()
        self._check(self.L.pmvs_pso_test(self.h, n, L, U, init, has, mi, P, fn, keys, gb, gf, it, parts))
        res = []
        for i in range(n):
            pp = [list(parts[(i * 64 + k) * 8:(i * 64 + k) * 8 + 8]) for k in range(problems[i]["P"])]
            res.append(dict(gbest=list(gb[3 * i:3 * i + 3]), gbestFitness=gf[i], iterations=it[i], particles=pp))
        return res


def pyramid_levels(cols, rows, lod_ratio, cfg_max_lod):
    """camera.cpp:63-64: (maxLOD, [(cols, rows) per level]) — host-only."""
    L = load()
    m = C.c_int()
    lc = (C.c_int32 * abi.MAX_LEVELS)()
    lr = (C.c_int32 * abi.MAX_LEVELS)()
    rc = L.pmvs_pyramid_levels(cols, rows, lod_ratio, cfg_max_lod, C.byref(m), lc, lr)
    if rc != 0:
        raise PmvsError(rc, "pmvs_pyramid_levels: bad arguments")
    return m.value, [(lc[i], lr[i]) for i in range(m.value + 1)]


def build_pyramid(grey0, lod_ratio, cfg_max_lod, with_edge=True, device=0):
    """The Camera ctor's pyramid (camera.cpp:63-92) built on the GPU; returns [(grey u8, edge f64|None)] like
    scene.build_pyramid."""
    import numpy as np
    L = load()
    rows, cols = grey0.shape
    max_lod, dims = pyramid_levels(cols, rows, lod_ratio, cfg_max_lod)
    levels = (abi.PmvsLevelOut * (max_lod + 1))()
    out = []
    for l, (c, r) in enumerate(dims):
        g = np.ascontiguousarray(grey0) if l == 0 else np.empty((r, c), dtype=np.uint8)
        e = np.empty((r, c), dtype=np.float64) if with_edge else None
        levels[l].cols, levels[l].rows, levels[l].pitch = c, r, g.strides[0]
        levels[l].grey = g.ctypes.data
        levels[l].edge = e.ctypes.data if with_edge else None
        out.append((g, e))
    rc = L.pmvs_build_pyramid(device, out[0][0].ctypes.data, cols, rows, out[0][0].strides[0], lod_ratio, max_lod, int(with_edge), levels)
    if rc != 0:
        raise PmvsError(rc, "pmvs_build_pyramid failed")
    return out


def neighbor_counts(centers, radius, device=0, first=0, count=None):
    """The PCMVS neighbour filter's pair scan (MVS::neighborPatchFiltering, mvs.cpp:470-499) on the GPU: for the rows
    [first, first+count) of centers [n,3] f64, the number of OTHER points within `radius` (norm <= radius). int32."""
    import numpy as np
    L = load()
    c = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
    n = len(c)
    count = n - first if count is None else count
    out = np.zeros(max(count, 0), dtype=np.int32)
    rc = L.pmvs_neighbor_counts(device, n, c.ctypes.data, float(radius), first, count, out.ctypes.data)
    if rc != 0:
        raise PmvsError(rc, "pmvs_neighbor_counts failed")
    return out
