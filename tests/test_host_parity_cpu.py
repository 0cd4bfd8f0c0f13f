"""Caller-side parity (last row): the host driver's C++ restatement of the reference functions around
Patch::refine() — Camera ctor / project, getExpansionPatchCenter, skipNeighborCell, runtimeFiltering, insertPatch /
deletePatch / cell maps, the four queue strategies, isNeighbor, reCentering, setNeighborRadius — against oracle/orc_host.py
(plain-float restatement of the cited reference lines). Geometry is compared BIT FOR BIT (both sides are unfused f64 in the
reference's expression order); reCentering to 1e-12 (the reference solves with OpenCV's SVD inverse). No GPU involved:
the hooks library never creates a device context."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import orc_host as oh
from pmvs_b200 import abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HOOKS = os.path.join(ROOT, "pais-mvs_b200", "lib", "libtmvs_host.so")
D3, D4 = C.c_double * 3, C.c_double * 4


@pytest.fixture(scope="module")
def hooks():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "pais-mvs_b200", "host")])
    L = C.CDLL(HOOKS)
    L.tmvs_hook_create.restype = C.c_void_p
    L.tmvs_hook_create.argtypes = [C.POINTER(abi.PmvsConfig)]
    L.tmvs_hook_destroy.argtypes = [C.c_void_p]
    L.tmvs_hook_set_neighbor_radius_value.argtypes = [C.c_void_p, C.c_double]
    L.tmvs_hook_add_camera.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.tmvs_hook_project.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    patch_args = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.tmvs_hook_put_patch.argtypes = patch_args
    L.tmvs_hook_put_patch.restype = None
    L.tmvs_hook_runtime_filtering.argtypes = patch_args
    L.tmvs_hook_insert_patch.argtypes = patch_args
    for f in ("tmvs_hook_set_cell_maps", "tmvs_hook_init_queue", "tmvs_hook_recentering"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = None
    for f in ("tmvs_hook_pop", "tmvs_hook_patch_count", "tmvs_hook_deleted_count"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.tmvs_hook_delete_patch.argtypes = [C.c_void_p, C.c_int]
    L.tmvs_hook_set_expanded.argtypes = [C.c_void_p, C.c_int]
    L.tmvs_hook_cell.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.tmvs_hook_map_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.tmvs_hook_expansion_center.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tmvs_hook_skip_neighbor_cell.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.tmvs_hook_is_neighbor.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    L.tmvs_hook_set_neighbor_radius.argtypes = [C.c_void_p]
    L.tmvs_hook_set_neighbor_radius.restype = C.c_double
    L.tmvs_hook_get_patch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def _look_at_quaternion(center, target):
    from pmvs_b200 import scene
    return scene.R_to_quat(scene._look_at(np.asarray(center, float), np.asarray(target, float)))


class Pair:
    """The same scene on both sides: hooks handle + oracle MVS."""

    def __init__(self, L, cfg, n_cams=5, cols=320, rows=240, seed=1, with_grey=True):
        self.L, self.cfg = L, cfg
        self.h = L.tmvs_hook_create(C.byref(cfg))
        rng = np.random.RandomState(seed)
        cams = []
        for i in range(n_cams):
            ang = math.radians(-20 + 40.0 * i / max(1, n_cams - 1))
            center = [10 * math.sin(ang) + 0.1 * rng.randn(), 0.3 * rng.randn(), -10 * math.cos(ang)]
            q = _look_at_quaternion(center, [0.2 * rng.randn(), 0.2 * rng.randn(), 0.0]) * (1.0 + 0.3 * rng.rand())   # un-normalised on purpose
            grey = None
            if with_grey:
                grey = rng.randint(1, 256, size=(rows, cols)).astype(np.uint8)
                grey[rng.rand(rows, cols) < 0.03] = 0                                    # background pixels (mvs.cpp:860)
            focal = 1.2 * cols
            qa, ca = D4(*q), D3(*center)
            idx = L.tmvs_hook_add_camera(self.h, focal, qa, ca, cols, rows, grey.ctypes.data if with_grey else None)
            assert idx == i
            cams.append(oh.Camera(focal, list(q), center, cols, rows, grey))
        self.o = oh.MVS(cfg, cams)
        self.rng = rng

    def close(self):
        self.L.tmvs_hook_destroy(self.h)

    @staticmethod
    def _args(p):
        n = len(p.cam_idx)
        ci = (C.c_int * max(1, n))(*p.cam_idx)
        ip = (C.c_double * max(1, 2 * n))(*[v for pt in p.img_point for v in pt]) if p.img_point else None
        return D3(*p.center), D3(*p.normal), p.fitness, p.priority, p.correlation, n, ci, ip

    def put(self, p):
        c, n, fit, pri, cor, nc, ci, ip = self._args(p)
        self.L.tmvs_hook_put_patch(self.h, p.id, c, n, fit, pri, cor, nc, ci, ip, int(p.expanded))
        self.o.patches[p.id] = p

    def random_patch(self, pid, spread=1.5, good=True):
        """A patch near the z = 0 plane whose image points are its true projections."""
        rng = self.rng
        center = [spread * (2 * rng.rand() - 1), spread * (2 * rng.rand() - 1), 0.3 * rng.randn()]
        nrm = np.array([0.3 * rng.randn(), 0.3 * rng.randn(), -1.0])
        nrm /= np.linalg.norm(nrm)
        cams = sorted(rng.choice(len(self.o.cameras), size=rng.randint(3, len(self.o.cameras) + 1), replace=False).tolist())
        pts = [self.o.cameras[ci].project(center, 0, self.cfg.lodRatio)[0] for ci in cams]
        fit = float(rng.rand() * 5 + 0.01) if good else float(rng.choice([0.0, 50.0, float("nan"), 2.0]))
        pri = float(rng.rand() * 100)
        cor = float(0.9 + 0.1 * rng.rand()) if good else float(rng.choice([0.5, 0.95, float("nan")]))
        return oh.Patch(pid, center, nrm.tolist(), fit, pri, cor, cams, pts)


@pytest.fixture()
def cfg():
    c = abi.readme_config()
    c.cellSize, c.maxCellPatchNum, c.minCamNum = 4, 3, 3
    c.maxFitness, c.minCorrelation, c.neighborRadiusScalar = 10.0, 0.9, 0.01
    return c


def test_camera_and_project_bit_exact(hooks, cfg):
    P = Pair(hooks, cfg, seed=3)
    rng = np.random.RandomState(0)
    out = (C.c_double * 2)()
    n_in = 0
    for k in range(2000):
        X = [3 * rng.randn(), 3 * rng.randn(), 2 * rng.randn()]
        if k % 97 == 0:
            X = list(P.o.cameras[k % 5].center)                         # z2 ~ 0: huge / inf / nan path of inImage
        for ci, cam in enumerate(P.o.cameras):
            got_in = hooks.tmvs_hook_project(P.h, ci, D3(*X), 0, out)
            (u, v), want_in = cam.project(X, 0, cfg.lodRatio)
            assert got_in == int(want_in)
            assert (out[0] == u or (math.isnan(out[0]) and math.isnan(u))) and (out[1] == v or (math.isnan(out[1]) and math.isnan(v)))
            n_in += got_in
    assert 1000 < n_in < 9000                                           # both outcomes exercised
    P.close()


def test_expansion_center_bit_exact_and_on_the_parent_plane(hooks, cfg):
    P = Pair(hooks, cfg, seed=4)
    for pid in range(40):
        P.put(P.random_patch(pid))
    out = D3()
    for pid in range(40):
        parent = P.o.patches[pid]
        for ci, cam in enumerate(P.o.cameras):
            for cx, cy in [(0, 0), (17, 23), (79, 59), (40, 30)]:
                hooks.tmvs_hook_expansion_center(P.h, ci, pid, cx, cy, out)
                want = P.o.expansion_patch_center(cam, parent, cx, cy)
                assert list(out) == want                                # mvs.cpp:809-836, bit for bit
                # closed form: the point lies on the parent's plane and projects to the centre of cell (cx, cy)
                d = [want[k] - parent.center[k] for k in range(3)]
                assert abs(oh.dot3(d, parent.normal)) < 1e-9
                (u, v), _ = cam.project(want, 0, cfg.lodRatio)
                assert abs(u - (cx + 0.5) * cfg.cellSize) < 1e-7 and abs(v - (cy + 0.5) * cfg.cellSize) < 1e-7
    P.close()


def test_is_neighbor_and_radius(hooks, cfg):
    rng = np.random.RandomState(5)
    flips = 0
    for k in range(3000):
        c1, c2 = rng.randn(3), rng.randn(3) * (0.01 if k % 2 else 1.0)
        if k % 2:
            c2 = c1 + c2
        n1, n2 = rng.randn(3), rng.randn(3)
        n1, n2 = n1 / np.linalg.norm(n1), n2 / np.linalg.norm(n2)
        a, b = oh.Patch(0, c1, n1), oh.Patch(1, c2, n2)
        d = [a.center[i] - b.center[i] for i in range(3)]
        dist = abs(oh.dot3(d, a.normal)) + abs(oh.dot3(d, b.normal))
        for r in (dist, math.nextafter(dist, 0.0), math.nextafter(dist, 10.0), 0.01):      # exactly at the threshold: `<=`
            got = hooks.tmvs_hook_is_neighbor(D3(*a.center), D3(*a.normal), D3(*b.center), D3(*b.normal), r)
            assert got == int(oh.is_neighbor(a, b, r))
            flips += got
    assert 0 < flips < 12000
    P = Pair(hooks, cfg, seed=6)
    for pid in range(25):
        P.put(P.random_patch(pid))
    assert hooks.tmvs_hook_set_neighbor_radius(P.h) == P.o.set_neighbor_radius()            # pow(volume, 1/3) * scalar, mvs.cpp:147-152
    P.close()


def _cells_equal(P):
    wh = (C.c_int * 2)()
    buf = (C.c_int * 64)()
    for ci, cm in enumerate(P.o.cell_maps):
        P.L.tmvs_hook_map_size(P.h, ci, wh)
        assert (wh[0], wh[1]) == (cm.width, cm.height)
        for y in range(cm.height):
            for x in range(cm.width):
                n = P.L.tmvs_hook_cell(P.h, ci, x, y, buf, 64)
                assert list(buf[:n]) == cm.map[y][x], (ci, x, y)
    assert P.L.tmvs_hook_cell(P.h, 0, -1, 0, buf, 64) == -1 and P.L.tmvs_hook_cell(P.h, 0, 0, P.o.cell_maps[0].height, buf, 64) == -1


def test_runtime_filtering_insert_delete_and_cell_maps(hooks, cfg):
    P = Pair(hooks, cfg, cols=160, rows=120, seed=7)
    for pid in range(60):
        P.put(P.random_patch(pid, spread=0.6))
    hooks.tmvs_hook_set_cell_maps(P.h)
    P.o.set_cell_maps()
    _cells_equal(P)
    verdicts = {True: 0, False: 0}
    next_id = 60
    for k in range(1500):
        p = P.random_patch(next_id, spread=(0.6 if k % 3 else 4.0), good=(k % 4 != 0))
        if k % 11 == 0:
            p.normal = [-v for v in p.normal]                           # faces away from its cameras (:867-875)
        if k % 13 == 0:
            p.drop = True
        if k % 17 == 0:
            p.priority = 10000.5
        if k % 19 == 0:
            p.cam_idx, p.img_point = p.cam_idx[:2], p.img_point[:2]     # fewer than minCamNum
        c, n, fit, pri, cor, nc, ci, ip = P._args(p)
        inside = all(P.o.cameras[i].project(p.center, 0, cfg.lodRatio)[1] for i in p.cam_idx)
        if not inside:                                                   # image points must index a cell (the reference reads unchecked)
            p.img_point = [(1.0, 1.0)] * len(p.cam_idx)
            c, n, fit, pri, cor, nc, ci, ip = P._args(p)
        want = P.o.runtime_filtering(p)
        assert hooks.tmvs_hook_runtime_filtering(P.h, p.id, c, n, fit, pri, cor, nc, ci, ip, int(p.drop)) == int(want)
        verdicts[want] += 1
        if k % 2 == 0:
            assert hooks.tmvs_hook_insert_patch(P.h, p.id, c, n, fit, pri, cor, nc, ci, ip, int(p.drop)) == int(P.o.insert_patch(p))
            next_id += 1
        if k % 7 == 3 and P.o.patches:
            victim = sorted(P.o.patches)[(k * 31) % len(P.o.patches)]
            hooks.tmvs_hook_delete_patch(P.h, victim)
            P.o.delete_patch(victim)
        assert hooks.tmvs_hook_patch_count(P.h) == len(P.o.patches) and hooks.tmvs_hook_deleted_count(P.h) == len(P.o.deleted)
    assert verdicts[True] > 100 and verdicts[False] > 100
    _cells_equal(P)
    # full cells were reached: the cell-count rule (:878-895) was exercised
    assert any(len(c) >= cfg.maxCellPatchNum for cm in P.o.cell_maps for row in cm.map for c in row)
    P.close()


def test_skip_neighbor_cell(hooks, cfg):
    cfg.minCorrelation = 0.95
    P = Pair(hooks, cfg, cols=160, rows=120, seed=8, with_grey=False)
    for pid in range(80):
        p = P.random_patch(pid, spread=0.5)
        p.correlation = 0.9 + 0.1 * P.rng.rand()                         # both sides of minCorrelation
        P.put(p)
    hooks.tmvs_hook_set_cell_maps(P.h)
    P.o.set_cell_maps()
    outcomes = {True: 0, False: 0}
    for radius in (1e-4, 0.02, 0.3):
        hooks.tmvs_hook_set_neighbor_radius_value(P.h, radius)
        P.o.neighbor_radius = radius
        for ci, cm in enumerate(P.o.cell_maps):
            for y in range(cm.height):
                for x in range(cm.width):
                    if not cm.map[y][x] and (x + y) % 9:
                        continue
                    for ref in (0, 17, 42):
                        want = P.o.skip_neighbor_cell(cm.map[y][x], P.o.patches[ref])
                        assert hooks.tmvs_hook_skip_neighbor_cell(P.h, ci, x, y, ref) == int(want)
                        outcomes[want] += 1
    assert outcomes[True] > 50 and outcomes[False] > 50
    P.close()


@pytest.mark.parametrize("strategy", [oh.BEST_FIRST, oh.WORST_FIRST, oh.BREATH_FIRST, oh.DEPTH_FIRST])
def test_queue_strategies_follow_the_reference_scans(hooks, cfg, strategy):
    """The indexed queue pops what the reference's linear scans pop (mvs.cpp:636-788): ties in priority go to the earliest
    queued entry, expanded / deleted entries are skipped, NaN priorities are never selected. Depth-first: the reference's
    scan never examines the first queued entry and erases it (returning -1); the indexed queue returns it as its last pop —
    the one documented difference."""
    cfg.expansionStrategy = strategy
    P = Pair(hooks, cfg, cols=160, rows=120, seed=9 + strategy, with_grey=False)
    rng = np.random.RandomState(100 + strategy)
    for pid in range(50):
        p = P.random_patch(pid, spread=0.5)
        p.priority = float(rng.randint(0, 8))                             # many ties
        if pid in (7, 30):
            p.priority = float("nan")
        if pid == 11:
            p.priority = oh.DBL_MAX
        P.put(p)
    hooks.tmvs_hook_set_cell_maps(P.h)
    P.o.set_cell_maps()
    hooks.tmvs_hook_init_queue(P.h)
    P.o.init_priority_queue()
    first_queued = P.o.queue[0]
    next_id, popped = 50, []
    for step in range(400):
        want = P.o.pop()
        got = hooks.tmvs_hook_pop(P.h)
        if strategy == oh.DEPTH_FIRST and want == -1 and got != -1:
            assert got == first_queued and hooks.tmvs_hook_pop(P.h) == -1
            break
        assert got == want, (step, popped[-5:])
        if want == -1:
            break
        popped.append(want)
        hooks.tmvs_hook_set_expanded(P.h, want)
        P.o.patches[want].expanded = True
        for _ in range(rng.randint(0, 3) if step < 120 else 0):          # children enter the queue through insertPatch
            p = P.random_patch(next_id, spread=0.5)
            p.priority = float(rng.randint(0, 8))
            c, n, fit, pri, cor, nc, ci, ip = P._args(p)
            assert hooks.tmvs_hook_insert_patch(P.h, p.id, c, n, fit, pri, cor, nc, ci, ip, 0) == int(P.o.insert_patch(p))
            next_id += 1
        if step % 5 == 2:                                                 # something still queued disappears (filters, parent deletion)
            live = [q for q in P.o.queue if q in P.o.patches]
            if live:
                victim = live[rng.randint(len(live))]
                hooks.tmvs_hook_delete_patch(P.h, victim)
                P.o.delete_patch(victim)
        if step % 9 == 4:                                                 # or gets expanded while queued
            live = [q for q in P.o.queue if q in P.o.patches]
            if live:
                victim = live[rng.randint(len(live))]
                hooks.tmvs_hook_set_expanded(P.h, victim)
                P.o.patches[victim].expanded = True
    assert len(popped) > 60 and len(set(popped)) == len(popped)
    if strategy in (oh.BEST_FIRST, oh.WORST_FIRST):
        assert 7 not in popped and 30 not in popped                       # NaN priority: never selected by < or >
    if strategy == oh.BEST_FIRST:
        assert 11 not in popped                                           # DBL_MAX < DBL_MAX is false (:682)
    P.close()


def test_recentering(hooks, cfg):
    P = Pair(hooks, cfg, seed=21, with_grey=False)
    truth = {}
    for pid in range(30):
        p = P.random_patch(pid)
        truth[pid] = list(p.center)
        p.img_point = [(u + 0.2 * P.rng.randn(), v + 0.2 * P.rng.randn()) for (u, v) in p.img_point]      # measurement noise
        p.center = [0.0, 0.0, 0.0]
        if pid == 5:
            p.cam_idx, p.img_point = p.cam_idx[:2], p.img_point[:2]       # < minCamNum: setEstimatedNormal drops it (patch.cpp:396-400)
        P.put(p)
    hooks.tmvs_hook_recentering(P.h)
    c, n, s = D3(), D3(), (C.c_double * 2)()
    for pid in range(30):
        p = P.o.patches[pid]
        oh.recentering(P.o, p)
        drop = hooks.tmvs_hook_get_patch(P.h, pid, c, n, s)
        assert drop == int(p.drop) == int(pid == 5)
        assert np.allclose(list(c), p.center, rtol=0, atol=1e-11)
        assert np.linalg.norm(np.array(list(c)) - truth[pid]) < 0.05      # and it is the triangulated point
        if not p.drop:
            assert np.allclose(list(n), p.normal, rtol=0, atol=1e-11) and np.allclose(list(s), p.normalS, rtol=0, atol=1e-11)
            assert abs(np.linalg.norm(list(n)) - 1) < 1e-14
    P.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_cell_id_container_matches_a_vector(hooks, seed):
    """CellIds (host/tmvs.h: cell ids inline in the cell record, heap beyond five) behaves like the std::vector<int> the
    reference keeps per cell (cellmap.h) under random push / erase-first / copy traffic across the inline <-> heap boundary."""
    L = hooks
    L.tmvs_hook_cellids_selftest.argtypes = [C.c_uint, C.c_int]
    L.tmvs_hook_cellids_selftest.restype = C.c_int
    assert L.tmvs_hook_cellids_selftest(seed, 20000) == 0
