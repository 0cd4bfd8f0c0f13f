"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's caller-side functions around Patch::refine()
(last row), in plain Python floats (IEEE f64, unfused, the reference's expression order) so that the host
driver (pais-mvs_b200/host) can be compared bit for bit. Only tests/ may import this module.

Each function cites the reference lines it follows (paths relative to the reference tree). The reference's containers are
kept as they are there: `queue` is a vector scanned linearly for every pop, cells are vectors of patch ids.
"""
import math

DBL_MAX = 1.7976931348623157e308
BEST_FIRST, WORST_FIRST, BREATH_FIRST, DEPTH_FIRST = 0, 1, 2, 3       # TMVS/mvs/mvs.h


def cv_round(v):
    """cvRound = round half to even (SSE2 cvtsd2si)."""
    return int(round(v))        # Python's round() is half-to-even on floats


def ieee_div(a, b):
    """a / b with IEEE results for b == 0 (Python raises instead)."""
    if b != 0.0:
        return a / b
    if a == 0.0 or math.isnan(a):
        return float("nan")
    return math.copysign(float("inf"), a) * math.copysign(1.0, b)


def dot3(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


class Camera:
    """Camera ctor, TMVS/mvs/camera.cpp:45-136 (NVM form: principal point from the image size, :101-106)."""

    def __init__(self, focal, quaternion, center, cols, rows, grey=None):
        self.focal = (float(focal), float(focal))
        self.principal = (float(cols >> 1), float(rows >> 1))
        self.center = [float(c) for c in center]
        self.cols, self.rows, self.grey = cols, rows, grey
        q = [float(v) for v in quaternion]
        qq = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])       # camera.cpp:6-35
        if qq > 0:
            qw, qx, qy, qz = q[0] / qq, q[1] / qq, q[2] / qq, q[3] / qq
        else:
            qw, qx, qy, qz = 1.0, 0.0, 0.0, 0.0
        self.R = [qw * qw + qx * qx - qz * qz - qy * qy, 2 * qx * qy - 2 * qz * qw, 2 * qy * qw + 2 * qz * qx,
                  2 * qx * qy + 2 * qw * qz, qy * qy + qw * qw - qz * qz - qx * qx, 2 * qz * qy - 2 * qx * qw,
                  2 * qx * qz - 2 * qy * qw, 2 * qy * qz + 2 * qw * qx, qz * qz + qw * qw - qy * qy - qx * qx]
        R, C = self.R, self.center
        self.t = [-(R[3 * r] * C[0] + R[3 * r + 1] * C[1] + R[3 * r + 2] * C[2]) for r in range(3)]      # :120
        self.optical_normal = [R[6], R[7], R[8]]                                      # R^T (0,0,1), :130-133

    def project(self, X, lod=0, lod_ratio=0.8):
        """Camera::project + inImage, camera.cpp:138-160, camera.h:116-131 (level 0 dimensions only here)."""
        R, t = self.R, self.t
        x2 = (R[0] * X[0] + R[1] * X[1] + R[2] * X[2]) + t[0]
        y2 = (R[3] * X[0] + R[4] * X[1] + R[5] * X[2]) + t[1]
        z2 = (R[6] * X[0] + R[7] * X[1] + R[8] * X[2]) + t[2]
        u = self.focal[0] * ieee_div(x2, z2) + self.principal[0]
        v = self.focal[1] * ieee_div(y2, z2) + self.principal[1]
        sc = math.pow(lod_ratio, lod)
        u *= sc
        v *= sc
        if math.isnan(u) or math.isnan(v):
            return (u, v), False
        return (u, v), not (u < 0 or u >= self.cols or v < 0 or v >= self.rows)

    def back_project(self, px, py):
        """R^T * (K^-1 (px,py,1) - t): the `p3d` of mvs.cpp:822-826 and patch.cpp:82-86."""
        R, t = self.R, self.t
        q = ((px - self.principal[0]) / self.focal[0] - t[0], (py - self.principal[1]) / self.focal[1] - t[1], 1.0 - t[2])
        return [R[c] * q[0] + R[3 + c] * q[1] + R[6 + c] * q[2] for c in range(3)]


class Patch:
    def __init__(self, pid, center, normal, fitness=DBL_MAX, priority=DBL_MAX, correlation=0.0, cam_idx=(), img_point=(),
                 expanded=False, drop=False):
        self.id = pid
        self.center = [float(v) for v in center]
        self.normal = [float(v) for v in normal]
        self.normalS = [0.0, 0.0]
        self.fitness, self.priority, self.correlation = fitness, priority, correlation
        self.cam_idx = list(cam_idx)
        self.img_point = [(float(a), float(b)) for a, b in img_point]
        self.expanded, self.drop = expanded, drop


def is_neighbor(p1, p2, neighbor_radius):
    """Patch::isNeighbor, TMVS/mvs/patch.cpp:6-23."""
    d = [p1.center[k] - p2.center[k] for k in range(3)]
    dist = 0.0
    dist += abs(dot3(d, p1.normal))
    dist += abs(dot3(d, p2.normal))
    return dist <= neighbor_radius


class CellMap:
    """TMVS/mvs/cellmap.cpp:5-41."""

    def __init__(self, cam, cell_size):
        self.width = int(math.ceil(float(cam.cols) / float(cell_size)))
        self.height = int(math.ceil(float(cam.rows) / float(cell_size)))
        self.map = [[[] for _ in range(self.width)] for _ in range(self.height)]

    def in_map(self, x, y):
        return not (x < 0 or y < 0 or x >= self.width or y >= self.height)

    def insert(self, x, y, pid):
        if not self.in_map(x, y):
            return False
        self.map[y][x].append(pid)
        return True

    def drop(self, x, y, pid):
        if not self.in_map(x, y):
            return False
        cell = self.map[y][x]
        if pid not in cell:
            return False
        cell.remove(pid)        # first occurrence, like find + erase
        return True


class MVS:
    """The caller-side state of TMVS/mvs/mvs.cpp: cameras, patches (std::map: id order), cell maps, queue."""

    def __init__(self, cfg, cameras):
        self.cfg, self.cameras = cfg, cameras
        self.patches = {}
        self.deleted = []
        self.cell_maps = []
        self.queue = []
        self.neighbor_radius = 0.0

    # mvs.cpp:116-133 (+ initCellMaps :74-88)
    def set_cell_maps(self):
        cs = self.cfg.cellSize
        self.cell_maps = [CellMap(c, cs) for c in self.cameras]
        for pid in sorted(self.patches):
            p = self.patches[pid]
            for i, ci in enumerate(p.cam_idx):
                self.cell_maps[ci].insert(int(p.img_point[i][0] / cs), int(p.img_point[i][1] / cs), pid)

    def init_priority_queue(self):          # mvs.cpp:90-95
        self.queue = sorted(self.patches)

    def set_neighbor_radius(self):          # mvs.cpp:147-152, getBoundingVolume :967-990
        mn, mx = [DBL_MAX] * 3, [-DBL_MAX] * 3
        for pid in sorted(self.patches):
            c = self.patches[pid].center
            for i in range(3):
                if c[i] < mn[i]:
                    mn[i] = c[i]
                if c[i] > mx[i]:
                    mx[i] = c[i]
        vol = abs((mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]))
        self.neighbor_radius = math.pow(vol, 1.0 / 3.0) * self.cfg.neighborRadiusScalar
        return self.neighbor_radius

    def runtime_filtering(self, p):         # mvs.cpp:838-898
        cfg = self.cfg
        cam_num = len(p.cam_idx)
        if p.drop:
            return False
        if cam_num < cfg.minCamNum:
            return False
        if p.fitness > cfg.maxFitness:
            return False
        if p.fitness == 0.0:
            return False
        if p.priority > 10000:
            return False
        if math.isnan(p.fitness) or math.isnan(p.priority) or math.isnan(p.correlation):
            return False
        if p.correlation < cfg.minCorrelation:
            return False
        for cam in self.cameras:            # :851-863
            (u, v), inside = cam.project(p.center, 0, cfg.lodRatio)
            if not inside:
                return False
            if cam.grey is not None:
                # cvRound may land one past the last column/row (the reference then reads out of the row): clamped
                x, y = min(cv_round(u), cam.cols - 1), min(cv_round(v), cam.rows - 1)
                if cam.grey[y][x] == 0:
                    return False
        count = 0                           # :867-875
        for ci in p.cam_idx:
            on = self.cameras[ci].optical_normal
            if dot3(p.normal, [-on[0], -on[1], -on[2]]) > 0:
                count += 1
        if count < cfg.minCamNum:
            return False
        if not self.cell_maps:              # :878
            return True
        full = 0
        for i, ci in enumerate(p.cam_idx):
            cx, cy = int(p.img_point[i][0] / cfg.cellSize), int(p.img_point[i][1] / cfg.cellSize)
            cell = self.cell_maps[ci].map[cy][cx]
            if p.id in cell:
                return True
            if len(cell) >= cfg.maxCellPatchNum:
                full += 1
        if full >= cam_num:
            return False
        return True

    def insert_patch(self, p):              # mvs.cpp:579-601
        if not self.runtime_filtering(p):
            return False
        if p.id not in self.patches:        # std::map::insert keeps an existing entry
            self.patches[p.id] = p
        self.queue.append(p.id)
        cs = self.cfg.cellSize
        for i, ci in enumerate(p.cam_idx):
            self.cell_maps[ci].insert(int(p.img_point[i][0] / cs), int(p.img_point[i][1] / cs), p.id)
        return True

    def delete_patch(self, pid):            # mvs.cpp:607-634
        p = self.patches.get(pid)
        if p is None:
            return
        if self.cell_maps:
            cs = self.cfg.cellSize
            for i, ci in enumerate(p.cam_idx):
                self.cell_maps[ci].drop(int(p.img_point[i][0] / cs), int(p.img_point[i][1] / cs), pid)
        self.deleted.append(p)
        del self.patches[pid]

    def skip_neighbor_cell(self, cell, ref):        # mvs.cpp:792-807
        if len(cell) >= self.cfg.maxCellPatchNum:
            return True
        for pid in cell:
            p = self.patches.get(pid)
            if p is None:
                continue
            if p.correlation > self.cfg.minCorrelation:
                return True
            if is_neighbor(ref, p, self.neighbor_radius):
                return True
        return False

    def expansion_patch_center(self, cam, parent, cx, cy):      # mvs.cpp:809-836
        cs = self.cfg.cellSize
        px, py = (cx + 0.5) * cs, (cy + 0.5) * cs
        p3d = cam.back_project(px, py)
        v13 = [parent.center[k] - cam.center[k] for k in range(3)]
        v12 = [p3d[k] - cam.center[k] for k in range(3)]
        u = dot3(parent.normal, v13) / dot3(parent.normal, v12)
        return [cam.center[k] + u * v12[k] for k in range(3)]

    # --- queue pops, mvs.cpp:636-788: linear scans that erase dead entries on the way -------------------------------
    def _dead(self, pid):
        p = self.patches.get(pid)
        return p is None or p.expanded

    def pop(self):
        s = self.cfg.expansionStrategy
        if s == WORST_FIRST:
            return self._pop_priority(worst=True)
        if s == BREATH_FIRST:
            return self._pop_breadth()
        if s == DEPTH_FIRST:
            return self._pop_depth()
        return self._pop_priority(worst=False)

    def _pop_priority(self, worst):         # :656-690 (best, strict <), :692-726 (worst, strict >)
        self.queue = [q for q in self.queue if not self._dead(q)]
        top_i, top = -1, (-DBL_MAX if worst else DBL_MAX)
        for i, pid in enumerate(self.queue):
            pr = self.patches[pid].priority
            if (pr > top) if worst else (pr < top):
                top, top_i = pr, i
        if top_i < 0:
            return -1
        return self.queue.pop(top_i)

    def _pop_breadth(self):                 # :728-755 (an empty queue is undefined behaviour there: -1 here)
        while self.queue and self._dead(self.queue[0]):
            self.queue.pop(0)
        return self.queue.pop(0) if self.queue else -1

    def _pop_depth(self):                   # :757-788 — the scan stops AT queue.begin() without examining it, then erases it:
        while len(self.queue) > 1 and self._dead(self.queue[-1]):          # the first queued entry is never returned
            self.queue.pop()
        if len(self.queue) > 1:
            return self.queue.pop()
        if self.queue:
            self.queue.pop()
        return -1


def estimated_normal(mvs, p):
    """Patch::setEstimatedNormal, TMVS/mvs/patch.cpp:390-413 (+ utility.h:17-22)."""
    if p.drop:
        return
    if len(p.cam_idx) < mvs.cfg.minCamNum:
        p.drop = True
        return
    n = [0.0, 0.0, 0.0]
    for ci in p.cam_idx:
        C = mvs.cameras[ci].center
        d = [C[k] - p.center[k] for k in range(3)]
        inv = 1.0 / math.sqrt(dot3(d, d))
        for k in range(3):
            n[k] += d[k] * inv
    inv = 1.0 / math.sqrt(dot3(n, n))
    p.normal = [n[k] * inv for k in range(3)]
    p.normalS = [math.acos(p.normal[2]), math.atan2(p.normal[1], p.normal[0])]


def recentering(mvs, p):
    """Patch::reCentering, TMVS/mvs/patch.cpp:67-112: least-squares intersection of the viewing rays. The 3x3 system is
    returned as well; the reference solves it with cv::Mat::inv(DECOMP_SVD) (OpenCV 2.4.2, not in the reference tree),
    this restatement with numpy's pinv — compare to ~1e-12, not bit for bit."""
    import numpy as np
    A = [0.0] * 9
    b = [0.0] * 3
    for i, ci in enumerate(p.cam_idx):
        cam = mvs.cameras[ci]
        C = cam.center
        p3d = cam.back_project(p.img_point[i][0], p.img_point[i][1])
        n = [p3d[k] - C[k] for k in range(3)]
        inv = 1.0 / math.sqrt(dot3(n, n))
        n = [v * inv for v in n]
        A[0] += 1 - n[0] * n[0]; A[1] += -n[0] * n[1]; A[2] += -n[0] * n[2]
        A[3] += -n[0] * n[1]; A[4] += 1 - n[1] * n[1]; A[5] += -n[1] * n[2]
        A[6] += -n[0] * n[2]; A[7] += -n[1] * n[2]; A[8] += 1 - n[2] * n[2]
        b[0] += (1 - n[0] * n[0]) * C[0] - n[0] * n[1] * C[1] - n[0] * n[2] * C[2]
        b[1] += -n[0] * n[1] * C[0] + (1 - n[1] * n[1]) * C[1] - n[1] * n[2] * C[2]
        b[2] += -n[0] * n[2] * C[0] - n[1] * n[2] * C[1] + (1 - n[2] * n[2]) * C[2]
    x = np.linalg.pinv(np.array(A).reshape(3, 3)) @ np.array(b)
    p.center = [float(v) for v in x]
    estimated_normal(mvs, p)
    return A, b


def expand_neighbor_cell(mvs, p, refine):
    """MVS::expandNeighborCell + expandCell, TMVS/mvs/mvs.cpp:529-577: every visible camera of the parent in turn, its four
    neighbour cells (left, up, right, down), each refined candidate inserted before the next cell is looked at.
    `refine(center, parent)` stands for `Patch(center, parent)` + refine() + removeInvisibleCamera() and returns a Patch."""
    cs = mvs.cfg.cellSize
    for i, ci in enumerate(p.cam_idx):
        cam, cmap = mvs.cameras[ci], mvs.cell_maps[ci]
        cx, cy = int(p.img_point[i][0] / cs), int(p.img_point[i][1] / cs)
        for nx, ny in ((cx - 1, cy), (cx, cy - 1), (cx + 1, cy), (cx, cy + 1)):
            if not cmap.in_map(nx, ny):
                continue
            if mvs.skip_neighbor_cell(cmap.map[ny][nx], p):
                continue
            mvs.insert_patch(refine(mvs.expansion_patch_center(cam, p, nx, ny), p))


def expansion_patches(mvs, refine, reference_loop_exit=False):
    """MVS::expansionPatches, TMVS/mvs/mvs.cpp:233-275. The reference tests `!queue.empty()` AFTER popping the next id
    (:241-243, :271), so the patch popped last is never expanded; reference_loop_exit=True keeps that, False expands it
    too (what the host driver does)."""
    mvs.set_cell_maps()
    mvs.init_priority_queue()
    mvs.set_neighbor_radius()
    pid = mvs.pop()
    while (len(mvs.queue) > 0) if reference_loop_exit else (pid >= 0):
        p = mvs.patches.get(pid)
        if p is None:               # cannot happen: pops return live ids (the reference would spin here, :246)
            break
        p.expanded = True
        if not mvs.runtime_filtering(p):            # :255-260
            mvs.delete_patch(pid)
            pid = mvs.pop()
            continue
        expand_neighbor_cell(mvs, p, refine)
        pid = mvs.pop()
    mvs.set_neighbor_radius()
