"""CPU restatement of the reference's `-f` post-process (TMVS/TMVS.cpp:124-170, TMVS/mvs/mvs.cpp:279-525) on plain
Python containers. TEST INFRASTRUCTURE ONLY (imported by tests/): the product path is pais-mvs_b200/host (C++) plus
pmvs_neighbor_counts on the GPU.

State: patches = {id: dict(center[3], normal[3], camIdx[], imgPoint[[x,y]...], correlation)}, cameras = [dict(center[3],
cols, rows)], cell maps as the reference's CellMap (TMVS/mvs/cellmap.cpp): per camera a width x height grid of id lists.
"""
import math

import numpy as np


class CellMaps:
    def __init__(self, cameras, cell_size, patches):          # MVS::setCellMaps, mvs.cpp:116-133 (+ initCellMaps :74-88)
        self.cs = cell_size
        self.dims = [(int(math.ceil(c["cols"] / cell_size)), int(math.ceil(c["rows"] / cell_size))) for c in cameras]   # cellmap.cpp:7-8
        self.cells = [dict() for _ in cameras]
        for pid in sorted(patches):
            self.insert(patches[pid])

    def key(self, p, i):
        return int(p["imgPoint"][i][0] / self.cs), int(p["imgPoint"][i][1] / self.cs)

    def in_map(self, cam, x, y):
        w, h = self.dims[cam]
        return not (x < 0 or y < 0 or x >= w or y >= h)

    def insert(self, p):
        for i, cam in enumerate(p["camIdx"]):
            x, y = self.key(p, i)
            if self.in_map(cam, x, y):                         # CellMap::insert ignores out-of-map cells
                self.cells[cam].setdefault((x, y), []).append(p["id"])

    def drop(self, p):
        for i, cam in enumerate(p["camIdx"]):
            x, y = self.key(p, i)
            cell = self.cells[cam].get((x, y))
            if cell and p["id"] in cell:
                cell.remove(p["id"])

    def cell(self, cam, x, y):
        return self.cells[cam].get((x, y), [])


def neighbor_radius(patches, scalar):                          # mvs.cpp:147-152, :974-997
    c = np.array([patches[k]["center"] for k in patches])
    vol = c.max(axis=0) - c.min(axis=0)
    return abs(vol[0] * vol[1] * vol[2]) ** (1.0 / 3.0) * scalar


def is_neighbor(a, b, radius):                                 # Patch::isNeighbor, patch.cpp:6-23
    d = [a["center"][k] - b["center"][k] for k in range(3)]
    dist = abs(sum(d[k] * a["normal"][k] for k in range(3))) + abs(sum(d[k] * b["normal"][k] for k in range(3)))
    return dist <= radius


def _delete(patches, maps, pid, deleted):                      # MVS::deletePatch, mvs.cpp:607-634
    p = patches.pop(pid, None)
    if p is not None:
        maps.drop(p)
        deleted.append(pid)


def cell_filtering(patches, maps, deleted):                    # mvs.cpp:279-325
    for cam in range(len(maps.dims)):
        w, h = maps.dims[cam]
        for x in range(w):
            for y in range(h):
                cell = maps.cell(cam, x, y)
                remove = []
                for j, pj in enumerate(cell):
                    corr_sum = 0.0
                    for k, pk in enumerate(cell):
                        if j != k and pk in patches:
                            corr_sum += patches[pk]["correlation"]
                    if pj in patches and patches[pj]["correlation"] * len(patches[pj]["camIdx"]) < corr_sum:
                        remove.append(pj)
                for pid in remove:
                    _delete(patches, maps, pid, deleted)


def visibility_filtering(patches, maps, cameras, min_cam_num, deleted):   # mvs.cpp:399-446
    for pid in sorted(patches):
        p = patches.get(pid)
        if p is None:
            continue
        visible = len(p["camIdx"])
        for i, cam in enumerate(p["camIdx"]):
            cc = cameras[cam]["center"]
            depth = math.sqrt(sum((p["center"][k] - cc[k]) ** 2 for k in range(3)))
            x, y = maps.key(p, i)
            if not maps.in_map(cam, x, y):
                continue
            for q in maps.cell(cam, x, y):
                if q == pid or q not in patches:
                    continue
                nd = math.sqrt(sum((patches[q]["center"][k] - cc[k]) ** 2 for k in range(3)))
                if depth > nd:
                    visible -= 1
                    break
        if visible < min_cam_num:
            _delete(patches, maps, pid, deleted)


def neighbor_cell_filtering(patches, maps, radius, ratio, deleted):       # mvs.cpp:327-397
    for cam in range(len(maps.dims)):
        w, h = maps.dims[cam]
        for x in range(w):
            for y in range(h):
                cell = maps.cell(cam, x, y)
                if not cell:
                    continue
                nx = [x, x - 1, x + 1, x - 1, x + 1, x + 1, x, x - 1, x]
                ny = [y, y - 1, y - 1, y + 1, y + 1, y, y + 1, y, y - 1]
                remove = []
                for pid in cell:
                    if pid not in patches:
                        continue
                    total = num = 0
                    for q in range(9):
                        if not maps.in_map(cam, nx[q], ny[q]):
                            continue
                        ncell = maps.cell(cam, nx[q], ny[q])
                        total += len(ncell)
                        num += sum(1 for r in ncell if r in patches and is_neighbor(patches[pid], patches[r], radius))
                    if total != 0 and num / total < ratio:      # 0/0 is NaN in the reference: never removed
                        remove.append(pid)
                for pid in remove:
                    _delete(patches, maps, pid, deleted)


def neighbor_patch_filtering(patches, radius, ratio, counts_fn, deleted, maps=None):   # mvs.cpp:448-525
    ids = sorted(patches)
    centers = np.array([patches[i]["center"] for i in ids], dtype=np.float64)
    counts = counts_fn(centers, radius)
    avg = 0.0
    for c in counts:
        avg += float(c)
    avg /= float(len(ids))
    for i, pid in enumerate(ids):
        if float(counts[i]) < avg * ratio:
            p = patches.pop(pid)
            if maps is not None:
                maps.drop(p)
            deleted.append(pid)
    return avg
