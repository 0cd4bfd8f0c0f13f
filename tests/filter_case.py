"""A small crafted MVS file for the `-f` tests: clusters of patches near a plane, several per cell, with depth offsets
(visibility filter), outliers (neighbour filters) and varied correlation (cell filter)."""
import math
import os

import numpy as np

from pmvs_b200 import mvsio


def make_case(directory, cfg, sc, n_base=260, seed=11):
    rng = np.random.RandomState(seed)
    os.makedirs(directory, exist_ok=True)
    for c in sc.cams:
        mvsio.write_pgm(os.path.join(directory, c.name + ".pgm"), c.levels[0][0])
    ext = 0.30 * sc.distance * min(sc.width, sc.height) / sc.focal
    patches = []
    for _ in range(n_base):
        base = np.array([(2 * rng.rand() - 1) * ext, (2 * rng.rand() - 1) * ext, sc.plane_z])
        for _k in range(rng.randint(1, 5)):
            X = base + np.array([rng.randn() * 0.004, rng.randn() * 0.004, rng.randn() * 0.03])
            if rng.rand() < 0.06:
                X[2] += rng.choice([-1.0, 1.0]) * (0.5 + rng.rand())          # outliers far off the surface
            theta, phi = 0.25 * rng.rand(), 2 * math.pi * rng.rand() - math.pi
            cams = sorted(rng.choice(len(sc.cams), size=rng.randint(3, len(sc.cams) + 1), replace=False).tolist())
            patches.append(dict(center=X.tolist(), normalS=[theta, phi], camIdx=cams, fitness=float(rng.rand() * 5),
                                correlation=float(0.5 + 0.5 * rng.rand())))
    path = os.path.join(directory, "in.mvs")
    mvsio.write_mvs(path, cfg, sc.cams, patches)
    mvsio.write_config(os.path.join(directory, "config.txt"), cfg)
    return path, patches


def load_state(path, sc):
    """MVS file -> the containers oracle/orc_filters.py works on (ids = file order, like the loader's nextId)."""
    cfg, cams, plist = mvsio.read_mvs(path)
    patches = {}
    for i, p in enumerate(plist):
        th, ph = p["normalS"]
        n = [math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)]
        img = [list(sc.cams[c].project(p["center"])) for c in p["camIdx"]]
        patches[i] = dict(id=i, center=list(p["center"]), normal=n, camIdx=list(p["camIdx"]), imgPoint=img, correlation=p["correlation"])
    cameras = [dict(center=list(c.center), cols=sc.width, rows=sc.height) for c in sc.cams]
    return cfg, cameras, patches


def centers_of(path):
    return np.array([p["center"] for p in mvsio.read_mvs(path)[2]], dtype=np.float64).reshape(-1, 3)
