"""Independent NumPy f64 implementation of PAIS::getFitness (TMVS/mvs/patch.cpp:914-1047) used to cross-check the
C++ oracle restatement (a second reading of the same reference lines, vectorised over the window instead of looped).
Test infrastructure only."""
import math

import numpy as np

DBL_MAX = 1.7976931348623157e308


def homographies(cams, cfg, ref, cam_idx, lod, center, normal):
    """patch.cpp:290-330"""
    d = -float(np.dot(center, normal))
    s = cfg.lodRatio ** lod
    L = np.diag([s, s, 1.0])
    n = np.asarray(normal, dtype=np.float64).reshape(3, 1)

    def bracket(cam):
        return d * (L @ cam.KR) - (L @ cam.KT.reshape(3, 1)) @ n.T

    Mref = bracket(cams[ref])
    det = np.linalg.det(Mref)
    inv = np.linalg.inv(Mref) if det != 0 else np.zeros((3, 3))
    return [np.eye(3) if ci == ref else bracket(cams[ci]) @ inv for ci in cam_idx]


def dist_table(cfg):
    """mvs.cpp:97-114"""
    ps, r, sigma = cfg.patchSize, cfg.patchRadius, cfg.distWeighting
    x, y = np.meshgrid(np.arange(ps), np.arange(ps), indexing="ij")
    g = np.exp(-((x - r) ** 2.0 + (y - r) ** 2.0) / (2 * sigma * sigma)) / (2 * math.pi * sigma * sigma)
    return g / g.sum()


def fitness(cams, cfg, hyp):
    r = cfg.patchRadius
    ref = hyp.refCamIdx
    lod = hyp.LOD
    cam_idx = list(hyp.camIdx[:hyp.nCam])
    rc = cams[ref]
    th, ph = hyp.theta, hyp.phi
    normal = np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
    if float(normal @ rc.optical_normal) > 0:
        return DBL_MAX
    center = np.array(hyp.ray[:]) * hyp.depth + rc.center
    H = homographies(cams, cfg, ref, cam_idx, lod, center, normal)
    pt = rc.project(center, lod, cfg.lodRatio)
    ref_img, ref_edge = rc.levels[lod]
    rows, cols = ref_img.shape
    if np.isnan(pt).any() or pt[0] < 0 or pt[0] >= cols or pt[1] < 0 or pt[1] >= rows:
        return DBL_MAX
    if pt[0] - r < 2 or pt[0] + r >= cols - 3 or pt[1] - r < 2 or pt[1] + r >= rows - 3:
        return DBL_MAX
    offs = np.arange(-r, r + 1, dtype=np.float64)
    X, Y = np.meshgrid(pt[0] + offs, pt[1] + offs, indexing="ij")          # outer x, inner y
    rx = np.rint(X).astype(int)
    ry = np.rint(Y).astype(int)
    keep = ref_img[ry, rx] != 0
    cs = []
    for ci, Hi in zip(cam_idx, H):
        img = cams[ci].levels[lod][0]
        h, w_ = img.shape
        W = Hi[2, 0] * X + Hi[2, 1] * Y + Hi[2, 2]
        with np.errstate(all="ignore"):
            ix = (Hi[0, 0] * X + Hi[0, 1] * Y + Hi[0, 2]) / W
            iy = (Hi[1, 0] * X + Hi[1, 1] * Y + Hi[1, 2]) / W
        bad = ~((ix >= 2) & (ix < w_ - 3) & (iy >= 2) & (iy < h - 3)) | (W == 0)
        if (bad & keep).any():
            return DBL_MAX
        ixs = np.where(bad, 2.0, ix)
        iys = np.where(bad, 2.0, iy)
        px = ixs.astype(int)
        py = iys.astype(int)
        g = img.astype(np.float64)
        c = (g[py, px] * (px + 1 - ixs) * (py + 1 - iys) + g[py, px + 1] * (ixs - px) * (py + 1 - iys) +
             g[py + 1, px] * (px + 1 - ixs) * (iys - py) + g[py + 1, px + 1] * (ixs - px) * (iys - py))
        cs.append(c)
    cs = np.stack(cs)
    mean = cs.sum(0) / len(cam_idx)
    sad = np.abs(cs - mean).sum(0) / len(cam_idx)
    wgt = np.ones_like(sad)
    if cfg.adaptiveDistanceEnable:
        wgt = wgt * dist_table(cfg)
    if cfg.adaptiveDifferenceEnable:
        wgt = wgt * np.exp(-sad * sad / cfg.diffWeighting)
    if cfg.adaptiveGradientEnable:
        with np.errstate(all="ignore"):
            wgt = wgt * np.exp(-1.0 / (ref_edge[ry, rx] * cfg.gradientWeighting))
    wgt = np.where(keep, wgt, 0.0)
    sw = wgt.sum()
    with np.errstate(all="ignore"):
        return float((wgt * np.where(keep, sad, 0.0)).sum() / sw) if sw != 0 else float("nan")
