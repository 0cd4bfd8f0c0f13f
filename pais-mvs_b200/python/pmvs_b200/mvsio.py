"""On-disk formats around the reconstruction command, for tests and tools: write a synthetic scene as NVM_V3 + PGM
images (TMVS/io/fileloader.cpp:251-325, README.md:59-86), parse MVS_V3 (TMVS/io/filewriter.cpp:26-102), PLY and PSR."""
import ctypes as C
import os
import struct

import numpy as np

from . import abi


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())


def write_nvm_scene(directory, sc, n_seeds=64, seed=7, jitter_px=0.3, nvm2=False):
    """Images as PGM plus an NVM file whose points are plane points seen by every camera (measurements relative to the
    image centre, with a little pixel noise so re-triangulation has work to do)."""
    os.makedirs(directory, exist_ok=True)
    rng = np.random.RandomState(seed)
    for c in sc.cams:
        write_pgm(os.path.join(directory, c.name + ".pgm"), c.levels[0][0])
    ext = 0.30 * sc.distance * min(sc.width, sc.height) / sc.focal
    lines = ["NVM_V3", "", str(len(sc.cams))]
    for c in sc.cams:
        q, ctr = c.quaternion, c.center
        if nvm2:
            lines.append("%s.pgm %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g" %
                         (c.name, c.focal[0], c.focal[1], c.principal[0], c.principal[1], q[0], q[1], q[2], q[3], ctr[0], ctr[1], ctr[2]))
        else:
            lines.append("%s.pgm %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g 0 0" %
                         (c.name, c.focal[0], q[0], q[1], q[2], q[3], ctr[0], ctr[1], ctr[2]))
    lines += ["", str(n_seeds)]
    for k in range(n_seeds):
        X = np.array([(2 * rng.rand() - 1) * ext, (2 * rng.rand() - 1) * ext, sc.plane_z])
        meas = []
        for i, c in enumerate(sc.cams):
            u = c.project(X) + jitter_px * (rng.rand(2) - 0.5)
            meas.append("%d %d %.6f %.6f" % (i, k, u[0] - sc.width // 2, u[1] - sc.height // 2))
        lines.append("%.17g %.17g %.17g 128 128 128 %d %s" % (X[0], X[1], X[2], len(sc.cams), " ".join(meas)))
    lines += ["", "0", ""]
    path = os.path.join(directory, "scene.nvm2" if nvm2 else "scene.nvm")
    open(path, "w").write("\n".join(lines))
    return path


def write_config(path, cfg, keys=None):
    """config.txt in the reference's key-value form (README.md:110-207)."""
    keys = keys or ["patchRadius", "reduceNormalRange", "adaptiveDistanceEnable", "adaptiveDifferenceEnable", "adaptiveGradientEnable",
                    "distWeighting", "diffWeighting", "visibleCorrelation", "depthRangeScalar", "particleNum", "maxIteration", "cellSize",
                    "maxCellPatchNum", "expansionStrategy", "textureVariation", "minLOD", "maxLOD", "lodRatio", "minCamNum",
                    "minCorrelation", "minRegionRatio", "maxFitness", "neighborRadiusScalar"]
    with open(path, "w") as f:
        f.write("# written by pmvs_b200.mvsio.write_config\n")
        for k in keys:
            f.write("%s %r\n" % (k, getattr(cfg, k)))


def read_mvs(path):
    """-> (PmvsConfig, cameras [dict], patches [dict])"""
    data = open(path, "rb").read()
    assert data.startswith(b"MVS_V3\n"), data[:8]
    off = 7
    cfg = abi.PmvsConfig.from_buffer_copy(data[off:off + 160])
    off += 160
    nl = data.index(b"\n", off)
    head = data[off:nl].split()
    assert head[0] == b"CAMERAS"
    off = nl + 1
    cams = []
    for _ in range(int(head[1])):
        (ln,) = struct.unpack_from("<i", data, off)
        off += 4
        name = data[off:off + ln].decode()
        off += ln
        v = struct.unpack_from("<12d", data, off)
        off += 96
        cams.append(dict(name=name, center=v[0:3], focal=v[3:5], principal=v[5:7], quaternion=v[7:11], radial=v[11]))
    nl = data.index(b"\n", off)
    head = data[off:nl].split()
    assert head[0] == b"PATCHES"
    off = nl + 1
    patches = []
    for _ in range(int(head[1])):
        v = struct.unpack_from("<5d", data, off)
        off += 40
        (n,) = struct.unpack_from("<i", data, off)
        off += 4
        idx = struct.unpack_from("<%di" % n, data, off)
        off += 4 * n
        fit, corr = struct.unpack_from("<2d", data, off)
        off += 16
        patches.append(dict(center=v[0:3], normalS=v[3:5], camIdx=list(idx), fitness=fit, correlation=corr))
    assert off == len(data)
    return cfg, cams, patches


def write_mvs(path, cfg, cams, patches):
    """MVS_V3 writer (TMVS/io/filewriter.cpp:26-102): cams = scene.Camera objects (image name `name`.pgm), patches = dicts
    with center, normalS, camIdx, fitness, correlation."""
    with open(path, "wb") as f:
        f.write(b"MVS_V3\n")
        f.write(bytes(cfg))
        f.write(b"CAMERAS %d\n" % len(cams))
        for c in cams:
            name = (c.name + ".pgm").encode()
            f.write(struct.pack("<i", len(name)) + name)
            f.write(struct.pack("<12d", *c.center, *c.focal, *c.principal, *c.quaternion, 0.0))
        f.write(b"PATCHES %d\n" % len(patches))
        for p in patches:
            f.write(struct.pack("<5d", *p["center"], *p["normalS"]))
            f.write(struct.pack("<i", len(p["camIdx"])))
            f.write(struct.pack("<%di" % len(p["camIdx"]), *p["camIdx"]))
            f.write(struct.pack("<2d", p["fitness"], p["correlation"]))


def read_ply(path):
    lines = open(path).read().split("\n")
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    body = lines[lines.index("end_header") + 1:]
    return np.array([[float(t) for t in l.split()] for l in body[:n]])


def read_psr(path):
    return np.fromfile(path, dtype="<f4").reshape(-1, 6)
