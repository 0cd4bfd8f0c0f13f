"""CPU tests of the host-side driver (pais-mvs_b200/host, C++): formats, config, geometry helpers — no GPU needed
(`tmvs --convert` loads and writes without creating a device context)."""
import math
import os
import subprocess

import numpy as np
import pytest

from pmvs_b200 import abi, mvsio, scene

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
TMVS = os.path.join(ROOT, "pais-mvs_b200", "bin", "tmvs")


@pytest.fixture(scope="module")
def tmvs_bin():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "pais-mvs_b200", "host")])
    return TMVS


@pytest.fixture(scope="module")
def dataset(tmp_path_factory, small_scene):
    cfg, sc = small_scene
    d = str(tmp_path_factory.mktemp("nvm"))
    path = mvsio.write_nvm_scene(d, sc, n_seeds=24)
    mvsio.write_config(os.path.join(d, "config.txt"), cfg)
    return d, path, cfg, sc


def test_nvm_load_and_mvs_roundtrip(tmvs_bin, dataset):
    d, path, cfg, sc = dataset
    out1, out2 = os.path.join(d, "a.mvs"), os.path.join(d, "b.mvs")
    subprocess.check_call([tmvs_bin, "--convert", path, out1, "--config", os.path.join(d, "config.txt")], cwd=d)
    c1, cams, patches = mvsio.read_mvs(out1)
    assert len(cams) == len(sc.cams) and len(patches) == 24
    assert bytes(c1)[:112] == bytes(cfg)[:112]                       # config.txt reproduced the README config (neighborRadius is derived)
    assert c1.patchSize == 2 * cfg.patchRadius + 1
    for got, cam in zip(cams, sc.cams):
        assert got["name"] == cam.name + ".pgm"
        assert np.allclose(got["center"], cam.center, rtol=0, atol=1e-15) and np.allclose(got["quaternion"], cam.quaternion, atol=1e-15)
        assert got["principal"] == (sc.width >> 1, sc.height >> 1)  # camera.cpp:101-106
    for p in patches:                                                # re-triangulated onto the plane, normal towards the cameras
        assert abs(p["center"][2] - sc.plane_z) < 0.02 and p["camIdx"] == list(range(len(sc.cams)))
        assert p["normalS"][0] < 0.35
    # MVS_V3 -> load -> MVS_V3 is the identity on the bytes
    subprocess.check_call([tmvs_bin, "--convert", out1, out2, "--config", os.path.join(d, "config.txt")], cwd=d)
    assert open(out1, "rb").read() == open(out2, "rb").read()


def test_nvm2_and_defaults(tmvs_bin, tmp_path, small_scene):
    cfg, sc = small_scene
    d = str(tmp_path)
    path = mvsio.write_nvm_scene(d, sc, n_seeds=5, nvm2=True)
    out = os.path.join(d, "a.mvs")
    subprocess.check_call([tmvs_bin, "--convert", path, out, "--config", os.path.join(d, "missing.txt")], cwd=d)
    c, cams, patches = mvsio.read_mvs(out)
    want = abi.default_config()                                      # TMVS.cpp:26-52
    assert (c.cellSize, c.patchRadius, c.patchSize, c.minCamNum, c.particleNum, c.maxIteration) == (4, 15, 31, 3, 5, 10)
    assert (c.distWeighting, c.diffWeighting, c.minRegionRatio, c.depthRangeScalar) == (5.0, 16384.0, 0.55, 1.0)
    assert bytes(c)[:112] == bytes(want)[:112]
    assert len(patches) == 5 and cams[0]["focal"] == (sc.focal, sc.focal)


def test_usage_and_out_of_scope_commands(tmvs_bin):
    assert subprocess.run([tmvs_bin], capture_output=True).returncode == 2
    r = subprocess.run([tmvs_bin, "-v", "x.mvs"], capture_output=True, text=True)        # viewer: out of scope
    assert r.returncode == 2 and "scope" in r.stderr
    r = subprocess.run([tmvs_bin, "-f", "missing.mvs"], capture_output=True, text=True)  # filtering: in scope, file missing
    assert r.returncode == 1 and "load failed" in r.stderr


def test_mvs_v2_ply_psr_and_loader_ctor(tmvs_bin, dataset, tmp_path):
    """MVS_V2 input (no config block, fileloader.cpp:431-434), and the PLY / PSR writers byte for byte against the format of
    filewriter.cpp:104-171: ostream default formatting (%g) of centre and normal, colour written r g b from the b,g,r
    pixel under the reference camera's projection — the loader ctor's setReferenceCameraIndex + setImagePoint
    (patch.cpp:45-59, :415-445, :627-653) — and six float32 per patch."""
    import orc_host as oh
    d, path, cfg, sc = dataset
    rng = np.random.RandomState(11)
    patches = []
    for k in range(40):
        center = [0.8 * (2 * rng.rand() - 1), 0.6 * (2 * rng.rand() - 1), sc.plane_z + 0.01 * rng.randn()]
        cams = sorted(rng.choice(len(sc.cams), size=rng.randint(3, len(sc.cams) + 1), replace=False).tolist())
        patches.append(dict(center=center, normalS=[0.35 * rng.rand(), 2 * math.pi * rng.rand() - math.pi], camIdx=cams,
                            fitness=float(rng.rand()), correlation=float(rng.rand())))
    v3 = os.path.join(d, "p.mvs")
    mvsio.write_mvs(v3, cfg, sc.cams, patches)
    # MVS_V2 = the same records without the 160-byte config block
    raw = open(v3, "rb").read()
    v2 = str(tmp_path / "p_v2.mvs")
    open(v2, "wb").write(b"MVS_V2\n" + raw[7 + 160:])
    conf = os.path.join(d, "config.txt")
    out_v2, ply, psr = str(tmp_path / "from_v2.mvs"), str(tmp_path / "p.ply"), str(tmp_path / "p.psr")
    for src, dst in ((v2, out_v2), (v3, ply), (v3, psr)):
        subprocess.check_call([tmvs_bin, "--convert", src, dst, "--config", conf, "--image-dir", d], cwd=d)
    got = open(out_v2, "rb").read()                                      # V2 in, config.txt applied, V3 out: identical records
    assert got[:7 + 112] == raw[:7 + 112] and got[7 + 120:] == raw[7 + 120:]      # (neighborRadius, bytes 112..119, is derived at run time)

    ocams = [oh.Camera(c.focal[0], list(c.quaternion), list(c.center), sc.width, sc.height, c.levels[0][0]) for c in sc.cams]
    want_lines, want_psr = [], []
    for p in patches:
        th, ph = p["normalS"]
        n = [math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)]           # utility.h:25-29
        ref, best = -1, -oh.DBL_MAX
        for ci in p["camIdx"]:
            on = ocams[ci].optical_normal
            corr = oh.dot3(n, [-on[0], -on[1], -on[2]])
            if corr > best:
                best, ref = corr, ci
        colour = (0, 0, 0)
        (u, v), inside = ocams[ref].project(p["center"], 0, cfg.lodRatio)
        if inside:
            g = int(ocams[ref].grey[min(oh.cv_round(v), sc.height - 1)][min(oh.cv_round(u), sc.width - 1)])
            colour = (g, g, g)                                                                   # PGM: r = g = b
        want_lines.append("%g %g %g %g %g %g %d %d %d" % (*p["center"], *n, *colour))
        want_psr.append(np.array([*p["center"], *n], dtype=np.float32))
    text = open(ply).read().split("\n")
    assert text[:13] == ["ply", "format ascii 1.0", "element vertex 40", "property float x", "property float y", "property float z",
                         "property float nx", "property float ny", "property float nz", "property uchar diffuse_red",
                         "property uchar diffuse_green", "property uchar diffuse_blue", "end_header"]
    assert text[13:53] == want_lines and text[53:] == [""]
    assert len({l.split()[-1] for l in want_lines}) > 10                  # real image colours, not a constant
    assert open(psr, "rb").read() == np.stack(want_psr).astype("<f4").tobytes()


def test_config_txt_parser(tmvs_bin, dataset, tmp_path):
    """config.txt as fileloader.cpp:474-565 reads it: '#' comment lines, blank lines, blanks or tabs between key and value,
    CRLF line ends, unknown keys ignored, later lines win, patchSize follows patchRadius; plus the documented
    gradientWeighting key (README.md:140-142) the reference parser forgot."""
    d, path, cfg, sc = dataset
    conf = str(tmp_path / "c.txt")
    open(conf, "wb").write(b"# comment\r\n\r\npatchRadius\t9\r\nparticleNum 7\nparticleNum   11\n#maxIteration 99\nmaxIteration 21 trailing words\n"
                           b"unknownKey 5\nadaptiveGradientEnable 1\ngradientWeighting 0.125\nadaptiveDistanceEnable 0\nlodRatio 0.75\n"
                           b"expansionStrategy 2\nminRegionRatio\t0.4\nkeyWithoutValue\nneighborRadiusScalar 0.02")
    out = str(tmp_path / "c.mvs")
    subprocess.check_call([tmvs_bin, "--convert", path, out, "--config", conf], cwd=d)
    c, _, _ = mvsio.read_mvs(out)
    want = abi.default_config()                                      # compiled defaults (TMVS.cpp:26-52) under the file's keys
    assert (c.patchRadius, c.patchSize, c.particleNum, c.maxIteration) == (9, 19, 11, 21)
    assert (c.adaptiveGradientEnable, c.adaptiveDistanceEnable, c.adaptiveDifferenceEnable) == (1, 0, want.adaptiveDifferenceEnable)
    assert (c.gradientWeighting, c.lodRatio, c.expansionStrategy, c.minRegionRatio, c.neighborRadiusScalar) == (0.125, 0.75, 2, 0.4, 0.02)
    assert (c.cellSize, c.minCamNum, c.maxCellPatchNum, c.distWeighting) == (want.cellSize, want.minCamNum, want.maxCellPatchNum, want.distWeighting)


def test_nvm_seeds_match_the_oracle(tmvs_bin, dataset):
    """NVM -> seeds as the reference builds them: camera line (fileloader.cpp:15-60), point line with measurements relative
    to the INTEGER image centre (:112-165, `cols / 2`), seed ctor's estimated normal, then MVS::loadNVM's reCentering
    (mvs.cpp:161-164, patch.cpp:67-112) — the host's result against oracle/orc_host.py's restatement, every seed."""
    import orc_host as oh
    d, path, cfg, sc = dataset
    out = os.path.join(d, "seeds_oracle.mvs")
    subprocess.check_call([tmvs_bin, "--convert", path, out, "--config", os.path.join(d, "config.txt")], cwd=d)
    _, _, got = mvsio.read_mvs(out)
    lines = [l for l in open(path).read().split("\n")]
    assert lines[0] == "NVM_V3"
    body = [l for l in lines[1:] if l.strip()]
    ncam = int(body[0])
    cams = []
    for l in body[1:1 + ncam]:
        t = l.split()
        cams.append(oh.Camera(float(t[1]), [float(v) for v in t[2:6]], [float(v) for v in t[6:9]], sc.width, sc.height))
    npt = int(body[1 + ncam])
    o = oh.MVS(cfg, cams)
    assert npt == len(got) == 24
    for k, l in enumerate(body[2 + ncam:2 + ncam + npt]):
        t = l.split()
        n = int(t[6])
        idx = [int(t[7 + 4 * i]) for i in range(n)]
        pts = [(float(t[9 + 4 * i]) + sc.width // 2, float(t[10 + 4 * i]) + sc.height // 2) for i in range(n)]
        p = oh.Patch(k, [float(v) for v in t[0:3]], [0.0, 0.0, 0.0], cam_idx=idx, img_point=pts)
        oh.estimated_normal(o, p)                 # seed ctor, patch.cpp:26-34
        oh.recentering(o, p)
        assert got[k]["camIdx"] == idx
        assert np.allclose(got[k]["center"], p.center, rtol=0, atol=1e-10)
        assert np.allclose(got[k]["normalS"], p.normalS, rtol=0, atol=1e-10)
