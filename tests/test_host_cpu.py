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
