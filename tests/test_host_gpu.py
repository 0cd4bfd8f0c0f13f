"""GPU test of the reconstruction command: `tmvs -r scene.nvm` (reference `TMVS.exe -r`, TMVS/TMVS.cpp:76-122) on a
synthetic NVM scene — seed refinement, round-based expansion, MVS/PLY/PSR outputs."""
import os
import subprocess

import numpy as np
import pytest

from pmvs_b200 import abi, mvsio, scene

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
TMVS = os.path.join(ROOT, "pais-mvs_b200", "bin", "tmvs")


def run_tmvs(d, path, extra=()):
    r = subprocess.run([TMVS, "-r", path, "--config", os.path.join(d, "config.txt"), "--out-dir", d, "-V"] + list(extra),
                       cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_reconstruct_plane(tmp_path):
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = 7, 15, 7 / 3.0, 2
    cfg.cellSize = 8
    sc = scene.SynthScene(cfg, nviews=5, width=320, height=240, seed=1234, tex_size=1024)
    d = str(tmp_path)
    path = mvsio.write_nvm_scene(d, sc, n_seeds=16)
    mvsio.write_config(os.path.join(d, "config.txt"), cfg)
    out = run_tmvs(d, path, ["--round", "64"])
    assert "time1" in out
    _, cams, init = mvsio.read_mvs(os.path.join(d, "init.mvs"))
    _, _, seeds = mvsio.read_mvs(os.path.join(d, "seed.mvs"))
    c, _, exp = mvsio.read_mvs(os.path.join(d, "exp.mvs"))
    assert len(init) == 16 and 8 <= len(seeds) <= 16
    assert len(exp) > 5 * len(seeds), (len(seeds), len(exp))           # the surface was grown from the seeds
    z = np.array([p["center"][2] for p in exp])
    th = np.array([p["normalS"][0] for p in exp])
    assert np.percentile(np.abs(z - sc.plane_z), 95) < 5e-3 and np.percentile(th, 95) < 0.05
    assert all(len(p["camIdx"]) >= cfg.minCamNum and 0 < p["fitness"] <= cfg.maxFitness and p["correlation"] >= cfg.minCorrelation for p in exp)
    assert c.neighborRadius > 0
    ply = mvsio.read_ply(os.path.join(d, "exp.ply"))
    psr = mvsio.read_psr(os.path.join(d, "exp.psr"))
    assert ply.shape == (len(exp), 9) and psr.shape == (len(exp), 6)
    ctr = np.array([p["center"] for p in exp])
    assert np.allclose(psr[:, :3], ctr.astype(np.float32)) and np.allclose(ply[:, :3], ctr, rtol=1e-4, atol=1e-4)
    # determinism: same run, same bytes (counter-based RNG keyed by patch id)
    d2 = os.path.join(d, "again")
    os.makedirs(d2)
    run_tmvs(d, path, ["--round", "64", "--out-dir", d2])
    assert open(os.path.join(d, "exp.mvs"), "rb").read() == open(os.path.join(d2, "exp.mvs"), "rb").read()
    # --slot-passes (one GPU pass per camera slot, the reference's visiting order) grows the same surface as the default
    # merged pass: same coverage to a few per cent, same accuracy
    d3 = os.path.join(d, "slots")
    os.makedirs(d3)
    out_s = run_tmvs(d, path, ["--round", "64", "--slot-passes", "--out-dir", d3])
    _, _, exp_s = mvsio.read_mvs(os.path.join(d3, "exp.mvs"))
    assert abs(len(exp_s) - len(exp)) <= 0.05 * len(exp), (len(exp_s), len(exp))
    zs = np.array([p["center"][2] for p in exp_s])
    assert np.percentile(np.abs(zs - sc.plane_z), 95) < 5e-3
    calls = lambda txt: int(txt.split("gpu calls ")[1].split()[0])
    assert calls(out) < calls(out_s)                                     # fewer, larger calls
    # warm start from the MVS file (TMVS.cpp:87-89): loads cameras + patches and re-refines them as seeds
    out3 = run_tmvs(d, os.path.join(d, "seed.mvs"), ["--no-expand", "--out-dir", d2])
    assert "seeds kept" in out3


def test_reconstruct_two_gpus_identical(tmp_path):
    """--gpus 2 shards every batch over two devices; the output is byte-identical to the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD, cfg.cellSize = 7, 15, 7 / 3.0, 2, 8
    sc = scene.SynthScene(cfg, nviews=5, width=320, height=240, seed=1234, tex_size=1024)
    d = str(tmp_path)
    path = mvsio.write_nvm_scene(d, sc, n_seeds=16)
    mvsio.write_config(os.path.join(d, "config.txt"), cfg)
    d1, d2 = os.path.join(d, "g1"), os.path.join(d, "g2")
    os.makedirs(d1)
    os.makedirs(d2)
    run_tmvs(d, path, ["--out-dir", d1, "--gpus", "1"])
    run_tmvs(d, path, ["--out-dir", d2, "--gpus", "2"])
    assert open(os.path.join(d1, "exp.mvs"), "rb").read() == open(os.path.join(d2, "exp.mvs"), "rb").read()
