"""End-to-end timing of `tmvs -r` on a synthetic NVM scene: how much of the wall clock is GPU refinement vs the host's
serial commit (the expected scaling limit). usage: python tools/tmvs_scale.py [width height round cell [extra tmvs -r switches ...]]"""
import contextlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
from pmvs_b200 import abi, mvsio, scene  # noqa: E402

w = int(sys.argv[1]) if len(sys.argv) > 1 else 800
h = int(sys.argv[2]) if len(sys.argv) > 2 else 600
rnd = sys.argv[3] if len(sys.argv) > 3 else "512"
cell = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfg = abi.readme_config()
cfg.maxLOD = 2
cfg.cellSize = cell
sc = scene.SynthScene(cfg, nviews=5, width=w, height=h, seed=1234)
keep = os.environ.get("TMVS_SCALE_DIR")            # keep the dataset and the outputs there (profiling tmvs directly afterwards)
if keep:
    os.makedirs(keep, exist_ok=True)
with (contextlib.nullcontext(keep) if keep else tempfile.TemporaryDirectory()) as d:
    path = mvsio.write_nvm_scene(d, sc, n_seeds=64)
    mvsio.write_config(os.path.join(d, "config.txt"), cfg)
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "pais-mvs_b200", "bin", "tmvs"), "-r", path, "--config", os.path.join(d, "config.txt"),
                        "--out-dir", d, "--round", rnd] + sys.argv[5:], cwd=d, capture_output=True, text=True)
    dt = time.time() - t0
    print(r.stdout[-700:], r.stderr[-400:])
    print("wall %.2f s (incl. image load + pyramid build)" % dt)
    # quality of the reconstruction: the scene is the plane z = plane_z seen by every camera
    import numpy as np
    _, _, exp = mvsio.read_mvs(os.path.join(d, "exp.mvs"))
    z = np.array([p["center"][2] for p in exp])
    th = np.array([p["normalS"][0] for p in exp])
    cells = {(int(u / cell), int(v / cell)) for u, v in (sc.cams[0].project(np.array(p["center"])) for p in exp)}
    print("quality: %d patches, |z - plane| p50 %.2e p95 %.2e, tilt p95 %.3f rad, %d cells of camera 0 covered (of %d)"
          % (len(exp), np.percentile(np.abs(z - sc.plane_z), 50), np.percentile(np.abs(z - sc.plane_z), 95), np.percentile(th, 95),
             len(cells), (w // cell) * (h // cell)))
    # the -f post-process on the reconstruction just written
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "pais-mvs_b200", "bin", "tmvs"), "-f", os.path.join(d, "exp.mvs"), "--config",
                        os.path.join(d, "config.txt"), "--out-dir", d], cwd=d, capture_output=True, text=True)
    print(r.stdout[-300:], r.stderr[-300:])
    print("tmvs -f wall %.2f s" % (time.time() - t0))
