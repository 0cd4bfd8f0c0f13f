"""`tmvs -r` end to end on 1, 2, 4, 8 GPUs of one box: one synthetic NVM scene with many views and a fine cell grid
(>= 16 views, ~10^6 accepted patches), the C++ driver run with --gpus N for each N. Prints, per N, the driver's phase timers (host
pop / generate / commit, GPU seconds, context creation) so the scaling limiter is named by measurement.
usage: python tools/tmvs_multi_gpu.py [views width height cell round gpus_csv [extra driver flags, e.g. --no-pipeline]]"""
import os
import re
import subprocess
import sys
import tempfile
import time
import json

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
from pmvs_b200 import abi, mvsio, scene  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
cell = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rnd = sys.argv[5] if len(sys.argv) > 5 else "8192"
gpus = [int(g) for g in (sys.argv[6] if len(sys.argv) > 6 else "1,2,4,8").split(",")]
extra = sys.argv[7:]
cfg = abi.readme_config()
cfg.maxLOD = 2
cfg.cellSize = cell
t0 = time.time()
sc = scene.SynthScene(cfg, nviews=views, width=w, height=h, seed=1234, arc_deg=30.0)
print("scene: %d views %dx%d, cellSize %d (%.1f s to synthesise) driver flags %s" % (views, w, h, cell, time.time() - t0, extra), flush=True)
rows = []
with tempfile.TemporaryDirectory() as d:
    path = mvsio.write_nvm_scene(d, sc, n_seeds=256)
    mvsio.write_config(os.path.join(d, "config.txt"), cfg)
    for g in gpus:
        t0 = time.time()
        r = subprocess.run([os.path.join(ROOT, "pais-mvs_b200", "bin", "tmvs"), "-r", path, "--config", os.path.join(d, "config.txt"),
                            "--out-dir", d, "--round", rnd, "--gpus", str(g)] + extra, cwd=d, capture_output=True, text=True)
        wall = time.time() - t0
        out = r.stdout
        m = re.search(r"expansion host seconds: pop ([\d.]+) generate ([\d.]+) commit ([\d.]+) auto_save ([\d.]+); gpu calls (\d+)", out)
        p = re.search(r"patches: (\d+) refined: (\d+) gpu_seconds: ([\d.]+)", out)
        ph = re.search(r"phase seconds: seeds ([\d.]+) \(context ([\d.]+)\) expansion ([\d.]+) output ([\d.]+)", out)
        if not (m and p and ph):
            print("gpus", g, "FAILED", out[-500:], r.stderr[-500:])
            continue
        row = dict(gpus=g, wall_s=wall, patches=int(p.group(1)), refined=int(p.group(2)), gpu_s=float(p.group(3)), host_pop_s=float(m.group(1)),
                   host_generate_s=float(m.group(2)), host_commit_s=float(m.group(3)), gpu_calls=int(m.group(5)), context_s=float(ph.group(2)),
                   seeds_s=float(ph.group(1)), expansion_s=float(ph.group(3)), output_s=float(ph.group(4)))
        row["patches_per_s_expansion"] = row["patches"] / row["expansion_s"]
        rows.append(row)
        print(json.dumps(row), flush=True)
if rows:
    b = rows[0]
    print("\n| GPUs | patches | expansion s | GPU s | host pop+generate+commit s | context s | patches/s (expansion) | speed-up |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        print("| %d | %d | %.2f | %.2f | %.2f | %.2f | %.0f | %.2fx |" % (r["gpus"], r["patches"], r["expansion_s"], r["gpu_s"],
              r["host_pop_s"] + r["host_generate_s"] + r["host_commit_s"], r["context_s"], r["patches_per_s_expansion"],
              b["expansion_s"] / r["expansion_s"]))
