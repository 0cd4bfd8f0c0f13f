"""Histogram of the image footprint one swarm touches in each non-reference view (experiment build, -DPMVS_FOOTPRINT=1):
for every patch of a configs[1] batch, the bounding box of all windows its first PSO run evaluated (the four projected window
corners of every hypothesis that took the unchecked loop). Answers SURVEY.md section 7's staging question: would a
per-patch shared-memory tile of T x T pixels per view serve the swarm?  usage (GPU box):
    make -C pais-mvs_b200/csrc OUT=$PWD/pais-mvs_b200/lib/libpmvs_foot.so EXTRA=-DPMVS_FOOTPRINT=1
    PMVS_LIB=$PWD/pais-mvs_b200/lib/libpmvs_foot.so python tools/footprints.py [patches]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from pmvs_b200 import abi  # noqa: E402
from pmvs_b200.api import PatchRefiner  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = bench.bench_config()
sc = bench.make_scene(cfg)
patches = sc.patches(n, seed=5678)
with PatchRefiner(cfg, sc.records, seed=42) as pr:
    out = pr.refine(patches, flags=abi.F_POST_REMOVE_INVISIBLE)
    buf = (C.c_int * (n * 32))()
    pr.L.pmvs_debug_footprints.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    assert pr.L.pmvs_debug_footprints(pr.h, n, buf) == 0
f = np.frombuffer(buf, dtype=np.int32).reshape(n, 8, 4)
ref = np.array([q.refCamIdx for q in out])
w, h = [], []
for i in range(n):
    for v in range(5):
        if v == ref[i] or f[i, v, 2] <= f[i, v, 0]:
            continue
        w.append(f[i, v, 2] - f[i, v, 0] + 1)      # + 1: the bilinear taps read one pixel beyond the floor
        h.append(f[i, v, 3] - f[i, v, 1] + 1)
w, h = np.array(w), np.array(h)
side = np.maximum(w, h)
res = {"patches": n, "view_boxes": int(len(side)), "window": cfg.patchSize,
       "side_percentiles": {str(p): int(np.percentile(side, p)) for p in (1, 10, 25, 50, 75, 90, 99, 100)},
       "fraction_within": {str(t): float((side <= t).mean()) for t in (40, 48, 56, 64, 80, 96, 128, 192, 256)},
       "quad_tile_KB_per_view": {str(t): t * t * 4 / 1024.0 for t in (48, 64, 96, 128)}}
ev = f[:, 7, :].astype(np.int64)          # per patch: evaluations {all, within 48, 64, 96} of the tile centred on the swarm's first window
ev128 = f[:, 6, 0].astype(np.int64)
tot = ev[:, 0].sum()
res["evaluations"] = int(tot)
res["evaluations_served_by_tile"] = {"48": float(ev[:, 1].sum() / tot), "64": float(ev[:, 2].sum() / tot), "96": float(ev[:, 3].sum() / tot),
                                     "128": float(ev128.sum() / tot)}
print(json.dumps(res, indent=1))
