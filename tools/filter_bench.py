"""Timing of the PCMVS pair scan (pmvs_neighbor_counts) on one GPU next to the CPU oracle (all host cores) on a bounded
sample: python tools/filter_bench.py [n]"""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
from pmvs_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
pts = np.random.RandomState(1).rand(n, 3)
r = 0.012
api.neighbor_counts(pts[:1024], r)
t0 = time.time()
got = api.neighbor_counts(pts, r)
t1 = time.time()
print("gpu: n=%d  %.1f ms  %.3g pairs/s (host buffers, incl. copies and allocation)" % (n, 1e3 * (t1 - t0), n * float(n) / (t1 - t0)))
import orc  # noqa: E402
m = min(n, 30000)
t0 = time.time()
want = orc.neighbor_counts(pts[:m], r)
t1 = time.time()
print("cpu oracle (%d threads): n=%d  %.1f ms  %.3g pairs/s" % (os.cpu_count(), m, 1e3 * (t1 - t0), m * float(m) / (t1 - t0)))
assert np.array_equal(api.neighbor_counts(pts[:m], r), want)
