"""The device pyramid builder (csrc/pmvs_pyramid.cuh) on one 4000x3000 image, 3 levels + edges: run under
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
to list its kernels with their device time and DRAM bytes (profiles/r2_pyramid_kernels.md)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
import numpy as np  # noqa: E402
from pmvs_b200 import api  # noqa: E402

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4000, 3000)
rng = np.random.RandomState(3)
img = rng.randint(1, 256, size=(h, w)).astype(np.uint8)
for _ in range(2):
    lv = api.build_pyramid(img, 0.8, 2, with_edge=True)
print([g.shape for g, e in lv])
