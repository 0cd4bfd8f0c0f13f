"""BASELINE.json's five configurations at their NAMED sizes on one B200: a handful of patches checked against the CPU
oracle (same tolerances as tests/test_gpu_parity.py) and the throughput of a larger batch through the host-buffer C-ABI
call. One JSON line per configuration (profiles/r1_configs.json). usage: python tools/config_sweep.py [1 2 3 4 5]
(config 5 synthesises 64 views of 4000x3000 on the host: several minutes and ~25 GB of host memory)."""
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in ("pais-mvs_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import orc  # noqa: E402
from pmvs_b200 import abi, scene  # noqa: E402
from pmvs_b200.api import PatchRefiner  # noqa: E402
from test_gpu_parity import compare_refine  # noqa: E402

from pmvs_b200.named_configs import CONFIGS  # noqa: E402


def run(k):
    c = CONFIGS[k]
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = c["r"], 2 * c["r"] + 1, c["r"] / 3.0, c["levels"] - 1
    cfg.adaptiveDistanceEnable, cfg.adaptiveDifferenceEnable, cfg.adaptiveGradientEnable = c["weights"]
    cfg.particleNum, cfg.maxIteration, cfg.visibleCorrelation = c["P"], c["I"], 0.7
    t0 = time.time()
    sc = scene.SynthScene(cfg, nviews=c["views"], width=c["w"], height=c["h"], seed=1234, with_edge=bool(c["weights"][2]), arc_deg=30.0)
    t_scene = time.time() - t0
    patches = sc.patches(c["n"], seed=5678)
    flags = abi.F_POST_REMOVE_INVISIBLE | abi.F_EXPAND_VISIBLE
    t0 = time.time()
    with PatchRefiner(cfg, sc.records, seed=42) as pr:
        t_create = time.time() - t0
        nw = min(256, c["n"])
        pr.refine((abi.PmvsPatchIn * nw)(*[patches[i] for i in range(nw)]), flags=flags)      # warm-up
        t0 = time.time()
        out = pr.refine(patches, flags=flags)
        dt = time.time() - t0
    kept = sum(1 for q in out if not q.drop)
    evals = sum(q.evaluations for q in out)
    views = sum(q.nCam for q in out if not q.drop) / max(kept, 1)
    o = orc.Oracle(cfg, sc.records, seed=42, use_ref_pso=orc.ref_lib() is not None)
    nchk = c["check"]
    sub = (abi.PmvsPatchIn * nchk)(*[patches[i] for i in range(nchk)])
    t0 = time.time()
    want = o.refine_batch(sub, flags=flags, patch_threads=min(nchk, os.cpu_count() or 1))
    t_cpu = time.time() - t0
    worst = compare_refine([out[i] for i in range(nchk)], want)
    line = dict(config=c["name"], patches=c["n"], seconds=dt, patches_per_s=c["n"] / dt, evaluations_per_patch=evals / c["n"], kept=kept,
                mean_views_kept=views, oracle_checked=nchk, worst_relative_deviation=worst, oracle_seconds=t_cpu,
                cpu_patches_per_s=nchk / t_cpu, cpu_threads=min(nchk, os.cpu_count() or 1), scene_seconds=t_scene, create_seconds=t_create)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    for k in ([int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]):
        run(k)
