"""GPU-vs-oracle diagnostic for one refine case: prints every differing field of every mismatching patch."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402
from pmvs_b200 import abi, scene  # noqa: E402
from pmvs_b200.api import PatchRefiner  # noqa: E402

cfg = abi.readme_config()
cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = 7, 15, 7 / 3.0, 2
sc = scene.SynthScene(cfg, nviews=5, width=400, height=300, seed=21, with_edge=True, tex_size=1024)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
patches = sc.patches(n, seed=5, ptype=abi.TYPE_SEED)
o = orc.Oracle(cfg, sc.records, seed=42)
want = o.refine_batch(patches, flags=abi.F_POST_REMOVE_INVISIBLE, patch_threads=8)
with PatchRefiner(cfg, sc.records, seed=42) as pr:
    got = pr.refine(patches, flags=abi.F_POST_REMOVE_INVISIBLE)
for i, (g, w) in enumerate(zip(got, want)):
    print("patch %d: it %d/%d evals %d/%d wev %d/%d fit %s / %s" % (i, g.psoIterations, w.psoIterations, g.evaluations, w.evaluations,
          g.windowEvaluations, w.windowEvaluations, g.fitness.hex(), w.fitness.hex()))
    print("   normalS gpu %r\n   normalS cpu %r" % (list(g.normalS), list(w.normalS)))
    print("   depth %s / %s   center dz %.3g" % (g.depth.hex(), w.depth.hex(), g.center[2] - w.center[2]))
