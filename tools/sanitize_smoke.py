"""Small GPU workload for compute-sanitizer (memcheck / racecheck): fitness + refine (expansion and seed patches, with
visibility removal) + swarm test + pyramid build on tiny inputs."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
import numpy as np  # noqa: E402
from pmvs_b200 import abi, api, scene  # noqa: E402
from pmvs_b200.api import PatchRefiner  # noqa: E402

cfg = abi.readme_config()
cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = 4, 9, 4 / 3.0, 1
cfg.particleNum, cfg.maxIteration, cfg.adaptiveGradientEnable = 6, 4, 1
sc = scene.SynthScene(cfg, nviews=7, width=160, height=120, seed=5, with_edge=True, tex_size=256, arc_deg=55.0, background=3)
lean = scene.camera_array(sc.cams)
for c in lean:
    for l in range(1, c.maxLOD + 1):
        c.level[l].grey = None
        c.level[l].edge = None
with PatchRefiner(cfg, lean) as pr:
    ps = sc.patches(6, seed=1, extent=1.6)
    hy = scene.hypotheses_from_patches(sc, ps, cfg, per_patch=2, spread=2.0)
    f = pr.fitness(hy)
    out = pr.refine(ps, flags=abi.F_POST_REMOVE_INVISIBLE | abi.F_EXPAND_VISIBLE)
    seeds = sc.patches(3, seed=2, ptype=abi.TYPE_SEED)
    out2 = pr.refine(seeds)
    r = pr.pso_test([dict(L=[-1, -1, 0], U=[1, 1, 2], init=None, maxIter=5, P=9, fn=0, key=7)])
# the flag-free fast loop (distance + difference weights only), the batched scalar part and the latency launch configurations
cfg2 = abi.readme_config()
cfg2.patchRadius, cfg2.patchSize, cfg2.distWeighting, cfg2.maxLOD = 4, 9, 4 / 3.0, 1
cfg2.particleNum, cfg2.maxIteration = 7, 3
sc2 = scene.SynthScene(cfg2, nviews=5, width=160, height=120, seed=6, tex_size=256)
with PatchRefiner(cfg2, sc2.records) as pr:
    out3 = pr.refine(sc2.patches(5, seed=3), flags=abi.F_POST_REMOVE_INVISIBLE)
    f3 = pr.fitness(scene.hypotheses_from_patches(sc2, sc2.patches(4, seed=4), cfg2, per_patch=2))
# many views: the rows loop with 8 (50 views), 4 (34), 2 (18) and 1 (12) lanes per pixel
for nv in (50, 34, 18, 12):
    sc3 = scene.SynthScene(cfg2, nviews=nv, width=160, height=120, seed=7, tex_size=256, arc_deg=30.0)
    with PatchRefiner(cfg2, sc3.records) as pr:
        f4 = pr.fitness(scene.hypotheses_from_patches(sc3, sc3.patches(3, seed=5), cfg2, per_patch=2))
        out4 = pr.refine(sc3.patches(2, seed=6))
# the -f pair scan
cnt = api.neighbor_counts(np.random.RandomState(0).rand(700, 3), 0.1)
pyr = api.build_pyramid(sc.cams[0].levels[0][0], cfg.lodRatio, 2, with_edge=True)
print("sanitize workload ok", len(f), sum(1 for q in out if not q.drop), sum(1 for q in out2 if not q.drop), r[0]["iterations"], len(pyr),
      sum(1 for q in out3 if not q.drop), len(f3), len(f4), sum(1 for q in out4 if not q.drop), int(cnt.sum()))
