"""The refine() test cases shared by the GPU parity tests (CUDA path vs oracle) and the CPU pin tests (oracle vs the
unmodified reference build): small scenes that exercise every branch of Patch::refine (TMVS/mvs/patch.cpp:114-176)."""
from pmvs_b200 import abi, scene

CASES = ["expand_r7", "seed_r7", "wide_arc", "occluded", "gradient_lod", "p5_defaults", "v12", "v16_p32", "r17_wide", "v34_grad"]


def build(case):
    """-> cfg, scene, patches, flags, patch type, n"""
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = 7, 15, 7 / 3.0, 2
    kw = dict(nviews=5, width=400, height=300, seed=21, with_edge=True, tex_size=1024)
    ptype, n, flags, extent = abi.TYPE_EXPAND, 40, abi.F_POST_REMOVE_INVISIBLE, None
    if case == "seed_r7":
        ptype, n = abi.TYPE_SEED, 12
    elif case == "wide_arc":            # cameras beyond the visibility cone: region-ratio / normal tests remove views
        kw.update(nviews=7, arc_deg=58.0)
        cfg.minRegionRatio = 0.55
        flags |= abi.F_EXPAND_VISIBLE
    elif case == "occluded":
        kw.update(nviews=6)
    elif case == "gradient_lod":
        cfg.adaptiveGradientEnable = 1
        cfg.textureVariation = 2500.0   # forces LOD > 0 (patch.cpp:546)
        extent = 2.3                    # border patches: drops and sentinels
    elif case == "p5_defaults":
        cfg = abi.default_config()
        cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = 7, 15, 7 / 3.0, 2
    elif case == "v12":
        kw.update(nviews=12, arc_deg=35.0)
        n = 16
    elif case == "v16_p32":             # BASELINE.json config 3 shape: 16 views, 32 particles x 50 iterations
        kw.update(nviews=16, arc_deg=35.0)
        cfg.particleNum, cfg.maxIteration = 32, 50
        n = 8
    elif case == "v34_grad":            # BASELINE.json config 4 shape: more than 32 views, gradient weighting (the many-view loop, 4 lanes per pixel)
        kw.update(nviews=34, arc_deg=30.0)
        cfg.adaptiveGradientEnable = 1
        n = 4
    elif case == "r17_wide":            # window wider than a warp: two column passes
        cfg.patchRadius, cfg.patchSize, cfg.distWeighting = 17, 35, 17 / 3.0
        n = 12
    else:
        assert case == "expand_r7", case
    sc = scene.SynthScene(cfg, **kw)
    if case == "occluded":              # one view shows unrelated texture: correlation test removes it (patch.cpp:703)
        other = scene.SynthScene(cfg, nviews=1, width=400, height=300, seed=999, with_edge=True, tex_size=512)
        for l in range(len(sc.cams[4].levels)):
            sc.cams[4].levels[l][0][:] = other.cams[0].levels[l][0]
    patches = sc.patches(n, seed=5, ptype=ptype, extent=extent)
    return cfg, sc, patches, flags, ptype, n
