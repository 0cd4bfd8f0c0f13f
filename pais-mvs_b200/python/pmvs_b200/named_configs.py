"""BASELINE.json's five configurations (`configs[0..4]`) as test / bench scenes. configs[0..2] at their named sizes;
configs[3] and [4] keep the named view count, patch radius, weights and swarm but may be built on a reduced image size
(`scale`) where a test has to synthesise them on the host in seconds (the work of one evaluation, O(V (2r+1)^2), does not
depend on the image size: SURVEY.md section 5)."""
from . import abi, scene

CONFIGS = {
    1: dict(name="configs[0]: 5 views 640x480, patchRadius=7, 1 pyramid level, adaptive weights off", views=5, w=640, h=480, r=7, levels=1,
            weights=(0, 0, 0), P=15, I=30, n=8192, check=64, arc=30.0),
    2: dict(name="configs[1]: 5 views 1600x1200, patchRadius=15, 3 pyramid levels, adaptive distance+difference on", views=5, w=1600, h=1200,
            r=15, levels=3, weights=(1, 1, 0), P=15, I=30, n=8192, check=64, arc=30.0),
    3: dict(name="configs[2]: 16 views 1600x1200, patchRadius=15, PSO 32 particles x 50 iters, visibleCorrelation=0.7", views=16, w=1600,
            h=1200, r=15, levels=3, weights=(1, 1, 0), P=32, I=50, n=8192, check=64, arc=30.0),
    4: dict(name="configs[3]: 32 views 1920x1080, patchRadius=15, adaptiveGradient on", views=32, w=1920, h=1080, r=15, levels=3,
            weights=(1, 1, 1), P=15, I=30, n=8192, check=64, arc=30.0),
    5: dict(name="configs[4]: 64 views 4000x3000, patchRadius=21, full adaptive weighting", views=64, w=4000, h=3000, r=21, levels=3,
            weights=(1, 1, 1), P=15, I=30, n=8192, check=64, arc=30.0),
}


def config_of(k):
    c = CONFIGS[k]
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = c["r"], 2 * c["r"] + 1, c["r"] / 3.0, c["levels"] - 1
    cfg.adaptiveDistanceEnable, cfg.adaptiveDifferenceEnable, cfg.adaptiveGradientEnable = c["weights"]
    cfg.particleNum, cfg.maxIteration, cfg.visibleCorrelation = c["P"], c["I"], 0.7
    return cfg


def build(k, scale=1.0, seed=1234):
    """-> config record, MvsConfig, scene (images scale x the named size)"""
    c = CONFIGS[k]
    cfg = config_of(k)
    sc = scene.SynthScene(cfg, nviews=c["views"], width=int(round(c["w"] * scale)), height=int(round(c["h"] * scale)), seed=seed,
                          with_edge=bool(c["weights"][2]), arc_deg=c["arc"])
    return c, cfg, sc
