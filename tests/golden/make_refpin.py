"""Golden vectors of the hot path produced by the UNMODIFIED reference build (oracle/_ref/libtmvs_ref.so, see
oracle/ref_patch_shim.cpp): `python tests/golden/make_refpin.py` rewrites refpin_kat.json (needs /root/reference to have
been compiled by `make -C oracle`). generate() is also what the test runs with the restatement in place of the reference."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
for p in (os.path.join(ROOT, "pais-mvs_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from pmvs_b200 import abi, scene  # noqa: E402


def pin_scene():
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = 7, 15, 7 / 3.0, 2
    cfg.adaptiveGradientEnable = 1
    sc = scene.SynthScene(cfg, nviews=6, width=320, height=240, seed=4242, with_edge=True, tex_size=1024, arc_deg=50.0)
    return cfg, sc


def digest(sc):
    h = hashlib.sha256()
    for cam in sc.cams:
        for g, e in cam.levels:
            h.update(g.tobytes())
    return h.hexdigest()


def rec(q):
    f = lambda v: float(v).hex()
    return {"drop": q.drop, "nCam": q.nCam, "camIdx": list(q.camIdx[:q.nCam]), "LOD": q.LOD, "refCamIdx": q.refCamIdx, "psoRuns": q.psoRuns,
            "center": [f(v) for v in q.center], "normal": [f(v) for v in q.normal], "fitness": f(q.fitness), "correlation": f(q.correlation),
            "priority": f(q.priority), "depthRange": [f(v) for v in q.depthRange]}


def generate(make):
    """make(cfg, records) -> object with fitness_batch / homographies / refine_batch (the reference build or the oracle)."""
    cfg, sc = pin_scene()
    impl = make(cfg, sc.records)
    out = {"scene_sha256": digest(sc), "fitness": [], "homographies": [], "refine": []}
    patches = sc.patches(20, seed=11, extent=2.0)
    for lod in (0, 1):
        hyps = scene.hypotheses_from_patches(sc, patches, cfg, lod=lod, per_patch=3, spread=1.5)
        out["fitness"].append([v.hex() for v in impl.fitness_batch(hyps)])
        out["homographies"].append([v.hex() for v in impl.homographies(hyps[0])])
    for ptype, n, flags in ((abi.TYPE_EXPAND, 16, abi.F_POST_REMOVE_INVISIBLE | abi.F_EXPAND_VISIBLE), (abi.TYPE_SEED, 6, abi.F_POST_REMOVE_INVISIBLE)):
        ps = sc.patches(n, seed=13, ptype=ptype, first_id=100)
        out["refine"].append([rec(q) for q in impl.refine_batch(ps, flags=flags)])
    return out


if __name__ == "__main__":
    import ref_tmvs
    kat = generate(lambda cfg, records: ref_tmvs.RefScene(cfg, records, seed=42))
    kat["generator"] = "unmodified reference (TMVS/mvs/patch.cpp et al. compiled against oracle/cvshim): oracle/_ref/libtmvs_ref.so"
    json.dump(kat, open(os.path.join(HERE, "refpin_kat.json"), "w"), indent=0)
    print("wrote refpin_kat.json:", sum(len(x) for x in kat["fitness"]), "fitness values,", sum(len(x) for x in kat["refine"]), "refined patches")
