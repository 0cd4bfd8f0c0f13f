"""CPU tests: the f64 restatement (oracle/pmvs_oracle.cpp) against the UNMODIFIED reference patch model.

oracle/_ref/libtmvs_ref.so is TMVS/mvs/{patch,abstractpatch,camera,cellmap,mvs}.cpp + TMVS/pso/*.cpp compiled in place from
/root/reference against the OpenCV stand-in oracle/cvshim (oracle/Makefile; harness oracle/ref_patch_shim.cpp, which only
injects the scene and interposes rand()). Every comparison below is BIT FOR BIT: PAIS::getFitness (patch.cpp:914-1047),
Patch::getHomographies (:290-330), getHomographyRegionRatio (:269-288), Camera::project (camera.cpp:138-160),
MVS::initPatchDistanceWeighting (mvs.cpp:97-114) and the whole Patch::refine() / removeInvisibleCamera() /
expandVisibleCamera() path (:114-176, :655-761) on the scenes the GPU parity tests use. The library travels to the GPU
box prebuilt; where it is absent (fresh checkout without /root/reference) the committed golden vectors it produced are
checked instead (tests/golden/refpin_kat.json, written by tests/golden/make_refpin.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import orc
import ref_tmvs
import refine_cases
from pmvs_b200 import abi, scene

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(ref_tmvs.lib() is None, reason="oracle/_ref/libtmvs_ref.so not built (no /root/reference on this box)")


def record_key(q):
    """Every field refine() writes, as exact bit patterns."""
    f = lambda v: float(v).hex()
    return (q.drop, q.nCam, tuple(q.camIdx[:q.nCam]), q.LOD, q.refCamIdx, q.psoRuns, tuple(f(v) for v in q.center), tuple(f(v) for v in q.normal),
            tuple(f(v) for v in q.normalS), tuple(f(v) for v in q.ray), f(q.depth), tuple(f(v) for v in q.depthRange), f(q.fitness), f(q.priority),
            f(q.correlation), q.nImgPoint, tuple(f(q.imgPoint[k][j]) for k in range(q.nImgPoint) for j in range(2)))


def scene_for(weights, nviews=5, radius=7, width=320, height=240, **kw):
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = radius, 2 * radius + 1, radius / 3.0, 2
    cfg.adaptiveDistanceEnable, cfg.adaptiveDifferenceEnable, cfg.adaptiveGradientEnable = weights
    return cfg, scene.SynthScene(cfg, nviews=nviews, width=width, height=height, seed=77, with_edge=True, tex_size=1024, **kw)


@needs_ref
@pytest.mark.parametrize("nviews,radius,weights", [(5, 7, (0, 0, 0)), (5, 15, (1, 1, 0)), (3, 4, (1, 1, 1)), (12, 7, (1, 1, 1)), (2, 7, (1, 0, 1)),
                                                   (20, 5, (1, 1, 0)), (33, 5, (1, 1, 1)), (64, 4, (1, 1, 1)), (6, 21, (1, 1, 1))])
def test_fitness_homographies_weights_bit_exact(nviews, radius, weights):
    cfg, sc = scene_for(weights, nviews=nviews, radius=radius, width=(640 if radius > 15 else 480) if radius > 7 else 320, height=(480 if radius > 15 else 360) if radius > 7 else 240)
    o = orc.Oracle(cfg, sc.records, seed=42)
    r = ref_tmvs.RefScene(cfg, sc.records, seed=42)
    assert [v.hex() for v in o.dist_weight(cfg.patchSize)] == [v.hex() for v in r.dist_weight()]
    patches = sc.patches(24, seed=3, extent=3.0)           # extent beyond the images: sentinels (DBL_MAX) included
    sentinels = finite = 0
    for lod in (0, 1, 2):
        hyps = scene.hypotheses_from_patches(sc, patches, cfg, lod=lod, per_patch=4, spread=1.5)
        fo, fr = o.fitness_batch(hyps), r.fitness_batch(hyps)
        assert [v.hex() for v in fo] == [v.hex() for v in fr]
        sentinels += sum(1 for v in fr if v == abi.DBL_MAX)
        finite += sum(1 for v in fr if v < abi.DBL_MAX)
        for h in hyps[:8]:
            assert [v.hex() for v in o.homographies(h)] == [v.hex() for v in r.homographies(h)]
    assert sentinels > 0 and finite > 0, (sentinels, finite)


@needs_ref
def test_project_and_region_ratio_bit_exact():
    cfg, sc = scene_for((1, 1, 0))
    o = orc.Oracle(cfg, sc.records, seed=42)
    r = ref_tmvs.RefScene(cfg, sc.records, seed=42)
    rng = np.random.RandomState(5)
    L = orc.lib()
    for _ in range(200):
        X = (rng.rand(3) - 0.5) * np.array([4.0, 4.0, 1.0])
        cam, lod = int(rng.randint(5)), int(rng.randint(3))
        out = (C.c_double * 2)()
        ok = L.orc_project(o.h, cam, (C.c_double * 3)(*X), lod, out)
        ok_r, out_r = r.project(cam, list(X), lod)
        assert (ok, [v.hex() for v in out]) == (ok_r, [v.hex() for v in out_r])
    patches = sc.patches(8, seed=3)
    hyps = scene.hypotheses_from_patches(sc, patches, cfg, per_patch=2)
    for h in hyps:
        H = o.homographies(h)
        for v in range(h.nCam):
            pt = [160.0 + 3.3 * v, 120.0 - 1.7 * v]
            Hv = H[9 * v:9 * v + 9]
            a = L.orc_region_ratio(o.h, (C.c_double * 2)(*pt), (C.c_double * 9)(*Hv))
            assert a.hex() == r.region_ratio(pt, Hv).hex()


@needs_ref
@pytest.mark.parametrize("case", refine_cases.CASES)
def test_refine_bit_exact(case):
    """The cases of tests/test_gpu_parity.py::test_refine_vs_oracle: restatement == unmodified reference, every output field."""
    cfg, sc, patches, flags, ptype, n = refine_cases.build(case)
    if case in ("v16_p32", "v12", "v34_grad"):
        patches = (abi.PmvsPatchIn * 6)(*patches[:6])      # the reference allocates ~6 V cv::Mat per evaluation: keep the CPU suite short
    o = orc.Oracle(cfg, sc.records, seed=42, use_ref_pso=False)          # restated solver AND restated patch model
    r = ref_tmvs.RefScene(cfg, sc.records, seed=42)
    want = r.refine_batch(patches, flags=flags)
    got = o.refine_batch(patches, flags=flags, patch_threads=4)
    for i, (g, w) in enumerate(zip(got, want)):
        assert record_key(g) == record_key(w), (case, i)
    if case in ("wide_arc", "occluded"):
        assert any((not q.drop) and q.nCam != p.nCam for q, p in zip(want, patches)) or any(q.drop for q in want)


def test_restatement_matches_reference_golden():
    """Golden vectors written by the unmodified reference build (tests/golden/make_refpin.py): checked on every box."""
    kat = json.load(open(os.path.join(GOLD, "refpin_kat.json")))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_refpin", os.path.join(GOLD, "make_refpin.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    got = m.generate(lambda cfg, records: orc.Oracle(cfg, records, seed=42, use_ref_pso=False))
    assert got["scene_sha256"] == kat["scene_sha256"]
    assert got["fitness"] == kat["fitness"]
    assert got["homographies"] == kat["homographies"]
    assert got["refine"] == kat["refine"]
    assert kat["generator"].startswith("unmodified reference")
