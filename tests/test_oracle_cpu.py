"""CPU tests: the oracle against the reference's own compiled solver, the committed golden vectors, an independent
NumPy reading of getFitness and closed-form cases. (The reference ships no tests or fixtures: SURVEY.md section 4.)"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import np_reference
import orc
from pmvs_b200 import abi, scene

GOLD = os.path.join(os.path.dirname(__file__), "golden")
D3 = C.c_double * 3


def fromhex(v):
    return float.fromhex(v)


def load_golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def golden_scene():
    m = load_golden_module()
    cfg, sc = m.golden_scene()
    return m, cfg, sc


def run_orc_pso(c):
    L = orc.lib()
    gb, gf, it = D3(), C.c_double(), C.c_int()
    init = D3(*c["init"]) if c["init"] else None
    L.orc_pso_test(c["fn"], D3(*c["L"]), D3(*c["U"]), c["maxIter"], c["P"], init, c["key"], c["glnpso"], gb, C.byref(gf),
                   C.byref(it), None)
    return list(gb), gf.value, it.value


def test_pso_restatement_matches_golden_from_unmodified_reference():
    kat = json.load(open(os.path.join(GOLD, "pso_kat.json")))
    assert len(kat["cases"]) >= 32
    for c in kat["cases"]:
        gb, gf, it = run_orc_pso(c)
        assert [v.hex() for v in gb] == c["gbest"], c
        assert gf.hex() == c["gbestFitness"]
        assert it == c["iterations"]


def test_pso_restatement_matches_live_reference_solver():
    R = orc.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref/libpso_ref.so not built (no /root/reference on this box)")
    L = orc.lib()
    rng = np.random.RandomState(3)
    for k in range(40):
        c = dict(fn=int(rng.choice([0, 1, 2, 3, 4])), P=int(rng.randint(2, 40)), maxIter=int(rng.randint(1, 40)),
                 glnpso=int(rng.randint(0, 2)), L=[-2.0, -1.0, 0.5], U=[1.0, 2.0, 2.5],
                 init=[0.0, 0.5, 1.0] if k % 3 else None, key=int(rng.randint(1, 1 << 30)))
        gb, gf, it = D3(), C.c_double(), C.c_int()
        init = D3(*c["init"]) if c["init"] else None
        R.ref_pso_solve_basic(D3(*c["L"]), D3(*c["U"]), L.orc_test_fn_ptr(), C.byref(C.c_int(c["fn"])), c["maxIter"], c["P"],
                              init, c["key"], c["glnpso"], gb, C.byref(gf), C.byref(it))
        gb2, gf2, it2 = run_orc_pso(c)
        assert list(gb) == gb2 and gf.value == gf2 and it.value == it2, c


def test_rng_stream_definition():
    from pmvs_b200 import rng
    L = orc.lib()
    for seed, pid, run, ctr in [(42, 0, 0, 0), (42, 7, 1, 5), (2 ** 63 + 5, 123456, 3, 2 ** 40)]:
        assert L.orc_rand31(seed, pid, run, ctr) == rng.rand31(rng.stream_key(seed, pid, run), ctr)
        assert L.orc_rand31(seed, pid, run, ctr) < 2 ** 31


def test_fitness_golden(golden_scene):
    m, cfg, sc = golden_scene
    kat = json.load(open(os.path.join(GOLD, "fitness_kat.json")))
    assert m.scene_digest(sc) == kat["scene_sha256"], "synthetic scene drifted from the golden vectors"
    patches = sc.patches(24, seed=77)
    for name, c in m.fitness_configs(cfg).items():
        o = orc.Oracle(c, sc.records, seed=42)
        for e in kat["configs"][name]:
            hy = scene.hypotheses_from_patches(sc, patches, c, lod=e["lod"], seed=5 + e["lod"], per_patch=3, spread=1.0 + e["lod"])
            got = o.fitness_batch(hy)
            assert [v.hex() for v in got] == e["fitness"], (name, e["lod"])


def test_refine_golden(golden_scene):
    m, cfg, sc = golden_scene
    kat = json.load(open(os.path.join(GOLD, "refine_kat.json")))
    assert m.scene_digest(sc) == kat["scene_sha256"]
    o = orc.Oracle(cfg, sc.records, seed=42)      # restated solver; golden came from the unmodified one
    for s in kat["sets"]:
        ps = sc.patches(s["n"], seed=s["seed"], ptype=s["type"], first_id=s["first_id"])
        out = o.refine_batch(ps, flags=abi.F_POST_REMOVE_INVISIBLE)
        for q, g in zip(out, s["records"]):
            assert [v.hex() for v in q.center] == g["center"]
            assert [v.hex() for v in q.normal] == g["normal"]
            assert q.fitness.hex() == g["fitness"] and q.correlation.hex() == g["correlation"]
            assert (q.drop, q.nCam, list(q.camIdx[:q.nCam]), q.LOD, q.refCamIdx) == (g["drop"], g["nCam"], g["camIdx"], g["LOD"], g["refCamIdx"])
            assert (q.psoRuns, q.psoIterations, q.evaluations) == (g["psoRuns"], g["psoIterations"], g["evaluations"])


def test_fitness_vs_numpy_and_sentinels(small_scene):
    cfg, sc = small_scene
    c = abi.PmvsConfig.from_buffer_copy(cfg)
    c.adaptiveGradientEnable = 1
    o = orc.Oracle(c, sc.records, seed=1)
    patches = sc.patches(30, seed=9, extent=2.2)          # reaches the image borders -> DBL_MAX sentinels
    hy = scene.hypotheses_from_patches(sc, patches, c, lod=0, seed=2, per_patch=4, spread=3.0)
    hy[1].theta = math.pi - 0.1                            # normal facing away: patch.cpp:939-941
    got = o.fitness_batch(hy)
    n_max = 0
    for h, v in zip(hy, got):
        v2 = np_reference.fitness(sc.cams, c, h)
        if v == abi.DBL_MAX or v2 == abi.DBL_MAX:
            assert v == v2
            n_max += 1
        else:
            assert abs(v - v2) <= 1e-9 * max(1.0, abs(v))
    assert got[1] == abi.DBL_MAX
    assert 0 < n_max < len(hy)


def test_closed_form_constant_image_and_masked_window():
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.maxLOD, cfg.distWeighting = 5, 11, 0, 5 / 3.0
    sc = scene.SynthScene(cfg, nviews=3, width=160, height=120, tex_size=256)
    for cam in sc.cams:
        cam.levels[0][0][:] = 77
    patches = sc.patches(4, seed=1)
    hy = scene.hypotheses_from_patches(sc, patches, cfg, per_patch=1)
    o = orc.Oracle(cfg, sc.records)
    assert all(0 <= v < 1e-12 for v in o.fitness_batch(hy))   # constant image: avgSad = 0 up to bilinear rounding
    for cam in sc.cams:
        cam.levels[0][0][:] = 0
    o = orc.Oracle(cfg, sc.records)
    assert all(math.isnan(v) for v in o.fitness_batch(hy))   # every sample masked: 0/0 (patch.cpp:986, :1046)


def test_true_geometry_minimises_cost(small_scene):
    cfg, sc = small_scene
    o = orc.Oracle(cfg, sc.records)
    patches = sc.patches(6, seed=4, normal_jitter_deg=0.0, depth_jitter=0.0)
    base = scene.hypotheses_from_patches(sc, patches, cfg, per_patch=1)
    f0 = o.fitness_batch(base)
    for sign in (-1, 1):
        for h in base:
            h.depth += sign * 0.05
        f1 = o.fitness_batch(base)
        for h in base:
            h.depth -= sign * 0.05
        assert all(a < b for a, b in zip(f0, f1))


def test_refine_recovers_plane(small_scene):
    cfg, sc = small_scene
    o = orc.Oracle(cfg, sc.records)
    out = o.refine_batch(sc.patches(8, seed=11), flags=abi.F_POST_REMOVE_INVISIBLE)
    kept = [q for q in out if not q.drop]
    assert len(kept) >= 6
    for q in kept:
        assert abs(q.center[2] - sc.plane_z) < 5e-3 and q.normal[2] > 0.995 and q.nCam >= cfg.minCamNum


def test_fit_ellipse_against_cv2():
    cv2 = pytest.importorskip("cv2")
    L = orc.lib()
    rng = np.random.RandomState(0)
    F8 = C.c_float * 8
    for _ in range(50):
        A = np.eye(2) + 0.4 * rng.randn(2, 2)
        t = rng.rand(2) * 200
        r = 15.0
        sq = np.array([[-r, -r], [-r, r], [r, r], [r, -r], [-r, 0], [0, r], [r, 0], [0, -r]])
        pts = (sq @ A.T + t).astype(np.float32)
        (cx, cy), (w, h), ang = cv2.fitEllipse(pts.reshape(-1, 1, 2))
        want = min(w, h) / max(w, h)
        got = L.orc_fit_ellipse_ratio(F8(*pts[:, 0]), F8(*pts[:, 1]), 8, None, None)
        assert abs(got - want) < 2e-3, (got, want)


def test_dist_weight_table(small_scene):
    cfg, sc = small_scene
    o = orc.Oracle(cfg, sc.records)
    w = np.array(o.dist_weight(cfg.patchSize)).reshape(cfg.patchSize, cfg.patchSize)
    assert abs(w.sum() - 1) < 1e-12 and np.allclose(w, np_reference.dist_table(cfg), rtol=1e-12, atol=0)


def test_pyramid_restatement_against_cv2():
    """The NumPy restatement of the Camera ctor's pyramid (oracle/orc_pyramid.py: INTER_AREA resize, fractional and integer
    scales, and the Sobel(ksize=1) edge image, camera.cpp:71-92) is BIT-IDENTICAL to OpenCV's own cv::resize / cv::Sobel."""
    cv2 = pytest.importorskip("cv2")
    import orc_pyramid
    rng = np.random.RandomState(5)
    from scipy.ndimage import gaussian_filter
    for (h, w) in ((389, 613), (240, 320), (121, 203)):
        img = np.clip(gaussian_filter(rng.rand(h, w), 1.5) * 900 - 320, 0, 255).astype(np.uint8)
        for f in (0.8, 0.8 ** 2, 0.8 ** 5, 0.7, 0.5, 0.25, 1 / 3.0):
            mine = orc_pyramid.resize_area(img, f)
            ref = cv2.resize(img, None, fx=f, fy=f, interpolation=cv2.INTER_AREA)
            assert mine.shape == ref.shape and np.array_equal(mine, ref), (h, w, f)
        e = orc_pyramid.edge_image(img)
        gx = cv2.Sobel(img, cv2.CV_64F, 1, 0, ksize=1)
        gy = cv2.Sobel(img, cv2.CV_64F, 0, 1, ksize=1)
        m = np.sqrt(gx * gx + gy * gy)
        assert np.array_equal(e, (m - m.min()) / (m.max() - m.min()))


def test_homographies_are_the_plane_induced_maps(small_scene):
    """getHomographies (patch.cpp:290-330) pinned geometrically, independent of how it is computed: for points X on the
    hypothesis plane, H_v maps the reference camera's pixel of X (at the patch's LOD) onto view v's pixel of X; the
    reference view's H is the identity (:317-319). And the 3x3 inverse it uses (cv::Mat::inv, closed form for n <= 3)
    against cv2.invert."""
    cfg, sc = small_scene
    o = orc.Oracle(cfg, sc.records, seed=42)
    patches = sc.patches(6, seed=21)
    rng = np.random.RandomState(4)
    for lod in (0, 1, 2):
        hyps = scene.hypotheses_from_patches(sc, patches, cfg, lod=lod, per_patch=2, spread=1.0)
        for h in hyps:
            Hs = np.array(o.homographies(h)).reshape(h.nCam, 3, 3)
            ref = sc.cams[h.refCamIdx]
            n = np.array([math.sin(h.theta) * math.cos(h.phi), math.sin(h.theta) * math.sin(h.phi), math.cos(h.theta)])
            center = np.array(list(h.ray)) * h.depth + ref.center
            a = np.cross(n, [1.0, 0.3, 0.2])
            a /= np.linalg.norm(a)
            b = np.cross(n, a)
            for _ in range(5):
                X = center + 0.2 * rng.randn() * a + 0.2 * rng.randn() * b          # on the plane
                xr = ref.project(X, lod, cfg.lodRatio)
                for k in range(h.nCam):
                    v = h.camIdx[k]
                    q = Hs[k] @ np.array([xr[0], xr[1], 1.0])
                    assert np.allclose(q[:2] / q[2], sc.cams[v].project(X, lod, cfg.lodRatio), rtol=0, atol=1e-7)
                    if v == h.refCamIdx:
                        assert np.array_equal(Hs[k], np.eye(3))
    cv2 = pytest.importorskip("cv2")
    L = orc.lib()
    for _ in range(50):
        A = rng.randn(3, 3) * 10 ** rng.uniform(-2, 3)
        out = (C.c_double * 9)()
        L.orc_inv3(A.ravel().ctypes.data_as(C.POINTER(C.c_double)), out)
        ok, ref_inv = cv2.invert(A)
        assert ok != 0 and np.allclose(np.array(list(out)).reshape(3, 3), ref_inv, rtol=1e-12, atol=0)
