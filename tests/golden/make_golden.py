"""Regenerates the golden vectors under tests/golden/. Run in the build container (needs /root/reference for the
unmodified reference PSO, compiled by oracle/Makefile into oracle/_ref/):

    python tests/golden/make_golden.py

pso_kat.json      outputs of the UNMODIFIED reference solver (TMVS/pso/psosolver.cpp, particle.cpp) on analytic
                  functions under the interposed counter-based rand() (oracle/ref_pso_shim.cpp).
fitness_kat.json  PAIS::getFitness values of the f64 restatement on the seeded synthetic scene, each required to be BIT-IDENTICAL
                  to the unmodified reference's own getFitness (oracle/_ref/libtmvs_ref.so: TMVS/mvs/patch.cpp compiled in place
                  against oracle/cvshim) and cross-checked against the independent NumPy implementation (tests/np_reference.py)
                  before being written.
refine_kat.json   Patch::refine()+removeInvisibleCamera() outputs of the restatement driven by the unmodified
                  reference solver, same scene; likewise required to equal the unmodified reference's refine() bit for bit.
Floats are stored as C99 hex strings (exact)."""
import ctypes as C
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import orc  # noqa: E402
import ref_tmvs  # noqa: E402
import np_reference  # noqa: E402
from pmvs_b200 import abi, scene  # noqa: E402


def hx(v):
    return float(v).hex()


PSO_CASES = []
for fn in (0, 1, 3, 4):
    for (P, it) in ((5, 10), (15, 30), (30, 60), (64, 50)):
        for gln in (1, 0):
            PSO_CASES.append(dict(fn=fn, P=P, maxIter=it, glnpso=gln, L=[-2.0, -1.5, 0.0], U=[2.0, 1.5, 3.0],
                                  init=[0.1, 0.1, 1.0] if (P + fn) % 2 else None, key=1000 + 17 * fn + P))


def golden_scene():
    cfg = abi.readme_config()
    cfg.patchRadius = 7
    cfg.patchSize = 15
    cfg.distWeighting = 7 / 3.0
    cfg.maxLOD = 2
    sc = scene.SynthScene(cfg, nviews=5, width=320, height=240, seed=1234, with_edge=True, tex_size=1024, background=6)
    return cfg, sc


def scene_digest(sc):
    h = hashlib.sha256()
    for c in sc.cams:
        for g, e in c.levels:
            h.update(g.tobytes())
    return h.hexdigest()


def fitness_configs(cfg):
    """name -> config variants exercised by the fitness KAT"""
    out = {}
    for name, (d, f, g) in dict(none=(0, 0, 0), dist=(1, 0, 0), distdiff=(1, 1, 0), all=(1, 1, 1)).items():
        c = abi.PmvsConfig.from_buffer_copy(cfg)
        c.adaptiveDistanceEnable, c.adaptiveDifferenceEnable, c.adaptiveGradientEnable = d, f, g
        out[name] = c
    return out


def main():
    R = orc.ref_lib()
    assert R is not None, "oracle/_ref/libpso_ref.so missing (needs /root/reference)"
    L = orc.lib()
    D3 = C.c_double * 3
    cases = []
    for c in PSO_CASES:
        gb, gf, it = D3(), C.c_double(), C.c_int()
        init = D3(*c["init"]) if c["init"] else None
        R.ref_pso_solve_basic(D3(*c["L"]), D3(*c["U"]), L.orc_test_fn_ptr(), C.byref(C.c_int(c["fn"])), c["maxIter"], c["P"],
                              init, c["key"], c["glnpso"], gb, C.byref(gf), C.byref(it))
        cases.append(dict(c, gbest=[hx(v) for v in gb], gbestFitness=hx(gf.value), iterations=it.value))
    json.dump(dict(source="unmodified TMVS/pso/psosolver.cpp + particle.cpp via oracle/ref_pso_shim.cpp", cases=cases),
              open(os.path.join(HERE, "pso_kat.json"), "w"), indent=1)

    cfg, sc = golden_scene()
    patches = sc.patches(24, seed=77)
    fk = dict(scene_sha256=scene_digest(sc), configs={})
    for name, c in fitness_configs(cfg).items():
        o = orc.Oracle(c, sc.records, seed=42)
        ref = ref_tmvs.RefScene(c, sc.records, seed=42)
        entries = []
        for lod in (0, 1, 2):
            hy = scene.hypotheses_from_patches(sc, patches, c, lod=lod, seed=5 + lod, per_patch=3, spread=1.0 + lod)
            f = o.fitness_batch(hy)
            assert [hx(v) for v in f] == [hx(v) for v in ref.fitness_batch(hy)], (name, lod)
            for h, v in zip(hy, f):
                v2 = np_reference.fitness(sc.cams, c, h)
                if v == abi.DBL_MAX or v2 == abi.DBL_MAX or v != v:
                    assert (v == v2) or (v != v and v2 != v2), (name, lod, v, v2)
                else:
                    assert abs(v - v2) <= 1e-9 * max(1.0, abs(v)), (name, lod, v, v2)
            entries.append(dict(lod=lod, fitness=[hx(v) for v in f]))
        fk["configs"][name] = entries
    json.dump(fk, open(os.path.join(HERE, "fitness_kat.json"), "w"), indent=1)

    o = orc.Oracle(cfg, sc.records, seed=42, use_ref_pso=True)
    ref = ref_tmvs.RefScene(cfg, sc.records, seed=42)
    rk = dict(scene_sha256=scene_digest(sc), sets=[])
    for ptype, n, seed in ((abi.TYPE_EXPAND, 12, 31), (abi.TYPE_SEED, 4, 32)):
        ps = sc.patches(n, seed=seed, ptype=ptype, first_id=100 * ptype)
        out = o.refine_batch(ps, flags=abi.F_POST_REMOVE_INVISIBLE)
        for q, w in zip(out, ref.refine_batch(ps, flags=abi.F_POST_REMOVE_INVISIBLE)):
            assert ([hx(v) for v in q.center], [hx(v) for v in q.normal], hx(q.fitness), hx(q.correlation), hx(q.priority), q.drop, q.nCam,
                    list(q.camIdx[:q.nCam]), q.LOD, q.refCamIdx, q.psoRuns) == \
                   ([hx(v) for v in w.center], [hx(v) for v in w.normal], hx(w.fitness), hx(w.correlation), hx(w.priority), w.drop, w.nCam,
                    list(w.camIdx[:w.nCam]), w.LOD, w.refCamIdx, w.psoRuns)
        recs = []
        for q in out:
            recs.append(dict(center=[hx(v) for v in q.center], normal=[hx(v) for v in q.normal], fitness=hx(q.fitness),
                             correlation=hx(q.correlation), priority=hx(q.priority), drop=q.drop, nCam=q.nCam,
                             camIdx=list(q.camIdx[:q.nCam]), LOD=q.LOD, refCamIdx=q.refCamIdx, psoRuns=q.psoRuns,
                             psoIterations=q.psoIterations, evaluations=q.evaluations))
        rk["sets"].append(dict(type=ptype, n=n, seed=seed, first_id=100 * ptype, records=recs))
    json.dump(rk, open(os.path.join(HERE, "refine_kat.json"), "w"), indent=1)
    print("golden vectors written")


if __name__ == "__main__":
    main()
