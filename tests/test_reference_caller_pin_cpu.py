"""CPU test: the caller side of the hot path against the UNMODIFIED reference.

MVS::expansionPatches (TMVS/mvs/mvs.cpp:233-275) runs unmodified inside oracle/_ref/libtmvs_ref.so with everything under it —
expandNeighborCell / expandCell (:529-577), getExpansionPatchCenter (:809-836), skipNeighborCell (:792-807), runtimeFiltering
(:838-898), insertPatch / deletePatch (:579-634), the queue pops (:656-788), the cell maps, setNeighborRadius (:147-152) — each
expansion patch refined by the unmodified Patch::refine() (patch.cpp). The same reconstruction is replayed with
oracle/orc_host.py's restatement of that loop around the f64 restatement of refine() (oracle/liborc.so): the final patch
containers must be identical — same ids, same cameras, every value bit for bit — for every expansion strategy. This pins
orc_host.py, which tests/test_host_parity_cpu.py and tests/test_expansion_cpu.py use as the oracle of the C++ host driver."""
import ctypes as C
import math

import pytest

import orc
import orc_host as oh
import ref_tmvs
from pmvs_b200 import abi, scene

pytestmark = pytest.mark.skipif(ref_tmvs.lib() is None, reason="oracle/_ref/libtmvs_ref.so not built (no /root/reference on this box)")


def oh_cameras(sc):
    cams = []
    for c in sc.cams:
        g = c.levels[0][0]
        oc = oh.Camera(float(c.focal[0]), list(c.quaternion), list(c.center), g.shape[1], g.shape[0], g)
        oc.focal, oc.principal = (float(c.focal[0]), float(c.focal[1])), (float(c.principal[0]), float(c.principal[1]))
        oc.R, oc.t = [float(v) for v in c.R.flatten()], [float(v) for v in c.t]
        oc.optical_normal, oc.center = [float(v) for v in c.optical_normal], [float(v) for v in c.center]
        cams.append(oc)
    return cams


def to_oh(q, pid):
    return oh.Patch(pid, list(q.center), list(q.normal), q.fitness, q.priority, q.correlation, list(q.camIdx[:q.nCam]),
                    [(q.imgPoint[k][0], q.imgPoint[k][1]) for k in range(q.nImgPoint)], drop=bool(q.drop))


@pytest.mark.parametrize("strategy", [oh.BEST_FIRST, oh.WORST_FIRST, oh.BREATH_FIRST, oh.DEPTH_FIRST])
def test_expansion_loop_identical_to_unmodified_reference(strategy):
    cfg = abi.readme_config()
    cfg.patchRadius, cfg.patchSize, cfg.distWeighting, cfg.maxLOD = 4, 9, 4 / 3.0, 1
    cfg.cellSize, cfg.maxCellPatchNum, cfg.expansionStrategy = 10, 2, strategy
    cfg.particleNum, cfg.maxIteration = 8, 12
    sc = scene.SynthScene(cfg, nviews=4, width=120, height=90, seed=31, with_edge=True, tex_size=512)
    seeds = sc.patches(4, seed=9, ptype=abi.TYPE_SEED, extent=0.25 * sc.distance * 90 / sc.focal)
    ref = ref_tmvs.RefScene(cfg, sc.records, seed=42)
    want = ref.run_reconstruction(seeds)
    assert len(want) > len(seeds) + 20, len(want)          # the expansion really ran

    # the same reconstruction with the restated caller (orc_host.py) around the restated refine() (liborc)
    o = orc.Oracle(cfg, sc.records, seed=42, use_ref_pso=False)
    mvs = oh.MVS(cfg, oh_cameras(sc))
    for s in seeds:                                          # seeds enter unrefined (mvs.cpp:196-231 below)
        mvs.patches[s.id] = oh.Patch(s.id, list(s.center), list(s.normal), cam_idx=list(s.camIdx[:s.nCam]))
    o.set_neighbor_radius(mvs.set_neighbor_radius())         # :202
    flags = abi.F_POST_REMOVE_INVISIBLE
    for pid in sorted(mvs.patches):
        if len(mvs.patches[pid].cam_idx) < cfg.minCamNum:
            mvs.delete_patch(pid)
            continue
        q = o.refine_batch((abi.PmvsPatchIn * 1)(seeds[pid]), flags=flags)[0]
        mvs.patches[pid] = to_oh(q, pid)
        if not mvs.runtime_filtering(mvs.patches[pid]):
            mvs.delete_patch(pid)
    mvs.set_neighbor_radius()
    next_id = [len(seeds)]

    def refine(center, parent):                              # Patch(center, parent) + refine() + removeInvisibleCamera(), mvs.cpp:566-577
        o.set_neighbor_radius(mvs.neighbor_radius)
        pin = abi.PmvsPatchIn()
        pin.center[:] = center
        pin.normal[:] = parent.normal
        pin.normalS[:] = ref.normal2spherical(parent.normal)     # setNormal(Vec3d), abstractpatch.cpp:42-45
        pin.type, pin.id, pin.nCam = abi.TYPE_EXPAND, next_id[0], len(parent.cam_idx)
        pin.camIdx[:pin.nCam] = parent.cam_idx
        next_id[0] += 1
        return to_oh(o.refine_batch((abi.PmvsPatchIn * 1)(pin), flags=flags | abi.F_EXPAND_VISIBLE)[0], pin.id)

    oh.expansion_patches(mvs, refine, reference_loop_exit=True)
    assert mvs.neighbor_radius.hex() == ref.neighbor_radius().hex()
    got = [mvs.patches[k] for k in sorted(mvs.patches)]
    assert [p.id for p in got] == [w["id"] for w in want]
    for p, w in zip(got, want):
        assert p.cam_idx == w["cam_idx"], p.id
        assert [v.hex() for v in p.center + p.normal + [p.fitness, p.priority, p.correlation]] == \
               [v.hex() for v in w["center"] + w["normal"] + [w["fitness"], w["priority"], w["correlation"]]], p.id
        assert [(a.hex(), b.hex()) for a, b in p.img_point] == [(a.hex(), b.hex()) for a, b in w["img_point"]], p.id
        assert p.expanded == w["expanded"], p.id
