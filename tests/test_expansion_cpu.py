"""The expansion driver's control flow on a GPU-less box (SURVEY.md 8a caller row, 8f row 2): MVS::expansionPatches of the host
driver with a deterministic stand-in for Patch::refine() (tmvs_hooks.cpp `planeRefine`, restated below), against
oracle/orc_host.py's restatement of the reference's SERIAL loop (mvs.cpp:233-275, :529-577: one parent at a time, every
candidate inserted before the next cell is looked at).

* `--round 1 --slot-passes` must reproduce the serial reference exactly — same patches, bit for bit — for every strategy:
  the generation-time skip test can only become truer as patches are inserted, and the commit re-checks the target cell.
* Larger rounds and the merged slot pass change the visiting order, not the outcome class: same coverage to a few per cent
  (this scene rejects one candidate in seven — 3 % background pixels in each of five views — so the retry of a parent
  whose expectation failed matters: without it the merged pass lost 6 % of the cells), no duplicate patches, fewer /
  larger refine calls, and no more refinements than the slot passes need."""
import ctypes as C
import math

import numpy as np
import pytest

import orc_host as oh
from test_host_parity_cpu import D3, Pair, cfg, hooks  # noqa: F401  (fixtures)

PLANE_Z = 0.0


def plane_refine(o, center, parent, pid):
    """tmvs_hooks.cpp planeRefine, same expression order."""
    Cc = o.cameras[parent.cam_idx[0]].center
    t = (PLANE_Z - Cc[2]) / (center[2] - Cc[2])
    c = [Cc[k] + t * (center[k] - Cc[k]) for k in range(3)]
    pts, drop = [], False
    for cam in o.cameras:
        (u, v), inside = cam.project(c, 0, o.cfg.lodRatio)
        pts.append((u, v))
        drop = drop or not inside
    if drop:
        return oh.Patch(pid, c, [0.0, 0.0, -1.0], oh.DBL_MAX, oh.DBL_MAX, 0.95, range(len(o.cameras)), [], drop=True)
    return oh.Patch(pid, c, [0.0, 0.0, -1.0], 1.0, 1.0 + (c[0] * 0.37 + c[1] * 0.11), 0.95, range(len(o.cameras)), pts)


def seeded_pair(L, cfg, seed, n_seeds=6, strategy=oh.BEST_FIRST):
    cfg.expansionStrategy = strategy
    cfg.maxCellPatchNum = 2
    P = Pair(L, cfg, cols=160, rows=120, seed=seed)
    rng = np.random.RandomState(seed)
    for pid in range(n_seeds):
        c = [0.8 * (2 * rng.rand() - 1), 0.6 * (2 * rng.rand() - 1), PLANE_Z]
        pts = [cam.project(c, 0, cfg.lodRatio)[0] for cam in P.o.cameras]
        P.put(oh.Patch(pid, c, [0.0, 0.0, -1.0], 1.0, 1.0 + 0.1 * pid, 0.95, range(len(P.o.cameras)), pts))
    return P


def host_expand(P, round_size, merge):
    L = P.L
    L.tmvs_hook_expand_plane.restype = C.c_long
    L.tmvs_hook_expand_plane.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.tmvs_hook_patch_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    refined = C.c_long(0)
    calls = L.tmvs_hook_expand_plane(P.h, PLANE_Z, round_size, int(merge), C.byref(refined))
    assert calls >= 0
    n = L.tmvs_hook_patch_count(P.h)
    ids = (C.c_int * n)()
    L.tmvs_hook_patch_ids(P.h, ids, n)
    c, nr, s = D3(), D3(), (C.c_double * 2)()
    centers = []
    for pid in ids:
        assert L.tmvs_hook_get_patch(P.h, pid, c, nr, s) == 0
        centers.append(tuple(c))
    return centers, calls, refined.value


def oracle_expand(P, **kw):
    counter = [10 ** 6]

    def refine(center, parent):
        counter[0] += 1
        return plane_refine(P.o, center, parent, counter[0])
    oh.expansion_patches(P.o, refine, **kw)
    return [tuple(P.o.patches[k].center) for k in sorted(P.o.patches)], counter[0] - 10 ** 6


def coverage(P, centers):
    cs = P.cfg.cellSize
    return {(int(u / cs), int(v / cs)) for (u, v) in (P.o.cameras[0].project(list(c), 0, P.cfg.lodRatio)[0] for c in centers)}


@pytest.mark.parametrize("strategy", [oh.BEST_FIRST, oh.WORST_FIRST, oh.BREATH_FIRST, oh.DEPTH_FIRST])
def test_round_of_one_is_the_serial_reference(hooks, cfg, strategy):
    P = seeded_pair(hooks, cfg, seed=31 + strategy, strategy=strategy)
    got, calls, refined = host_expand(P, 1, merge=False)
    want, want_refined = oracle_expand(P)
    assert len(want) > 300                                            # the plane was grown from six seeds
    if strategy == oh.DEPTH_FIRST:
        # the reference's backward scan never returns the first queued patch (mvs.cpp:761-788); the driver pops it last:
        # everything up to there is the serial reference, plus at most that one patch's children
        assert set(want) <= set(got) and len(got) - len(want) <= 4 * len(P.o.cameras)
        assert want_refined <= refined <= want_refined + 4 * len(P.o.cameras)
        P.close()
        return
    assert sorted(got) == sorted(want)                                # same patches, bit for bit
    assert refined == want_refined                                    # and not one refinement more than the serial loop
    # the reference's own loop exit leaves the patch popped last unexpanded (mvs.cpp:241-243): at most its children differ
    Q = seeded_pair(hooks, cfg, seed=31 + strategy, strategy=strategy)
    quirk, _ = oracle_expand(Q, reference_loop_exit=True)
    assert set(quirk) <= set(want) and len(want) - len(quirk) <= 4 * len(P.o.cameras)
    P.close()
    Q.close()


def test_rounds_and_merged_pass_grow_the_same_surface(hooks, cfg):
    base = seeded_pair(hooks, cfg, seed=77)
    serial, _ = oracle_expand(base)
    cov0 = coverage(base, serial)
    results = {}
    for rnd, merge in ((16, False), (16, True), (64, True), (1024, True), (64, 2)):      # merge 2: merged pass without pipelined rounds
        P = seeded_pair(hooks, cfg, seed=77)
        got, calls, refined = host_expand(P, rnd, merge)
        results[(rnd, merge)] = (len(got), calls, refined)
        assert len(set(got)) == len(got)
        cov = coverage(P, got)
        print("round %d merged %d: %d patches (serial %d), %d cells covered (serial %d, symmetric difference %d), %d calls, %d refinements"
              % (rnd, merge, len(got), len(serial), len(cov), len(cov0), len(cov ^ cov0), calls, refined))
        assert abs(len(cov) - len(cov0)) <= 0.03 * len(cov0), (rnd, merge, len(cov), len(cov0))
        assert abs(len(got) - len(serial)) <= 0.03 * len(serial)
        P.close()
    # one merged pass per round: fewer calls than one per camera slot, and no more refinements
    assert results[(16, True)][1] < results[(16, False)][1] and results[(16, True)][2] <= results[(16, False)][2]
    assert results[(1024, True)][1] < results[(64, True)][1] < results[(16, True)][1]
    # pipelined rounds generate round k+1 before commit k and prune afterwards: about the work of the unpipelined loop
    assert results[(64, True)][2] <= 1.1 * results[(64, 2)][2]
    base.close()


@pytest.mark.parametrize("merge", [1, 2])
def test_expansion_is_deterministic(hooks, cfg, merge):
    runs = []
    for _ in range(2):
        P = seeded_pair(hooks, cfg, seed=5)
        runs.append(host_expand(P, 64, merge))
        P.close()
    assert runs[0] == runs[1]
