"""GPU parity of the `-f` post-process: pmvs_neighbor_counts (the PCMVS filter's pair scan,
mvs.cpp:470-499) against the CPU oracle, bit-exact, and `tmvs -f` end to end against the Python restatement."""
import os
import subprocess

import numpy as np
import pytest

import orc
import orc_filters
from filter_case import centers_of, load_state, make_case
from pmvs_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
TMVS = os.path.join(ROOT, "pais-mvs_b200", "bin", "tmvs")


def test_neighbor_counts_known_answers_and_ties():
    c = np.array([[0, 0, 0], [3, 4, 0], [0, 0, 5], [2, 3, 6], [5.0000001, 0, 0], [0, 0, 0]], dtype=np.float64)
    assert api.neighbor_counts(c, 5.0).tolist() == [3, 3, 3, 1, 1, 3]
    assert api.neighbor_counts(c, 7.0).tolist() == orc.neighbor_counts(c, 7.0).tolist()
    assert api.neighbor_counts(c[:1], 1.0).tolist() == [0]
    # integer lattice: thousands of pairs exactly at the radius (3-4-5, 5-12-13 ...), inside the kernel's guard band
    g = np.stack(np.meshgrid(np.arange(14.0), np.arange(14.0), np.arange(6.0), indexing="ij"), axis=-1).reshape(-1, 3)
    for r in (5.0, 13.0, 3.0, np.sqrt(50.0), 1e-300, 0.0):
        assert np.array_equal(api.neighbor_counts(g, r), orc.neighbor_counts(g, r)), r
    # the same lattice scaled by an inexact factor: distances land one ulp either side of the radius
    for s in (0.1, 1.0 / 3.0, 1e-7, 3.7e5):
        assert np.array_equal(api.neighbor_counts(g * s, 5.0 * s), orc.neighbor_counts(g * s, 5.0 * s)), s


@pytest.mark.parametrize("n", [2, 255, 256, 257, 5000, 30000])
def test_neighbor_counts_random_and_sharded(n):
    rng = np.random.RandomState(n)
    pts = rng.rand(n, 3) * np.array([2.0, 1.0, 0.2])
    pts[rng.randint(0, n, size=max(1, n // 50))] = pts[0]                       # duplicates
    r = 0.6 * (0.4 / n) ** (1.0 / 3.0) * 3
    want = orc.neighbor_counts(pts, r)
    got = api.neighbor_counts(pts, r)
    assert np.array_equal(got, want) and want.max() > 0
    cut = [0, n // 3, n // 3, (2 * n) // 3 + 1, n]                              # ragged shards incl. an empty one
    parts = [api.neighbor_counts(pts, r, first=a, count=b - a) for a, b in zip(cut[:-1], cut[1:])]
    assert np.array_equal(np.concatenate(parts), want)


def test_neighbor_counts_full_size_properties():
    # 200k patches (a full reconstruction): symmetric relation -> even total; equals a k-d tree count on tie-free data
    from scipy.spatial import cKDTree
    rng = np.random.RandomState(5)
    pts = rng.rand(200000, 3)
    r = 0.012
    got = api.neighbor_counts(pts, r)
    assert int(got.sum()) % 2 == 0
    want = cKDTree(pts).query_ball_point(pts, r, return_length=True) - 1
    assert np.array_equal(got, want)


def test_tmvs_filter_end_to_end(tmp_path, small_scene):
    cfg, sc = small_scene
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "pais-mvs_b200", "host")])
    d = str(tmp_path)
    path, _ = make_case(d, cfg, sc, n_base=400, seed=23)
    r = subprocess.run([TMVS, "-f", path, "--config", os.path.join(d, "config.txt"), "--out-dir", d], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    fcfg, cameras, patches = load_state(path, sc)
    radius = orc_filters.neighbor_radius(patches, fcfg.neighborRadiusScalar)
    maps = orc_filters.CellMaps(cameras, fcfg.cellSize, patches)
    deleted = []
    orc_filters.cell_filtering(patches, maps, deleted)
    orc_filters.visibility_filtering(patches, maps, cameras, fcfg.minCamNum, deleted)
    orc_filters.neighbor_cell_filtering(patches, maps, radius, 0.25, deleted)
    pc_deleted = []
    avg = orc_filters.neighbor_patch_filtering(patches, radius, 0.25, orc.neighbor_counts, pc_deleted, maps)
    all_centers = centers_of(path)
    assert len(pc_deleted) > 0 and len(patches) > 0
    assert np.array_equal(centers_of(os.path.join(d, "PCMVS_filter.mvs")), all_centers[sorted(patches)])
    assert np.array_equal(centers_of(os.path.join(d, "PCMVS_filter_deleted.mvs")), all_centers[pc_deleted])
    assert np.array_equal(centers_of(os.path.join(d, "PMVS_filter_deleted.mvs")), all_centers[deleted])
    assert ("average neighbor number: %f" % avg) in r.stdout
