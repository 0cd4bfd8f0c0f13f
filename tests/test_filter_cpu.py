"""CPU tests of the `-f` post-process: the oracle's pair scan against an independent NumPy count,
the host's three PMVS filters (C++) against the Python restatement of mvs.cpp:279-446, and the loud failure of the
PCMVS stage without a CUDA device (there is no CPU path)."""
import os
import subprocess

import numpy as np
import pytest

import orc
import orc_filters
from filter_case import centers_of, load_state, make_case
from pmvs_b200 import mvsio

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
TMVS = os.path.join(ROOT, "pais-mvs_b200", "bin", "tmvs")


def np_counts(c, r):
    d = c[:, None, :] - c[None, :, :]
    s = d[..., 0] * d[..., 0]
    s = s + d[..., 1] * d[..., 1]
    s = s + d[..., 2] * d[..., 2]
    inside = ~(np.sqrt(s) > r)
    np.fill_diagonal(inside, False)
    return inside.sum(axis=1).astype(np.int32)


def test_oracle_neighbor_counts_known_answers():
    # 3-4-5 and 2-3-6-7 lattice distances are exact: points AT the radius count (mvs.cpp:496 breaks on dist > radius)
    c = np.array([[0, 0, 0], [3, 4, 0], [0, 0, 5], [2, 3, 6], [5.0000001, 0, 0], [0, 0, 0]], dtype=np.float64)
    assert orc.neighbor_counts(c, 5.0).tolist() == [3, 3, 3, 1, 1, 3]      # A-B, A-C at exactly 5; A-E just beyond
    assert orc.neighbor_counts(c, 7.0).tolist()[3] == 4                         # D-A at exactly 7
    assert orc.neighbor_counts(c[:1], 1.0).tolist() == [0]
    rng = np.random.RandomState(3)
    for n, r in ((1, 0.1), (2, 0.5), (257, 0.2), (1500, 0.08)):
        pts = rng.rand(n, 3)
        assert np.array_equal(orc.neighbor_counts(pts, r), np_counts(pts, r))


def test_host_pmvs_filters_match_restatement(tmp_path, small_scene):
    cfg, sc = small_scene
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "pais-mvs_b200", "host")])
    d = str(tmp_path)
    path, _ = make_case(d, cfg, sc)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")                     # the PCMVS stage must fail loudly, not fall back
    r = subprocess.run([TMVS, "-f", path, "--config", os.path.join(d, "config.txt"), "--out-dir", d], cwd=d, capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "no CPU path" in r.stderr, (r.returncode, r.stderr)
    assert not os.path.exists(os.path.join(d, "PCMVS_filter.mvs"))

    fcfg, cameras, patches = load_state(path, sc)
    radius = orc_filters.neighbor_radius(patches, fcfg.neighborRadiusScalar)
    maps = orc_filters.CellMaps(cameras, fcfg.cellSize, patches)
    n0 = len(patches)
    stages = []
    deleted = []
    orc_filters.cell_filtering(patches, maps, deleted)
    stages.append(sorted(patches))
    orc_filters.visibility_filtering(patches, maps, cameras, fcfg.minCamNum, deleted)
    stages.append(sorted(patches))
    orc_filters.neighbor_cell_filtering(patches, maps, radius, 0.25, deleted)
    stages.append(sorted(patches))
    assert n0 > len(stages[0]) > len(stages[1]) > len(stages[2]) > 0           # every filter removes something in this case
    all_centers = centers_of(path)
    for k, ids in enumerate(stages):
        got = centers_of(os.path.join(d, "PMVS_filter%d.mvs" % (k + 1)))
        assert np.array_equal(got, all_centers[ids]), "PMVS_filter%d" % (k + 1)
        assert len(mvsio.read_ply(os.path.join(d, "PMVS_filter%d.ply" % (k + 1)))) == len(ids)
    got_del = centers_of(os.path.join(d, "PMVS_filter_deleted.mvs"))
    assert np.array_equal(got_del, all_centers[deleted])                        # same patches, same deletion order
