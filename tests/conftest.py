import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a GPU-less box: they are skipped with a reason there, and on the GPU box the
    # library refuses to run without its CUDA path (no fallback exists).
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def small_scene():
    """5 views 320x240, r=7, 3 levels, README sample config (README.md:110-207)."""
    from pmvs_b200 import abi, scene
    cfg = abi.readme_config()
    cfg.patchRadius = 7
    cfg.patchSize = 15
    cfg.distWeighting = 7 / 3.0
    cfg.maxLOD = 2
    sc = scene.SynthScene(cfg, nviews=5, width=320, height=240, seed=1234, with_edge=True, tex_size=1024)
    return cfg, sc
