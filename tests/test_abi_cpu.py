"""CPU tests of the drop-in boundary: the C-ABI library builds/loads, exports every symbol the header declares, the
ctypes mirrors match the C layouts, and compute entry points fail loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from pmvs_b200 import abi, lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "pmvs_b200.h")


@pytest.fixture(scope="module")
def library():
    lib.build()
    return lib.load()


def test_exports_every_declared_symbol(library):
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(pmvs_[a-z_0-9]+)\s*\(", src))
    assert declared == set(lib.SYMBOLS)
    for name in declared:
        assert hasattr(library, name), name
    assert b"sm_100a" in library.pmvs_version()


def test_struct_layouts_match_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "pmvs_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(PmvsConfig), sizeof(PmvsLevel), sizeof(PmvsCamera), sizeof(PmvsHypothesis), sizeof(PmvsPatchIn), sizeof(PmvsPatchOut));
  printf("%zu %zu %zu %zu %zu\n", offsetof(PmvsConfig, reduceNormalRange), offsetof(PmvsConfig, particleNum), offsetof(PmvsCamera, level), offsetof(PmvsPatchOut, camIdx), offsetof(PmvsPatchOut, imgPoint));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    got = [int(v) for v in out]
    want = [C.sizeof(abi.PmvsConfig), C.sizeof(abi.PmvsLevel), C.sizeof(abi.PmvsCamera), C.sizeof(abi.PmvsHypothesis),
            C.sizeof(abi.PmvsPatchIn), C.sizeof(abi.PmvsPatchOut), abi.PmvsConfig.reduceNormalRange.offset,
            abi.PmvsConfig.particleNum.offset, abi.PmvsCamera.level.offset, abi.PmvsPatchOut.camIdx.offset,
            abi.PmvsPatchOut.imgPoint.offset]
    assert got == want
    assert got[0] == 160          # MvsConfig as dumped into MVS_V3 files (TMVS/io/filewriter.cpp:71-102)


def test_no_cpu_fallback(library, small_scene):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pmvs_b200.api import PatchRefiner
    cfg, sc = small_scene
    with pytest.raises(lib.PmvsError) as e:
        PatchRefiner(cfg, sc.records)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under pais-mvs_b200/ may reference it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "pais-mvs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"liborc|orc\.py|import orc|oracle/|pmvs_oracle", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
