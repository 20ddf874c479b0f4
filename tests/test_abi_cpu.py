"""CPU-side checks of the C ABI: the shared library loads without a GPU, exports every symbol that
include/mgb200.h declares, the ctypes table covers exactly that set, and compute entry points fail
loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from multigrid_prj_b200 import _lib, load
    lib = load()
    names = declared_functions()
    assert len(names) > 50
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/mgb200.h but not exported by libmgb200.so"
    assert sorted(_lib.SYMBOLS) == names, set(names) ^ set(_lib.SYMBOLS)


def test_struct_layouts_match_header():
    """the ctypes mirrors must have the size the C compiler gives the structs"""
    import subprocess
    import tempfile
    from multigrid_prj_b200 import _lib
    code = ('#include "mgb200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(mgb_gmg_config), '
            'sizeof(mgb_amg_config), sizeof(mgb_gmg_stats));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(code)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(_lib.GmgConfigStruct), C.sizeof(_lib.AmgConfigStruct), C.sizeof(_lib.GmgStatsStruct)]


def test_no_cpu_fallback_without_a_device():
    from multigrid_prj_b200 import Amg, Gmg, GmgConfig, MgbError, load
    if load().mgb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(MgbError, match="no CUDA device"):
        Gmg(GmgConfig(n=33, levels=3))
    ptr = np.array([0, 1, 2], np.int64); col = np.array([0, 1], np.int64); val = np.ones(2)
    with pytest.raises(MgbError, match="no CUDA device"):
        Amg(ptr, col, val, np.ones(2), levels=1)


def test_argument_validation_needs_no_device():
    from multigrid_prj_b200 import MgbError
    from multigrid_prj_b200.gmg import partition
    assert partition(1025, 10, 1, 0, 3) == (False, 0, 129)
    with pytest.raises(MgbError):
        partition(200, 2, 1, 0, 0)           # the reference's default N=200 violates (N-1) % 2^(L-1) == 0


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under multigrid_prj_b200/ may reference it"""
    pkg = os.path.join(ROOT, "multigrid_prj_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "oracle/" not in txt and "liboracle" not in txt, os.path.join(dirpath, f)
