"""The AMG setup stages of the library (host side, O(nnz)) against the stored outputs of the reference's own classes
(tests/golden/amg_*.npz), WITHOUT a GPU: C/F splitting, direct interpolation and the Galerkin product are exported one
by one (mgb_amg_select_coarse_nodes / mgb_amg_build_prolongation / mgb_amg_build_coarse_matrix, the calls the reference's
debugtest.cpp makes on RestrictionOperator, AMG/include/AMG.hpp:150-369) and must reproduce every P_l and A_{l+1} of the
reference bit for bit.  (The same comparison through mgb_amg_create_from_csr needs a device: tests/test_amg_gpu.py.)"""
import ctypes as C

import numpy as np
import pytest

import oracle
from amg_fixtures import load_case


def _handle(lib, m):
    from multigrid_prj_b200._lib import check
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    ptr, col, val = np.asarray(m.ptr, np.int64), np.asarray(m.col, np.int64), np.asarray(m.val, np.float64)
    h = C.c_void_p()
    check(lib.mgb_csr_create(m.n_rows, m.n_cols, p(ptr), p(col), p(val), C.byref(h)))
    return h


def _fetch(lib, h):
    from multigrid_prj_b200._lib import check
    nr, nc, nz = C.c_size_t(), C.c_size_t(), C.c_size_t()
    check(lib.mgb_csr_info(h, C.byref(nr), C.byref(nc), C.byref(nz)))
    ptr, col, val = np.zeros(nr.value + 1, np.int64), np.zeros(max(nz.value, 1), np.int64), np.zeros(max(nz.value, 1))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib.mgb_csr_get(h, p(ptr), p(col), p(val)))
    return oracle.Csr(nr.value, nc.value, ptr, col[:nz.value], val[:nz.value])


@pytest.mark.parametrize("name", ["mesh2", "mesh_pipe", "mesh1"])
def test_setup_stages_reproduce_the_reference_hierarchy(name):
    from multigrid_prj_b200 import load
    from multigrid_prj_b200._lib import check
    lib = load()
    c = load_case(name)
    hA = _handle(lib, c["A"][0])
    for l in range(c["levels"] - 1):
        n = c["A"][l].n_rows
        mask = np.zeros(n, np.uint8)
        ncoarse = C.c_size_t()
        check(lib.mgb_amg_select_coarse_nodes(hA, 0.2, -1, mask.ctypes.data_as(C.c_void_p), C.byref(ncoarse)))
        assert ncoarse.value == c["A"][l + 1].n_rows, f"level {l}: coarse count"
        hP = C.c_void_p()
        check(lib.mgb_amg_build_prolongation(hA, 0.2, mask.ctypes.data_as(C.c_void_p), C.byref(hP)))
        assert _fetch(lib, hP).same_as(c["P"][l]), f"P{l}"
        hC = C.c_void_p()
        check(lib.mgb_amg_build_coarse_matrix(hA, hP, C.byref(hC)))
        assert _fetch(lib, hC).same_as(c["A"][l + 1]), f"A{l + 1}"
        lib.mgb_csr_destroy(hA); lib.mgb_csr_destroy(hP)
        hA = hC
    lib.mgb_csr_destroy(hA)


def test_setup_stage_errors():
    from multigrid_prj_b200 import load
    lib = load()
    n = C.c_size_t()
    assert lib.mgb_amg_select_coarse_nodes(None, 0.2, -1, None, C.byref(n)) != 0
    assert b"null" in lib.mgb_last_error()
    assert lib.mgb_amg_build_coarse_matrix(None, None, None) != 0
