"""Device-side P1 assembly (mgb_fem_assemble_p1 / mgb_fem_synthetic) against the reference's own
assembly: the stored level-0 system of mesh1.msh (tests/golden/amg_mesh1.npz, written by the compiled reference) and, for
the synthetic triangulation of config 5, an independent scipy assembly of the same mesh."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from amg_fixtures import load_case
from msh_reader import read_msh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mesh1_matrix_bit_identical_to_the_reference():
    from multigrid_prj_b200 import System
    x, y, bnd, tri = read_msh(os.path.join(ROOT, "tests", "golden", "mesh", "mesh1.msh"))
    c = load_case("mesh1")
    A = c["A"][0]
    with System.assemble_p1(x, y, bnd, tri, exact_order=True) as s:
        ptr, col, val, rhs = s.get()
    assert ptr.size - 1 == A.n_rows == 6241
    assert np.array_equal(ptr, A.ptr) and np.array_equal(col, A.col)
    assert np.array_equal(val, A.val), f"max rel diff {np.max(np.abs(val - A.val) / np.abs(A.val)):.3e}"
    # the load vector goes through sin / cos / sqrt (device libm vs glibc): last bits only
    assert np.allclose(rhs, c["rhs"][0], rtol=1e-12, atol=1e-13 * np.abs(c["rhs"][0]).max())
    with System.assemble_p1(x, y, bnd, tri, exact_order=False) as s:
        _, _, val2, rhs2 = s.get()
    assert np.allclose(val2, A.val, rtol=1e-13, atol=0) and np.allclose(rhs2, c["rhs"][0], rtol=1e-12, atol=1e-13 * np.abs(rhs).max())


def scipy_p1(x, y, bnd, tri):
    """textbook P1 assembly with the reference's scaling (weights 2*area/3: A and b are 2x the usual values)"""
    px, py = x[tri], y[tri]
    det = (px[:, 1] - px[:, 0]) * (py[:, 2] - py[:, 0]) - (px[:, 2] - px[:, 0]) * (py[:, 1] - py[:, 0])
    area2 = np.abs(det)
    gx = np.stack([py[:, 1] - py[:, 2], py[:, 2] - py[:, 0], py[:, 0] - py[:, 1]], 1) / det[:, None]
    gy = np.stack([px[:, 2] - px[:, 1], px[:, 0] - px[:, 2], px[:, 1] - px[:, 0]], 1) / det[:, None]
    K = (gx[:, :, None] * gx[:, None, :] + gy[:, :, None] * gy[:, None, :]) * area2[:, None, None]
    interior = bnd == 0
    dof = np.full(x.size, -1, np.int64)
    dof[interior] = np.arange(interior.sum())
    rows = np.repeat(tri[:, :, None], 3, 2).ravel(); cols = np.repeat(tri[:, None, :], 3, 1).ravel(); vals = K.ravel()
    n = int(interior.sum())
    keep = interior[rows] & interior[cols]
    A = sp.coo_matrix((vals[keep], (dof[rows[keep]], dof[cols[keep]])), shape=(n, n)).tocsr()
    A.sum_duplicates(); A.sort_indices()
    r = np.sqrt(x * x + y * y)
    g = np.sin(5 * r)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(r > 0, -5 * (np.cos(5 * r) / r - 5 * np.sin(5 * r)), 0.0)
    rhs = np.zeros(n)
    lift = interior[rows] & ~interior[cols]
    np.add.at(rhs, dof[rows[lift]], -vals[lift] * g[cols[lift]])
    w = np.repeat(area2 / 3.0, 3)
    tn = tri.ravel()
    ok = interior[tn]
    np.add.at(rhs, dof[tn[ok]], f[tn[ok]] * w[ok])
    return A, rhs


@pytest.mark.parametrize("side", [5, 66, 257])
def test_synthetic_system_equals_scipy_assembly(side):
    from multigrid_prj_b200 import System
    x, y, bnd, tri = System.synthetic_mesh(side)
    assert bnd.sum() == 4 * side - 4 and tri.shape == (2 * (side - 1) ** 2, 3)
    h = 2.0 / (side - 1)
    jj, ii = np.meshgrid(np.arange(side), np.arange(side))
    assert np.abs(x - (jj * h).ravel()).max() <= 0.2 * h + 1e-15 and np.abs(y - (ii * h).ravel()).max() <= 0.2 * h + 1e-15
    A, rhs = scipy_p1(x, y, bnd, tri)
    with System.synthetic(side) as s:
        n, nnz = s.info()
        ptr, col, val, b = s.get()
    assert n == (side - 2) ** 2
    D = sp.csr_matrix((val, col, ptr), shape=(n, n))
    # (entries that cancel to exactly 0 in one assembly and to 1e-17 in the other are compared by value, not by pattern)
    # two different formulas for the element gradients (the reference's normal-vector form on the device, the textbook form in
    # scipy_p1) on jittered elements: differences of nearby coordinates cost log2(side) bits, so the bar scales with side
    tol = 1e-14 * side * side
    diff = abs(D - A)
    assert diff.max() <= tol * abs(A).max(), diff.max() / abs(A).max()
    assert np.allclose(b, rhs, rtol=0, atol=tol * np.abs(rhs).max())
    # both diagonals occur: the triangulation is unstructured
    assert 5.5 < nnz / n < 7.2 or side < 10
