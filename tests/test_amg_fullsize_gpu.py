"""AMG at a size the CPU oracle does not reach in seconds (1 M DoF synthetic unstructured triangulation, BASELINE
config 5 shape): size-independent properties instead of element-wise comparison with the oracle --
fast kernels (SELL-32, chunked, streaming loads) against the exact-order kernels (which ARE pinned to the reference bit
for bit on the small meshes), linearity of the smoothers, adjointness of restriction and prolongation, monotone
convergence of the correction-scheme cycle, and the persistent tail / CUDA graph leaving the iterates unchanged."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
LEVELS = 6


@pytest.fixture(scope="module")
def system():
    from amg_bench import synthetic_system
    A, rhs = synthetic_system(1001)
    return A.tocsr(), rhs


def _amg(system, **kw):
    from multigrid_prj_b200 import Amg
    A, rhs = system
    kw.setdefault("levels", LEVELS)
    return Amg(A.indptr, A.indices, A.data, rhs, **kw)


def test_fast_kernels_agree_with_exact_order_kernels(system):
    from multigrid_prj_b200 import amg as M
    A, rhs = system
    n = A.shape[0]
    x0 = np.random.default_rng(8).standard_normal(n)
    got = {}
    for fast in (False, True):
        with _amg(system, fast=fast, smoother=M.GS_MULTICOLOUR) as a:
            out = {}
            a.set_vector(0, 0, x0)
            out["res_norm"] = a.residual(0)
            out["res"] = a.vector(0, 2)
            a.restrict(1)
            out["restrict"] = a.vector(1, 0)
            a.prolong(0)
            out["prolong"] = a.vector(0, 0)
            a.set_vector(0, 0, x0); a.smooth(0, M.JACOBI, 2)
            out["jacobi"] = a.vector(0, 0)
            a.set_vector(0, 0, x0); a.smooth(0, M.GS_MULTICOLOUR, 2)
            out["gs"] = a.vector(0, 0)
            got[fast] = out
    # against scipy for the order-free operators
    r = rhs - A @ x0
    assert abs(got[True]["res_norm"] - np.linalg.norm(r)) <= 1e-12 * np.linalg.norm(r)
    assert np.allclose(got[True]["res"], r, rtol=1e-11, atol=1e-12 * np.abs(r).max())
    for k in ("res", "restrict", "prolong", "jacobi", "gs"):
        u, v = got[True][k], got[False][k]
        assert np.allclose(u, v, rtol=1e-11, atol=1e-12 * np.abs(v).max()), k
    assert abs(got[True]["res_norm"] - got[False]["res_norm"]) <= 1e-12 * got[False]["res_norm"]


def test_restriction_is_the_transpose_of_prolongation(system):
    """<R x, y> = <x, P y>: R is stored explicitly as P^T (atomics-free gather), this pins the two copies to each other"""
    A, rhs = system
    n = A.shape[0]
    rng = np.random.default_rng(9)
    with _amg(system, fast=True, levels=2) as a:
        nc = a.info(1)["n"]
        x, y = rng.standard_normal(n), rng.standard_normal(nc)
        a.set_vector(0, 0, x)
        a.restrict(1)
        Rx = a.vector(1, 0)
        a.set_vector(0, 0, np.zeros(n)); a.set_vector(1, 0, y)
        a.prolong(0)
        Py = a.vector(0, 0)
    lhs, rhs_ = float(Rx @ y), float(x @ Py)
    assert abs(lhs - rhs_) <= 1e-11 * max(abs(lhs), np.linalg.norm(x) * np.linalg.norm(Py))


@pytest.mark.parametrize("kind", ["jacobi", "gs"])
def test_smoothers_are_affine(system, kind):
    """S(x) = M x + c: S(a u + (1 - a) v) = a S(u) + (1 - a) S(v) for any a"""
    from multigrid_prj_b200 import amg as M
    A, rhs = system
    n = A.shape[0]
    rng = np.random.default_rng(10)
    u, v, al = rng.standard_normal(n), rng.standard_normal(n), 0.3
    k = M.JACOBI if kind == "jacobi" else M.GS_MULTICOLOUR
    with _amg(system, fast=True, levels=2, jacobi_omega=0.8) as a:
        def S(x):
            a.set_vector(0, 0, x); a.smooth(0, k, 2)
            return a.vector(0, 0)
        lhs = S(al * u + (1 - al) * v)
        rhs_ = al * S(u) + (1 - al) * S(v)
    assert np.allclose(lhs, rhs_, rtol=1e-10, atol=1e-11 * np.abs(rhs_).max())


def test_cycle_converges_and_is_unchanged_by_tail_and_graph(system):
    A, rhs = system
    hists, xs = [], []
    for tail, graph in ((-1, -1), (0, 0)):          # one launch per operator / persistent tail + CUDA graph (defaults)
        with _amg(system, fast=True, tail_max_rows=tail, cycle_graph=graph) as a:
            hists.append(a.solve(tol=0.0, maxit=8))
            xs.append(a.vector(0, 0))
            if tail == 0:
                assert a.stats()["graph_launches"] == 8
    h0, h1 = hists
    assert np.all(h0[1:] < h0[:-1]) and h0[-1] < 0.5 * h0[0], h0
    assert np.allclose(h0, h1, rtol=1e-9)
    assert np.allclose(xs[0], xs[1], rtol=1e-10, atol=1e-11 * np.abs(xs[0]).max())
    r = rhs - A @ xs[1]
    assert abs(np.linalg.norm(r) - h1[-1]) <= 1e-9 * h1[-1]
