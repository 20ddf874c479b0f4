"""The device-side AMG setup (csrc/amg_setup.cu; SURVEY.md section 8f item 1: parity-EXEMPT, validated by properties):
strength / PMIS splitting / direct interpolation / Galerkin product built on the GPU, checked on the host with scipy, and
the cycle on that hierarchy against the cycle on the reference's (host-built, bit-exact) hierarchy."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
EPS = 0.2


@pytest.fixture(scope="module")
def system():
    from multigrid_prj_b200 import System
    s = System.synthetic(301)
    ptr, col, val, rhs = s.get()
    n = ptr.size - 1
    yield s, sp.csr_matrix((val, col, ptr), shape=(n, n)), rhs
    s.close()


def level_matrices(a):
    out = []
    for l in range(a.levels):
        i = a.info(l)
        ptr, col, val = a.matrix(l, 0)
        A = sp.csr_matrix((val, col, ptr), shape=(i["n"], i["n"]))
        P = None
        if l + 1 < a.levels:
            pp, pc, pv = a.matrix(l, 1)
            P = sp.csr_matrix((pv, pc, pp), shape=(i["n"], i["n_coarse"]))
        out.append((A, P))
    return out


def strong_graph(A):
    """S[i, j] = 1 iff j != i and |a_ij| >= eps * max_{k != i} |a_ik|   (AMG/include/AMG.hpp:105-130)"""
    B = abs(A).tolil(); B.setdiag(0); B = B.tocsr(); B.eliminate_zeros()
    big = B.max(axis=1).toarray().ravel()
    C = B.tocoo()
    keep = C.data >= EPS * big[C.row]
    return sp.csr_matrix((np.ones(keep.sum()), (C.row[keep], C.col[keep])), shape=A.shape)


def test_hierarchy_properties(system):
    from multigrid_prj_b200 import Amg
    s, A0, rhs = system
    with Amg.from_system(s, levels=8) as a:
        assert 3 <= a.levels <= 8
        lv = level_matrices(a)
        assert abs(lv[0][0] - A0).max() == 0
        sizes = [m[0].shape[0] for m in lv]
        assert all(sizes[i + 1] < 0.6 * sizes[i] for i in range(len(sizes) - 1)), sizes
        for l, (A, P) in enumerate(lv[:-1]):
            n, nc = P.shape
            S = strong_graph(A)
            rows_len = np.diff(P.indptr)
            is_c = a.schedule(l, 2) == 1
            # unit rows on coarse points, numbered in ascending fine index (AMG.hpp:201-228)
            c_rows = np.flatnonzero(is_c)
            assert c_rows.size == nc and np.array_equal(P.indices[P.indptr[c_rows]], np.arange(nc))
            assert np.all(rows_len[c_rows] == 1) and np.all(P.data[P.indptr[c_rows]] == 1.0)
            # PMIS: no two C points depend strongly on EACH OTHER (a point that depends on a new C point turns fine at once; a
            # one-sided coupling may leave both coarse); every coupled F point depends strongly on a C point
            Ssym = ((S + S.T) > 0).astype(np.float64).tocsr()
            Smut = S.multiply(S.T).tocsr()
            cvec = is_c.astype(np.float64)
            assert (Smut @ cvec)[is_c].max() == 0, f"level {l}: mutually coupled coarse points"
            coupled = np.asarray(Ssym.sum(axis=1)).ravel() > 0
            f_rows = ~is_c & coupled
            assert ((S @ cvec)[f_rows] > 0).all(), f"level {l}: fine point without a strong coarse neighbour"
            # direct interpolation: weights a_ij / sum_{k in S_i cap C} a_ik over the strong coarse neighbours (AMG.hpp:230-300)
            assert np.allclose(np.asarray(P.sum(axis=1)).ravel()[f_rows], 1.0, rtol=1e-12)
            f0 = np.flatnonzero(f_rows)[:200]
            cidx = np.cumsum(is_c) - 1
            for i in f0:
                js = S.indices[S.indptr[i]:S.indptr[i + 1]]
                js = js[is_c[js]]
                w = np.array([A[i, j] for j in js]); w = w / w.sum()
                assert np.array_equal(P.indices[P.indptr[i]:P.indptr[i + 1]], cidx[js]) and np.allclose(P.data[P.indptr[i]:P.indptr[i + 1]], w, rtol=1e-13)
            # Galerkin operator (AMG.hpp:303-369)
            Ac = (P.T @ A @ P).tocsr()
            got = lv[l + 1][0]
            assert abs(got - Ac).max() <= 1e-12 * abs(Ac).max(), f"level {l + 1}"
            assert np.all(np.diff(got.indices)[np.setdiff1d(np.arange(got.nnz - 1), got.indptr[1:-1] - 1)] > 0), "rows sorted by column"
            # restricted right-hand side b_c = P^T b (AMG.cpp:100-109)
            assert np.allclose(a.vector(l + 1, 1), P.T @ a.vector(l, 1), rtol=1e-12, atol=1e-13 * np.abs(a.vector(l, 1)).max())


def test_cycle_converges_at_least_as_fast_as_on_the_reference_hierarchy(system):
    from multigrid_prj_b200 import Amg
    from multigrid_prj_b200 import amg as M
    s, A0, rhs = system
    K = 60
    with Amg.from_system(s, levels=8) as a:
        hist = a.solve(tol=0.0, maxit=K)
        x = a.vector(0, 0)
    assert abs(np.linalg.norm(rhs - A0 @ x) - hist[-1]) <= 1e-9 * hist[0]          # the reported norm is the true residual
    assert np.all(np.diff(hist) < 0)                                               # monotone
    with Amg(A0.indptr, A0.indices, A0.data, rhs, levels=8, fast=True) as h:      # the reference's hierarchy, multicolour GS
        hist_h = h.solve(tol=0.0, maxit=K)
    # asymptotic factor: the last 20 cycles
    rate_dev = (hist[-1] / hist[-21]) ** (1.0 / 20)
    rate_host = (hist_h[-1] / hist_h[-21]) ** (1.0 / 20)
    print(f"reduction per V(2,2) cycle (asymptotic): device hierarchy {rate_dev:.3f}, reference hierarchy {rate_host:.3f}; "
          f"after {K} cycles {hist[-1] / hist[0]:.2e} vs {hist_h[-1] / hist_h[0]:.2e}")
    print("device hierarchy, first cycles:", np.array2string(hist[1:9] / hist[:8], precision=3))
    print("reference hierarchy, first cycles:", np.array2string(hist_h[1:9] / hist_h[:8], precision=3))
    assert rate_dev <= 1.1 * rate_host and hist[-1] <= 1.1 * hist_h[-1]


def test_same_hierarchy_from_host_csr_and_from_device_system(system):
    from multigrid_prj_b200 import Amg
    s, A0, rhs = system
    with Amg.from_system(s, levels=6) as a, Amg(A0.indptr, A0.indices, A0.data, rhs, levels=6, device_path=True) as b:
        assert a.levels == b.levels
        for l in range(a.levels):
            assert a.info(l) == b.info(l)
        ha, hb = a.solve(tol=0.0, maxit=5), b.solve(tol=0.0, maxit=5)
        assert np.array_equal(a.vector(0, 0), b.vector(0, 0)) and np.array_equal(ha, hb)


def test_l1_jacobi_and_multicolour_sweeps_on_the_device_built_level(system):
    from multigrid_prj_b200 import Amg
    from multigrid_prj_b200 import amg as M
    s, A0, rhs = system
    n = A0.shape[0]
    x0 = np.random.default_rng(3).standard_normal(n)
    with Amg.from_system(s, levels=3) as a:
        d = A0.diagonal()
        dl1 = np.asarray(abs(A0).sum(axis=1)).ravel()            # a_ii + sum_{j != i} |a_ij| (a_ii > 0)
        x = x0.copy()
        for sweeps in (1, 2, 3):
            a.set_vector(0, 0, x0); a.smooth(0, M.L1_JACOBI, sweeps)
            x = x0.copy()
            for _ in range(sweeps):
                x = x + (rhs - A0 @ x) / dl1
            assert np.allclose(a.vector(0, 0), x, rtol=1e-11, atol=1e-12 * np.abs(x).max()), sweeps
        # multicolour GS on level 0: the colouring is valid and the sweep equals the colour-by-colour replay
        colour = a.schedule(0, 1)
        C = A0.tocoo()
        off = C.row != C.col
        assert not np.any(colour[C.row[off]] == colour[C.col[off]])
        a.set_vector(0, 0, x0); a.smooth(0, M.GS_MULTICOLOUR, 1)
        x = x0.copy()
        for c in range(colour.max() + 1):
            rows = np.flatnonzero(colour == c)
            x[rows] = x[rows] + (rhs[rows] - (A0[rows] @ x)) / d[rows]
        assert np.allclose(a.vector(0, 0), x, rtol=1e-11, atol=1e-12 * np.abs(x).max())
        # levels >= 1 carry no colouring on this path: asking for multicolour GS there fails loudly
        from multigrid_prj_b200 import MgbError
        with pytest.raises(MgbError, match="without a colouring"):
            a.smooth(1, M.GS_MULTICOLOUR, 1)
