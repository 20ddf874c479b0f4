"""CPU checks of the checker's OWN additions (oracle/gmg_oracle.c, second half): the textbook cycles, the full-multigrid
pass and the Krylov solvers are not in the reference, so they are pinned against an independent statement -- the 5-point
operator assembled with scipy and solved directly -- and against the properties that define them."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle

W = 10.0


def assembled(N, alpha):
    """A of GeometricMultigrid/include/linear_system.hpp:21-42 on the fine level: identity rows on the boundary"""
    h = W / (N - 1)
    c = alpha / (h * h)
    idx = np.arange(N * N).reshape(N, N)
    inner = idx[1:-1, 1:-1].ravel()
    rows = [inner] * 5
    cols = [inner, inner - N, inner - 1, inner + 1, inner + N]
    vals = [np.full(inner.size, 4 * c)] + [np.full(inner.size, -c)] * 4
    bd = np.setdiff1d(idx.ravel(), inner)
    rows.append(bd); cols.append(bd); vals.append(np.ones(bd.size))
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N * N, N * N))


@pytest.fixture(scope="module")
def orc():
    return oracle.gmg()


@pytest.fixture(scope="module")
def problem(orc):
    N, alpha = 65, 1.0
    b = orc.rhs(N, W, 1)
    return N, alpha, b, spla.spsolve(assembled(N, alpha).tocsc(), b)


@pytest.mark.parametrize("cycle", [1, 2, 3])
@pytest.mark.parametrize("kind,rm", [(oracle.RBGS, 2), (oracle.RBGS, 1), (oracle.GS, 2)])
def test_textbook_cycles_converge_to_the_direct_solution(orc, problem, cycle, kind, rm):
    N, alpha, b, x = problem
    t = orc.textbook(N, W, alpha, 6, kind, cycle, 4, nu_pre=2, nu=2, restrict_mode=rm)
    u = np.zeros(N * N)
    for _ in range(25):
        t.iteration(u, b, pre_kind=kind, n_pre=0)
    assert np.linalg.norm(u - x) <= 1e-9 * np.linalg.norm(x)


def test_w_cycle_contracts_faster_than_v_cycle(orc, problem):
    N, alpha, b, x = problem
    err = {}
    for cycle in (1, 2):
        t = orc.textbook(N, W, alpha, 6, oracle.RBGS, cycle, 5, nu_pre=1, nu=1, restrict_mode=2)
        u = np.zeros(N * N)
        for _ in range(6):
            t.iteration(u, b, pre_kind=oracle.RBGS, n_pre=0)
        err[cycle] = np.linalg.norm(u - x)
    assert err[2] < err[1]


def test_fmg_pass_is_a_good_first_iterate(orc, problem):
    N, alpha, b, x = problem
    t = orc.textbook(N, W, alpha, 6, oracle.RBGS, 1, 4, nu_pre=2, nu=2, restrict_mode=2)
    u = t.fmg(np.zeros(N * N), b)
    v = np.zeros(N * N)
    orc.textbook(N, W, alpha, 6, oracle.RBGS, 1, 4, nu_pre=2, nu=2, restrict_mode=2).iteration(v, b, pre_kind=oracle.RBGS, n_pre=0)
    assert np.linalg.norm(u - x) < 0.2 * np.linalg.norm(v - x)      # one FMG pass beats one V cycle from zero by far
    assert np.linalg.norm(u - x) < 2e-2 * np.linalg.norm(x)


@pytest.mark.parametrize("method,precond,kind,omega,maxit,expect", [
    (0, 0, oracle.JACOBI, 1.0, 400, None),      # plain CG
    (0, 1, oracle.JACOBI, 0.8, 40, 12),         # CG + weighted-Jacobi V(2,2): a symmetric preconditioner
    (1, 0, oracle.RBGS, 1.0, 400, None),        # plain BiCGSTAB
    (1, 1, oracle.RBGS, 1.0, 40, 8),            # BiCGSTAB + red-black V(2,2)
])
def test_krylov_solvers_reach_the_direct_solution(orc, problem, method, precond, kind, omega, maxit, expect):
    N, alpha, b, x = problem
    L = 6
    t = orc.textbook(N, W, alpha, L, kind, 1, L - 1, nu_pre=2, nu=2, restrict_mode=2, coarse_tol=1e-12, omega=omega)
    u, hist = t.krylov(method, precond, b, np.zeros(N * N), tol=1e-11, maxit=maxit)
    assert hist[-1] <= 1e-11, hist
    assert np.linalg.norm(u - x) <= 1e-9 * np.linalg.norm(x)
    if expect:
        assert hist.size - 1 <= expect, hist.size
    if method == 0 and precond == 0:
        # unpreconditioned CG on the symmetric interior operator: the recurrence residual equals the true one
        ss, _ = orc.residual(N, W, alpha, 0, u, b, store=False)
        assert abs(np.sqrt(ss / orc.sumsq(N, W, alpha, 0, b)) - hist[-1]) <= 1e-12
